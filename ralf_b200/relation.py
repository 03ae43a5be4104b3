"""Gen-R: generation conditioned on pairwise element relationships, host side (its last item).

What the reference does for ``cond_type="relation"`` and where:
  * relation vocabulary / detectors / sparse edge tensors   image2layout/train/helpers/relationships.py:11-166
  * the per-id relationship table (offline)                  image2layout/preprocess/precompute_relationship.py:31-131
  * constraint sequence of the user-constraint encoder       models/layoutformerpp/task_preprocessor.py:488-590
  * per-step decoding-space restriction + backtrack target   models/layoutformerpp/relation_restriction.py:354-825
  * backtracking sampler                                      models/retrieval_augmented_autoreg.py:335-507
  * violation count                                           models/layoutformerpp/violate.py:142-236

The reference keeps one deep-copied decode state per position and slices that list when it backtracks.  Every field of
that state is a function of the token prefix alone, so here the restriction is a pure function
``RelationConstraint.mask(prefix, constraints)``: the element boxes are re-read from the prefix (at most 50 tokens) on
every call and a rewind needs no bookkeeping.  The admissible bins of a variable are integer intervals; they are kept
as ``(lo, hi)`` pairs and intersected arithmetically instead of through python sets.  Everything else (which token the
model is asked next, how the intervals follow from the target box, the order of the host RNG draws) is the reference's
behaviour, quirks included -- tests/test_relation_cpu.py pins it to fixtures dumped from the reference.

Model arithmetic stays on the GPU: ``sample_with_backtracking`` is handed a ``logits_of(prefix)`` callable, which in the
product is the KV-cached ``engine.DecodeSession`` (rewind = overwrite the cache rows), and only looks at one row of V
logits per step on the host, like the reference does.
"""
from __future__ import annotations

import io
import pickle
import random
from enum import IntEnum
from math import ceil, floor
from typing import Any, Callable, Optional, Sequence

import torch
from torch import Tensor

from .task import ConditionalInputs, TaskPreprocessor
from .tokenizer import GEO_KEYS, LayoutSequenceTokenizer


# The integer values are wire format: they fix the order of the relation tokens in the constraint vocabulary
# (task_preprocessor.py:36-38,119-124) and they are what the reference pickles into its relationship table.
class RelSize(IntEnum):
    UNKNOWN = 0
    SMALLER = 1
    EQUAL = 2
    LARGER = 3


class RelLoc(IntEnum):
    UNKNOWN = 4
    LEFT = 5
    TOP = 6
    RIGHT = 7
    BOTTOM = 8
    CENTER = 9


RelElement = IntEnum("RelElement", {chr(ord("A") + i): 10 + i for i in range(11)}, module=__name__)  # A = 10 ... K = 20

MIRRORED = {RelLoc.LEFT: RelLoc.RIGHT, RelLoc.RIGHT: RelLoc.LEFT, RelLoc.TOP: RelLoc.BOTTOM, RelLoc.BOTTOM: RelLoc.TOP,
            RelLoc.CENTER: RelLoc.CENTER, RelLoc.UNKNOWN: RelLoc.UNKNOWN, RelSize.SMALLER: RelSize.LARGER,
            RelSize.LARGER: RelSize.SMALLER, RelSize.EQUAL: RelSize.EQUAL, RelSize.UNKNOWN: RelSize.UNKNOWN}
SIZE_ALPHA = 0.1   # helpers/relationships.py:55
EDGE_RATIO = 0.1   # helpers/task.py:17
CANVAS = "canvas"
REFERENCE_ENUM_MODULE = "image2layout.train.helpers.relationships"
REFERENCE_TABLE_NAME = "pku_cgl_relationships_dic_using_canvas_sort_label_lexico.pt"


# --------------------------------------------------------------------------------------------------------------------
# detectors: "how does box 2 sit relative to box 1"; boxes are (cx, cy, w, h).  The arithmetic runs in whatever type the
# caller passes (0-d fp32 tensors in compute_relation, python floats in the table builder), as in the reference.
# --------------------------------------------------------------------------------------------------------------------
def size_relation(b1: Sequence, b2: Sequence) -> RelSize:
    a1, a2 = b1[2] * b1[3], b2[2] * b2[3]
    if (1 - SIZE_ALPHA) * a1 < a2 < (1 + SIZE_ALPHA) * a1:
        return RelSize.EQUAL
    return RelSize.LARGER if a1 < a2 else RelSize.SMALLER


def _edges(b: Sequence):
    cx, cy, w, h = b
    return cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2  # left, top, right, bottom


def loc_relation(b1: Sequence, b2: Sequence) -> RelLoc:
    l1, t1, r1, bo1 = _edges(b1)
    l2, t2, r2, bo2 = _edges(b2)
    for hit, rel in ((bo2 <= t1, RelLoc.TOP), (bo1 <= t2, RelLoc.BOTTOM), (r2 <= l1, RelLoc.LEFT), (r1 <= l2, RelLoc.RIGHT)):
        if hit:
            return rel
    return RelLoc.CENTER  # the boxes overlap


def canvas_relation(b: Sequence) -> RelLoc:
    cy = b[1]
    return RelLoc.TOP if cy < 1.0 / 3 else (RelLoc.CENTER if cy < 2.0 / 3 else RelLoc.BOTTOM)


def compute_relation(batch: dict, edge_ratio: float = EDGE_RATIO) -> dict:
    """helpers/relationships.py:110-166: a random ``edge_ratio`` subset of the pairwise relations as dense tensors
    ``edge_indexes`` [B, P, 2] (-1 padded) and ``edge_attributes`` [B, P] (bit set ``1 << size | 1 << loc``), node 0 being
    the canvas.  One ``random.random()`` draw per valid pair in (i, j) lexicographic order, like the reference."""
    B, S = batch["label"].shape
    lead = {"center_x": 0.5, "center_y": 0.5, "width": 1.0, "height": 1.0}
    geo = {k: torch.cat([torch.full((B, 1), v), batch[k]], dim=1) for k, v in lead.items()}
    count = batch["mask"].sum(dim=1) + 1  # + the canvas node
    unknown = (1 << RelSize.UNKNOWN) | (1 << RelLoc.UNKNOWN)
    P = (S + 1) * (S + 2) // 2
    index = torch.full((B, P, 2), -1, dtype=torch.long)
    attr = torch.full((B, P), unknown, dtype=torch.long)
    for b in range(B):
        n, filled = int(count[b]), 0
        for i in range(min(n, S + 1)):
            for j in range(i + 1, min(n, S + 1)):
                if random.random() > edge_ratio:
                    continue
                bi = [geo[k][b][i] for k in GEO_KEYS]
                bj = [geo[k][b][j] for k in GEO_KEYS]
                loc = canvas_relation(bj) if i == 0 else loc_relation(bi, bj)
                index[b, filled, 0], index[b, filled, 1] = i, j
                attr[b, filled] = (1 << size_relation(bi, bj)) | (1 << loc)
                filled += 1
    return {"edge_indexes": index, "edge_attributes": attr}


# --------------------------------------------------------------------------------------------------------------------
# the relationship table: id -> rows [label name, RelElement, relation, label name | "canvas", RelElement | "pad"]
# --------------------------------------------------------------------------------------------------------------------
def describe_relationships(batch: dict, label_names: Sequence[str]) -> dict:
    """preprocess/precompute_relationship.py:56-131 for one collated batch.  Elements are named (label, ordinal among the
    same label); element pairs are visited from the last valid element backwards; per id the rows are
    [pairwise locations..., pairwise sizes..., element-vs-canvas locations...].
    ``label_names[label id]`` is the name written into the rows (the reference script indexes its own per-dataset list,
    precompute_relationship.py:24-27)."""
    letters = list(RelElement)
    table = {}
    for b in range(batch["label"].size(0)):
        seen: dict = {}
        tag = []
        for lab, ok in zip(batch["label"][b].tolist(), batch["mask"][b].tolist()):
            if not ok:
                tag.append(None)
                continue
            seen[lab] = seen.get(lab, 0) + 1
            tag.append([label_names[lab], letters[seen[lab] - 1]])
        order = [i for i, ok in enumerate(batch["mask"][b].tolist()) if ok][::-1]
        box = {i: [batch[k][b, i].item() for k in GEO_KEYS] for i in order}
        loc_rows, size_rows, canvas_rows = [], [], []
        for n, i in enumerate(order):
            for j in order[n + 1:]:
                loc_rows.append([*tag[i], loc_relation(box[i], box[j]), *tag[j]])
                size_rows.append([*tag[i], size_relation(box[i], box[j]), *tag[j]])
            canvas_rows.append([*tag[i], canvas_relation(box[i]), CANVAS, "pad"])
        key = batch["id"][b]
        table[key.item() if torch.is_tensor(key) else key] = loc_rows + size_rows + canvas_rows
    return table


class _ReferenceEnums(pickle.Unpickler):
    """Reads a table pickled by the reference: its enum classes resolve to the ones above (same names and values)."""

    def find_class(self, module: str, name: str):
        if module == REFERENCE_ENUM_MODULE and name in ("RelSize", "RelLoc", "RelElement"):
            return globals()[name]
        return super().find_class(module, name)


class _PickleShim:
    __name__ = "ralf_b200.relation._PickleShim"
    Unpickler = _ReferenceEnums

    @staticmethod
    def load(f, **kw):
        return _ReferenceEnums(f, **kw).load()


def load_relation_table(path: str) -> dict:
    """``torch.load`` of the reference's ``cache/pku_cgl_relationships_dic_using_canvas_sort_label_lexico.pt``
    (task_preprocessor.py:498-506) without the reference package on the path; keys are normalised to ``str(id)``."""
    table = torch.load(path, pickle_module=_PickleShim, weights_only=False)
    return {str(k): v for k, v in table.items()}


def save_relation_table(path: str, table: dict) -> None:
    """Written with this module's enum classes (``load_relation_table`` reads it back; the reference cannot)."""
    buf = io.BytesIO()
    torch.save({str(k): v for k, v in table.items()}, buf)
    with open(path, "wb") as f:
        f.write(buf.getvalue())


# --------------------------------------------------------------------------------------------------------------------
# constraint sequence for the user-constraint encoder
# --------------------------------------------------------------------------------------------------------------------
class RelationPreprocessor(TaskPreprocessor):
    """task_preprocessor.py:488-590:
        <bos> relationship <end_of_task> l0 <sep> l1 ... <relation_sep> r0 <sep> r1 <sep> ... <eos> <pad>...
    where l_n are the (shuffled) element labels and r_m = [label, ordinal, relation, label | canvas, ordinal | pad] is a
    ``relation_size`` % sample of the id's table rows.  Host RNG draws, in the reference's order: construction shuffles
    every table entry (``random.sample``); a call draws one ``torch.randperm`` per sample (an element shuffle whose result
    the reference discards), the label sequence's own ``torch.randperm`` per sample, then one ``random.sample`` per sample
    that has relations."""

    def __init__(self, tokenizer: LayoutSequenceTokenizer, table: Any, relation_size: int = 10) -> None:
        super().__init__(tokenizer, "uncond")
        self.task = "relation"
        if isinstance(table, str):
            table = load_relation_table(table)
        self.table = {str(k): random.sample(v, len(v)) for k, v in table.items()}
        self.relation_size = relation_size
        self.labels = TaskPreprocessor(tokenizer, "c")
        loc, size = list(RelLoc), list(RelSize)
        self._enum_name = {**{e: f"rel_elem_{i}" for i, e in enumerate(RelElement)},
                           **{e: f"rel_loc_{i}" for i, e in enumerate(loc)},
                           **{e: f"rel_size_{i}" for i, e in enumerate(size)}}
        self._enum_of = {self.name_to_id(n): e for e, n in self._enum_name.items() if n in self.tokens}

    @property
    def TASK(self) -> str:
        return "relationship"

    def set_relation_size(self, relation_size: int) -> None:
        self.relation_size = relation_size

    def token_id(self, item: Any) -> int:
        return self.name_to_id(self._enum_name[item] if isinstance(item, IntEnum) else item)

    def token_of(self, i: int) -> Any:
        """Inverse of ``token_id``: enum member for relation tokens, name otherwise."""
        return self._enum_of[i] if i in self._enum_of else self.id_to_name(i)

    def decode_tokens(self, seq: Tensor) -> list:
        return [[self.token_of(int(t)) for t in row] for row in seq.tolist()]

    def __call__(self, cond: ConditionalInputs) -> dict:
        ids = cond.id.cpu().tolist() if torch.is_tensor(cond.id) else cond.id
        rows = [self.table[str(i)] for i in ids]
        pad, eos = self.name_to_id("pad"), self.name_to_id("eos")
        seq = cond.seq
        seq[seq == eos] = pad  # in place, like parse_seq_into_vars (:157)
        C = self.tokenizer.N_var_per_element
        first_var = seq[:, 1:].reshape(seq.size(0), -1, C)[:, :, 0]
        for n in (first_var != pad).sum(dim=1):
            torch.randperm(int(n))  # the reference shuffles a copy here and only keeps its shape statistics (:534-537)
        cond_task, cond.task = cond.task, "c"
        try:
            lab = self.labels(cond)
        finally:
            cond.task = cond_task
        lab_seq, lab_pad = lab["seq"].cpu(), lab["pad_mask"].cpu()
        lab_seq[:, 1] = self.name_to_id(self.TASK)
        lab_seq[lab_seq == eos] = self.name_to_id("relation_sep")
        out, width = [], -1
        sep = self.name_to_id("sep")
        for b, table_rows in enumerate(rows):
            head = lab_seq[b][~lab_pad[b]].tolist()
            if not table_rows:
                out.append(head + [eos])  # NB the reference leaves such rows out of the width computation (:556-560)
                continue
            picked = random.sample(table_rows, max(len(table_rows) * self.relation_size // 100, 1))
            tail = []
            for r in picked:
                tail += [self.token_id(x) for x in r] + [sep]
            tail[-1] = eos
            out.append(head + tail)
            width = max(width, len(out[-1]))
        full = torch.full((len(out), width), pad, dtype=torch.long)
        for b, r in enumerate(out):
            full[b, :len(r)] = torch.tensor(r, dtype=torch.long)  # a relation-free row longer than `width` raises, as there
        full = full.to(cond.image.device)
        return {"seq": full, "pad_mask": full == pad}


# --------------------------------------------------------------------------------------------------------------------
# decoding-space restriction
# --------------------------------------------------------------------------------------------------------------------
WIDTH, HEIGHT, CX, CY = 1, 2, 3, 4  # slot of a token inside its element (slot 0 = label)
FULL = None  # "no restriction from this relation"


class RelationConstraint:
    """relation_restriction.py:354-825 (TransformerSortByDictRelationConstraint) as a function of the token prefix.

    ``prepare(const_seq_row)`` parses one row of the RelationPreprocessor output into per-element constraint lists
    (the reference's format: ``("canvas", RelLoc)`` or ``(relation, earlier element index)``, relations re-expressed from
    the later element's point of view).  ``mask(prefix, constraints)`` -> (forbidden [V] bool, back_idx): which tokens
    the next position may NOT take, and the position the sampler should rewind to when nothing is left."""

    def __init__(self, preprocessor: RelationPreprocessor) -> None:
        self.pre = preprocessor
        tok = preprocessor.tokenizer
        self.nbin = int(tok.N_bbox_per_var)
        self.canvas = self.nbin - 1  # bins are 0..127
        self.token_mask = tok.token_mask.clone().bool()
        self.V = self.token_mask.size(-1)
        self.start = {}
        for slot in (WIDTH, HEIGHT, CX, CY):
            allowed = self.token_mask[slot].nonzero().flatten()
            self.start[slot] = int(allowed[0])
            assert int(allowed[-3]) + 1 - self.start[slot] == self.nbin  # the bins, then two special tokens
        self.types: Tensor = torch.zeros(0, dtype=torch.long)

    # ---- parsing ------------------------------------------------------------------------------------------------
    def prepare(self, seq: Tensor) -> list:
        pre = self.pre
        seq = seq.cpu()
        eos_at = int(torch.argmax((seq == pre.name_to_id("eos")).float()))
        rel_at = int(torch.argmax((seq == pre.name_to_id("relation_sep")).float()))
        body = seq[:eos_at]
        types = body[3:rel_at][::2]  # <bos> task <end_of_task> l0 <sep> l1 ...
        self.types = types
        rel = body[rel_at + 1:]
        rel = rel[rel != pre.name_to_id("sep")].reshape(-1, 5).tolist()
        ordinal = {e: i for i, e in enumerate(RelElement)}

        def position(label: int, letter: int) -> int:
            return int((types == label).nonzero()[ordinal[pre.token_of(letter)]])

        cons: list = [[] for _ in range(types.size(0))]
        canvas_id = pre.name_to_id(CANVAS)
        for la, ia, r, lb, ib in rel:
            kind = pre.token_of(r)
            pa = position(la, ia)
            if lb == canvas_id:
                cons[pa].append((CANVAS, kind))
                continue
            pb = position(lb, ib)
            if pb > pa:  # store at the later element, seen from there
                pa, pb, kind = pb, pa, MIRRORED[kind]
            assert pa > pb, f"{pa=} {pb=} {kind=}"
            cons[pa].append((kind, pb))
        return cons

    # ---- admissible intervals -----------------------------------------------------------------------------------
    def _canvas_band(self, rel: RelLoc, h: int):
        c, half = self.canvas, h / 2
        if rel == RelLoc.TOP:
            return ceil(half), floor(c / 3 - half)
        if rel == RelLoc.CENTER:
            return ceil(1 * c / 3 + half), floor(2 * c / 3 - half)
        if rel == RelLoc.BOTTOM:
            return ceil(2 * c / 3 + half), floor(c - half)
        raise ValueError(f"Unknown rel_type: {rel}")

    def _interval(self, slot: int, rel: IntEnum, cur: list, tgt: list):
        """Bins [lo, hi) the variable ``slot`` of the current element may take so that the target box ``tgt`` =
        [w, h, cx, cy] (an earlier element) stands in relation ``rel`` to it; ``cur`` = the current element's bins so far."""
        c, n, a = self.canvas, self.nbin, SIZE_ALPHA
        tw, th, tcx, tcy = tgt
        if slot == CX:
            w = cur[0]
            if rel == RelLoc.LEFT:
                return floor(tcx + tw / 2 + w / 2), ceil(c - w / 2)
            if rel == RelLoc.RIGHT:
                return floor(w / 2), ceil(tcx - tw / 2 - w / 2)
            if rel == RelLoc.CENTER:
                return ceil(tcx - tw / 2 + w / 2), floor(tcx + tw / 2 - w / 2)
            return floor(w / 2), ceil(c - w / 2)
        if slot == CY:
            half = cur[1] / 2
            if rel == RelLoc.TOP:
                return floor(tcy + th / 2 + half), ceil(c - half)
            if rel == RelLoc.BOTTOM:
                return floor(half), ceil(tcy - th / 2 - half)
            if rel == RelLoc.CENTER:
                return ceil(tcy - th / 2 - half), floor(tcy + th / 2 + half)
            return floor(cur[1] / 2), ceil(c - half)
        area = tw * th
        if slot == WIDTH:
            if rel == RelLoc.LEFT:
                return 0, ceil(c - tcx - tw / 2)
            if rel == RelLoc.RIGHT:
                return 0, ceil(tcx - tw / 2)
            if rel == RelLoc.CENTER:
                return 0, floor(c - tcx + tw / 2) if tcx < n // 2 else floor(tcx + tw / 2)
            if rel == RelSize.SMALLER:
                area /= 1 - a
                return min(ceil(area / c), c), ceil(area)
            if rel == RelSize.LARGER:
                area /= 1 + a
                return 0, floor(area / c)
            if rel == RelSize.EQUAL:
                return floor(area / (1 + a) / c), ceil(area / (1 - a) / c)
            return FULL
        assert slot == HEIGHT
        w = cur[0]
        if rel == RelLoc.TOP:
            return 0, ceil(tcy - th / 2)
        if rel == RelLoc.BOTTOM:
            return 0, floor(tcy - th / 2)
        if rel == RelLoc.CENTER:
            return 0, floor(c - tcy + th / 2) if tcy < n // 2 else floor(tcy + th / 2)
        if rel == RelSize.SMALLER:
            area /= 1 - a
            return (c if w == 0 else min(ceil(area / w), c)), n
        if rel == RelSize.LARGER:
            area /= 1 + a
            return 0, (n if w == 0 else min(floor(area / w), n))
        if rel == RelSize.EQUAL:
            w = 1 if w == 0 else w
            return floor(area / (1 + a) / w), ceil(area / (1 - a) / w)
        return FULL

    # ---- the restriction ----------------------------------------------------------------------------------------
    def mask(self, prefix: Sequence[int], cons: list):
        """``prefix`` = <bos> + the tokens emitted so far.  Returns (forbidden, back_idx)."""
        s = len(prefix) - 1
        slot, k = s % 5, s // 5
        forbidden = torch.ones(self.V, dtype=torch.bool)
        n_elem = len(cons)
        if slot == 0:
            if k >= 1 and k == n_elem:  # every element is complete: only <eos>
                forbidden[self.pre.tokenizer.name_to_id("eos")] = False
            else:
                forbidden[int(self.types[k])] = False
            return forbidden, None
        body = prefix[1:]
        boxes = []  # per started element: bins [w, h, cx, cy] read so far
        for e in range(k + 1):
            geo = body[5 * e + 1:5 * e + 5]
            boxes.append([int(t) - self.start[i + 1] for i, t in enumerate(geo)])
        cur, mine = boxes[k], cons[k]
        lo, hi, back = 0, self.nbin, None

        def narrow(span) -> None:
            nonlocal lo, hi
            if span is FULL or span[0] >= span[1]:
                return  # an empty interval does not restrict (the reference's _intersect returns the other operand)
            lo, hi = max(lo, span[0]), min(hi, span[1])

        if k == 0:  # the first element can only be tied to the canvas, and only its cy is restricted
            for kind, tgt in mine:
                if slot == CY and kind == CANVAS:
                    narrow(self._canvas_band(tgt, cur[1]))
        elif not mine:
            return ~self.token_mask[s], None
        else:
            for kind, tgt in mine:
                if kind == CANVAS:
                    back = None
                    if slot == CY:
                        narrow(self._canvas_band(tgt, cur[1]))
                    continue
                back = tgt * 5 + slot
                narrow(self._interval(slot, kind, cur, boxes[tgt]))
        if lo < hi:
            forbidden[self.start[slot] + lo:self.start[slot] + hi] = False
        return forbidden, back


# --------------------------------------------------------------------------------------------------------------------
# sampling
# --------------------------------------------------------------------------------------------------------------------
NEG_INF = -float("inf")


def _cfg(cfg: Any, key: str, default: Any = None) -> Any:
    if cfg is None:
        return default
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


def draw_token(logits: Tensor, sampling_cfg: Any, temperature: Optional[float] = None) -> int:
    """helpers/sampling.py:18-68 for one row of host logits [V] (the backtracking loop is per sample and per step on
    the host in the reference too).  Same torch calls, so a seeded CPU generator reproduces the reference's draws."""
    name = _cfg(sampling_cfg, "name", "deterministic") or "deterministic"
    x = logits.view(1, -1)
    if name == "deterministic":
        return int(torch.argmax(x, dim=1))
    t = _cfg(sampling_cfg, "temperature", 1.0) if temperature is None else temperature
    y = x / t
    if name == "top_k":
        kth = torch.topk(y, int(_cfg(sampling_cfg, "top_k", 5)), dim=1).values[:, -1:]
        y = y.masked_fill(y < kth, NEG_INF)
    elif name == "top_p":
        p = float(_cfg(sampling_cfg, "top_p", 0.9))
        assert 0.0 < p <= 1.0
        srt, order = torch.sort(y, descending=True, dim=1)
        cum = torch.cumsum(torch.softmax(srt, dim=1), dim=1)
        drop = cum > p
        drop[:, 0] = False  # the most likely token always stays
        srt = srt.masked_fill(drop, NEG_INF)
        y = srt.gather(dim=1, index=order.argsort(dim=1))
    elif name == "gumbel":
        u = torch.rand_like(y)
        y = y + -torch.log(-torch.log(u + 1e-30) + 1e-30)
    elif name != "random":
        raise NotImplementedError(name)
    return int(torch.multinomial(torch.softmax(y, dim=1), num_samples=1))


def sample_with_backtracking(logits_of: Callable[[list], Tensor], constraint: RelationConstraint, cons: list,
                             forced_row: Tensor, *, bos_id: int, eos_id: int, max_token_length: int, sampling_cfg: Any,
                             prob_gate: float = 0.3, max_backtracks: int = 30, max_resets: int = 3) -> list:
    """One canvas of retrieval_augmented_autoreg.py:335-470.  ``logits_of(prefix)`` -> host fp32 [V] logits of the next
    token; ``forced_row`` int [max_token_length] = this canvas' row of ``task.forced_token_table("relation", ...)``
    (the label slots).  Returns the prefix including <bos> (and the final <eos> when one was drawn).

    Per step: position mask, label restriction, relation mask; if nothing (or, outside a retry, nothing scoring at
    least ``prob_gate``) is left, rewind to the constraint's target (or a random earlier position once the same step has
    failed three times), at most ``max_backtracks`` times before starting over; after ``max_resets`` restarts the
    relation mask is dropped.  The token after a rewind is drawn at temperature 1.5."""
    token_mask = constraint.token_mask
    prefix = [bos_id]
    failed_at: list = []
    retrying, backtracks, resets, idx = False, 0, 0, 0
    while True:
        s = len(prefix) - 1
        logits = logits_of(prefix).detach().to(torch.float32).cpu().clone()
        logits[~token_mask[s]] = NEG_INF
        f = int(forced_row[s]) if s < len(forced_row) else -1
        if f >= 0:
            keep = logits[f].clone()
            logits[:] = NEG_INF
            logits[f] = keep
        raw = logits.clone()
        forbidden, back_idx = constraint.mask(prefix, cons)
        assert bool(token_mask[s][~forbidden].all())
        logits[forbidden] = NEG_INF
        gated_best = logits[logits >= prob_gate].max().item() if bool((logits >= prob_gate).any()) else NEG_INF
        if resets > max_resets:
            logits, retrying = raw, False
        elif (not retrying and gated_best == NEG_INF) or logits.max().item() == NEG_INF:
            failed_at.append(idx)
            retrying = True
            if back_idx is not None and failed_at.count(idx) < 3:
                idx = back_idx
            else:
                idx = random.randint(2, max(2, idx - 1))
            prefix = prefix[:idx]
            backtracks += 1
            if backtracks > max_backtracks:
                failed_at, retrying, backtracks = [], False, 0
                resets += 1
                prefix, idx = [bos_id], 0
            continue
        temperature = None
        if retrying:
            retrying, temperature = False, 1.5
        token = draw_token(logits, sampling_cfg, temperature)
        prefix.append(token)
        if token == eos_id or len(prefix) == max_token_length + 1:
            return prefix
        idx += 1


def pad_like_reference(prefixes: list, max_token_length: int) -> Tensor:
    """Sampler outputs (each <bos> + tokens, ragged) -> int64 [B, max_token_length] without <bos>.  The reference
    right-pads every row to max_token_length + 2 with ``torch.full(..., fill_value=True).type_as(tokens)``, i.e. with the
    VALUE 1, and strips the first and last column (:472-485); rows that ended early therefore end in <eos> 1 1 1 ...,
    which the tokenizer cuts at <eos>."""
    seq = torch.ones((len(prefixes), max_token_length + 2), dtype=torch.long)
    for b, row in enumerate(prefixes):
        seq[b, :len(row)] = torch.tensor(row, dtype=torch.long)
    return seq[:, 1:-1].contiguous()


def violation_count(out: dict, prepared: list) -> dict:
    """violate.py:142-236: a relation counts as violated when the detector, run on the generated boxes, disagrees."""
    total = bad = 0
    for b, cons in enumerate(prepared):
        def box(i: int) -> list:
            return [out[k][b, i].item() for k in GEO_KEYS]

        for i, mine in enumerate(cons):
            for kind, tgt in mine:
                total += 1
                if kind == CANVAS:
                    bad += canvas_relation(box(i)) != tgt
                elif isinstance(kind, RelSize):
                    bad += size_relation(box(i), box(tgt)) != kind
                else:
                    bad += loc_relation(box(i), box(tgt)) != kind
    return {"total": total, "viorated": int(bad)}
