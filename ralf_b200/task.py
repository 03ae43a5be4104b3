"""Constrained-generation tasks on the host side.

Mirrors, for the tasks ``uncond / c / cwh / partial / refinement``:
  * ``get_condition``                      image2layout/train/helpers/task.py:45-183
  * the LayoutFormer++ constraint sequences  models/layoutformerpp/task_preprocessor.py:354-484
  * ``DECODE_SPACE_RESTRICTION``           models/layoutformerpp/decoding_space_restriction.py:5-106
  * ``calculate_violation``                models/layoutformerpp/violate.py:24-139

The reference restricts the decoding space with per-sample python loops and ``.item()`` syncs at every step; every
restriction it implements reduces to "at step i sample b must emit token f" or "is free", so here it becomes ONE int32
table ``forced[B, S]`` (-1 = free) computed before the loop and consumed by the device sampling kernel
(``ralf_sample_next``).  ``relation`` (Gen-R) shares ``get_condition`` / the label restriction with the tasks here; its
relationship table, constraint sequence, per-step relation masks and backtracking sampler live in ``ralf_b200/relation.py``.

RNG contract: the pieces of the reference that draw random numbers on the host (refinement noise, element shuffles)
are drawn here with the same torch calls in the same order, so a seeded run reproduces the reference's constraint
sequences bit for bit (tests/test_task_cpu.py checks this against fixtures dumped from the reference).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Optional

import torch
from torch import Tensor

from .tokenizer import GEO_KEYS, LayoutSequenceTokenizer

REFINEMENT_NOISE_STD = 0.01  # helpers/task.py:16
COND_TYPES = ("c", "cwh", "partial", "refinement", "relation", None, "none", "uncond")
TASK_VARS = {  # helpers/task.py:33-42
    "c": ["label"],
    "cwh": ["label", "width", "height"],
    "refinement": ["label", "width", "height", "center_x", "center_y"],
    "partial": ["label", "width", "height", "center_x", "center_y"],
    "relation": ["label"],
}
TASK_TOKEN = {"c": "label", "cwh": "label_size", "refinement": "refinement", "partial": "completion",
              "uncond": "uncondition", "none": "uncondition", None: "uncondition"}
TASK_TOKENS = ["end_of_task", "label", "label_size", "relationship", "refinement", "completion", "uncondition"]
PREPROCESS_SPECIAL = ["sep", "relation_sep", "canvas"]
N_REL_LOC, N_REL_SIZE = 6, 4  # helpers/relationships.py:11-24
UNCOND = (None, "none", "uncond")


@dataclass
class ConditionalInputs:
    """models/common/base_model.py:17-109 (the discrete-layout + retrieval-augmented container)."""

    image: Tensor
    id: Any = None
    task: Optional[str] = None
    seq: Optional[Tensor] = None
    mask: Optional[Tensor] = None
    seq_observed: Any = None
    edge_indexes: Optional[Tensor] = None     # relation only: [B, P, 2] node pairs, node 0 = canvas (relation.compute_relation)
    edge_attributes: Optional[Tensor] = None  # relation only: [B, P] bit sets 1 << RelSize | 1 << RelLoc
    retrieved: dict = field(default_factory=dict)

    def __post_init__(self) -> None:
        r = self.retrieved
        if r and torch.is_tensor(r.get("image")) and r["image"].size(2) < 4:  # base_model.py:87-100
            r["image"] = torch.cat([r["image"], r["saliency"]], dim=2)

    def to(self, x: Any) -> "ConditionalInputs":
        for name in ("image", "id", "seq", "mask", "edge_indexes", "edge_attributes"):
            v = getattr(self, name)
            if torch.is_tensor(v):
                setattr(self, name, v.to(x))
        self.retrieved = {k: (v.to(x) if torch.is_tensor(v) else v) for k, v in self.retrieved.items()}
        return self


def get_condition(batch: dict, cond_type: Optional[str], tokenizer: LayoutSequenceTokenizer):
    """helpers/task.py:45-183.  Returns (cond, batch); like the reference, ``refinement`` overwrites the geometry of
    ``batch`` with the perturbed values."""
    if cond_type not in COND_TYPES:
        raise AssertionError(f"cond_type={cond_type!r} is not one of {COND_TYPES}")
    image = batch["image"] if batch["image"].size(1) == 4 else torch.cat([batch["image"], batch["saliency"]], dim=1)
    pad_id = tokenizer.name_to_id("pad")
    mask_id = tokenizer.name_to_id("mask") if "mask" in tokenizer.special_tokens else -1
    has_bos = "bos" in tokenizer.special_tokens
    assert has_bos, "the sequence tokenizer of the autoregressive models always carries <bos>"
    enc = tokenizer.encode(batch)
    seq, mask = enc["seq"], enc["mask"]
    B, S = seq.shape
    C = tokenizer.N_var_per_element
    extra: dict = {}
    if cond_type in UNCOND:
        seq, mask = None, None
    elif cond_type == "partial":
        # keep <bos> + the first element (shifted to the front for order-sensitive models): task.py:88-110
        new_seq = torch.full_like(seq, mask_id)
        new_mask = torch.zeros_like(mask)
        n_keep = 1 + C
        new_seq[:, :n_keep] = seq[:, :n_keep]
        new_mask[:, :n_keep] = True
        seq, mask = new_seq, new_mask
    elif cond_type in ("c", "cwh", "relation"):
        if cond_type == "relation":  # drawn before anything else of this branch, like task.py:112-113
            from .relation import compute_relation

            extra.update(compute_relation(batch))
        slot = (torch.arange(S) - 1) % C
        slot[0] = -1
        keep = torch.zeros(S, dtype=torch.bool)
        keep[0] = True
        for name in TASK_VARS[cond_type]:
            keep |= slot == tokenizer.var_order.index(name)
        keep = keep[None].expand(B, S)
        seq = seq.clone()
        seq[~keep] = mask_id
        seq[~mask] = pad_id  # the number of elements is known
        mask = (mask & keep) | ~mask
    elif cond_type == "refinement":
        noisy = {"label": batch["label"], "mask": batch["mask"]}
        for key in GEO_KEYS:  # same draw order as the reference (task.py:141-147)
            noise = torch.normal(0, REFINEMENT_NOISE_STD, size=batch[key].size())
            noisy[key] = torch.clamp(batch[key] + noise, min=0.0, max=1.0)
            noisy[key][~batch["mask"]] = 0.0
            batch[key] = noisy[key].clone()
        seq = tokenizer.encode(noisy)["seq"]
        extra["seq_observed"] = noisy
    try:
        ids = torch.tensor(list(map(int, batch["id"])), dtype=torch.long)
    except Exception:
        ids = batch.get("id")
    retrieved = batch.get("retrieved", {})
    if isinstance(retrieved, list):
        assert len(retrieved) == 1
        retrieved = batch["retrieved"] = retrieved[0]
    return ConditionalInputs(image=image, id=ids, task=cond_type, seq=seq, mask=mask, retrieved=retrieved, **extra), batch


# --------------------------------------------------------------------------------------------------------------------
# constraint sequences for the user-constraint encoder
# --------------------------------------------------------------------------------------------------------------------
class TaskPreprocessor:
    """One class for the reference's Unconditional / Label / LabelSize / Refinement / Partial preprocessors
    (task_preprocessor.py:354-484; ``global_task_embedding=False`` like every shipped config).

    Output: {"seq": [B, L] = <bos> TASK <end_of_task> e0 <sep> e1 ... <eos> <pad>..., "pad_mask": seq == <pad>} where
    e_n are the task's variables of element n (TASK_VARS) and L = 3 + (len(vars) + 1) * max_b(#elements)."""

    def __init__(self, tokenizer: LayoutSequenceTokenizer, task: Optional[str] = "uncond") -> None:
        if task == "relation":
            raise ValueError("the relation task needs its relationship table: use ralf_b200.relation.RelationPreprocessor")
        self.tokenizer = tokenizer
        self.task = task
        self.tokens = TASK_TOKENS + PREPROCESS_SPECIAL + [f"rel_elem_{i}" for i in range(tokenizer.max_seq_length)] + \
            [f"rel_loc_{i}" for i in range(N_REL_LOC)] + [f"rel_size_{i}" for i in range(N_REL_SIZE)]

    @property
    def TASK(self) -> str:
        return TASK_TOKEN[self.task]

    @property
    def N_total(self) -> int:
        return self.tokenizer.N_total + len(self.tokens)

    def name_to_id(self, name: str) -> int:
        if name in self.tokenizer.special_tokens:
            return self.tokenizer.name_to_id(name)
        if name in self.tokens:
            return self.tokens.index(name) + self.tokenizer.N_total
        return self.tokenizer.label_names.index(name)

    def id_to_name(self, i: int) -> str:
        n = self.tokenizer.N_total
        if i >= n:
            return self.tokens[i - n]
        for t in self.tokenizer.special_tokens:
            if self.tokenizer.name_to_id(t) == i:
                return t
        if i < self.tokenizer.N_label:
            return self.tokenizer.label_names[i]
        return str(i)

    def decode_tokens(self, seq: Tensor) -> list:
        return [[self.id_to_name(int(t)) for t in row] for row in seq.tolist()]

    def __call__(self, cond: ConditionalInputs) -> dict:
        tid = self.name_to_id
        dev = cond.image.device
        if self.task in UNCOND:
            B = cond.image.size(0)
            ids = [tid("bos"), tid(self.TASK), tid("end_of_task"), tid("eos")]
            seq = torch.tensor(ids, dtype=torch.long, device=dev)[None].expand(B, -1).contiguous()
            return {"seq": seq, "pad_mask": seq == tid("pad")}
        assert cond.task == self.task, f"task={cond.task!r} does not match the preprocessor ({self.task!r})"
        pad, eos = tid("pad"), tid("eos")
        shuffle = self.task in ("c", "partial")  # Label / Partial preprocessors shuffle the elements (:386-412,:466-484)
        src = cond.seq
        if self.task == "partial":
            assert bool((src[~cond.mask] == -1).all())
            src = src.clone()  # the reference works on a deep copy here
            src[~cond.mask] = pad
        src[src == eos] = pad  # in place on purpose: the reference's parse_seq_into_vars mutates cond.seq (:157)
        B = src.size(0)
        C = self.tokenizer.N_var_per_element
        elems = src[:, 1:].reshape(B, -1, C)  # [B, E, C]
        label_col = self.tokenizer.var_order.index("label")
        counts = (elems[:, :, label_col] != pad).sum(dim=1)
        cols = [self.tokenizer.var_order.index(v) for v in TASK_VARS[self.task]]
        nvar = len(cols)
        if shuffle:
            perms = [torch.randperm(int(c)) for c in counts]  # one draw per sample, in batch order, like the reference
        else:
            perms = [torch.arange(int(c)) for c in counts]
        n_valid = [int(((elems[b, :, label_col] != pad) & (elems[b, :, label_col] != eos)).sum()) for b in range(B)]
        L = 3 + (nvar + 1) * max(n_valid)
        out = torch.full((B, L), pad, dtype=torch.long)
        head = [tid("bos"), tid(self.TASK), tid("end_of_task")]
        for b in range(B):
            e = elems[b].cpu()
            order = perms[b].tolist() + list(range(int(counts[b]), e.size(0)))
            row = list(head)
            for n in range(n_valid[b]):
                if n:
                    row.append(tid("sep"))
                row += [int(e[order[n], c]) for c in cols]
            row.append(eos)
            out[b, :len(row)] = torch.tensor(row, dtype=torch.long)
        out = out.to(dev)
        return {"seq": out, "pad_mask": out == pad}


# --------------------------------------------------------------------------------------------------------------------
# decoding-space restriction as a forced-token table
# --------------------------------------------------------------------------------------------------------------------
def forced_token_table(cond_type: Optional[str], cond_seq: Optional[Tensor], pad_id: int, eos_id: int,
                       max_length: int, n_var: int = 5) -> Optional[Tensor]:
    """int32 [B, max_length]: entry [b, i] is the token sample b must emit at decode step i, or -1 when the step is free.

    restrict_reliable_label_or_size (c, cwh; :5-39) and restrict_only_category (refinement; :42-84) both do, for
    sampling index s = i + 1 into ``cond_seq`` [B, max_length + 1]:
        s <  first <pad> position of the row:  token given and not <pad>/-1 -> only that token;  else free
        s >= first <pad> position            :  only <eos>
    (refinement and relation: label slots only).  ``partial`` teacher-forces tokens 1..5 of ``cond_seq`` and starts at step 5
    (retrieval_augmented_autoreg.py:257-259), i.e. the same thing as forcing steps 0..4."""
    if cond_type in UNCOND:
        return None
    B, S1 = cond_seq.shape
    assert S1 == max_length + 1
    seq = cond_seq.to(torch.long).cpu()
    forced = torch.full((B, max_length), -1, dtype=torch.int32)
    if cond_type == "partial":
        forced[:, :n_var] = seq[:, 1:1 + n_var].to(torch.int32)
        return forced
    if cond_type not in ("c", "cwh", "refinement", "relation"):
        raise NotImplementedError(f"cond_type={cond_type!r}")
    is_pad = seq == pad_id
    first_pad = torch.where(is_pad.any(dim=1), is_pad.float().argmax(dim=1), torch.full((B,), S1 + 1))
    s = torch.arange(1, S1)[None]  # sampling indices
    given = seq[:, 1:]
    before = s < first_pad[:, None]
    f = torch.where(before, torch.where((given == pad_id) | (given == -1), torch.full_like(given, -1), given),
                    torch.full_like(given, eos_id))
    if cond_type in ("refinement", "relation"):  # restrict_only_category: label slots only (:42-84, :97-105)
        f = torch.where(((s - 1) % n_var == 0).expand_as(f), f, torch.full_like(f, -1))
    return f.to(torch.int32)


def calculate_violation(cond_type: Optional[str], cond: ConditionalInputs, out_seq: Tensor,
                        tokenizer: LayoutSequenceTokenizer, output: Optional[dict] = None,
                        prepared_rel_constraints: Optional[list] = None) -> dict:
    """violate.py:24-139: how many given tokens the output failed to reproduce; ``relation`` counts the relationships the
    decoded layout ``output`` breaks (violate.py:142-236)."""
    if cond_type == "relation":
        from .relation import violation_count

        assert len(prepared_rel_constraints) == cond.seq.size(0)
        return violation_count(output, prepared_rel_constraints)
    if cond_type in UNCOND or cond_type == "partial":
        return {"total": 1, "viorated": 0}
    pad_id, eos_id = tokenizer.name_to_id("pad"), tokenizer.name_to_id("eos")
    total = bad = 0
    given_all, mask_all = cond.seq[:, 1:].cpu(), cond.mask[:, 1:].cpu()
    out_seq = out_seq.cpu()
    for b in range(given_all.size(0)):
        m = mask_all[b]
        g = given_all[b][m]
        g = g[(g != pad_id) & (g != eos_id)]
        if cond_type == "refinement":
            o = out_seq[b][:g.size(0)][::5]
            g = g[::5]
        else:
            o = out_seq[b][m]
            o = o[(o != pad_id) & (o != eos_id)]
        assert g.size(0) == o.size(0), "diff_elems should be 0"
        bad += int((g != o).sum())
        total += g.size(0)
    return {"total": total, "viorated": bad}
