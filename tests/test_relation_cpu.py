"""cond_type="relation" (Gen-R) host logic against tests/golden/relation_cgl_128.npz, which holds what the UNMODIFIED
reference produced under recorded `random` / torch seeds (tests/golden/make_golden.py:run_relation): relationship table,
compute_relation edges, constraint sequences, prepare(), every per-step relation mask + backtrack index, the tokens the
backtracking sampler ended with, decoded layouts and violation counts.  The model side of the sampler (next-token logits
of a prefix) is the CPU oracle here; on the GPU it is engine.DecodeSession (tests/test_tasks_gpu.py)."""
import copy
import os
import random

import numpy as np
import pytest
import torch

from ralf_b200 import relation as R
from ralf_b200 import task as T
from tests import helpers

GEO = ["center_x", "center_y", "width", "height"]


@pytest.fixture(scope="module")
def fx():
    z, meta = helpers.load_golden("relation_cgl_128")
    tok = helpers.make_tokenizer("cgl")
    batch = {k: torch.from_numpy(z[k]) for k in ["label", "mask", *GEO]}
    B = meta["B"]
    batch["id"] = [str(i) for i in range(B)]
    batch["image"] = torch.zeros(B, 3, 8, 8)
    batch["saliency"] = torch.zeros(B, 1, 8, 8)
    table = R.describe_relationships(batch, meta["label_names"])
    return z, meta, tok, batch, table


def _preprocessor(tok, table, meta):
    random.seed(meta["ctor_seed"])
    return R.RelationPreprocessor(tok, copy.deepcopy(table))


def _rows(pre, rows):
    return np.array([[pre.token_id(x) for x in r] for r in rows], dtype=np.int64).reshape(-1, 5)


def _decode_constraints(arr):
    n = int(arr[:, 0].max()) + 1 if len(arr) else 0
    cons = [[] for _ in range(n)]
    for e, kind, tgt in arr.tolist():
        if kind == -1:
            cons[e].append((R.CANVAS, R.RelLoc(tgt)))
        else:
            cons[e].append((R.RelSize(kind) if kind < 4 else R.RelLoc(kind), tgt))
    return cons


def _encode_constraints(cons):
    rows = []
    for e, mine in enumerate(cons):
        for kind, tgt in mine:
            rows.append([e, -1, int(tgt)] if kind == R.CANVAS else [e, int(kind), int(tgt)])
    return np.array(rows, dtype=np.int64).reshape(-1, 3)


def test_enums_are_the_reference_wire_values():
    assert [int(x) for x in R.RelSize] == [0, 1, 2, 3] and [x.name for x in R.RelSize] == ["UNKNOWN", "SMALLER", "EQUAL", "LARGER"]
    assert [int(x) for x in R.RelLoc] == [4, 5, 6, 7, 8, 9]
    assert [x.name for x in R.RelLoc] == ["UNKNOWN", "LEFT", "TOP", "RIGHT", "BOTTOM", "CENTER"]
    assert [int(x) for x in R.RelElement] == list(range(10, 21)) and R.RelElement.A.name == "A" and R.RelElement.K.name == "K"


def test_relationship_table_matches_reference(fx):
    z, meta, tok, batch, table = fx
    pre = _preprocessor(tok, table, meta)
    assert list(table) == [str(i) for i in range(meta["B"])]
    for key, rows in table.items():
        np.testing.assert_array_equal(_rows(pre, rows), z[f"table_{key}"])
    for key, rows in pre.table.items():  # the constructor's shuffle draws from `random` like the reference's
        np.testing.assert_array_equal(_rows(pre, rows), z[f"table_shuffled_{key}"])


def test_reference_pickle_loads_without_the_reference_package(fx, tmp_path):
    z, meta, tok, batch, table = fx
    loaded = R.load_relation_table(os.path.join(helpers.GOLDEN, "relation_table_reference_pickle.pt"))
    assert loaded == table
    assert type(loaded["0"][0][1]) is R.RelElement
    p = str(tmp_path / "t.pt")
    R.save_relation_table(p, table)
    assert R.load_relation_table(p) == table


def test_get_condition_relation_matches_reference(fx):
    z, meta, tok, batch, table = fx
    random.seed(meta["rng_seed"]["deterministic"])
    torch.manual_seed(meta["rng_seed"]["deterministic"])
    cond, _ = T.get_condition(copy.deepcopy(batch), "relation", tok)
    np.testing.assert_array_equal(cond.seq.numpy(), z["cond_seq"])
    np.testing.assert_array_equal(cond.mask.numpy(), z["cond_mask"])
    np.testing.assert_array_equal(cond.edge_indexes.numpy(), z["edge_indexes"])
    np.testing.assert_array_equal(cond.edge_attributes.numpy(), z["edge_attributes"])
    assert cond.task == "relation"


def _oracle_logits_fn(sd, memory_row, pad_id):
    from oracle import ralf_oracle as O

    mem = memory_row[None]

    def logits_of(prefix):
        tgt = torch.tensor([prefix], dtype=torch.long)
        with torch.no_grad():
            return O.decoder_logits(sd, tgt, mem, tgt == pad_id)[0, -1]

    return logits_of


@pytest.mark.parametrize("mode", ["deterministic", "random"])
def test_backtracking_sampler_matches_reference(fx, mode):
    """Same seeds -> same constraint sequence (element shuffles + relation sample), same prepare(), same mask and
    backtrack index at EVERY call the reference made (so the rewinds took the same path), same final tokens."""
    z, meta, tok, batch, table = fx
    torch.set_num_threads(8)
    pre = _preprocessor(tok, table, meta)
    seed = meta["rng_seed"][mode]
    random.seed(seed)
    torch.manual_seed(seed)
    cond, _ = T.get_condition(copy.deepcopy(batch), "relation", tok)
    const = pre(cond)
    p = f"bt_{mode}_"
    np.testing.assert_array_equal(const["seq"].numpy(), z[p + "const_seq"])
    np.testing.assert_array_equal(const["pad_mask"].numpy(), z[p + "const_pad_mask"])
    np.testing.assert_array_equal(cond.seq.numpy(), z[p + "cond_seq_after"])  # <eos> -> <pad> rewritten in place
    sp = meta["special"]
    forced = T.forced_token_table("relation", cond.seq, sp["pad"], sp["eos"], tok.max_token_length)
    sd = helpers.synth_weights("ralf_cgl", meta["seed"])
    memory = torch.from_numpy(z["memory" if mode == "deterministic" else p + "memory"])
    fn = R.RelationConstraint(pre)
    cfg = {"name": mode, "temperature": 1.0, "top_k": 5, "top_p": 0.9}
    calls = []
    real_mask = fn.mask

    def spy(prefix, cons):
        m, back = real_mask(prefix, cons)
        calls.append((len(prepared) - 1, list(prefix), m, -1 if back is None else back))
        return m, back

    fn.mask = spy
    prepared, rows = [], []
    for b in range(meta["B"]):
        cons = fn.prepare(const["seq"][b])
        prepared.append(cons)
        np.testing.assert_array_equal(_encode_constraints(cons), z[p + f"prepared_{b}"])
        prefix = R.sample_with_backtracking(_oracle_logits_fn(sd, memory[b], sp["pad"]), fn, cons, forced[b],
                                            bos_id=sp["bos"], eos_id=sp["eos"], max_token_length=tok.max_token_length,
                                            sampling_cfg=cfg)
        rows.append(prefix)
    assert len(calls) == len(z[p + "call_sample"]), "the sampler took a different path through the rewinds"
    ref_mask = np.unpackbits(z[p + "call_mask"], axis=1)[:, :tok.N_total].astype(bool)
    for i, (b, prefix, m, back) in enumerate(calls):
        assert b == z[p + "call_sample"][i]
        assert prefix == z[p + "call_prefix"][i, :z[p + "call_len"][i]].tolist(), f"call {i}"
        assert back == z[p + "call_back"][i], f"call {i}"
        np.testing.assert_array_equal(m.numpy(), ref_mask[i], err_msg=f"call {i}")
    seq = R.pad_like_reference(rows, tok.max_token_length)
    out = tok.decode(seq)
    for k in ["label", "mask", *GEO]:
        np.testing.assert_array_equal(out[k].numpy(), z[p + f"gen_{k}"], err_msg=k)
    vio = T.calculate_violation("relation", cond, seq, tok, output=out, prepared_rel_constraints=prepared)
    assert [vio["total"], vio["viorated"]] == z[p + "violation"].tolist()


def test_relation_masks_for_every_relation_kind(fx):
    """1800 forward-walked prefixes: all table rows as constraints (relation size 100 %) and crafted constraint lists
    covering every RelSize / RelLoc kind, several per element, boxes from tiny to canvas-sized."""
    z, meta, tok, batch, table = fx
    pre = _preprocessor(tok, table, meta)
    random.seed(420)
    torch.manual_seed(420)
    cond, _ = T.get_condition(copy.deepcopy(batch), "relation", tok)
    pre.set_relation_size(100)
    const = pre(cond)
    np.testing.assert_array_equal(const["seq"].numpy(), z["sweep_const_seq"])
    fn = R.RelationConstraint(pre)
    B = meta["B"]
    cons_of, types_of = {}, {}
    for b in range(B):
        cons_of[b] = fn.prepare(const["seq"][b])
        types_of[b] = fn.types.clone()
        np.testing.assert_array_equal(_encode_constraints(cons_of[b]), z[f"sweep_prepared_{b}"])
    for t in range(int(z["craft_count"])):
        cons_of[B + t] = _decode_constraints(z[f"craft_prepared_{t}"])
        n = len(types_of[t % B])
        cons_of[B + t] += [[] for _ in range(n - len(cons_of[B + t]))]
        types_of[B + t] = types_of[t % B]
    ref_mask = np.unpackbits(z["sweep_mask"], axis=1)[:, :tok.N_total].astype(bool)
    kinds = set()
    for i in range(len(z["sweep_len"])):
        s = int(z["sweep_sample"][i])
        fn.types = types_of[s]
        prefix = z["sweep_prefix"][i, :z["sweep_len"][i]].tolist()
        m, back = fn.mask(prefix, cons_of[s])
        assert (-1 if back is None else back) == z["sweep_back"][i], f"call {i}"
        np.testing.assert_array_equal(m.numpy(), ref_mask[i], err_msg=f"call {i} sample {s} prefix {prefix}")
        kinds |= {k for mine in cons_of[s] for k, _ in mine}
    assert {R.CANVAS, *R.RelSize, *R.RelLoc} <= kinds


def test_draw_token_matches_reference_sampler(fx):
    z, meta, tok, batch, table = fx
    rows = torch.from_numpy(z["draw_logits"])
    for mode in ["deterministic", "random", "top_k", "top_p", "gumbel"]:
        cfg = {"name": mode, "temperature": 0.8, "top_k": 5, "top_p": 0.9}
        torch.manual_seed(431)
        got = [R.draw_token(rows[i].clone(), cfg, temperature=1.5 if i % 2 else None) for i in range(rows.size(0))]
        assert got == z[f"draw_{mode}"].tolist(), mode


def test_label_only_restriction_without_backtracking(fx):
    """use_backtrack=False (retrieval_augmented_autoreg.py:244-300): the batched decode under the label restriction;
    host side (forced-token table, violation count on the reference's own output); the GPU decode is in tests/test_tasks_gpu.py."""
    z, meta, tok, batch, table = fx
    pre = _preprocessor(tok, table, meta)
    random.seed(meta["nobt_seed"])
    torch.manual_seed(meta["nobt_seed"])
    cond, _ = T.get_condition(copy.deepcopy(batch), "relation", tok)
    const = pre(cond)
    sp = meta["special"]
    forced = T.forced_token_table("relation", cond.seq, sp["pad"], sp["eos"], tok.max_token_length)
    ref_forced = T.forced_token_table("refinement", cond.seq, sp["pad"], sp["eos"], tok.max_token_length)
    assert torch.equal(forced, ref_forced)  # DECODE_SPACE_RESTRICTION maps both to restrict_only_category
    fn = R.RelationConstraint(pre)
    prepared = [fn.prepare(const["seq"][b]) for b in range(meta["B"])]
    out = {k: torch.from_numpy(z[f"nobt_gen_{k}"]) for k in ["label", "mask", *GEO]}
    vio = T.calculate_violation("relation", cond, None, tok, output=out, prepared_rel_constraints=prepared)
    assert [vio["total"], vio["viorated"]] == z["nobt_violation"].tolist()
    # every given label is reproduced by the reference's own output (the restriction the forced table encodes)
    for b in range(meta["B"]):
        k = int(out["mask"][b].sum())
        given = cond.seq[b, 1::5][:k]
        assert torch.equal(out["label"][b][:k], given[:k])


def test_relation_preprocessor_reference_properties():
    """The reference's own test for this preprocessor (tests/train/helpers/test_task_preprocessor.py:28-58,117-137):
    check_get_condition, check_output, and prepare() on every row, over random batches."""
    from oracle import synth

    tok = helpers.make_tokenizer()
    names = ["logo", "text", "underlay", "embellishment"]
    for seed in range(8):
        batch = synth.synth_batch(5, 8, 8, 10, 1, 4, seed=seed)
        random.seed(seed)
        torch.manual_seed(seed)
        pre = R.RelationPreprocessor(tok, R.describe_relationships(batch, names), relation_size=[10, 50, 100][seed % 3])
        cond, _ = T.get_condition(copy.deepcopy(batch), "relation", tok)
        assert -1 not in cond.seq[cond.mask].tolist()
        assert len(set(cond.seq[~cond.mask].tolist())) <= 1
        out = pre(cond)
        seq, pad_mask = out["seq"], out["pad_mask"]
        assert pre.N_total > 0 and seq.min() >= 0 and seq.max() < pre.N_total
        assert set(seq[pad_mask].tolist()) <= {pre.name_to_id("pad")}
        assert pre.name_to_id("pad") not in seq[~pad_mask].tolist()
        assert (seq[:, 0] == pre.name_to_id("bos")).all() and (seq[:, 1] == pre.name_to_id("relationship")).all()
        torch.nn.Embedding(pre.N_total, 8)(seq)
        fn = R.RelationConstraint(pre)
        for b in range(seq.size(0)):
            cons = fn.prepare(seq[b])
            assert len(cons) == int(batch["mask"][b].sum())
            for e, mine in enumerate(cons):
                for kind, tgt in mine:
                    assert (kind == R.CANVAS and isinstance(tgt, R.RelLoc)) or (isinstance(kind, (R.RelLoc, R.RelSize)) and 0 <= tgt < e)


def test_ground_truth_layout_violates_none_of_its_own_relations():
    """Size-independent property tying table builder, constraint sequence, prepare() and the violation count together:
    with every table row used as a constraint (relation size 100 %), the layout the table was computed from breaks none
    of them.  Single-label canvases, where the label shuffle of the constraint sequence cannot move positions."""
    tok = helpers.make_tokenizer()
    names = ["logo", "text", "underlay", "embellishment"]
    g = torch.Generator().manual_seed(11)
    B, E = 6, 10
    n = torch.randint(1, E + 1, (B,), generator=g)
    mask = torch.arange(E)[None] < n[:, None]
    batch = {"mask": mask, "label": torch.randint(0, 4, (B, 1), generator=g).expand(B, E) * mask,
             "id": [str(i) for i in range(B)], "image": torch.zeros(B, 3, 8, 8), "saliency": torch.zeros(B, 1, 8, 8)}
    for k in GEO:
        batch[k] = torch.rand(B, E, generator=g) * mask
    table = R.describe_relationships(batch, names)
    random.seed(1)
    torch.manual_seed(1)
    pre = R.RelationPreprocessor(tok, table, relation_size=100)
    cond, _ = T.get_condition(copy.deepcopy(batch), "relation", tok)
    const = pre(cond)
    fn = R.RelationConstraint(pre)
    prepared = [fn.prepare(const["seq"][b]) for b in range(B)]
    vio = R.violation_count({k: batch[k] for k in GEO}, prepared)
    assert vio["total"] == sum(len(v) for v in table.values()) and vio["viorated"] == 0, vio
    # and a layout with one element moved does break some
    moved = {k: batch[k].clone() for k in GEO}
    moved["center_y"] = 1.0 - moved["center_y"]
    assert R.violation_count(moved, prepared)["viorated"] > 0


def test_decode_session_cache_bookkeeping(monkeypatch):
    """engine.DecodeSession's rewind logic with the kernels replaced by a toy decoder whose logits depend on EVERY cached
    row (position-weighted), so a stale or missing K/V row after a rewind changes the answer.  Random walks with rewinds,
    as sample_with_backtracking produces them; the device kernels themselves are covered in tests/test_tasks_gpu.py."""
    from ralf_b200 import engine as E

    V, steps, pad = 23, 12, 21
    calls = []

    def fake_embed(seq, col, S, emb, scale, pe, pos0):
        assert S == 1 and col == pos0
        return seq[:, col:col + 1].to(torch.float32) + 1.0  # "embedding" = token id + 1

    class FakeEngine:
        dev = torch.device("cpu")
        w = {"decoder.emb": None, "pe1d": None}

        def _decode_step(self, x, t, kc, vc, kvm, pad_mask, B, Mlen):
            calls.append(t)
            kc[0][0, t, 0] = x[0, 0]                       # append this position's "key"
            keys = kc[0][0, :t + 1, 0]
            live = (pad_mask[0, :t + 1] == 0).to(torch.float32)
            weights = torch.arange(1, t + 2, dtype=torch.float32)
            h = (keys * weights * live).sum()              # every cached row matters, and so does its position
            return (h * torch.arange(1, V + 1, dtype=torch.float32))[None]

    def direct(prefix):
        keys = torch.tensor(prefix, dtype=torch.float32) + 1.0
        live = torch.tensor([p != pad for p in prefix], dtype=torch.float32)
        h = (keys * torch.arange(1, len(prefix) + 1, dtype=torch.float32) * live).sum()
        return h * torch.arange(1, V + 1, dtype=torch.float32)

    monkeypatch.setattr(E.ops, "embed", fake_embed)
    rng = random.Random(3)
    for trial in range(20):
        session = E.DecodeSession(FakeEngine(), kvm=[None], Mlen=1, steps=steps, pad_id=pad)
        prefix = [20]
        for _ in range(60):
            calls.clear()
            before = list(session.cached)
            got = session.logits_of(prefix)
            assert torch.equal(got, direct(prefix)), (trial, prefix)
            common = 0
            while common < min(len(before), len(prefix)) and before[common] == prefix[common]:
                common += 1
            if before == prefix:
                assert calls == []                         # same question twice: answered from the last logits
            else:
                assert calls == list(range(min(common, len(prefix) - 1), len(prefix)))  # only the new suffix is decoded
            move = rng.random()
            if move < 0.25 and len(prefix) > 1:            # rewind (sample_with_backtracking: prefix[:idx])
                prefix = prefix[:rng.randint(1, len(prefix) - 1)]
            elif move < 0.35 and len(prefix) > 2:          # rewind and change a token in the middle
                cut = rng.randint(1, len(prefix) - 1)
                prefix = prefix[:cut] + [rng.randrange(V)]
            elif len(prefix) < steps:
                prefix = prefix + [rng.randrange(V)]
            else:
                prefix = [20]
    with pytest.raises(AssertionError):
        session.logits_of(list(range(steps + 1)))


class _OracleEngine:
    """Stands in for ralf_b200.engine.Engine on the CPU: the oracle computes what the kernels would.  Lets the model
    class's relation plumbing (sample -> sample_relation -> DecodeSession -> sampler -> decode -> violation) run end to
    end without a GPU; tests only."""

    npass = 3
    dev = torch.device("cpu")
    w = {"decoder.emb": None, "pe1d": None}

    def __init__(self, sd, pad_id):
        self.sd, self.pad_id = sd, pad_id

    def encode(self, image, retrieved, seq_const, pad_mask):
        from oracle import ralf_oracle as O

        with torch.no_grad():
            mem = O.encode_ralf_memory(self.sd, image, {k: v.float() for k, v in retrieved.items() if torch.is_tensor(v)},
                                       seq_const, pad_mask.bool())
        return mem, mem.reshape(-1, mem.shape[-1])

    def cross_kv(self, mem_s, kv24=False):
        return [mem_s]

    def _logits(self, tokens, memory, pad):
        from oracle import ralf_oracle as O

        with torch.no_grad():
            return O.decoder_logits(self.sd, tokens, memory, pad)[:, -1]

    def _decode_step(self, x, t, kc, vc, kvm, pad_mask, B, Mlen):
        kc[0][:, t, 0] = x[:, 0]  # the fake embedding below is the token id itself
        tokens = kc[0][:, :t + 1, 0].round().long()
        return self._logits(tokens, kvm[0].view(B, Mlen, -1), pad_mask[:, :t + 1].bool())

    def generate_graphed(self, mem_s, B, Mlen, token_mask, bos_id, pad_id, steps):
        return self.generate(mem_s, B, Mlen, token_mask, bos_id, pad_id, steps)  # no graphs on the CPU: same loop

    def generate(self, mem_s, B, Mlen, token_mask, bos_id, pad_id, steps, forced=None, sampling=None, rng=None, **kw):
        assert (sampling or {}).get("name", "deterministic") == "deterministic"
        seq = torch.full((B, 1), bos_id)
        memory = mem_s.view(B, Mlen, -1)
        for t in range(steps):
            lg = self._logits(seq, memory, seq == pad_id).clone()
            lg[:, ~token_mask[t].bool()] = -float("inf")
            for b in range(B):
                f = int(forced[b, t]) if forced is not None else -1
                if f >= 0:
                    keep = lg[b, f].clone()
                    lg[b] = -float("inf")
                    lg[b, f] = keep
            seq = torch.cat([seq, lg.argmax(dim=1, keepdim=True)], dim=1)
        return seq[:, 1:]


def _relation_model(fx, monkeypatch):
    from oracle import synth
    from ralf_b200 import engine as E
    from ralf_b200 import generator as G

    z, meta, tok, batch, table = fx
    full = synth.synth_batch(meta["B"], meta["H"], meta["W"], 10, 16, 4, seed=meta["seed"])
    for k in ["label", "mask", *GEO]:
        full[k] = batch[k]
    random.seed(meta["ctor_seed"])
    model = G.RALF(features=None, tokenizer=tok, dataset_name="cgl", max_seq_length=10, top_k=16,
                   auxilary_task="relation", relation_table=copy.deepcopy(table))
    sd = helpers.synth_weights("ralf_cgl", meta["seed"])
    model.load_state_dict(sd, strict=True)
    fake = _OracleEngine(sd, meta["special"]["pad"])
    monkeypatch.setattr(model, "engine", lambda: fake)
    monkeypatch.setattr(E.ops, "embed", lambda seq, col, S, emb, scale, pe, pos0: seq[:, col:col + 1].to(torch.float32))
    return model.eval(), full


@pytest.mark.parametrize("mode", ["deterministic", "random"])
def test_model_sample_relation_end_to_end_on_oracle_engine(fx, monkeypatch, mode):
    """model.sample(cond_type="relation", use_backtrack=True) through the drop-in class with the oracle standing in for the
    kernels: the reference's decoded layouts and violation counts (encoder memory included this time)."""
    z, meta, tok, batch, table = fx
    torch.set_num_threads(8)
    model, full = _relation_model(fx, monkeypatch)
    seed = meta["rng_seed"][mode]
    random.seed(seed)
    torch.manual_seed(seed)
    cond, _ = T.get_condition(copy.deepcopy(full), "relation", tok)
    out, vio = model.sample(cond=cond, sampling_cfg={"name": mode, "temperature": 1.0, "top_k": 5, "top_p": 0.9},
                            cond_type="relation", return_violation=True, use_backtrack=True, return_decoded_cond=True)
    p = f"bt_{mode}_"
    for k in ["label", "mask", *GEO]:
        np.testing.assert_array_equal(out[k].numpy(), z[p + f"gen_{k}"], err_msg=k)
    assert [vio["total"], vio["viorated"]] == z[p + "violation"].tolist()
    assert out["decoded_tokens"][0][:3] == ["bos", "relationship", "end_of_task"]


def test_model_sample_relation_without_backtracking_on_oracle_engine(fx, monkeypatch):
    z, meta, tok, batch, table = fx
    torch.set_num_threads(8)
    model, full = _relation_model(fx, monkeypatch)
    random.seed(meta["nobt_seed"])
    torch.manual_seed(meta["nobt_seed"])
    cond, _ = T.get_condition(copy.deepcopy(full), "relation", tok)
    out, vio = model.sample(cond=cond, sampling_cfg={"name": "deterministic"}, cond_type="relation",
                            return_violation=True, use_backtrack=False)
    for k in ["label", "mask", *GEO]:
        np.testing.assert_array_equal(out[k].numpy(), z[f"nobt_gen_{k}"], err_msg=k)
    assert [vio["total"], vio["viorated"]] == z["nobt_violation"].tolist()


def test_model_preprocess_with_the_relation_task(fx, monkeypatch):
    """Training entry point (train.py:432) of a model built with auxilary_task="relation" (configs/ralf_cgl/relation.sh):
    preprocess() = get_condition -> RelationPreprocessor -> tokenizer.encode, same host RNG order as the reference, so
    under the fixture's seed it yields the fixture's constraint sequence; teacher-forcing tensors keep their contract."""
    z, meta, tok, batch, table = fx
    model, full = _relation_model(fx, monkeypatch)
    seed = meta["rng_seed"]["deterministic"]
    random.seed(seed)
    torch.manual_seed(seed)
    inputs, targets = model.preprocess(copy.deepcopy(full))
    np.testing.assert_array_equal(inputs["seq_layout_const"].numpy(), z["bt_deterministic_const_seq"])
    np.testing.assert_array_equal(inputs["seq_layout_const_pad_mask"].numpy(), z["bt_deterministic_const_pad_mask"])
    enc = tok.encode({k: full[k] for k in ["label", "mask", *GEO]})
    assert torch.equal(inputs["seq"], enc["seq"][:, :-1]) and torch.equal(targets["seq"], enc["seq"][:, 1:])
    assert inputs["image"].shape[1] == 4 and set(inputs["retrieved"]) >= {"label", "mask", *GEO}


def test_bench_model_api_leg_runs_on_the_oracle_engine(monkeypatch):
    """bench.py's e2e_model_api leg (search -> fetch -> model.sample through the reference scripts' call sequence) end to end
    on the CPU with the oracle standing in for the kernels: the code path itself, not a number."""
    import bench
    from oracle import synth
    from ralf_b200 import engine as E
    from ralf_b200 import generator as G
    from ralf_b200 import ops
    from ralf_b200.retrieval import GpuRetriever
    from tests import oracle_knn

    def oracle_knn_topk(gallery, queries, k, *, index_base=0, gallery_max_norm=0.0, exact=False, workspace=None):
        i, s = oracle_knn.topk(gallery.numpy(), queries.numpy(), k)
        return torch.from_numpy(i) + index_base, torch.from_numpy(s), torch.ones(queries.shape[0], dtype=torch.int32)

    monkeypatch.setattr(ops, "knn_topk", oracle_knn_topk)
    monkeypatch.setattr(ops, "gather_layouts", lambda packed, idx: packed[idx])
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    tok = helpers.make_tokenizer()
    model = G.RALF(features=None, tokenizer=tok, dataset_name="cgl", max_seq_length=10, top_k=16, pretrained=False)
    sd = helpers.synth_weights("ralf_cgl", 3)
    model.load_state_dict(sd, strict=True)
    monkeypatch.setattr(model, "engine", lambda: _OracleEngine(sd, tok.name_to_id("pad")))
    g = torch.Generator().manual_seed(0)
    n, E_ = 200, 10
    cnt = torch.randint(1, E_ + 1, (n,), generator=g)
    mask = torch.arange(E_)[None] < cnt[:, None]
    lay = {"mask": mask, "label": torch.randint(0, 4, (n, E_), generator=g) * mask}
    for k in GEO:
        lay[k] = torch.rand(n, E_, generator=g) * mask
    retr = GpuRetriever(torch.randn(n, 512, generator=g), lay, device="cpu")
    torch.set_num_threads(8)
    out = bench.model_api_e2e(model.eval(), retr, torch.rand(2, 4, 64, 64, generator=g), torch.randn(2, 512, generator=g),
                              torch.device("cpu"), n=2, iters=1)
    assert out["unit"] == "layouts/s" and out["value"] > 0 and out["canvases_per_call"] == 2


def test_sample_graph_switch_only_takes_plain_greedy(monkeypatch):
    """RALF_SAMPLE_GRAPH routes model.sample() through Engine.generate_graphed for unconstrained greedy decoding only;
    constrained tasks (forced-token table) and stochastic samplers keep the eager loop."""
    from oracle import synth
    from ralf_b200 import generator as G

    calls = []

    class Eng(_OracleEngine):
        def generate_graphed(self, mem_s, B, Mlen, token_mask, bos_id, pad_id, steps):
            calls.append("graphed")
            return _OracleEngine.generate(self, mem_s, B, Mlen, token_mask, bos_id, pad_id, steps)

        def generate(self, *a, **k):
            calls.append("eager")
            k.pop("sampling", None)
            return _OracleEngine.generate(self, *a, **k)

    tok = helpers.make_tokenizer()
    model = G.RALF(features=None, tokenizer=tok, dataset_name="cgl", max_seq_length=10, top_k=16, pretrained=False,
                   use_multitask=True)
    sd = helpers.synth_weights("ralf_cgl", 3)
    model.load_state_dict(sd, strict=True)
    eng = Eng(sd, tok.name_to_id("pad"))
    monkeypatch.setattr(model, "engine", lambda: eng)
    monkeypatch.setattr(G, "_SAMPLE_GRAPH", True)
    torch.set_num_threads(8)
    batch = synth.synth_batch(2, 64, 64, 10, 16, 4, seed=8)
    cond, _ = T.get_condition(copy.deepcopy(batch), "uncond", tok)
    plain = model.eval().sample(cond=cond, cond_type="uncond", return_seq=True)
    assert calls == ["graphed"]
    monkeypatch.setattr(G, "_SAMPLE_GRAPH", False)
    calls.clear()
    eager = model.sample(cond=cond, cond_type="uncond", return_seq=True)
    assert calls == ["eager"] and torch.equal(plain["seq"], eager["seq"])
    monkeypatch.setattr(G, "_SAMPLE_GRAPH", True)
    calls.clear()
    torch.manual_seed(0)
    cond_c, _ = T.get_condition(copy.deepcopy(batch), "c", tok)
    model.sample(cond=cond_c, cond_type="c")
    assert calls == ["eager"]  # a forced-token table: not the plain loop
