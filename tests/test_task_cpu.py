"""CPU: host side of the constrained tasks (ralf_b200/task.py) and the oracle's restatements of the
decoding-space restriction and of helpers/sampling.py, against fixtures dumped from the UNMODIFIED reference
(tests/golden/tasks_cgl_256.npz, sampling_filters.npz; tests/golden/make_golden.py)."""
import copy
import json

import numpy as np
import pytest
import torch

from tests import helpers

TASKS = ["c", "cwh", "partial", "refinement"]


def _task_inputs(task):
    from ralf_b200 import task as T

    z, meta = helpers.load_golden("tasks_cgl_256")
    tok = helpers.make_tokenizer()
    batch = helpers.synth_batch({**meta, "E": 10, "K": 16})
    torch.manual_seed(meta["rng_seed"][task])
    cond, batch = T.get_condition(copy.deepcopy(batch), task, tok)
    return z, tok, cond, batch


@pytest.mark.parametrize("task", TASKS)
def test_condition_and_constraint_sequence_match_reference(task):
    from ralf_b200 import task as T

    z, tok, cond, batch = _task_inputs(task)
    np.testing.assert_array_equal(cond.seq.numpy(), z[f"{task}_cond_seq"])
    np.testing.assert_array_equal(cond.mask.numpy(), z[f"{task}_cond_mask"])
    if task == "refinement":
        for k in ["center_x", "center_y", "width", "height"]:
            np.testing.assert_array_equal(batch[k].numpy(), z[f"{task}_noisy_{k}"])
    const = T.TaskPreprocessor(tok, task)(cond)
    np.testing.assert_array_equal(const["seq"].numpy(), z[f"{task}_const_seq"])
    np.testing.assert_array_equal(const["pad_mask"].numpy(), z[f"{task}_const_pad_mask"])
    np.testing.assert_array_equal(cond.seq.numpy(), z[f"{task}_cond_seq_after"])  # in-place <eos> -> <pad> like the reference


@pytest.mark.parametrize("task", TASKS)
def test_forced_table_equals_literal_restriction(task):
    """The forced-token table must act on ANY logits exactly like the reference's per-sample masking loop."""
    from oracle import ralf_oracle as O
    from ralf_b200 import task as T

    z, tok, _, _ = _task_inputs(task)
    cond_seq = torch.from_numpy(z[f"{task}_cond_seq_after"])
    pad, eos = tok.name_to_id("pad"), tok.name_to_id("eos")
    forced = T.forced_token_table(task, cond_seq, pad, eos, tok.max_token_length)
    g = torch.Generator().manual_seed(0)
    for i in range(5 if task == "partial" else 0, tok.max_token_length):
        logits = torch.randn((cond_seq.shape[0], tok.N_total), generator=g)
        logits[:, ~tok.token_mask[i]] = float("-inf")
        ref = O.restrict_logits(task, i + 1, cond_seq, logits.clone(), pad, eos)
        mine = logits.clone()
        for b in range(mine.shape[0]):
            f = int(forced[b, i])
            if f >= 0:
                v = mine[b, f].clone()
                mine[b] = float("-inf")
                mine[b, f] = v
        assert torch.equal(ref, mine), (task, i)
    if task == "partial":
        np.testing.assert_array_equal(forced[:, :5].numpy(), z["partial_cond_seq"][:, 1:6])
        assert (forced[:, 5:] == -1).all()


@pytest.mark.parametrize("task", TASKS)
def test_oracle_constrained_greedy_matches_reference(task):
    from oracle import ralf_oracle as O
    from ralf_b200 import task as T

    torch.set_num_threads(8)
    z, tok, cond, batch = _task_inputs(task)
    _, meta = helpers.load_golden("tasks_cgl_256")
    sd = helpers.synth_weights("ralf_cgl", meta["seed"])
    const = T.TaskPreprocessor(tok, task)(cond)
    pad, bos, eos = tok.name_to_id("pad"), tok.name_to_id("bos"), tok.name_to_id("eos")
    with torch.no_grad():
        retrieved = {k: v.float() for k, v in batch["retrieved"].items()}
        mem = O.encode_ralf_memory(sd, cond.image, retrieved, const["seq"], const["pad_mask"])
        seq = O.constrained_greedy_sample(sd, mem, tok.token_mask, task, cond.seq, bos, pad, eos, tok.max_token_length)
    np.testing.assert_array_equal(seq.numpy(), z[f"{task}_gen_seq"])
    dec = tok.decode(seq)
    for k in ["label", "mask", "center_x", "center_y", "width", "height"]:
        np.testing.assert_array_equal(dec[k].numpy(), z[f"{task}_gen_{k}"])
    vio = T.calculate_violation(task, cond, seq, tok)
    assert [vio["total"], vio["viorated"]] == z[f"{task}_violation"].tolist()


def test_oracle_sampling_filters_match_reference():
    from oracle import ralf_oracle as O

    z = np.load(helpers.GOLDEN + "/sampling_filters.npz")
    logits = torch.from_numpy(z["logits"])
    for i, cfg in enumerate(json.loads(str(z["cfgs"]))):
        ref = z[f"probs_{i}"]
        p = O.filtered_probs(logits.clone(), cfg["name"], cfg.get("temperature", 1.0), cfg.get("top_k", 5),
                             cfg.get("top_p", 0.9)).numpy()
        assert ((p > 0) == (ref > 0)).all(), cfg
        np.testing.assert_allclose(p, ref, rtol=1e-6, atol=1e-9)


def test_inverse_cdf_draw_distribution():
    from oracle import ralf_oracle as O

    p = torch.tensor([[0.5, 0.0, 0.3, 0.2]]).repeat(2000, 1)
    u = torch.rand(2000, generator=torch.Generator().manual_seed(1))
    tok, accept = O.inverse_cdf_draw(torch.log(p), u)
    freq = torch.bincount(tok, minlength=4).double() / 2000
    assert freq[1] == 0 and (freq - p[0].double()).abs().max() < 0.04
    assert all(int(t) in a for t, a in zip(tok, accept))


@pytest.mark.parametrize("seed", [0, 5, 17])
def test_task_edge_cases_match_reference(seed):
    """Single-element, full (10 elements) and single-label layouts plus random ones, all four tasks: condition and
    constraint sequences equal the reference's (tests/golden/tasks_edge_cases.npz; 160 such seed x task combinations
    were compared live against the reference when the fixture was made, without a mismatch)."""
    from oracle import synth
    from ralf_b200 import task as T

    z = np.load(helpers.GOLDEN + "/tasks_edge_cases.npz")
    tok = helpers.make_tokenizer()
    batch = synth.synth_batch(6, 8, 8, 10, 1, 4, seed=1000 + seed)
    if seed % 5 == 0:
        batch["mask"][0] = torch.arange(10) < 1
        batch["mask"][1] = True
        batch["label"][2] = 1
        for k in ["label", "center_x", "center_y", "width", "height"]:
            batch[k] = batch[k] * batch["mask"]
    for task in TASKS:
        torch.manual_seed(seed)
        cond, _ = T.get_condition(copy.deepcopy(batch), task, tok)
        np.testing.assert_array_equal(cond.seq.numpy(), z[f"{seed}_{task}_cond_seq"])
        np.testing.assert_array_equal(cond.mask.numpy(), z[f"{seed}_{task}_cond_mask"])
        const = T.TaskPreprocessor(tok, task)(cond)
        np.testing.assert_array_equal(const["seq"].numpy(), z[f"{seed}_{task}_const_seq"])


@pytest.mark.parametrize("seed", [1, 3, 7])
def test_violation_counts_match_reference_on_corrupted_outputs(seed):
    """violate.py:24-139 on sequences with ~15 % corrupted tokens (fixture: the reference's own counts)."""
    from oracle import synth
    from ralf_b200 import task as T

    z = np.load(helpers.GOLDEN + "/violation_cases.npz")
    tok = helpers.make_tokenizer()
    batch = synth.synth_batch(5, 8, 8, 10, 1, 4, seed=2000 + seed)
    for task in ("c", "cwh", "refinement"):
        torch.manual_seed(seed)
        cond, _ = T.get_condition(copy.deepcopy(batch), task, tok)
        T.TaskPreprocessor(tok, task)(cond)
        vio = T.calculate_violation(task, cond, torch.from_numpy(z[f"{seed}_{task}_seq"]), tok)
        assert [vio["total"], vio["viorated"]] == z[f"{seed}_{task}_violation"].tolist(), (seed, task)
        assert vio["viorated"] > 0


@pytest.mark.parametrize("task", ["c", "cwh", "partial", "refinement"])
def test_model_sample_constrained_tasks_end_to_end_on_oracle_engine(task, monkeypatch):
    """The drop-in class's sample() for the constrained tasks, wired end to end on the CPU with the oracle standing in for
    the kernels (tests/test_relation_cpu.py:_OracleEngine): get_condition -> task preprocessor -> forced-token table ->
    encode -> restricted greedy decode -> tokenizer.decode -> violation count = the reference's sample() golden."""
    import copy

    from ralf_b200 import engine as E
    from ralf_b200 import generator as G
    from ralf_b200 import task as T
    from tests.test_relation_cpu import _OracleEngine

    z, meta = helpers.load_golden("tasks_cgl_256")
    tok = helpers.make_tokenizer()
    model = G.RALF(features=None, tokenizer=tok, dataset_name="cgl", max_seq_length=10, top_k=16, auxilary_task="uncond",
                   use_multitask=True, pretrained=False)
    sd = helpers.synth_weights("ralf_cgl", meta["seed"])
    model.load_state_dict(sd, strict=True)
    monkeypatch.setattr(model, "engine", lambda: _OracleEngine(sd, tok.name_to_id("pad")))
    monkeypatch.setattr(E.ops, "embed", lambda seq, col, S, emb, scale, pe, pos0: seq[:, col:col + 1].to(torch.float32))
    torch.set_num_threads(8)
    batch = helpers.synth_batch({**meta, "E": 10, "K": 16})
    torch.manual_seed(meta["rng_seed"][task])
    cond, _ = T.get_condition(copy.deepcopy(batch), task, tok)
    out, vio = model.eval().sample(cond=cond, sampling_cfg={"name": "deterministic"}, cond_type=task, return_violation=True,
                                   return_seq=True)
    np.testing.assert_array_equal(out["seq"].numpy(), z[f"{task}_gen_seq"])
    for k in ["label", "mask", "center_x", "center_y", "width", "height"]:
        np.testing.assert_array_equal(out[k].numpy(), z[f"{task}_gen_{k}"])
    assert [vio["total"], vio["viorated"]] == z[f"{task}_violation"].tolist()
