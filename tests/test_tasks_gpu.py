"""GPU: constrained tasks and stochastic sampling through the drop-in model class.

* c / cwh / partial / refinement: greedy token ids, decoded layouts and violation counts bit-exact against the
  reference's own sample() (tests/golden/tasks_cgl_256.npz).
* ralf_sample_next: filters (top_k / top_p / temperature) + inverse-CDF draw against the oracle restatement of
  helpers/sampling.py on the fixture logits; forced-token table; degenerate samplers equal greedy."""
import copy
import json

import numpy as np
import pytest
import torch

from tests import helpers

pytestmark = pytest.mark.gpu
TASKS = ["c", "cwh", "partial", "refinement"]


def _model(dev, **kw):
    from ralf_b200 import generator as G

    z, meta = helpers.load_golden("tasks_cgl_256")
    tok = helpers.make_tokenizer()
    model = G.RALF(features=None, tokenizer=tok, dataset_name="cgl", max_seq_length=10, top_k=16, **kw)
    model.load_state_dict(helpers.synth_weights("ralf_cgl", meta["seed"]), strict=True)
    return model.eval().to(dev), tok, z, meta


@pytest.mark.parametrize("task", TASKS)
def test_constrained_task_matches_reference_golden(cuda_device, task):
    from ralf_b200 import task as T

    model, tok, z, meta = _model(cuda_device, auxilary_task="uncond", use_multitask=True)
    batch = helpers.synth_batch({**meta, "E": 10, "K": 16})
    torch.manual_seed(meta["rng_seed"][task])
    cond, _ = T.get_condition(copy.deepcopy(batch), task, tok)
    cond = cond.to(cuda_device)
    out, vio = model.sample(cond=cond, sampling_cfg={"name": "deterministic"}, cond_type=task, return_violation=True,
                            return_seq=True)
    np.testing.assert_array_equal(out["seq"].numpy(), z[f"{task}_gen_seq"])  # bit-exact token ids
    for k in ["label", "mask", "center_x", "center_y", "width", "height"]:
        np.testing.assert_array_equal(out[k].numpy(), z[f"{task}_gen_{k}"])
    assert [vio["total"], vio["viorated"]] == z[f"{task}_violation"].tolist()


def _run_kernel(logits, allowed, dev, **kw):
    from ralf_b200 import ops

    B, V = logits.shape
    seq = torch.full((B, 4), -7, dtype=torch.int64, device=dev)
    pad_mask = torch.zeros((B, 4), dtype=torch.uint8, device=dev)
    emb = torch.randn((V, 256), device=dev)
    pe = torch.randn((8, 256), device=dev)
    x_next = torch.zeros((B, 256), device=dev)
    ops.sample_next(logits.to(dev).contiguous(), allowed.to(dev).to(torch.uint8).contiguous(), seq, 2, pad_mask, 516, emb,
                    16.0, pe, x_next, **kw)
    torch.cuda.synchronize()
    tok = seq[:, 2].cpu()
    assert (seq[:, [0, 1, 3]] == -7).all()
    np.testing.assert_array_equal(pad_mask[:, 2].cpu().numpy(), (tok == 516).numpy().astype(np.uint8))
    ref_x = emb[tok.to(dev)] * 16.0 + pe[2]
    assert torch.allclose(x_next, ref_x, rtol=1e-6, atol=1e-5)  # the kernel fuses the multiply-add
    return tok


def test_sample_next_filters_and_draw_match_oracle(cuda_device):
    from oracle import ralf_oracle as O

    z = np.load(helpers.GOLDEN + "/sampling_filters.npz")
    logits = torch.from_numpy(z["logits"])
    N, V = logits.shape
    allowed = torch.ones(V, dtype=torch.bool)
    g = torch.Generator().manual_seed(11)
    n_checked = 0
    for cfg in json.loads(str(z["cfgs"])):
        for rep in range(8):
            u = torch.rand(N, generator=g)
            if rep == 0:
                u[:4] = torch.tensor([0.0, 0.9999, 0.5, 1e-7])
            tok = _run_kernel(logits, allowed, cuda_device, mode=cfg["name"], temperature=cfg.get("temperature", 1.0),
                              top_k=cfg.get("top_k", 5), top_p=cfg.get("top_p", 0.9), uniform=u.to(cuda_device))
            x = O.filtered_logits(logits.clone(), cfg["name"], cfg.get("temperature", 1.0), cfg.get("top_k", 5),
                                  cfg.get("top_p", 0.9))
            pick, accept = O.inverse_cdf_draw(x, u)
            kept = torch.isfinite(x)
            exact = 0
            for b in range(N):
                assert bool(kept[b, tok[b]]), (cfg, b, int(tok[b]))          # never outside the reference's filter
                assert int(tok[b]) in accept[b], (cfg, b, int(tok[b]), int(pick[b]), float(u[b]))
                exact += int(tok[b]) == int(pick[b])
            assert exact >= N - 1, (cfg, exact)  # boundary cases (|cdf - u| < 1e-5) are the only tolerated differences
            n_checked += N
    assert n_checked == 8 * 8 * N


def test_sample_next_vocabulary_mask_forced_and_degenerate_samplers(cuda_device):
    g = torch.Generator().manual_seed(3)
    B, V = 64, 519
    logits = torch.randn((B, V), generator=g)
    allowed = torch.rand(V, generator=g) < 0.3
    allowed[:4] = True
    masked = torch.where(allowed[None], logits, torch.full_like(logits, float("-inf")))
    greedy = masked.argmax(dim=1)
    tok = _run_kernel(logits, allowed, cuda_device, mode="deterministic")
    np.testing.assert_array_equal(tok.numpy(), greedy.numpy())
    u = torch.rand(B, generator=g).to(cuda_device)
    # top_k = 1 and a vanishing top_p keep only the maximum: equal to greedy for any uniform
    for kw in (dict(mode="top_k", top_k=1, temperature=0.7), dict(mode="top_p", top_p=1e-6, temperature=1.5)):
        np.testing.assert_array_equal(_run_kernel(logits, allowed, cuda_device, uniform=u, **kw).numpy(), greedy.numpy())
    # stochastic draws never leave the allowed vocabulary
    for mode in ("random", "top_k", "top_p", "gumbel"):
        noise = torch.rand((B, V), generator=g).to(cuda_device) if mode == "gumbel" else None
        t = _run_kernel(logits, allowed, cuda_device, mode=mode, uniform=u, noise=noise, top_k=5, top_p=0.9)
        assert allowed[t].all(), mode
    # forced table: column `step` overrides everything where >= 0
    forced = torch.full((B, 3), -1, dtype=torch.int32)
    forced[::2, 1] = torch.randint(0, V, (B // 2,), generator=g).to(torch.int32)
    t = _run_kernel(logits, allowed, cuda_device, mode="deterministic", forced=forced.to(cuda_device), step=1)
    want = torch.where(forced[:, 1] >= 0, forced[:, 1].long(), greedy)
    np.testing.assert_array_equal(t.numpy(), want.numpy())


def test_random_sampling_frequencies(cuda_device):
    """mode random over a 6-token distribution: empirical frequencies of 20 k draws within 4 sigma of softmax."""
    V, B = 6, 20000
    row = torch.tensor([0.0, 1.0, -1.0, 2.0, 0.5, -3.0])
    logits = row[None].repeat(B, 1)
    u = torch.rand(B, generator=torch.Generator().manual_seed(5)).to(cuda_device)
    tok = _run_kernel(logits, torch.ones(V, dtype=torch.bool), cuda_device, mode="random", temperature=1.0, uniform=u)
    p = torch.softmax(row.double(), 0)
    freq = torch.bincount(tok, minlength=V).double() / B
    assert ((freq - p).abs() < 4 * torch.sqrt(p * (1 - p) / B) + 1e-4).all(), (freq, p)


def test_model_sample_top_k_runs_and_respects_token_mask(cuda_device):
    """The reference's shipping inference path (scripts/bin/inference.sh: sampling=top_k): tokens stay inside the
    per-position vocabulary and top_k = 1 reproduces the greedy golden sequence."""
    from ralf_b200 import task as T

    model, tok, _, meta = _model(cuda_device, auxilary_task="uncond")
    z, _ = helpers.load_golden("ralf_cgl_256")
    batch = helpers.synth_batch({**meta, "E": 10, "K": 16})
    cond, _ = T.get_condition(copy.deepcopy(batch), "uncond", tok)
    cond = cond.to(cuda_device)
    gen = torch.Generator(device=cuda_device).manual_seed(0)
    out = model.sample(cond=cond, sampling_cfg={"name": "top_k", "top_k": 5, "temperature": 1.0}, cond_type="uncond",
                       return_seq=True, generator=gen)
    seq = out["seq"]
    tm = tok.token_mask
    assert all(bool(tm[i, seq[b, i]]) for b in range(seq.shape[0]) for i in range(seq.shape[1]))
    out1 = model.sample(cond=cond, sampling_cfg={"name": "top_k", "top_k": 1, "temperature": 1.0}, cond_type="uncond",
                        return_seq=True, generator=gen)
    np.testing.assert_array_equal(out1["seq"].numpy(), z["gen_seq"])


# ---- relation (Gen-R): tests/golden/relation_cgl_128.npz, host side pinned in tests/test_relation_cpu.py ----------------
def _relation_setup(dev):
    import random

    from oracle import synth
    from ralf_b200 import generator as G
    from ralf_b200 import relation as R

    z, meta = helpers.load_golden("relation_cgl_128")
    tok = helpers.make_tokenizer()
    batch = synth.synth_batch(meta["B"], meta["H"], meta["W"], 10, 16, 4, seed=meta["seed"])
    for k in ["label", "mask", "center_x", "center_y", "width", "height"]:
        batch[k] = torch.from_numpy(z[k])
    table = R.describe_relationships(batch, meta["label_names"])
    random.seed(meta["ctor_seed"])  # the preprocessor shuffles its table at construction (task_preprocessor.py:507)
    model = G.RALF(features=None, tokenizer=tok, dataset_name="cgl", max_seq_length=10, top_k=16,
                   auxilary_task="relation", relation_table=table)
    model.load_state_dict(helpers.synth_weights("ralf_cgl", meta["seed"]), strict=True)
    return model.eval().to(dev), tok, batch, z, meta


def test_decode_session_rewinds_match_oracle(cuda_device):
    """DecodeSession (KV cache kept across rewinds) against the cache-free oracle decoder on the prefixes the reference's
    backtracking sampler actually visited, rewinds included."""
    from oracle import ralf_oracle as O
    from ralf_b200.engine import KV24, DecodeSession

    model, tok, batch, z, meta = _relation_setup(cuda_device)
    eng = model.engine()
    sd = helpers.synth_weights("ralf_cgl", meta["seed"])
    memory = torch.from_numpy(z["memory"])
    B, Mlen = memory.shape[0], memory.shape[1]
    hi = memory.to(cuda_device).to(torch.bfloat16)
    lo = (memory.to(cuda_device) - hi.float()).to(torch.bfloat16)
    mem_s = torch.stack([hi.reshape(B * Mlen, -1), lo.reshape(B * Mlen, -1)])  # the split operand layout
    kvm = eng.cross_kv(mem_s, kv24=KV24 and eng.npass == 3)
    pad = meta["special"]["pad"]
    torch.set_num_threads(8)
    checked = 0
    for b in range(B):
        session = DecodeSession(eng, [k[b * Mlen:(b + 1) * Mlen] for k in kvm], Mlen, tok.max_token_length, pad)
        idx = np.nonzero(z["bt_deterministic_call_sample"] == b)[0][:60]
        for n, i in enumerate(idx):
            prefix = z["bt_deterministic_call_prefix"][i, :z["bt_deterministic_call_len"][i]].tolist()
            got = session.logits_of(prefix)
            if n % 6 and n != len(idx) - 1:
                continue  # every call goes through the cache; every 6th is compared
            tgt = torch.tensor([prefix])
            with torch.no_grad():
                ref = O.decoder_logits(sd, tgt, memory[b:b + 1], tgt == pad)[0, -1]
            err = float((got - ref).abs().max() / ref.abs().max())
            assert err < 1e-3, f"canvas {b} call {i} prefix length {len(prefix)}: {err:.3e}"
            checked += 1
    assert checked >= 3 * B


@pytest.mark.parametrize("mode", ["deterministic", "random"])
def test_relation_backtracking_matches_reference_golden(cuda_device, mode):
    """model.sample(cond_type="relation") end to end under the recorded seeds: decoded layouts and the violation count of
    the reference's sample_relation (retrieval_augmented_autoreg.py:335-507)."""
    import random

    from ralf_b200 import task as T

    model, tok, batch, z, meta = _relation_setup(cuda_device)
    seed = meta["rng_seed"][mode]
    random.seed(seed)
    torch.manual_seed(seed)
    cond, _ = T.get_condition(copy.deepcopy(batch), "relation", tok)
    cond = cond.to(cuda_device)
    cfg = {"name": mode, "temperature": 1.0, "top_k": 5, "top_p": 0.9}
    out, vio = model.sample(cond=cond, sampling_cfg=cfg, cond_type="relation", return_violation=True, use_backtrack=True)
    p = f"bt_{mode}_"
    for k in ["label", "mask", "center_x", "center_y", "width", "height"]:
        np.testing.assert_array_equal(out[k].numpy(), z[p + f"gen_{k}"], err_msg=k)
    assert [vio["total"], vio["viorated"]] == z[p + "violation"].tolist()


def test_relation_without_backtracking_matches_reference_golden(cuda_device):
    """use_backtrack=False: the batched device decode under the label restriction (:244-300), relations scored afterwards."""
    import random

    from ralf_b200 import task as T

    model, tok, batch, z, meta = _relation_setup(cuda_device)
    random.seed(meta["nobt_seed"])
    torch.manual_seed(meta["nobt_seed"])
    cond, _ = T.get_condition(copy.deepcopy(batch), "relation", tok)
    cond = cond.to(cuda_device)
    out, vio = model.sample(cond=cond, sampling_cfg={"name": "deterministic"}, cond_type="relation",
                            return_violation=True, use_backtrack=False)
    for k in ["label", "mask", "center_x", "center_y", "width", "height"]:
        np.testing.assert_array_equal(out[k].numpy(), z[f"nobt_gen_{k}"], err_msg=k)
    assert [vio["total"], vio["viorated"]] == z["nobt_violation"].tolist()
