"""GPU parity of the engine (ralf_b200/engine.py -> C ABI -> sm_100a kernels) against (a) the reference's own
outputs in tests/golden/*.npz and (b) the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): token ids bit-exact; logits within 1e-3 relative.  "Relative" is measured
against the logit scale max|logits| (element-wise relative error is ill-defined near zero --
two fp32 PyTorch paths already differ by 4.8e-2 element-wise)."""
import numpy as np
import pytest
import torch

from tests import helpers

pytestmark = pytest.mark.gpu

CASES = [("ralf_cgl_256", "ralf_cgl", True), ("ralf_cgl_350x240", "ralf_cgl", True),
         ("autoreg_cgl_350x240", "autoreg_cgl", False)]
LOGIT_RTOL = 1e-3


def _relerr(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() / np.abs(b).max()


def _engine(schema, seed, is_ralf, dev):
    from ralf_b200.engine import Engine

    return Engine(helpers.synth_weights(schema, seed), dev, is_ralf=is_ralf)


def test_resnet_fpn_matches_oracle(cuda_device):
    from oracle import ralf_oracle as O

    z, meta = helpers.load_golden("ralf_cgl_256")
    sd = helpers.synth_weights("ralf_cgl", meta["seed"])
    eng = _engine("ralf_cgl", meta["seed"], True, cuda_device)
    img = helpers.image4(helpers.synth_batch(meta))
    with torch.no_grad():
        f = O.resnet_fpn(sd, img)
        ref = f.flatten(2).transpose(1, 2) + O.pos_emb_2d(f.shape[2], f.shape[3], 256)[None]
    tokens, h, w = eng.resnet_fpn(img.to(cuda_device))
    assert (h, w) == tuple(f.shape[2:])
    err = _relerr(tokens.view(ref.shape).cpu().numpy(), ref.numpy())
    assert err < 1e-4, err


@pytest.mark.parametrize("name,schema,is_ralf", CASES)
def test_engine_matches_reference_golden(cuda_device, name, schema, is_ralf):
    z, meta = helpers.load_golden(name)
    eng = _engine(schema, meta["seed"], is_ralf, cuda_device)
    batch = helpers.synth_batch(meta)
    tok = helpers.make_tokenizer()
    B = meta["B"]
    mem, mem_s = eng.encode(helpers.image4(batch), batch.get("retrieved"), torch.from_numpy(z["seq_layout_const"]),
                            torch.from_numpy(z["seq_layout_const_pad_mask"]))
    assert tuple(mem.shape) == z["memory"].shape
    e_mem = _relerr(mem.cpu().numpy(), z["memory"])
    assert e_mem < LOGIT_RTOL, e_mem
    Mlen = mem.shape[1]
    logits = eng.decoder_logits(torch.from_numpy(z["seq_in"]), torch.from_numpy(z["tgt_key_padding_mask"]), mem_s, B, Mlen)
    e_log = _relerr(logits.cpu().numpy(), z["logits"])
    assert e_log < LOGIT_RTOL, e_log
    sp = meta["special"]
    seq, step_logits = eng.generate(mem_s, B, Mlen, tok.token_mask, sp["bos"], sp["pad"], tok.max_token_length,
                                    return_logits=True)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(seq.cpu().numpy(), z["gen_seq"])  # bit-exact token ids
    ref = z["gen_step_logits"]
    fin = np.isfinite(ref)
    e_step = _relerr(step_logits.cpu().numpy()[fin], ref[fin])
    assert e_step < LOGIT_RTOL, e_step
    print(f"{name}: memory {e_mem:.2e} logits {e_log:.2e} greedy-step logits {e_step:.2e}")


def test_engine_batch_matches_oracle_with_margin(cuda_device):
    """B=8 random canvases vs the CPU oracle: token ids equal; a divergence is tolerated only where the oracle's
    own top-2 margin at that step is below the logit tolerance (first-divergence rule)."""
    from oracle import ralf_oracle as O
    from oracle import synth

    sd = helpers.synth_weights("ralf_cgl", 7)
    eng = _engine("ralf_cgl", 7, True, cuda_device)
    tok = helpers.make_tokenizer()
    batch = synth.synth_batch(8, 256, 256, 10, 16, 4, seed=123)
    z, meta = helpers.load_golden("ralf_cgl_256")
    sc = torch.from_numpy(z["seq_layout_const"])[:1].expand(8, -1).contiguous()
    sp_ = torch.zeros_like(sc, dtype=torch.bool)
    sp = meta["special"]
    torch.set_num_threads(8)
    with torch.no_grad():
        mem_o = O.encode_ralf_memory(sd, helpers.image4(batch), {k: v.float() for k, v in batch["retrieved"].items()}, sc, sp_)
        seq_o, lg_o = O.greedy_sample(sd, mem_o, tok.token_mask, sp["bos"], sp["pad"], tok.max_token_length, return_logits=True)
    mem, mem_s = eng.encode(helpers.image4(batch), batch["retrieved"], sc, sp_)
    assert _relerr(mem.cpu().numpy(), mem_o.numpy()) < LOGIT_RTOL
    seq = eng.generate(mem_s, 8, mem.shape[1], tok.token_mask, sp["bos"], sp["pad"], tok.max_token_length).cpu()
    for b in range(8):
        diff = (seq[b] != seq_o[b]).nonzero()
        if len(diff) == 0:
            continue
        t = int(diff[0])
        top2 = torch.topk(lg_o[b, t], 2).values
        margin = float(top2[0] - top2[1]) / float(lg_o[b, t][torch.isfinite(lg_o[b, t])].abs().max())
        assert margin < LOGIT_RTOL, f"canvas {b} diverges at step {t} with oracle margin {margin:.3e}"


def test_engine_matches_reference_golden_pku(cuda_device):
    """BASELINE configs[2] names PKU (3 labels, vocabulary 518): memory within tolerance of the reference's own output;
    token ids equal, a divergence tolerated only where the oracle's top-2 margin at that step is below the logit
    tolerance (first-divergence rule)."""
    from oracle import ralf_oracle as O
    from oracle import synth
    from ralf_b200.engine import Engine

    z, meta = helpers.load_golden("ralf_pku_128")
    sd = synth.synth_state_dict(helpers.load_schema("ralf_pku"), seed=meta["weights_seed"])
    tok = helpers.make_tokenizer("pku")
    B = meta["B"]
    batch = synth.synth_batch(B, meta["H"], meta["W"], meta["E"], meta["K"], tok.N_label, seed=meta["seed"])
    sc, spm = torch.from_numpy(z["seq_layout_const"]), torch.from_numpy(z["seq_layout_const_pad_mask"])
    eng = Engine(sd, cuda_device, is_ralf=True)
    mem, mem_s = eng.encode(helpers.image4(batch), batch["retrieved"], sc, spm)
    e_mem = _relerr(mem.cpu().numpy(), z["memory"])
    assert e_mem < LOGIT_RTOL, e_mem
    sp = meta["special"]
    seq = eng.generate(mem_s, B, mem.shape[1], tok.token_mask, sp["bos"], sp["pad"], tok.max_token_length).cpu()
    ref = torch.from_numpy(z["gen_seq"])
    if not torch.equal(seq, ref):
        torch.set_num_threads(8)
        with torch.no_grad():
            _, lg_o = O.greedy_sample(sd, torch.from_numpy(z["memory"]), tok.token_mask, sp["bos"], sp["pad"],
                                      tok.max_token_length, return_logits=True)
        for b in range(B):
            diff = (seq[b] != ref[b]).nonzero()
            if len(diff) == 0:
                continue
            t = int(diff[0])
            top2 = torch.topk(lg_o[b, t], 2).values
            margin = float(top2[0] - top2[1]) / float(lg_o[b, t][torch.isfinite(lg_o[b, t])].abs().max())
            assert margin < LOGIT_RTOL, f"canvas {b} diverges at step {t} with oracle margin {margin:.3e}"


def test_engine_matches_oracle_at_the_bench_shape_e12(cuda_device):
    """The bench decodes max_seq_length = 12 (60 tokens), one element more than the reference can be built with (its
    constraint vocabulary has 11 element letters), so this shape is pinned to the oracle: memory within tolerance,
    greedy tokens equal under the first-divergence rule, through the drop-in class with its own E = 12 schema."""
    from oracle import ralf_oracle as O
    from oracle import synth
    from ralf_b200 import generator as G

    E, B = 12, 4
    tok = helpers.make_tokenizer(max_seq_length=E)
    model = G.RALF(features=None, tokenizer=tok, dataset_name="cgl", max_seq_length=E, top_k=16, auxilary_task="uncond")
    schema = {k: {"shape": list(v.shape), "dtype": str(v.dtype).replace("torch.", "")} for k, v in model.state_dict().items()}
    sd = synth.synth_state_dict(schema, seed=0)  # what bench.py loads
    model.load_state_dict(sd, strict=True)
    model = model.eval().to(cuda_device)
    batch = synth.synth_batch(B, 256, 256, E, 16, 4, seed=31)
    const = model.preprocessor(G.ConditionalInputs(image=helpers.image4(batch)))
    sp = model.special_token_ids
    torch.set_num_threads(8)
    with torch.no_grad():
        mem_o = O.encode_ralf_memory(sd, helpers.image4(batch), {k: v.float() for k, v in batch["retrieved"].items()},
                                     const["seq"], const["pad_mask"])
        seq_o, lg_o = O.greedy_sample(sd, mem_o, tok.token_mask, sp["bos"], sp["pad"], tok.max_token_length, return_logits=True)
    assert seq_o.shape[1] == 60
    eng = model.engine()
    mem, mem_s = eng.encode(helpers.image4(batch), batch["retrieved"], const["seq"], const["pad_mask"])
    assert _relerr(mem.cpu().numpy(), mem_o.numpy()) < LOGIT_RTOL
    seq = eng.generate(mem_s, B, mem.shape[1], tok.token_mask, sp["bos"], sp["pad"], tok.max_token_length).cpu()
    for b in range(B):
        diff = (seq[b] != seq_o[b]).nonzero()
        if len(diff) == 0:
            continue
        t = int(diff[0])
        top2 = torch.topk(lg_o[b, t], 2).values
        margin = float(top2[0] - top2[1]) / float(lg_o[b, t][torch.isfinite(lg_o[b, t])].abs().max())
        assert margin < LOGIT_RTOL, f"canvas {b} diverges at step {t} with oracle margin {margin:.3e}"


def test_engine_matches_reference_golden_at_maximum_length(cuda_device):
    """max_seq_length = 11 (the reference's upper limit, 55 tokens), every canvas and exemplar full: memory and
    teacher-forced logits within tolerance of the reference's own output, greedy tokens equal (first-divergence rule)."""
    from oracle import ralf_oracle as O
    from ralf_b200.engine import Engine
    from tests.test_oracle_golden import _e11_inputs

    z, meta, batch, sd = _e11_inputs()
    tok = helpers.make_tokenizer(max_seq_length=11)
    B = meta["B"]
    sc, spm = torch.from_numpy(z["seq_layout_const"]), torch.from_numpy(z["seq_layout_const_pad_mask"])
    eng = Engine(sd, cuda_device, is_ralf=True)
    mem, mem_s = eng.encode(helpers.image4(batch), batch["retrieved"], sc, spm)
    assert _relerr(mem.cpu().numpy(), z["memory"]) < LOGIT_RTOL
    seq_in = torch.from_numpy(z["seq_in"])
    logits = eng.decoder_logits(seq_in, seq_in == meta["special"]["pad"], mem_s, B, mem.shape[1])
    assert _relerr(logits.cpu().numpy(), z["logits"]) < LOGIT_RTOL
    sp = meta["special"]
    seq = eng.generate(mem_s, B, mem.shape[1], tok.token_mask, sp["bos"], sp["pad"], tok.max_token_length).cpu()
    ref = torch.from_numpy(z["gen_seq"])
    if not torch.equal(seq, ref):
        torch.set_num_threads(8)
        with torch.no_grad():
            _, lg_o = O.greedy_sample(sd, torch.from_numpy(z["memory"]), tok.token_mask, sp["bos"], sp["pad"],
                                      tok.max_token_length, return_logits=True)
        for b in range(B):
            diff = (seq[b] != ref[b]).nonzero()
            if len(diff) == 0:
                continue
            t = int(diff[0])
            top2 = torch.topk(lg_o[b, t], 2).values
            margin = float(top2[0] - top2[1]) / float(lg_o[b, t][torch.isfinite(lg_o[b, t])].abs().max())
            assert margin < LOGIT_RTOL, f"canvas {b} diverges at step {t} with oracle margin {margin:.3e}"


@pytest.mark.parametrize("B,Mlen", [(3, 40), (16, 532), (37, 532)])
def test_fused_decode_chain_equals_per_op_decode(cuda_device, monkeypatch, B, Mlen):
    """ralf_decode_chain (the row-local ops of a decoder-layer step fused into 3 kernels per layer, csrc/decode_chain.cu)
    against the per-op launches it replaces, over a whole greedy loop from a random memory: same LayerNorm arithmetic and
    the same products; by default also the same MMA order (bit-identical, measured 0.0); with RALF_CHAIN_ACC=3 (one
    accumulator per bf16x3 pass, summed in the epilogue) the fp32 summation order differs: step logits agree to <= 1e-5 of
    scale (measured 3.6e-6 .. 5.5e-6) and the token ids are identical.  B covers a partial 16-canvas tile, one full
    tile and two full + one partial."""
    from ralf_b200 import engine as E

    eng = _engine("ralf_cgl", 7, True, cuda_device)
    tok = helpers.make_tokenizer()
    g = torch.Generator().manual_seed(100 + B)
    mem = torch.randn(B * Mlen, 256, generator=g).to(cuda_device)
    from ralf_b200 import ops

    mem_s = ops.split_bf16(mem)
    outs = {}
    for chain in (False, True):
        monkeypatch.setattr(E, "DECODE_CHAIN", chain)
        seq, lg = eng.generate(mem_s, B, Mlen, tok.token_mask, 517, 516, tok.max_token_length, return_logits=True)
        torch.cuda.synchronize()
        outs[chain] = (seq.cpu().numpy(), lg.cpu().numpy())
    np.testing.assert_array_equal(outs[True][0], outs[False][0])
    a, b = outs[True][1], outs[False][1]
    assert np.isfinite(a).all()
    err = np.abs(a - b).max() / np.abs(b).max()
    assert err <= 1e-5, err
    print(f"B={B}: fused-vs-per-op step logits {err:.2e}")


@pytest.mark.parametrize("B,Mlen", [(3, 40), (37, 532), (130, 64)])
def test_decode_with_layernorm_in_the_residual_gemm_equals_per_op_decode(cuda_device, monkeypatch, B, Mlen):
    """RALF_DECODE_RESLN (opt-in): the decode step with every LayerNorm but the first computed in the epilogue of the residual
    GEMM in front of it (ralf_gemm_res_ln, cluster kernel) against the per-op launches, over a whole greedy loop: same
    products and residual adds (bit-identical x), LayerNorm statistics summed in a different order -> step logits within
    1e-5 of scale and identical token ids.  130 canvases: two row tiles, the second partial."""
    from ralf_b200 import engine as E
    from ralf_b200 import ops

    eng = _engine("ralf_cgl", 7, True, cuda_device)
    tok = helpers.make_tokenizer()
    g = torch.Generator().manual_seed(200 + B)
    mem_s = ops.split_bf16(torch.randn(B * Mlen, 256, generator=g).to(cuda_device))
    outs = {}
    for fused in (False, True):
        monkeypatch.setattr(E, "DECODE_RESLN", fused)
        seq, lg = eng.generate(mem_s, B, Mlen, tok.token_mask, 517, 516, tok.max_token_length, return_logits=True)
        torch.cuda.synchronize()
        outs[fused] = (seq.cpu().numpy(), lg.cpu().numpy())
    np.testing.assert_array_equal(outs[True][0], outs[False][0])
    a, b = outs[True][1], outs[False][1]
    assert np.isfinite(a).all()
    err = np.abs(a - b).max() / np.abs(b).max()
    assert err <= 1e-5, err


def test_engine_matches_reference_golden_at_batch_32(cuda_device):
    """B = 32 golden of the UNMODIFIED reference (fixture layout: test_oracle_golden.py): memory and logits of
    the first 4 canvases within 1e-3 of scale, logit summaries of all 32 within 1e-3, greedy token ids and decoded layouts
    of all 32 bit-exact."""
    z, meta = helpers.load_golden("ralf_cgl_b32_128")
    eng = _engine("ralf_cgl", meta["seed"], True, cuda_device)
    batch = helpers.synth_batch(meta)
    tok = helpers.make_tokenizer()
    B, full = meta["B"], meta["full"]
    mem, mem_s = eng.encode(helpers.image4(batch), batch["retrieved"], torch.from_numpy(z["seq_layout_const"]),
                            torch.from_numpy(z["seq_layout_const_pad_mask"]))
    assert _relerr(mem[:full].cpu().numpy(), z["memory_head"]) < LOGIT_RTOL
    assert _relerr(mem.norm(dim=-1).cpu().numpy(), z["memory_row_norm"]) < LOGIT_RTOL
    Mlen = mem.shape[1]
    lg = eng.decoder_logits(torch.from_numpy(z["seq_in"]), torch.from_numpy(z["tgt_key_padding_mask"]), mem_s, B, Mlen)
    assert _relerr(lg[:full].cpu().numpy(), z["logits_head"]) < LOGIT_RTOL
    scale = np.abs(z["logits_head"]).max()
    assert np.abs(lg.max(-1).values.cpu().numpy() - z["logits_max"]).max() <= LOGIT_RTOL * scale
    assert np.abs(torch.logsumexp(lg, -1).cpu().numpy() - z["logits_lse"]).max() <= LOGIT_RTOL * np.abs(z["logits_lse"]).max()
    sp = meta["special"]
    seq = eng.generate(mem_s, B, Mlen, tok.token_mask, sp["bos"], sp["pad"], tok.max_token_length)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(seq.cpu().numpy(), z["gen_seq"])
    dec = tok.decode(seq.cpu())
    for k in ["label", "mask", "center_x", "center_y", "width", "height"]:
        np.testing.assert_array_equal(dec[k].numpy(), z["gen_" + k])
