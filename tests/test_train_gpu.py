"""GPU: training-step parity.  Gradients from the kernel tape (ralf_b200/autograd.py)
vs torch.autograd on the CPU oracle for the same seeded weights/batch; optimizer vs torch.optim.AdamW.
Dropout is off on both sides here (masks are generator-specific; tests/test_dropout_gpu.py covers the dropout ops and
the whole step with dropout on); BatchNorm uses batch statistics when the trunk trains."""
import math

import pytest
import torch
import torch.nn.functional as F

from tests import helpers

pytestmark = pytest.mark.gpu
GRAD_RTOL = 2e-3


def _model(dev, seed):
    from ralf_b200 import generator as G

    m = G.RALF(features=None, tokenizer=helpers.make_tokenizer(), dataset_name="cgl", max_seq_length=10, top_k=16,
               auxilary_task="uncond")
    m.load_state_dict(helpers.synth_weights("ralf_cgl", seed), strict=True)
    return m.to(dev)


def test_adamw_clip_matches_torch(cuda_device):
    from ralf_b200 import autograd as ag

    g = torch.Generator().manual_seed(0)
    params = [("a.weight", torch.randn(300, 40, generator=g)), ("b.bias", torch.randn(77, generator=g))]
    ps = ag.ParamStore([(n, p.to(cuda_device)) for n, p in params], [["a.weight"], ["b.bias"]], cuda_device)
    ref = [torch.nn.Parameter(p.clone().to(cuda_device)) for _, p in params]
    opt = torch.optim.AdamW([{"params": [ref[0]], "weight_decay": 1e-2, "lr": 1e-3},
                             {"params": [ref[1]], "weight_decay": 0.0, "lr": 1e-3}], betas=(0.9, 0.999), eps=1e-8)
    for step in range(1, 4):
        grads = [torch.randn(p.shape, generator=g).to(cuda_device) * 3 for _, p in params]
        for (n, _), gr, r in zip(params, grads, ref):
            ps.g(n).copy_(gr)
            r.grad = gr.clone()
        norm = ag.grad_norm(ps.flat_g)
        tn = torch.nn.utils.clip_grad_norm_(ref, 0.5)
        assert abs(float(norm) - float(tn)) <= 1e-5 * float(tn)
        ag.adamw_step(ps, [(1e-3, 1e-2), (1e-3, 0.0)], step, 0.5, norm)
        opt.step()
        for (n, _), r in zip(params, ref):
            assert (ps.p(n) - r.data).abs().max().item() <= 1e-6


def _oracle_loss_and_grads(sd0, batch, inputs, targets, pad_id, train_trunk, dtype, is_ralf=True, bn_train=None):
    """torch.autograd over the CPU oracle in `dtype` (float64 = ground truth); model.train() semantics for BatchNorm
    when the trunk trains, dropout off, FIDNet frozen (evaluated in fp32 like the product does)."""
    from oracle import ralf_oracle as O

    sd = {k: (v.clone().to(dtype) if v.dtype == torch.float32 else v.clone()) for k, v in sd0.items()}
    leaves = {}
    frozen = ("layout_encoer",) if train_trunk else ("encoder.extractor", "layout_encoer")
    for k, v in sd.items():
        if v.is_floating_point() and not k.startswith(frozen) and not k.endswith(".pe") and "running_" not in k:
            v.requires_grad_(True)
            leaves[k] = v
    orig_pos = O.pos_emb_2d
    O.pos_emb_2d = lambda h, w, d=256: orig_pos(h, w, d).to(dtype)
    O.BN_TRAIN = train_trunk if bn_train is None else bn_train  # reference fixture: trunk trainable, BatchNorm in eval mode
    try:
        if is_ralf:
            with torch.no_grad():
                cls = []
                for k in range(16):
                    lay = {key: batch["retrieved"][key][:, k] for key in ["center_x", "center_y", "width", "height", "label", "mask"]}
                    cls.append(O.fidnet_features(sd0, lay))
            refs = [O._feed_forward(sd, "layout_adapter", c.to(dtype)) for c in cls]
            ref = O._pe1d(sd, "pos_emb_1d", torch.stack(refs, 1))
        memory = O.encode_image(sd, inputs["image"].to(dtype))
    finally:
        O.BN_TRAIN = False
        O.pos_emb_2d = orig_pos
    if is_ralf:
        ca = O.fusion_attention(sd, memory, ref)
        mem = O._feed_forward(sd, "head", torch.cat([memory, ca, ref], 1))
    else:  # Autoreg baseline (models/autoreg.py:590-622): the image tokens go into the memory as they are
        mem = memory
    uc = O.constraint_encoder(sd, inputs["seq_layout_const"], inputs["seq_layout_const_pad_mask"])
    t = sd["task_emb.weight"]
    mem = torch.cat([mem + t[sd["flag_img"]], uc + t[sd["flag_user_const"]]], 1)
    S = inputs["seq"].shape[1]
    h = O._pe1d(sd, "decoder.pos_emb", sd["decoder.emb.weight"][inputs["seq"]])
    causal = torch.triu(torch.full((S, S), float("-inf"), dtype=dtype), 1)
    for i in range(6):
        h = O._dec_layer_prenorm(sd, f"decoder.transformer.layers.{i}", h, mem, causal, inputs["tgt_key_padding_mask"])
    logits = F.linear(O._ln(sd, "decoder.head.0", h), sd["decoder.head.1.weight"])
    loss = F.cross_entropy(logits.permute(0, 2, 1), targets["seq"], label_smoothing=0.1, ignore_index=pad_id)
    loss.backward()
    return float(loss.detach()), {k: v.grad.double() for k, v in leaves.items() if v.grad is not None}


def _tensor_errors(mine, ref):
    out = []
    for name, gref in ref.items():
        g = mine[name].double()
        out.append((name, (g - gref).abs().max().item() / (gref.abs().max().item() + 1e-30),
                    ((g - gref).norm() / (gref.norm() + 1e-30)).item(), gref.abs().max().item()))
    return out


@pytest.mark.parametrize("train_trunk", [False, True])
def test_training_gradients_match_oracle(cuda_device, train_trunk):
    """Gradients vs the float64 oracle.  ReLU kinks make max-abs comparison ill-conditioned: torch's OWN fp32
    gradients differ from its fp64 ones by up to 7 % (trunk frozen) / 14 % (trunk training, BatchNorm over a
    2-sample batch) on individual tensors for this very input (measured).  The bar is therefore:
    loss 1e-5; global gradient norm 1e-3; per-tensor relative L2 error <= the same bound the fp32 oracle meets
    against fp64 (3e-2) and a median per-tensor max-rel error <= 2e-3 -- and we must not be worse than the fp32
    oracle itself on the median by more than the spread below.  The median is CHAOTIC in the fp32 rounding ORDER of
    otherwise identical arithmetic on this 2-sample batch (BatchNorm statistics over two samples + ReLU kinks amplify a
    last-bit difference of one activation into a different gradient path).  Measured on one B200 with the library's A/B
    switches (profiles/run_r2_call_gradchaos.sh, profiles/r2_train.md), trunk training: 1.43e-4 (three MMAs per k-step,
    round-1 attention kernel) / 1.47e-4 (same, half-TMEM attention) / 1.68e-4, 1.74e-4 (256- / 2048-row BatchNorm partial
    sums) / 3.20e-4 (the default since the folded two-MMA GEMM: x_hi.w_hi and the cross terms accumulate separately) /
    3.70e-4 (folded, 256-row partial sums); trunk frozen: 0.86e-4 ... 6.3e-4 over the same switches, torch's own fp32
    gradients 4.5e-4.  Loss (1e-7) and gradient norm (2e-6) do not move.  The bound on the median is therefore 1e-3 (five
    times under the 2e-3 bar, above the measured spread); a real defect shows up in the loss / norm / rel-L2 asserts."""
    from oracle import synth
    from ralf_b200.train import TrainEngine

    torch.set_num_threads(8)
    model = _model(cuda_device, seed=21)
    sd = helpers.synth_weights("ralf_cgl", 21)
    batch = synth.synth_batch(2, 128, 128, 10, 16, 4, seed=9)
    inputs, targets = model.preprocess(batch)
    pad = model.tokenizer.name_to_id("pad")
    loss64, g64 = _oracle_loss_and_grads(sd, batch, inputs, targets, pad, train_trunk, torch.float64)
    _, g32 = _oracle_loss_and_grads(sd, batch, inputs, targets, pad, train_trunk, torch.float32)
    te = TrainEngine(model, train_trunk=train_trunk, dropout=0.0)  # the fp64 oracle has no dropout; see test_dropout_gpu.py
    te.ps.flat_g.zero_()
    loss, tape, _ = te.forward_loss(inputs, targets)
    tape.backward()
    torch.cuda.synchronize()
    assert abs(float(loss) - loss64) <= 1e-5 * abs(loss64), (float(loss), loss64)
    assert set(g64) == set(te.ps.offsets), set(te.ps.offsets) ^ set(g64)
    mine = {n: te.ps.g(n).cpu() for n in g64}
    ours = _tensor_errors(mine, g64)
    torch32 = _tensor_errors(g32, g64)
    import json, os, statistics
    os.makedirs(os.path.join(helpers.ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(helpers.ROOT, "gpurun_out", f"train_grad_errors_trunk{int(train_trunk)}.json"), "w") as f:
        json.dump({"ours_vs_fp64": sorted(ours, key=lambda e: -e[2])[:40], "torch_fp32_vs_fp64": sorted(torch32, key=lambda e: -e[2])[:40]}, f, indent=0)
    gn = math.sqrt(sum(float((mine[n].double() ** 2).sum()) for n in g64))
    gn64 = math.sqrt(sum(float((g64[n] ** 2).sum()) for n in g64))
    med_ours = statistics.median(e[1] for e in ours)
    med_t32 = statistics.median(e[1] for e in torch32)
    worst_l2 = max(e[2] for e in ours)
    print(f"train_trunk={train_trunk}: loss {float(loss):.6f} vs {loss64:.6f}; grad norm {gn:.6f} vs {gn64:.6f}; "
          f"median max-rel ours {med_ours:.2e} / torch-fp32 {med_t32:.2e}; worst rel-L2 ours {worst_l2:.2e} / "
          f"torch-fp32 {max(e[2] for e in torch32):.2e}")
    assert abs(gn - gn64) <= 1e-3 * gn64
    assert worst_l2 <= 3e-2, sorted(ours, key=lambda e: -e[2])[:5]
    assert med_ours <= 2e-3 and med_ours <= max(4 * med_t32, 1e-3)


def test_autoreg_baseline_training_gradients_match_oracle(cuda_device):
    """The Autoreg baseline class (BASELINE configs[0]'s model; configs/autoreg_*/*.sh train it) through the same tape:
    loss and every parameter gradient vs the float64 oracle, bar as in test_training_gradients_match_oracle; then two
    optimisation steps through the reference-style loop surface (train_loss().backward())."""
    import statistics

    from oracle import synth
    from ralf_b200 import generator as G
    from ralf_b200.train import TrainEngine

    torch.set_num_threads(8)
    model = G.ConcateAuxilaryTaskAutoreg(features=None, tokenizer=helpers.make_tokenizer(), auxilary_task="uncond")
    sd = helpers.synth_weights("autoreg_cgl", 23)
    model.load_state_dict(sd, strict=True)
    model = model.to(cuda_device)
    batch = synth.synth_batch(2, 128, 128, 10, 16, 4, seed=10)
    batch.pop("retrieved")
    inputs, targets = model.preprocess(batch)
    pad = model.tokenizer.name_to_id("pad")
    loss64, g64 = _oracle_loss_and_grads(sd, batch, inputs, targets, pad, True, torch.float64, is_ralf=False)
    te = TrainEngine(model, train_trunk=True, dropout=0.0)
    te.ps.flat_g.zero_()
    loss, tape, _ = te.forward_loss(inputs, targets)
    tape.backward()
    torch.cuda.synchronize()
    assert abs(float(loss) - loss64) <= 1e-5 * abs(loss64), (float(loss), loss64)
    assert set(g64) == set(te.ps.offsets), set(te.ps.offsets) ^ set(g64)
    ours = _tensor_errors({n: te.ps.g(n).cpu() for n in g64}, g64)
    gn = math.sqrt(sum(float((te.ps.g(n).double() ** 2).sum()) for n in g64))
    gn64 = math.sqrt(sum(float((g64[n] ** 2).sum()) for n in g64))
    assert abs(gn - gn64) <= 1e-3 * gn64
    assert max(e[2] for e in ours) <= 3e-2, sorted(ours, key=lambda e: -e[2])[:5]
    assert statistics.median(e[1] for e in ours) <= 2e-3
    losses = [float(te.train_step(inputs, targets)) for _ in range(3)]
    assert losses[-1] < losses[0], losses


def test_train_steps_reduce_loss_and_update_state_dict(cuda_device):
    from oracle import synth
    from ralf_b200.train import TrainEngine

    model = _model(cuda_device, seed=22)
    batch = synth.synth_batch(4, 128, 128, 10, 16, 4, seed=10)
    inputs, targets = model.preprocess(batch)
    te = TrainEngine(model, lr=1e-3, max_grad_norm=0.0)
    before = model.state_dict()["decoder.head.1.weight"].clone()
    losses = [float(te.train_step(inputs, targets)) for _ in range(6)]
    assert losses[-1] < losses[0], losses
    assert not torch.equal(before, model.state_dict()["decoder.head.1.weight"])
    assert float(te.last_grad_norm) > 0


def test_graph_replayed_train_step_equals_eager(cuda_device):
    """The captured CUDA graph of the whole step (forward, backward, clip, AdamW with device-side step scalars) must
    reproduce the eager steps: same kernels in the same order on the same data."""
    from oracle import synth
    from ralf_b200.train import TrainEngine

    batch = synth.synth_batch(4, 128, 128, 10, 16, 4, seed=12)
    ma, mb = _model(cuda_device, seed=23), _model(cuda_device, seed=23)
    inputs, targets = ma.preprocess(batch)
    ta, tb = TrainEngine(ma, lr=1e-3), TrainEngine(mb, lr=1e-3)
    la = [float(ta.train_step(inputs, targets)) for _ in range(4)]
    tb.capture(inputs, targets)                      # = step 1 (warm-up is a real step)
    lb = [float(tb.train_step_graph(inputs, targets)) for _ in range(3)]
    assert ta.step_count == tb.step_count == 4
    # Bit-identical: same kernels, same order, same data.  (Round 1 saw 1.5e-4 drift here; root cause = three reductions
    # through float atomics whose order varies run to run -- LayerNorm dgamma/dbeta partials in shared memory, the max-pool
    # scatter, the embedding scatter-add -- amplified by Adam's normalisation of near-zero gradients.  All three are now
    # deterministic gathers / fixed-order reductions, so eager-vs-eager and graph-vs-eager agree to the bit.)
    assert la[1:] == lb, (la, lb)
    assert torch.equal(ta.ps.flat_p, tb.ps.flat_p)


def test_train_step_is_run_to_run_deterministic(cuda_device):
    """Two eager runs of the same steps on twin models give bit-identical losses, gradients and parameters (no atomics
    anywhere in forward / backward / optimiser), dropout on."""
    from oracle import synth
    from ralf_b200.train import TrainEngine

    batch = synth.synth_batch(4, 128, 128, 10, 16, 4, seed=13)
    runs = []
    for _ in range(2):
        m = _model(cuda_device, seed=24)
        inputs, targets = m.preprocess(batch)
        te = TrainEngine(m, lr=1e-3, seed=5)
        losses = [float(te.train_step(inputs, targets)) for _ in range(3)]
        runs.append((losses, te.ps.flat_g.clone(), te.ps.flat_p.clone()))
    assert runs[0][0] == runs[1][0]
    assert torch.equal(runs[0][1], runs[1][1]) and torch.equal(runs[0][2], runs[1][2])


def test_eval_after_fused_train_step_sees_the_updated_weights(cuda_device):
    """evaluate(); model.train(); train_step(); evaluate() -- the second evaluation must run on the
    UPDATED weights (the cached inference Engine is dropped by TrainEngine after every step, eager or graph-replayed),
    i.e. equal what a freshly built Engine computes, and differ from the first."""
    from oracle import synth
    from ralf_b200.engine import Engine
    from ralf_b200.train import TrainEngine

    model = _model(cuda_device, seed=26)
    batch = synth.synth_batch(4, 128, 128, 10, 16, 4, seed=15)
    inputs, targets = model.preprocess(batch)
    inputs = {k: (v.to(cuda_device) if torch.is_tensor(v) else v) for k, v in inputs.items()}
    targets = {k: v.to(cuda_device) for k, v in targets.items()}

    def eval_loss():
        model.eval()
        with torch.no_grad():
            _, lv = model.train_loss(inputs, targets, test=True)
        return float(lv["nll_loss"])

    l0 = eval_loss()
    model.train()
    te = TrainEngine(model, lr=1e-2, max_grad_norm=0.0)
    te.train_step(inputs, targets)
    l1 = eval_loss()
    fresh = Engine(model.state_dict(), cuda_device, is_ralf=True, top_k=model.top_k)
    model._engine = fresh
    assert eval_loss() == l1 and l1 != l0
    model.train()
    te.capture(inputs, targets)
    te.train_step_graph(inputs, targets)
    l2 = eval_loss()
    model._engine = Engine(model.state_dict(), cuda_device, is_ralf=True, top_k=model.top_k)
    assert eval_loss() == l2 and l2 != l1


def test_reference_train_loop_runs_unchanged(cuda_device):
    """The reference's own loop (train.py:440-454): optimizer built from optim_groups, zero_grad, train_loss,
    loss.backward(), clip_grad_norm_, optimizer.step().  Must give the same losses / parameters as the fused
    TrainEngine.train_step on a twin model (same kernels for forward/backward; torch's AdamW vs ours)."""
    from oracle import synth
    from ralf_b200.train import TrainEngine

    batch = synth.synth_batch(4, 128, 128, 10, 16, 4, seed=14)
    ma, mb = _model(cuda_device, seed=25), _model(cuda_device, seed=25)
    ma.train()
    mb.train()
    inputs, targets = ma.preprocess(batch)
    inputs = {k: (v.to(cuda_device) if torch.is_tensor(v) else v) for k, v in inputs.items()}
    targets = {k: v.to(cuda_device) for k, v in targets.items()}
    opt = torch.optim.AdamW(ma.optim_groups(base_lr=1e-3, weight_decay=1e-4, custom_lr={"encoder.extractor.body": 1e-4}),
                            betas=(0.9, 0.999))
    la = []
    for _ in range(3):
        ma.zero_grad()
        outputs, losses = ma.train_loss(inputs, targets)
        loss = sum(losses.values())
        assert loss.requires_grad and outputs["logits"].shape[:2] == inputs["seq"].shape
        loss.backward()
        gn = torch.nn.utils.clip_grad_norm_(ma.parameters(), 0.1)
        opt.step()
        la.append(float(loss))
    tb = TrainEngine(mb, lr=1e-3, weight_decay=1e-4, body_lr_scale=0.1, max_grad_norm=0.1)
    lb = [float(tb.train_step(inputs, targets)) for _ in range(3)]
    assert float(gn) > 0 and all(p.grad is not None for p in ma.parameters() if p.requires_grad)
    assert max(abs(a - b) for a, b in zip(la, lb)) <= 1e-4 * abs(la[0]), (la, lb)
    pa = torch.cat([p.detach().reshape(-1) for n, p in sorted(ma.named_parameters()) if p.requires_grad])
    pb = torch.cat([p.detach().reshape(-1) for n, p in sorted(mb.named_parameters()) if p.requires_grad])
    assert (pa - pb).abs().max().item() <= 5e-3 * pa.abs().max().item()
    # evaluate() path: eval + no_grad gives a plain value and sees the UPDATED parameters
    ma.eval()
    with torch.no_grad():
        _, lv = ma.train_loss(inputs, targets, test=True)
    assert not lv["nll_loss"].requires_grad and math.isfinite(float(lv["nll_loss"]))
