"""GPU: dropout of the training step.  The masks are counter-based (csrc/common.cuh), so they cannot be
bit-identical to torch's Philox stream; instead the tests export the mask the kernels use (ralf_dropout_mask) and check
every dropout op against a float64 torch reference that applies THAT mask, the mask statistics, and the whole training
step with dropout on against a finite difference of its own loss."""
import math

import numpy as np
import pytest
import torch

from tests import helpers

pytestmark = pytest.mark.gpu
P = 0.1


def _seed(dev, v=1234567):
    return torch.tensor([v], dtype=torch.int64, device=dev)


def test_mask_statistics_and_determinism(cuda_device):
    from ralf_b200 import ops

    n = 1 << 20
    m1 = ops.dropout_mask(_seed(cuda_device), 3, P, n)
    assert torch.equal(m1, ops.dropout_mask(_seed(cuda_device), 3, P, n))
    keep = m1.float().mean().item()
    assert abs(keep - (1 - P)) < 4 * math.sqrt(P * (1 - P) / n) + 1e-4, keep
    for other in (ops.dropout_mask(_seed(cuda_device), 4, P, n), ops.dropout_mask(_seed(cuda_device, 1234568), 3, P, n)):
        agree = (other == m1).float().mean().item()   # independent masks agree with prob. keep^2 + drop^2
        assert abs(agree - ((1 - P) ** 2 + P ** 2)) < 2e-3, agree
    # no short-range structure: lag-1 autocorrelation of the drop indicator ~ 0
    d = 1.0 - m1.float()
    ac = ((d[1:] - P) * (d[:-1] - P)).mean().item() / (P * (1 - P))
    assert abs(ac) < 5e-3, ac


def test_elementwise_dropout_matches_exported_mask(cuda_device):
    from ralf_b200 import ops

    g = torch.Generator().manual_seed(0)
    M, C = 77, 256
    x = torch.randn((M, C), generator=g).to(cuda_device)
    res = torch.randn((M, C), generator=g).to(cuda_device)
    seed = _seed(cuda_device)
    mask = ops.dropout_mask(seed, 9, P, M * C).view(M, C).float()
    want = res + x * mask / (1 - P)
    out = torch.empty_like(x)
    outs = torch.empty((2, M, C), dtype=torch.bfloat16, device=cuda_device)
    ops.dropout(seed, 9, P, x_f32=x, res=res, out_f32=out, out_split=outs)
    assert torch.allclose(out, want, rtol=1e-6, atol=1e-6)
    assert torch.allclose(ops.unsplit(outs), want, rtol=2e-5, atol=2e-5)
    # split source, in place (the FFN activation) and the backward use (gradient masked in place)
    xs = ops.split_bf16(x)
    ref = ops.unsplit(xs) * mask / (1 - P)
    ops.dropout(seed, 9, P, x_split=xs, out_split=xs)
    assert torch.allclose(ops.unsplit(xs), ref, rtol=2e-5, atol=2e-5)
    gbuf = res.clone()
    ops.dropout(seed, 9, P, x_f32=gbuf, out_f32=gbuf)
    assert torch.allclose(gbuf, res * mask / (1 - P), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("causal,use_mask,Tq,Tk,dh,H", [(False, False, 70, 70, 32, 8), (True, True, 50, 50, 32, 8),
                                                        (False, False, 50, 131, 32, 8)])
def test_attention_dropout_forward_backward(cuda_device, causal, use_mask, Tq, Tk, dh, H):
    from ralf_b200 import autograd as ag
    from ralf_b200 import ops

    g = torch.Generator().manual_seed(1)
    B, Dm = 3, H * dh
    q = torch.randn((B * Tq, Dm), generator=g).to(cuda_device)
    k = torch.randn((B * Tk, Dm), generator=g).to(cuda_device)
    v = torch.randn((B * Tk, Dm), generator=g).to(cuda_device)
    dO = torch.randn((B * Tq, Dm), generator=g).to(cuda_device)
    pad = None
    if use_mask:
        pad = torch.zeros((B, Tk), dtype=torch.uint8, device=cuda_device)
        pad[1, Tk - 7:] = 1
        pad[2, Tk - 20:] = 1
    seed, site = _seed(cuda_device, 99), 5
    out = ops.attention(q, k, v, B, H, Tq, Tk, dh, mask=pad, causal=causal, dropout=(seed, site, P))
    mask = ops.dropout_mask(seed, site, P, B * H * Tq * Tk).view(B, H, Tq, Tk).double()
    # float64 reference with the SAME mask
    q64, k64, v64 = (t.double().clone().requires_grad_(True) for t in (q, k, v))
    sp = lambda t, T: t.view(B, T, H, dh).transpose(1, 2)
    s = sp(q64, Tq) @ sp(k64, Tk).transpose(-1, -2) * dh ** -0.5
    if causal:
        s = s.masked_fill(torch.triu(torch.ones(Tq, Tk, dtype=torch.bool, device=cuda_device), 1), float("-inf"))
    if pad is not None:
        s = s.masked_fill(pad.bool()[:, None, None, :], float("-inf"))
    o = ((torch.softmax(s, -1) * mask / (1 - P)) @ sp(v64, Tk)).transpose(1, 2).reshape(B * Tq, Dm)
    o.backward(dO.double())
    rel = lambda a, b: (a.double() - b).abs().max().item() / b.abs().max().item()
    assert rel(ops.unsplit(out), o.detach()) < 1e-4
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    ag._attention_bwd(q, q.stride(0), k, v, k.stride(0), pad, B, H, Tq, Tk, dh, causal, out, dO, dq, dq.stride(0), dk, dv,
                      dk.stride(0), dropout=(seed, site, P))
    assert rel(dq, q64.grad) < 1e-4 and rel(dk, k64.grad) < 1e-4 and rel(dv, v64.grad) < 1e-4


def test_train_step_with_dropout_gradient_matches_finite_difference(cuda_device):
    """Whole step, dropout on: the directional derivative of the loss along the computed gradient, with the step's masks
    held fixed (same seed, same sites), must equal |g|^2 -- checks every dropout site's forward/backward pairing."""
    from oracle import synth
    from ralf_b200 import generator as G
    from ralf_b200.train import TrainEngine

    model = G.RALF(features=None, tokenizer=helpers.make_tokenizer(), dataset_name="cgl", max_seq_length=10, top_k=16)
    model.load_state_dict(helpers.synth_weights("ralf_cgl", 31), strict=True)
    model.to(cuda_device).train()
    batch = synth.synth_batch(4, 128, 128, 10, 16, 4, seed=15)
    inputs, targets = model.preprocess(batch)
    te = TrainEngine(model, dropout=P, seed=7)
    te.step_count = 1
    te._set_step_seed()

    def loss_at(delta=None):
        if delta is not None:
            te.ps.flat_p.add_(delta)
        te.refresh_operands()
        te.ps.flat_g.zero_()
        loss, tape, _ = te.forward_loss(inputs, targets)
        if delta is not None:
            te.ps.flat_p.sub_(delta)
        return loss, tape

    loss0, tape = loss_at()
    tape.backward()
    g = te.ps.flat_g.clone()
    l_nodrop = None
    gn2 = float((g.double() ** 2).sum())
    assert math.isfinite(float(loss0)) and gn2 > 0
    eps = 2e-3 / math.sqrt(gn2)                  # step of 2e-3 in parameter norm along the gradient
    lp, _ = loss_at(g * eps)
    lm, _ = loss_at(-g * eps)
    fd = (float(lp) - float(lm)) / (2 * eps)
    assert abs(fd - gn2) <= 0.03 * gn2, (fd, gn2)
    # and dropout really is on: the loss differs from the dropout-free loss of the same parameters, masks change per step
    te2 = TrainEngine(model, dropout=0.0)
    l_nodrop, _, _ = te2.forward_loss(inputs, targets)
    assert abs(float(l_nodrop) - float(loss0)) > 1e-4
    te.step_count = 2
    te._set_step_seed()
    l_other, _ = loss_at()
    assert abs(float(l_other) - float(loss0)) > 1e-5


def test_device_mask_equals_cpu_restatement(cuda_device):
    """The mask the kernels derive on the device equals oracle/dropout_oracle.py (itself checked on the CPU against the
    host compilation of the same __host__ __device__ functions, tests/test_dropout_rng_cpu.py)."""
    from oracle import dropout_oracle as D
    from ralf_b200 import ops

    n = 50000
    for seed, site, p in [(1234567, 3, 0.1), (-5, 77, 0.25), (2 ** 62 + 11, 0, 0.1), (0, 200, 0.5)]:
        dev = ops.dropout_mask(_seed(cuda_device, seed), site, p, n).cpu().numpy()
        np.testing.assert_array_equal(dev, D.keep_mask(seed, site, p, n))
