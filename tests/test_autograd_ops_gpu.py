"""GPU: each tape op's backward (ralf_b200/autograd.py, train_conv.py) against torch.autograd on the same values
(torch on the GPU is the test reference here, never the product path)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
TOL = 2e-4
# the torch reference must be true fp32: TF32 convolutions / matmuls are ~1e-3 off
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def _rel(a, b):
    return (a.double() - b.double()).abs().max().item() / (b.double().abs().max().item() + 1e-30)


def _ps(dev, params):
    from ralf_b200 import autograd as ag

    named = [(n, p.to(dev)) for n, p in params.items()]
    return ag.ParamStore(named, [[n for n, _ in named]], dev)


def _node(x, need_grad=True, split=True, f32=True):
    from ralf_b200 import autograd as ag
    from ralf_b200 import ops

    return ag.Node(x.shape[0], x.shape[1], x.contiguous() if f32 else None, ops.split_bf16(x) if split else None, need_grad)


@pytest.mark.parametrize("M,K,N,act,use_res", [(100, 256, 1024, "relu", False), (100, 1024, 256, None, True),
                                              (128, 256, 768, None, False), (37, 256, 519, None, False)])
def test_linear_backward(cuda_device, M, K, N, act, use_res):
    from ralf_b200 import autograd as ag

    g = torch.Generator().manual_seed(M + N)
    W = torch.randn(N, K, generator=g) / math.sqrt(K)
    b = torch.randn(N, generator=g) * 0.1
    x = torch.randn(M, K, generator=g).to(cuda_device)
    r = torch.randn(M, N, generator=g).to(cuda_device)
    dy = torch.randn(M, N, generator=g).to(cuda_device)
    ps = _ps(cuda_device, {"w.weight": W, "w.bias": b})
    ps.register_gemm_weight("w", "w.weight")
    ps.refresh_operands()
    tape = ag.Tape()
    xn = _node(x)
    rn = _node(r, split=False) if use_res else None
    y = ag.linear(tape, ps, xn, "w", "w.bias", act=act, res=rn, want_f32=True)
    # torch reference
    xt = x.clone().requires_grad_(True)
    Wt = W.to(cuda_device).requires_grad_(True)
    bt = b.to(cuda_device).requires_grad_(True)
    rt = r.clone().requires_grad_(True)
    yt = F.linear(xt, Wt, bt)
    if act == "relu":
        yt = torch.relu(yt)
    if use_res:
        yt = yt + rt
    assert _rel(y.f32, yt.detach()) < 1e-4
    yt.backward(dy)
    y.grad = dy.clone()
    tape.backward()
    torch.cuda.synchronize()
    assert _rel(ps.g("w.weight"), Wt.grad) < TOL, "dW"
    assert _rel(ps.g("w.bias"), bt.grad) < TOL, "db"
    assert _rel(xn.grad, xt.grad) < TOL, "dx"
    if use_res:
        assert _rel(rn.grad, rt.grad) < TOL, "dres"


def test_layernorm_and_gelu_backward(cuda_device):
    from ralf_b200 import autograd as ag

    g = torch.Generator().manual_seed(3)
    M, Dm = 100, 256
    x = (torch.randn(M, Dm, generator=g) * 2 + 0.5).to(cuda_device)
    gm = (torch.rand(Dm, generator=g) + 0.5)
    bt = torch.randn(Dm, generator=g) * 0.1
    dy = torch.randn(M, Dm, generator=g).to(cuda_device)
    prev = torch.randn(M, Dm, generator=g).to(cuda_device)
    ps = _ps(cuda_device, {"n.weight": gm, "n.bias": bt})
    tape = ag.Tape()
    xn = _node(x, split=False)
    xn.grad = prev.clone()   # an earlier consumer already contributed (residual stream)
    y = ag.layernorm(tape, ps, xn, "n", want_f32=True)
    xt = x.clone().requires_grad_(True)
    gt, btt = gm.to(cuda_device).requires_grad_(True), bt.to(cuda_device).requires_grad_(True)
    yt = F.layer_norm(xt, (Dm,), gt, btt, 1e-5)
    assert _rel(y.f32, yt.detach()) < 1e-5
    yt.backward(dy)
    y.grad = dy.clone()
    tape.backward()
    assert _rel(xn.grad, xt.grad + prev) < TOL
    assert _rel(ps.g("n.weight"), gt.grad) < TOL and _rel(ps.g("n.bias"), btt.grad) < TOL
    # gelu
    tape = ag.Tape()
    zn = _node(x, split=False)
    yg = ag.gelu(tape, zn)
    zt = x.clone().requires_grad_(True)
    F.gelu(zt).backward(dy)
    yg.grad = dy.clone()
    tape.backward()
    assert _rel(zn.grad, zt.grad) < TOL


@pytest.mark.parametrize("B,T,causal,masked", [(2, 50, True, True), (3, 64, False, False), (4, 4, False, True)])
def test_self_attention_backward(cuda_device, B, T, causal, masked):
    from ralf_b200 import autograd as ag

    g = torch.Generator().manual_seed(T)
    H, dh = 8, 32
    Dm = H * dh
    qkv = torch.randn(B * T, 3 * Dm, generator=g).to(cuda_device)
    dy = torch.randn(B * T, Dm, generator=g).to(cuda_device)
    mask = torch.zeros(B, T, dtype=torch.bool)
    if masked:
        mask[:, T - 2:] = True
        mask[0, 1] = True
    tape = ag.Tape()
    qn = _node(qkv, split=False)
    y = ag.self_attention(tape, qn, B, T, H, dh, mask=mask.to(cuda_device).to(torch.uint8) if masked else None, causal=causal)
    qt = qkv.clone().requires_grad_(True)
    q, k, v = [t.view(B, T, H, dh).transpose(1, 2) for t in qt.split(Dm, dim=1)]
    s = (q @ k.transpose(-1, -2)) / math.sqrt(dh)
    if causal:
        s = s + torch.triu(torch.full((T, T), float("-inf"), device=cuda_device), 1)
    if masked:
        s = s.masked_fill(mask.to(cuda_device)[:, None, None, :], float("-inf"))
    o = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * T, Dm)
    from ralf_b200 import ops
    assert _rel(ops.unsplit(y.s), o.detach()) < 1e-4
    o.backward(dy)
    y.grad = dy.clone()
    tape.backward()
    assert _rel(qn.grad, qt.grad) < 5e-4


def test_cross_attention_backward(cuda_device):
    from ralf_b200 import autograd as ag

    g = torch.Generator().manual_seed(5)
    B, Tq, Tk, H, dh = 2, 50, 148, 8, 32
    Dm = H * dh
    qx = torch.randn(B * Tq, Dm, generator=g).to(cuda_device)
    kv = torch.randn(B * Tk, 2 * Dm, generator=g).to(cuda_device)
    dy = torch.randn(B * Tq, Dm, generator=g).to(cuda_device)
    tape = ag.Tape()
    qn, kn = _node(qx, split=False), _node(kv, split=False)
    y = ag.cross_attention(tape, qn, kn, 0, Dm, B, Tq, Tk, H, dh)
    qt, kt = qx.clone().requires_grad_(True), kv.clone().requires_grad_(True)
    q = qt.view(B, Tq, H, dh).transpose(1, 2)
    k = kt[:, :Dm].reshape(B, Tk, H, dh).transpose(1, 2)
    v = kt[:, Dm:].reshape(B, Tk, H, dh).transpose(1, 2)
    o = (torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(dh), -1) @ v).transpose(1, 2).reshape(B * Tq, Dm)
    o.backward(dy)
    y.grad = dy.clone()
    tape.backward()
    assert _rel(qn.grad, qt.grad) < 5e-4 and _rel(kn.grad, kt.grad) < 5e-4


@pytest.mark.parametrize("k,stride,relu,use_res,cin,cout", [(1, 1, True, False, 64, 128), (3, 1, True, False, 64, 64),
                                                           (3, 2, True, False, 64, 64), (1, 2, False, False, 64, 128),
                                                           (1, 1, True, True, 64, 256)])
def test_conv_bn_backward(cuda_device, k, stride, relu, use_res, cin, cout):
    from ralf_b200 import autograd as ag
    from ralf_b200.train_conv import Trunk

    g = torch.Generator().manual_seed(k * 10 + stride)
    B, H, W = 2, 12, 10
    pad = 1 if k == 3 else 0
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    w = torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)
    gm, bt = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1
    x = torch.randn(B, cin, H, W, generator=g).to(cuda_device)
    r = torch.randn(B, cout, Ho, Wo, generator=g).to(cuda_device)
    dy = torch.randn(B, cout, Ho, Wo, generator=g).to(cuda_device)

    class FakeModel:
        def named_buffers(self):
            return [("c.bn.running_mean", torch.zeros(cout, device=cuda_device)),
                    ("c.bn.running_var", torch.ones(cout, device=cuda_device)),
                    ("c.bn.num_batches_tracked", torch.zeros((), dtype=torch.long, device=cuda_device))]

    ps = _ps(cuda_device, {"c.conv.weight": w, "c.bn.weight": gm, "c.bn.bias": bt})
    tr = Trunk.__new__(Trunk)
    tr.ps, tr.dev, tr.buffers = ps, cuda_device, dict(FakeModel().named_buffers())
    tr.cw, tr.cwT, tr.kconvs = {}, {}, []
    if k == 1:
        ps.register_gemm_weight("c.conv", "c.conv.weight")
    else:
        tr.kconvs.append(("c.conv.weight", cout, cin, k * k))
        tr.cw["c.conv.weight"] = torch.zeros((2, cout, cin * k * k), dtype=torch.bfloat16, device=cuda_device)
        tr.cwT["c.conv.weight"] = torch.zeros((2, cin * k * k, cout), dtype=torch.bfloat16, device=cuda_device)
    ps.refresh_operands()
    tr.refresh_operands()
    nhwc = lambda t: t.permute(0, 2, 3, 1).reshape(-1, t.shape[1]).contiguous()
    tape = ag.Tape()
    xn = _node(nhwc(x), f32=False)
    rn = _node(nhwc(r), f32=False) if use_res else None
    y, geom = tr.conv_bn(tape, xn, (B, H, W), "c.conv", "c.bn", k, stride, pad, relu, res=rn)
    assert geom == (B, Ho, Wo)
    xt = x.clone().requires_grad_(True)
    wt, gt, btt = [t.to(cuda_device).requires_grad_(True) for t in (w, gm, bt)]
    rt = r.clone().requires_grad_(True)
    z = F.conv2d(xt, wt, stride=stride, padding=pad)
    yt = F.batch_norm(z, torch.zeros(cout, device=cuda_device), torch.ones(cout, device=cuda_device), gt, btt, True, 0.1, 1e-5)
    if use_res:
        yt = yt + rt
    if relu:
        yt = torch.relu(yt)
    from ralf_b200 import ops
    assert _rel(ops.unsplit(y.s), nhwc(yt.detach())) < 1e-4, "forward"
    yt.backward(dy)
    y.grad = nhwc(dy).clone()
    tape.backward()
    torch.cuda.synchronize()
    assert _rel(ps.g("c.conv.weight"), wt.grad) < 5e-4, "dW"
    assert _rel(ps.g("c.bn.weight"), gt.grad) < 5e-4 and _rel(ps.g("c.bn.bias"), btt.grad) < 5e-4, "dgamma/dbeta"
    assert _rel(xn.grad, nhwc(xt.grad)) < 5e-4, "dx"
    if use_res:
        assert _rel(rn.grad, nhwc(rt.grad)) < 5e-4, "dres"
    assert _rel(tr.buffers["c.bn.running_var"], 0.9 + 0.1 * z.detach().var(dim=(0, 2, 3), unbiased=True)) < 1e-4


def test_pool_upsample_embed_backward(cuda_device):
    from ralf_b200 import autograd as ag
    from ralf_b200 import ops
    from ralf_b200.autograd import _L, _stream, check

    g = torch.Generator().manual_seed(8)
    B, C, H, W = 2, 64, 14, 10
    x = torch.randn(B, C, H, W, generator=g).to(cuda_device)
    nhwc = lambda t: t.permute(0, 2, 3, 1).reshape(-1, t.shape[1]).contiguous()
    xs = ops.split_bf16(nhwc(x))
    xq = ops.unsplit(xs).view(B, H, W, C).permute(0, 3, 1, 2).contiguous().requires_grad_(True)  # the value the kernel sees
    yt = F.max_pool2d(xq, 3, 2, 1)
    dy = torch.randn(yt.shape, generator=g).to(cuda_device)
    yt.backward(dy)
    dx = torch.full((B * H * W, C), float("nan"), device=cuda_device)  # written completely: no zero-initialisation needed
    taps = torch.empty(dy.numel(), dtype=torch.uint8, device=cuda_device)
    check(_L().ralf_maxpool3x3s2_bwd(xs.data_ptr(), xs.stride(0), nhwc(dy).data_ptr(), B, H, W, C, dx.data_ptr(),
                                     taps.data_ptr(), _stream()), "mp")
    assert _rel(dx, nhwc(xq.grad)) < 1e-5
    # nearest upsample backward (11x8 -> 22x15 like the 350x240 canvas, and 8x8 -> 16x16)
    for (h5, w5, h4, w4) in [(11, 8, 22, 15), (8, 8, 16, 16)]:
        s = torch.randn(B, C, h5, w5, generator=g).to(cuda_device).requires_grad_(True)
        up = F.interpolate(s, size=(h4, w4), mode="nearest")
        d = torch.randn(up.shape, generator=g).to(cuda_device)
        up.backward(d)
        out = torch.empty(B * h5 * w5, C, device=cuda_device)
        check(_L().ralf_upsample_nearest_bwd(nhwc(d).data_ptr(), C, B, h5, w5, h4, w4, C, out.data_ptr(), _stream()), "up")
        assert _rel(out, nhwc(s.grad)) < 1e-5
    # embedding backward
    emb = torch.randn(30, 256, generator=g)
    ps = _ps(cuda_device, {"e.weight": emb})
    tok = torch.randint(0, 30, (4, 7), generator=g).to(cuda_device)
    pe = torch.randn(50, 256, generator=g).to(cuda_device)
    tape = ag.Tape()
    y = ag.embed(tape, ps, tok, 7, "e.weight", 16.0, pe)
    et = emb.to(cuda_device).requires_grad_(True)
    yt = et[tok] * 16.0 + pe[:7][None]
    dyy = torch.randn(4 * 7, 256, generator=g).to(cuda_device)
    yt.reshape(-1, 256).backward(dyy)
    y.grad = dyy.clone()
    tape.backward()
    assert _rel(y.f32, yt.detach().reshape(-1, 256)) < 1e-6 and _rel(ps.g("e.weight"), et.grad) < 1e-5
