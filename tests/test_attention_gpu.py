"""GPU parity of the decode-step attention kernels (ralf_attention_decode / _append) and LayerNorm against float64
references of the same op (nn.MultiheadAttention arithmetic: softmax(q k^T / sqrt(dh)) v per head)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_decode(q, k, v, B, H, dh, Tk, mask=None):
    qd = q.double().view(B, H, 1, dh)
    kd = k.double().view(B, Tk, H, dh).permute(0, 2, 1, 3)
    vd = v.double().view(B, Tk, H, dh).permute(0, 2, 1, 3)
    s = (qd @ kd.transpose(-1, -2)) * dh ** -0.5
    if mask is not None:
        s = s.masked_fill(mask.bool()[:, None, None, :Tk], float("-inf"))
    return (torch.softmax(s, -1) @ vd).permute(0, 2, 1, 3).reshape(B, H * dh)


@pytest.mark.parametrize("B,Tk,H,dh", [(3, 532, 8, 32), (5, 680, 8, 32), (2, 64, 8, 32), (130, 334, 8, 32),
                                       (2, 100, 4, 64), (4, 33, 8, 32)])
def test_cross_attention_decode_matches_fp64(cuda_device, B, Tk, H, dh):
    """Memory cross-attention of one decode step on the layer-major K/V cache ([B*Tk, 2*D], K | V): the streaming
    single-pass kernel (Tk >= 64, no mask) and the two-pass kernel (short Tk)."""
    from ralf_b200 import ops

    g = torch.Generator(device=cuda_device).manual_seed(B * 1000 + Tk)
    D = H * dh
    kv = torch.randn(B * Tk, 2 * D, device=cuda_device, generator=g)
    q = torch.randn(B, D, device=cuda_device, generator=g) * 2.0
    out = ops.attention_decode(q, kv[:, :D], kv[:, D:], Tk, Tk, B, H, dh)
    ref = _ref_decode(q, kv[:, :D], kv[:, D:], B, H, dh, Tk)
    got = ops.unsplit(out).double()
    assert (got - ref).abs().max().item() <= 2e-5 * ref.abs().max().item() + 1e-6


def test_self_attention_decode_append_with_padding_mask(cuda_device):
    from ralf_b200 import ops

    B, S, H, dh = 6, 20, 8, 32
    D = H * dh
    g = torch.Generator(device=cuda_device).manual_seed(3)
    kc = torch.zeros(B, S, D, device=cuda_device)
    vc = torch.zeros(B, S, D, device=cuda_device)
    mask = torch.zeros(B, S, dtype=torch.uint8, device=cuda_device)
    mask[1, 2] = 1
    mask[4, 0:2] = 1
    ks, vs = [], []
    for pos in range(7):
        qkv = torch.randn(B, 3 * D, device=cuda_device, generator=g)
        ks.append(qkv[:, D:2 * D].clone())
        vs.append(qkv[:, 2 * D:].clone())
        out = ops.attention_decode_append(qkv, kc, vc, pos, B, H, dh, mask=mask)
        k = torch.stack(ks, 1).reshape(B * (pos + 1), D)
        v = torch.stack(vs, 1).reshape(B * (pos + 1), D)
        if pos >= 2:  # rows whose visible keys are all masked are undefined in the reference too; skip early steps
            ref = _ref_decode(qkv[:, :D], k, v, B, H, dh, pos + 1, mask=mask)
            got = ops.unsplit(out).double()
            assert (got - ref).abs().max().item() <= 2e-5 * ref.abs().max().item() + 1e-6
    assert torch.equal(kc[:, :7], torch.stack(ks, 1)) and torch.equal(vc[:, :7], torch.stack(vs, 1))


@pytest.mark.parametrize("M,D", [(1000, 256), (7, 256), (64, 128), (33, 1024)])
def test_layernorm_matches_fp64(cuda_device, M, D):
    from ralf_b200 import ops

    g = torch.Generator(device=cuda_device).manual_seed(M + D)
    x = torch.randn(M, D, device=cuda_device, generator=g) * 3 + 1
    gamma = torch.randn(D, device=cuda_device, generator=g)
    beta = torch.randn(D, device=cuda_device, generator=g)
    y, ys = ops.layernorm(x, gamma, beta, want_f32=True)
    ref = torch.nn.functional.layer_norm(x.double(), (D,), gamma.double(), beta.double(), 1e-5)
    assert (y.double() - ref).abs().max().item() <= 1e-5
    assert (ops.unsplit(ys).double() - ref).abs().max().item() <= 2e-4


@pytest.mark.parametrize("B,Tq,Tk,H", [(3, 256, 256, 8), (2, 200, 200, 8), (5, 128, 64, 8), (1, 300, 150, 8),
                                       (130, 256, 256, 8)])
def test_encoder_attention_tcgen05_matches_fp64(cuda_device, B, Tq, Tk, H):
    """Image-encoder self-attention shape class (head_dim 32, no mask, <= 256 keys): tcgen05 kernel with split-bf16
    operands and P kept in TMEM, against the float64 softmax(q k^T / sqrt(dh)) v of the same fp32 inputs."""
    from ralf_b200 import ops

    dh = 32
    D = H * dh
    g = torch.Generator(device=cuda_device).manual_seed(B + Tq + Tk)
    q = torch.randn(B * Tq, D, device=cuda_device, generator=g) * 1.5
    kv = torch.randn(B * Tk, 2 * D, device=cuda_device, generator=g) * 1.5
    out = ops.unsplit(ops.attention(q, kv[:, :D], kv[:, D:], B, H, Tq, Tk, dh)).double()
    qd = q.double().view(B, Tq, H, dh).permute(0, 2, 1, 3)
    kd = kv[:, :D].double().view(B, Tk, H, dh).permute(0, 2, 1, 3)
    vd = kv[:, D:].double().view(B, Tk, H, dh).permute(0, 2, 1, 3)
    ref = (torch.softmax(qd @ kd.transpose(-1, -2) * dh ** -0.5, -1) @ vd).permute(0, 2, 1, 3).reshape(B * Tq, D)
    err = (out - ref).abs().max().item() / ref.abs().max().item()
    assert err <= 3e-5, err


@pytest.mark.parametrize("B,Tk", [(3, 532), (130, 334), (2, 40)])
def test_kv24_cache_gemm_and_cross_attention(cuda_device, B, Tk):
    """24-bit K/V cache: ralf_gemm(out_kv24) must store exactly round-to-24-bit of its fp32 result, and the kv24 decode
    kernel must equal float64 attention over those stored values."""
    from ralf_b200 import ops

    H, dh, D = 8, 32, 256
    g = torch.Generator(device=cuda_device).manual_seed(B + Tk)
    mem = ops.split_bf16(torch.randn(B * Tk, D, device=cuda_device, generator=g))
    w = ops.split_bf16(torch.randn(2 * D, D, device=cuda_device, generator=g) / 16)
    bias = torch.randn(2 * D, device=cuda_device, generator=g)
    ref32, _ = ops.gemm(mem, w, bias=bias)
    kv24 = torch.empty(B * Tk, 1536, dtype=torch.uint8, device=cuda_device)
    ops.gemm(mem, w, bias=bias, want_f32=False, out_kv24=kv24)
    hi = kv24[:, :1024].contiguous().view(torch.int16).to(torch.int32) & 0xFFFF   # [rows, 512]: K | V
    lo = kv24[:, 1024:].to(torch.int32)
    got = ((hi << 16) | (lo << 8)).view(torch.float32)
    want = ((ref32.view(torch.int32) + 0x80) & ~0xFF).view(torch.float32)
    assert torch.equal(got, want)
    q = torch.randn(B, D, device=cuda_device, generator=g) * 2
    out = ops.unsplit(ops.attention_decode_kv24(q, kv24, Tk, Tk, B, H)).double()
    ref = _ref_decode(q, want[:, :D], want[:, D:], B, H, dh, Tk)
    assert (out - ref).abs().max().item() <= 2e-5 * ref.abs().max().item() + 1e-6


@pytest.mark.parametrize("B,Tk", [(3, 532), (130, 334), (2, 40)])
def test_kv16_cache_gemm_and_cross_attention(cuda_device, B, Tk):
    """16-bit per-head-scaled K/V cache (ralf_gemm out_kv_fmt = 16, the default of the decode loop): every stored value
    is the offset-binary 16-bit integer rint(x * 32767 / amax) + 32768 of the GEMM's fp32 result x, amax = the largest
    magnitude of its (row, head); the stored scale is amax / 32767; and the kv16 decode kernel equals float64 attention
    over the DEQUANTISED values.  Also bounds the format's own error: |dequant - x| <= amax / 65534 (half a step)."""
    from ralf_b200 import ops

    H, dh, D = 8, 32, 256
    g = torch.Generator(device=cuda_device).manual_seed(B + Tk)
    mem = ops.split_bf16(torch.randn(B * Tk, D, device=cuda_device, generator=g))
    w = ops.split_bf16(torch.randn(2 * D, D, device=cuda_device, generator=g) / 16)
    bias = torch.randn(2 * D, device=cuda_device, generator=g)
    ref32, _ = ops.gemm(mem, w, bias=bias)
    kv16 = torch.empty(B * Tk, ops.KV_ROW_BYTES[16], dtype=torch.uint8, device=cuda_device)
    ops.gemm(mem, w, bias=bias, want_f32=False, out_kv24=kv16)
    qv = (kv16[:, :1024].contiguous().view(torch.int16).to(torch.int32) & 0xFFFF) - 32768      # [rows, 512]: K | V
    sc = kv16[:, 1024:].contiguous().view(torch.float32).view(-1, 8, 2)                          # [rows, head, (K, V)]
    sc = torch.cat([sc[:, :, 0], sc[:, :, 1]], dim=1).contiguous()                               # [rows, 16]: K heads | V heads
    x = ref32.view(B * Tk, 16, 32)
    amax = x.abs().amax(dim=-1)
    assert torch.equal(sc, amax * (1.0 / 32767.0))
    want_q = torch.round(x * (32767.0 / amax)[..., None]).to(torch.int32).view(B * Tk, 512)
    assert (qv - want_q).abs().max().item() <= 1  # rint of a product rounded in fp32: at most one step apart, almost always 0
    assert (qv != want_q).float().mean().item() < 1e-3
    deq = (qv.view(B * Tk, 16, 32).float() * sc[..., None]).view(B * Tk, 512)
    assert ((deq - ref32).abs() <= (amax / 65534.0 * 1.01 + 1e-12)[..., None].expand(-1, -1, 32).reshape(B * Tk, 512)).all()
    q = torch.randn(B, D, device=cuda_device, generator=g) * 2
    out = ops.unsplit(ops.attention_decode_kv24(q, kv16, Tk, Tk, B, H)).double()
    ref = _ref_decode(q, deq[:, :D], deq[:, D:], B, H, dh, Tk)
    assert (out - ref).abs().max().item() <= 2e-5 * ref.abs().max().item() + 1e-6
    full = _ref_decode(q, ref32[:, :D], ref32[:, D:], B, H, dh, Tk)  # vs attention over the unquantised fp32 K/V
    assert (out - full).abs().max().item() <= 2e-4 * full.abs().max().item()


@pytest.mark.parametrize("B,Tq,Tk,H", [(3, 256, 16, 8), (2, 330, 16, 8), (5, 100, 7, 8), (130, 256, 16, 8), (1, 31, 1, 4)])
def test_fusion_attention_fewkeys_matches_fp64(cuda_device, B, Tq, Tk, H):
    """Fusion Attention shape class (common/attention.py:49-71: 8 heads x 64, image tokens over the 16 retrieved layouts):
    the tile kernel with staged, coalesced stores against float64 attention; q / k / v are column slices of wider
    projections like in the engine (q [B*Tq, 512], kv [B*Tk, 1024])."""
    from ralf_b200 import ops

    dh, Dm = 64, H * 64
    g = torch.Generator(device=cuda_device).manual_seed(B * 1000 + Tq + Tk)
    q = torch.randn(B * Tq, Dm, device=cuda_device, generator=g)
    kv = torch.randn(B * Tk, 2 * Dm, device=cuda_device, generator=g)
    out = ops.attention(q, kv[:, :Dm], kv[:, Dm:], B, H, Tq, Tk, dh)
    sp = lambda t, T: t.double().view(B, T, H, dh).transpose(1, 2)
    ref = (torch.softmax(sp(q, Tq) @ sp(kv[:, :Dm], Tk).transpose(-1, -2) * dh ** -0.5, -1) @ sp(kv[:, Dm:], Tk))
    ref = ref.transpose(1, 2).reshape(B * Tq, Dm)
    err = (ops.unsplit(out).double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 2e-5, err


@pytest.mark.parametrize("N,T,H", [(48, 11, 4), (2080, 13, 4), (7, 16, 4)])
def test_fidnet_attention_fewkeys_with_padding_matches_fp64(cuda_device, N, T, H):
    """FIDNetV3 shape class (fid/model.py:26-33): many short sequences, 4 heads x 64, key-padding mask, fused QKV."""
    from ralf_b200 import ops

    dh, Dm = 64, H * 64
    g = torch.Generator(device=cuda_device).manual_seed(N + T)
    qkv = torch.randn(N * T, 3 * Dm, device=cuda_device, generator=g)
    n_valid = torch.randint(1, T + 1, (N,), device=cuda_device, generator=g)
    pad = (torch.arange(T, device=cuda_device)[None] >= n_valid[:, None]).to(torch.uint8).contiguous()
    out = ops.attention(qkv[:, :Dm], qkv[:, Dm:2 * Dm], qkv[:, 2 * Dm:], N, H, T, T, dh, mask=pad)
    sp = lambda t: t.double().view(N, T, H, dh).transpose(1, 2)
    s = sp(qkv[:, :Dm]) @ sp(qkv[:, Dm:2 * Dm]).transpose(-1, -2) * dh ** -0.5
    s = s.masked_fill(pad.bool()[:, None, None, :], float("-inf"))
    ref = (torch.softmax(s, -1) @ sp(qkv[:, 2 * Dm:])).transpose(1, 2).reshape(N * T, Dm)
    err = (ops.unsplit(out).double() - ref).abs().max().item() / ref.abs().max().item()
    assert err < 2e-5, err


_CHILD = r'''
import json, sys, torch
sys.path.insert(0, sys.argv[1])
from ralf_b200 import ops
dev = torch.device("cuda:0")
dh, H = 32, 8
D = H * dh
for B, Tq, Tk in json.loads(sys.argv[3]):
    g = torch.Generator(device=dev).manual_seed(B + Tq + Tk)
    q = torch.randn(B * Tq, D, device=dev, generator=g) * 1.5
    kv = torch.randn(B * Tk, 2 * D, device=dev, generator=g) * 1.5
    out = ops.unsplit(ops.attention(q, kv[:, :D], kv[:, D:], B, H, Tq, Tk, dh)).double()
    qd = q.double().view(B, Tq, H, dh).permute(0, 2, 1, 3)
    kd = kv[:, :D].double().view(B, Tk, H, dh).permute(0, 2, 1, 3)
    vd = kv[:, D:].double().view(B, Tk, H, dh).permute(0, 2, 1, 3)
    ref = (torch.softmax(qd @ kd.transpose(-1, -2) * dh ** -0.5, -1) @ vd).permute(0, 2, 1, 3).reshape(B * Tq, D)
    err = (out - ref).abs().max().item() / ref.abs().max().item()
    assert err <= 3e-5, (B, Tq, Tk, err)
    torch.save(out.cpu(), sys.argv[2] + f"_{B}_{Tq}_{Tk}.pt")
print("ok")
'''


def _attention_in_child(env: dict, shapes: list) -> dict:
    """The kernel selectors (RALF_ATTN_TC, RALF_ATTN_TC_BIG) are read once per process: run the fp64 check of
    test_encoder_attention_tcgen05_matches_fp64 in a child with `env` and hand back its outputs per shape."""
    import json
    import os
    import subprocess
    import sys
    import tempfile

    from tests import helpers

    with tempfile.TemporaryDirectory() as tmp:
        r = subprocess.run([sys.executable, "-c", _CHILD, helpers.ROOT, os.path.join(tmp, "o"), json.dumps(shapes)],
                           env=dict(os.environ, **env), capture_output=True, text=True, timeout=300)
        assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]
        return {tuple(sh): torch.load(os.path.join(tmp, "o_%d_%d_%d.pt" % tuple(sh))) for sh in shapes}


def test_encoder_attention_tcgen05_two_threads_per_row_variant(cuda_device):
    """RALF_ATTN_TC=2 selects attention_tc2_kernel (256 threads: two threads per query row, key columns split in halves):
    same fp64 bar as the default kernel, and bit-identical to it wherever one half holds all the keys (Tk <= 128: the
    row sum has a single term)."""
    shapes = [[3, 256, 256], [2, 200, 200], [5, 128, 64], [1, 300, 150], [130, 256, 256], [2, 130, 128]]
    one = _attention_in_child({"RALF_ATTN_TC": "1"}, shapes)
    two = _attention_in_child({"RALF_ATTN_TC": "2"}, shapes)
    assert torch.equal(one[(5, 128, 64)], two[(5, 128, 64)]) and torch.equal(one[(2, 130, 128)], two[(2, 130, 128)])
    assert not torch.equal(one[(130, 256, 256)], two[(130, 256, 256)])  # (left) + (right) row sums: the other kernel ran


def test_encoder_attention_tcgen05_more_than_256_keys(cuda_device):
    """256 < Tk <= 480 (the reference's real 350 x 240 canvases give 330 image tokens) on the tensor cores with P written in
    place over S in TMEM (the default since round 2); RALF_ATTN_TC_BIG=0 sends these shapes to the CUDA-core kernel."""
    shapes = [[2, 330, 330], [1, 300, 257], [3, 480, 480], [2, 128, 400], [1, 200, 272], [130, 330, 330]]
    big = _attention_in_child({"RALF_ATTN_TC_BIG": "1"}, shapes)
    base = _attention_in_child({"RALF_ATTN_TC_BIG": "0"}, shapes)
    assert any(not torch.equal(big[k], base[k]) for k in big)  # a different kernel produced them
    for k in big:
        assert (big[k] - base[k]).abs().max().item() <= 6e-5 * base[k].abs().max().item()


_CHILD_FEWKEYS = r'''
import json, sys, torch
sys.path.insert(0, sys.argv[1])
from ralf_b200 import ops
dev = torch.device("cuda:0")
for kind, B, Tq, Tk, H in json.loads(sys.argv[3]):
    Dm = H * 64
    g = torch.Generator(device=dev).manual_seed(B + Tq + Tk + H)
    if kind == "fid":   # fused QKV of B sequences of Tq tokens, key-padding mask
        qkv = torch.randn(B * Tq, 3 * Dm, device=dev, generator=g)
        nv = torch.randint(1, Tq + 1, (B,), device=dev, generator=g)
        pad = (torch.arange(Tq, device=dev)[None] >= nv[:, None]).to(torch.uint8).contiguous()
        out = ops.attention(qkv[:, :Dm], qkv[:, Dm:2 * Dm], qkv[:, 2 * Dm:], B, H, Tq, Tq, 64, mask=pad)
    else:               # image tokens over the retrieved layouts
        q = torch.randn(B * Tq, Dm, device=dev, generator=g)
        kv = torch.randn(B * Tk, 2 * Dm, device=dev, generator=g)
        out = ops.attention(q, kv[:, :Dm], kv[:, Dm:], B, H, Tq, Tk, 64)
    torch.cuda.synchronize()
    torch.save(out.cpu(), sys.argv[2] + f"_{kind}_{B}_{Tq}_{Tk}_{H}.pt")
print("ok")
'''


def test_fewkeys_attention_smem_kernel_is_bit_identical_to_the_tile_kernel(cuda_device):
    """attention_kvsmem_kernel (RALF_ATTN_FEWKEYS=2, default: K/V of a CTA's groups staged once in shared memory) does the
    arithmetic of attention_fewkeys_kernel (=1) in the same order: outputs equal bit for bit on the FIDNetV3 and fusion
    Attention shape classes, partial groups / tiles included (the fp64 bars are test_fusion_... / test_fidnet_... above,
    which run the default kernel)."""
    import json
    import os
    import subprocess
    import sys
    import tempfile

    from tests import helpers

    shapes = [["fid", 2080, 11, 11, 4], ["fid", 7, 16, 16, 4], ["fid", 1, 13, 13, 4], ["fus", 130, 256, 16, 8],
              ["fus", 2, 330, 16, 8], ["fus", 5, 100, 7, 8], ["fus", 1, 31, 1, 4], ["fus", 3, 9, 16, 8]]
    outs = {}
    for mode in ("1", "2"):
        with tempfile.TemporaryDirectory() as tmp:
            r = subprocess.run([sys.executable, "-c", _CHILD_FEWKEYS, helpers.ROOT, os.path.join(tmp, "o"), json.dumps(shapes)],
                               env=dict(os.environ, RALF_ATTN_FEWKEYS=mode), capture_output=True, text=True, timeout=300)
            assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]
            outs[mode] = [torch.load(os.path.join(tmp, "o_%s_%d_%d_%d_%d.pt" % tuple(sh))) for sh in shapes]
    for sh, a, b in zip(shapes, outs["1"], outs["2"]):
        assert torch.equal(a, b), sh
