"""GPU parity of the tcgen05 GEMM (ralf_gemm) against a float64 contraction of the same operands."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(a, w, bias=None, act=None, res=None, post_relu=False):
    y = a.double() @ w.double().t()
    if bias is not None:
        y = y + bias.double()
    if act == "relu":
        y = y.relu()
    elif act == "gelu":
        y = torch.nn.functional.gelu(y)
    if res is not None:
        y = y + res.double()
    if post_relu:
        y = y.relu()
    return y


@pytest.mark.parametrize("M,N,K,bn", [
    (128, 64, 64, 64), (128, 128, 256, 128), (300, 768, 256, 0), (1000, 1024, 256, 0), (77, 519, 256, 0),
    (256, 256, 1024, 64), (4096, 256, 2304, 0), (130, 256, 200, 64), (4, 256, 256, 0), (513, 3072, 256, 256),
    (128, 256, 256, 32), (128, 1024, 256, 32), (100, 519, 256, 32), (40000, 256, 128, 128), (33000, 64, 576, 64),
])
def test_gemm_bf16x3_matches_fp64(cuda_device, M, N, K, bn):
    from ralf_b200 import ops

    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, generator=g).to(cuda_device)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(cuda_device)
    bias = torch.randn(N, generator=g).to(cuda_device)
    out, _ = ops.gemm(ops.split_bf16(a), ops.split_bf16(w), bias=bias, block_n=bn)
    torch.cuda.synchronize()
    ref = _ref(a, w, bias)
    err = (out.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= 2e-5 * scale, (err, scale)


def test_gemm_plain_bf16(cuda_device):
    from ralf_b200 import ops

    g = torch.Generator(device="cpu").manual_seed(1)
    a = torch.randn(512, 512, generator=g).to(cuda_device)
    w = (torch.randn(384, 512, generator=g) / 512 ** 0.5).to(cuda_device)
    a_s, w_s = ops.split_bf16(a), ops.split_bf16(w)
    out, _ = ops.gemm(a_s, w_s, npass=1)
    ref = a_s[0].double() @ w_s[0].double().t()  # bf16-rounded operands, exact products
    err = (out.double() - ref).abs().max().item()
    assert err <= 1e-5 * ref.abs().max().item(), err


@pytest.mark.parametrize("act", [None, "relu", "gelu"])
def test_gemm_epilogue(cuda_device, act):
    from ralf_b200 import ops

    g = torch.Generator(device="cpu").manual_seed(5)
    M, N, K = 333, 256, 256
    a = torch.randn(M, K, generator=g).to(cuda_device)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(cuda_device)
    bias = torch.randn(N, generator=g).to(cuda_device)
    res = torch.randn(M, N, generator=g).to(cuda_device)
    out, outs = ops.gemm(ops.split_bf16(a), ops.split_bf16(w), bias=bias, act=act, res=res, want_split=True)
    ref = _ref(a, w, bias, act, res)
    scale = ref.abs().max().item()
    assert (out.double() - ref).abs().max().item() <= 2e-5 * scale
    assert (ops.unsplit(outs).double() - ref).abs().max().item() <= 4e-5 * scale
    # split residual + relu after the residual (ResNet bottleneck tail), row-broadcast residual
    res_s = ops.split_bf16(res)
    out2, _ = ops.gemm(ops.split_bf16(a), ops.split_bf16(w), bias=bias, res_split=res_s, post_relu=True)
    ref2 = _ref(a, w, bias, None, ops.unsplit(res_s), post_relu=True)
    assert (out2.double() - ref2).abs().max().item() <= 2e-5 * scale
    tab = torch.randn(37, N, generator=g).to(cuda_device)
    out3, _ = ops.gemm(ops.split_bf16(a), ops.split_bf16(w), res=tab, res_row_mod=37)
    ref3 = _ref(a, w) + tab.double()[torch.arange(M) % 37]
    assert (out3.double() - ref3).abs().max().item() <= 2e-5 * scale


def test_gemm_row_remap_and_col_offset(cuda_device):
    from ralf_b200 import ops

    g = torch.Generator(device="cpu").manual_seed(9)
    B, T, N, K = 6, 10, 256, 512
    a = torch.randn(B * T, K, generator=g).to(cuda_device)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(cuda_device)
    out = torch.zeros(B * (T + 1), 2 * N, device=cuda_device)
    ops.gemm(ops.split_bf16(a), ops.split_bf16(w), out_f32=out, out_col0=N, rows_per_group=T,
             group_stride=T + 1, group_offset=1)
    ref = _ref(a, w).float().view(B, T, N)
    got = out.view(B, T + 1, 2 * N)
    assert (got[:, 1:, N:] - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()
    assert got[:, 0].abs().max().item() == 0 and got[:, :, :N].abs().max().item() == 0


@pytest.mark.parametrize("B,H,W,C,N,KH", [(2, 64, 64, 64, 64, 3), (3, 32, 32, 128, 128, 3), (2, 16, 16, 256, 256, 3),
                                          (3, 8, 8, 512, 512, 3), (2, 22, 15, 256, 256, 3), (3, 11, 8, 512, 64, 3),
                                          (1, 44, 30, 128, 128, 3), (5, 16, 16, 64, 96, 1)])
def test_conv_gemm_implicit_matches_conv2d_fp64(cuda_device, B, H, W, C, N, KH):
    """ralf_conv_gemm (3x3 / stride 1 / pad 1 taps as 5-D TMA boxes with zero fill) against torch conv2d in float64,
    including tiles that are not full (22x15, 11x8 feature maps of the 350x240 canvases) and multi-image tiles (8x8)."""
    from ralf_b200 import ops

    g = torch.Generator(device=cuda_device).manual_seed(B * H + W + C)
    x = torch.randn(B, H, W, C, device=cuda_device, generator=g)             # NHWC
    w = torch.randn(N, C, KH, KH, device=cuda_device, generator=g) / (C * KH * KH) ** 0.5
    bias = torch.randn(N, device=cuda_device, generator=g)
    xs = ops.split_bf16(x.reshape(B * H * W, C))
    ws = ops.split_bf16(w.permute(0, 2, 3, 1).reshape(N, KH * KH * C).contiguous())
    y, _ = ops.gemm(xs, ws, bias=bias, act="relu", conv=(B, H, W, C, KH, KH))
    ref = torch.nn.functional.conv2d(ops.unsplit(xs).view(B, H, W, C).permute(0, 3, 1, 2).double(),
                                     ops.unsplit(ws).view(N, KH, KH, C).permute(0, 3, 1, 2).double(), bias.double(),
                                     padding=KH // 2).relu().permute(0, 2, 3, 1).reshape(B * H * W, N)
    err = (y.double() - ref).abs().max().item() / ref.abs().max().item()
    assert err <= 2e-5, err


@pytest.mark.parametrize("B,H,W", [(2, 256, 256), (3, 64, 96), (1, 350, 240), (2, 128, 128)])
def test_stem_s2d_gemm_matches_conv2d_fp64(cuda_device, B, H, W):
    """7x7 / stride 2 / pad 3 stem on the 4-channel canvas (common/image.py:69-77) through the space-to-depth implicit
    GEMM (ralf_stem_s2d + ralf_stem_gemm) against torch conv2d in float64; then the vectorised 3x3/s2 max-pool."""
    from ralf_b200 import ops

    g = torch.Generator(device=cuda_device).manual_seed(H + W)
    img = torch.rand(B, 4, H, W, device=cuda_device, generator=g)
    w = torch.randn(64, 4, 7, 7, device=cuda_device, generator=g) / 14.0
    bias = torch.randn(64, device=cuda_device, generator=g)
    w2 = torch.zeros(64, 4, 4, 2, 2, 4, device=cuda_device)
    for kh in range(7):
        for kw in range(7):
            w2[:, (kh + 1) // 2, (kw + 1) // 2, (kh + 1) % 2, (kw + 1) % 2, :] = w[:, :, kh, kw]
    ws = ops.split_bf16(w2.reshape(64, 256))
    a, Ho, Wo = ops.stem_s2d(img)
    y, ys = ops.gemm(a, ws, bias=bias, act="relu", stem=(B, Ho, Wo), want_split=True)
    # reference on the SAME rounded operands (split bf16 keeps ~17 bits of every input / weight)
    img_r = ops.unsplit(ops.split_bf16(img.reshape(-1, 8))).view_as(img)
    w_r = ops.unsplit(ops.split_bf16(w.reshape(64, -1))).view_as(w)
    ref = torch.nn.functional.conv2d(img_r.double(), w_r.double(), bias.double(), stride=2, padding=3).relu()
    assert ref.shape[2] == Ho and ref.shape[3] == Wo
    err = (y.double().view(B, Ho, Wo, 64).permute(0, 3, 1, 2) - ref).abs().max().item() / ref.abs().max().item()
    assert err <= 2e-5, err
    p, Hp, Wp = ops.maxpool3x3s2(ys, B, Ho, Wo, 64)
    pref = torch.nn.functional.max_pool2d(ops.unsplit(ys).view(B, Ho, Wo, 64).permute(0, 3, 1, 2), 3, 2, 1)
    assert torch.equal(ops.unsplit(p).view(B, Hp, Wp, 64).permute(0, 3, 1, 2), pref)


@pytest.mark.parametrize("M,N,K", [(64, 64, 131072), (256, 64, 32768), (64, 576, 40000), (519, 256, 3200), (128, 40, 9000)])
def test_gemm_splitk_matches_fp64(cuda_device, M, N, K):
    """Weight-gradient shape class (tiny M x N, K = all rows): split-K slices + deterministic reduction."""
    from ralf_b200 import ops

    g = torch.Generator(device=cuda_device).manual_seed(M + N)
    a = ops.split_bf16(torch.randn(M, K, device=cuda_device, generator=g))
    w = ops.split_bf16(torch.randn(N, K, device=cuda_device, generator=g))
    out = torch.full((M, N + 8), 7.0, device=cuda_device)
    ops.gemm(a, w, out_f32=out[:, :N], splitk=True)
    ref = ops.unsplit(a).double() @ ops.unsplit(w).double().t()
    assert (out[:, :N].double() - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()
    assert torch.all(out[:, N:] == 7.0)
    out2 = torch.empty_like(out)
    ops.gemm(a, w, out_f32=out2[:, :N], splitk=True)
    assert torch.equal(out[:, :N], out2[:, :N])  # deterministic


@pytest.mark.parametrize("M,N,act", [(5, 256, None), (128, 768, None), (300, 1024, "relu"), (1024, 768, None), (1030, 519, None)])
def test_gemm_with_fused_layernorm_matches_fp64(cuda_device, M, N, act):
    """ralf_gemm_ln (LayerNorm computed inside the GEMM's A-operand prologue, one CTA per n-tile x 128-row tile; opt-in for
    the decode chain via RALF_FUSE_LN=1) against float64 LayerNorm + Linear, including partial row tiles and a ragged N."""
    from ralf_b200 import ops

    g = torch.Generator(device=cuda_device).manual_seed(M + N)
    x = torch.randn(M, 256, device=cuda_device, generator=g) * 3 + 0.5
    gamma = torch.randn(256, device=cuda_device, generator=g)
    beta = torch.randn(256, device=cuda_device, generator=g)
    w = torch.randn(N, 256, device=cuda_device, generator=g) / 16
    bias = torch.randn(N, device=cuda_device, generator=g)
    out, _ = ops.gemm_ln(x, gamma, beta, ops.split_bf16(w), bias=bias, act=act)
    ref = torch.nn.functional.layer_norm(x.double(), (256,), gamma.double(), beta.double(), 1e-5) @ w.double().T + bias.double()
    if act == "relu":
        ref = ref.relu()
    assert (out.double() - ref).abs().max().item() <= 3e-5 * ref.abs().max().item()


@pytest.mark.parametrize("M,N,K,residual,act", [
    (19072, 256, 64, True, None),     # ResNet layer-1 conv3: one k-block, four chunk buffers per warp
    (19072, 256, 64, False, None),    # layer-1 downsample: no residual
    (20000, 512, 128, True, None),    # layer-2 conv3: two ring stages, three chunk buffers; M % 128 = 32 (clipped boxes)
    (18950, 1024, 256, True, "relu"),  # layer-3 conv3 shape with an activation; M % 32 != 0
    (37888, 128, 64, True, None),     # two tiles per CTA and one n-tile
    (20000, 256, 256, False, "relu"),  # no residual, K >= 192: three operand stages, ONE chunk buffer per warp
    (19500, 384, 512, False, None),   # the same variant at K = 512 (taken only with RALF_TEPI_KMAX_NORES >= 512)
])
def test_gemm_tma_epilogue_is_bit_identical_to_register_epilogue(cuda_device, M, N, K, residual, act):
    """Bottleneck-tail shape class (K <= 256, N % 128 == 0, M >= 148 tiles, split output): gemm_bf16_tepi_kernel moves the
    residual and the output with bulk tensor copies.  Same MMAs, same epilogue arithmetic -> must equal the register-staged
    kernel (forced here with block_n = 64) bit for bit, and float64 within the bf16x3 bound."""
    from ralf_b200 import ops

    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(cuda_device)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(cuda_device)
    bias = torch.randn(N, generator=g).to(cuda_device)
    res_s = ops.split_bf16(torch.randn(M, N, generator=g).to(cuda_device)) if residual else None
    a_s, w_s = ops.split_bf16(a), ops.split_bf16(w)
    guard = 4096
    flat = torch.full((2 * M * N + 2 * guard,), 7.0, dtype=torch.bfloat16, device=cuda_device)  # canaries around the output
    out_t = flat[guard:guard + 2 * M * N].view(2, M, N)
    kw = dict(bias=bias, act=act, res_split=res_s, post_relu=residual, want_f32=False)
    ops.gemm(a_s, w_s, out_split=out_t, **kw)
    out_r = torch.empty(2, M, N, dtype=torch.bfloat16, device=cuda_device)
    ops.gemm(a_s, w_s, out_split=out_r, block_n=64, **kw)
    torch.cuda.synchronize()
    assert torch.equal(out_t.view(torch.int16), out_r.view(torch.int16))
    assert (flat[:guard] == 7.0).all() and (flat[-guard:] == 7.0).all()
    ref = _ref(a, w, bias, act, ops.unsplit(res_s) if residual else None, post_relu=residual)
    assert (ops.unsplit(out_t).double() - ref).abs().max().item() <= 4e-5 * ref.abs().max().item()


@pytest.mark.parametrize("B,H,W,C,N,KH", [(2, 64, 64, 128, 128, 3), (3, 32, 32, 256, 256, 3), (5, 16, 16, 512, 64, 3),
                                          (2, 64, 64, 256, 512, 1), (3, 16, 16, 1024, 128, 1), (2, 22, 15, 256, 256, 3),
                                          (1, 44, 30, 128, 64, 1), (160, 64, 64, 256, 512, 1)])
def test_conv_gemm_stride2_matches_conv2d_fp64_and_im2col(cuda_device, B, H, W, C, N, KH):
    """ResNet's stride-2 3x3 (pad 1) and 1x1 downsample convolutions as implicit GEMMs: the taps are TMA boxes with
    element strides 2.  Against torch conv2d in float64, and bit-identical to im2col + plain GEMM (same k order, same
    MMAs).  22x15 / 44x30: odd sizes and partial tiles; the last case is large enough for the TMA-epilogue kernel."""
    from ralf_b200 import ops

    g = torch.Generator(device=cuda_device).manual_seed(B * H + W + C + KH)
    x = torch.randn(B, H, W, C, device=cuda_device, generator=g)
    w = torch.randn(N, C, KH, KH, device=cuda_device, generator=g) / (C * KH * KH) ** 0.5
    bias = torch.randn(N, device=cuda_device, generator=g)
    xs = ops.split_bf16(x.reshape(B * H * W, C))
    ws = ops.split_bf16(w.permute(0, 2, 3, 1).reshape(N, KH * KH * C).contiguous())
    _, y = ops.gemm(xs, ws, bias=bias, act="relu", want_f32=False, want_split=True, conv=(B, H, W, C, KH, KH, 2))
    cols, Ho, Wo = ops.im2col(xs, B, H, W, C, KH, KH, 2, KH // 2)
    assert y.shape[1] == B * Ho * Wo
    _, y2 = ops.gemm(cols, ws, bias=bias, act="relu", want_f32=False, want_split=True, block_n=64 if N % 128 else 0)
    if N % 128 == 0:
        assert torch.equal(y.view(torch.int16), y2.view(torch.int16))
    ref = torch.nn.functional.conv2d(ops.unsplit(xs).view(B, H, W, C).permute(0, 3, 1, 2).double(),
                                     ops.unsplit(ws).view(N, KH, KH, C).permute(0, 3, 1, 2).double(), bias.double(),
                                     stride=2, padding=KH // 2).relu().permute(0, 2, 3, 1).reshape(B * Ho * Wo, N)
    err = (ops.unsplit(y).double() - ref).abs().max().item() / ref.abs().max().item()
    assert err <= 4e-5, err


@pytest.mark.parametrize("M,K", [(1024, 256), (1024, 1024), (1, 256), (300, 1024), (129, 256)])
def test_gemm_residual_layernorm_cluster_kernel(cuda_device, M, K):
    """ralf_gemm_res_ln (decode loop): x_new = a . w^T + bias + x must equal the plain residual GEMM bit for bit, and its
    LayerNorm output (row statistics exchanged between the eight n-tile CTAs of a cluster through distributed shared
    memory) must equal float64 LayerNorm of x_new to split-operand precision and the stand-alone LayerNorm kernel to fp32
    rounding.  M = 1 is the DecodeSession shape, 300 / 129 leave partial row tiles."""
    from ralf_b200 import ops

    g = torch.Generator(device="cpu").manual_seed(M + K)
    a = ops.split_bf16(torch.randn(M, K, generator=g).to(cuda_device))
    w = ops.split_bf16((torch.randn(256, K, generator=g) / K ** 0.5).to(cuda_device))
    bias = torch.randn(256, generator=g).to(cuda_device)
    x = (torch.randn(M, 256, generator=g) * 2 + 0.5).to(cuda_device)
    gamma = torch.randn(256, generator=g).to(cuda_device)
    beta = torch.randn(256, generator=g).to(cuda_device)
    want_x, _ = ops.gemm(a, w, bias=bias, res=x, block_n=32)
    _, want_h = ops.layernorm(want_x, gamma, beta)
    xin = x.clone()
    guard = torch.full((2, M + 8, 256), 3.0, dtype=torch.bfloat16, device=cuda_device)
    got_x, got_h = ops.gemm_res_ln(a, w, xin, gamma, beta, bias=bias, ln_split=guard[:, :M])
    torch.cuda.synchronize()
    assert got_x.data_ptr() == xin.data_ptr() and torch.equal(got_x, want_x)   # in place, bit-identical to the per-op GEMM
    assert (guard[:, M:] == 3.0).all()
    ref = torch.nn.functional.layer_norm(want_x.double(), (256,), gamma.double(), beta.double(), 1e-5)
    assert (ops.unsplit(got_h).double() - ref).abs().max().item() <= 2e-4
    # against the stand-alone kernel: one step of the split representation (2^-17 of the magnitude) at most
    assert (ops.unsplit(got_h) - ops.unsplit(want_h)).abs().max().item() <= 2e-5 * ops.unsplit(want_h).abs().max().item()


@pytest.mark.parametrize("M,N,K,residual,act,inplace", [
    (32768, 768, 256, False, None, False),   # image-encoder qkv
    (20000, 256, 256, True, None, True),     # out-projection + residual, written over the residual (M % 128 = 32)
    (22528, 256, 64, True, "relu", False),   # one k-block: four chunk buffers
    (19000, 384, 128, False, "relu", False),
])
def test_gemm_tma_epilogue_fp32_output_is_bit_identical(cuda_device, M, N, K, residual, act, inplace):
    """fp32 flavour of the TMA-epilogue kernel (fp32 output, optional fp32 residual; 32 x 32 fp32 chunks, SWIZZLE_128B) against
    the register-staged kernel (block_n = 64), bit for bit; canaries around the output."""
    from ralf_b200 import ops

    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    a_s = ops.split_bf16(torch.randn(M, K, generator=g).to(cuda_device))
    w_s = ops.split_bf16((torch.randn(N, K, generator=g) / K ** 0.5).to(cuda_device))
    bias = torch.randn(N, generator=g).to(cuda_device)
    res = torch.randn(M, N, generator=g).to(cuda_device) if residual else None
    want, _ = ops.gemm(a_s, w_s, bias=bias, act=act, res=res, block_n=64)
    guard = 4096
    flat = torch.full((M * N + 2 * guard,), 7.0, device=cuda_device)
    out = flat[guard:guard + M * N].view(M, N)
    if inplace:
        out.copy_(res)
        res = out
    ops.gemm(a_s, w_s, bias=bias, act=act, res=res, out_f32=out)
    torch.cuda.synchronize()
    assert torch.equal(out, want)
    assert (flat[:guard] == 7.0).all() and (flat[-guard:] == 7.0).all()
