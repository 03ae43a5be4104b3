"""Torch-tensor front end of the C ABI (device pointers + current stream; no torch types cross the ABI)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import GemmArgs

ACT = {None: 0, "none": 0, "relu": 1, "gelu": 2}

# Number of OUR kernels launched through the C ABI (bench.py reports the delta over the timed region).
_LAUNCHES = 0
_KERNELS_PER_CALL = {"ralf_knn_topk": 4, "ralf_knn_topk_exact": 2, "ralf_knn_fixup_exact": 3, "ralf_knn_merge": 2, "ralf_ce_label_smooth": 2}


def launch_count() -> int:
    return _LAUNCHES


def check(rc: int, what: str) -> None:  # noqa: F811  (wraps _lib.check with launch accounting)
    global _LAUNCHES
    _lib.check(rc, what)
    _LAUNCHES += _KERNELS_PER_CALL.get(what, 1)


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def split_bf16(x: torch.Tensor, lo: bool = True) -> torch.Tensor:
    """fp32 [..., K] -> split bf16 [2, ..., K]: plane 0 = bf16(x), plane 1 = bf16(x - plane 0)."""
    x = x.float()
    hi = x.to(torch.bfloat16)
    out = torch.empty((2, *x.shape), dtype=torch.bfloat16, device=x.device)
    out[0] = hi
    if lo:
        out[1] = (x - hi.float()).to(torch.bfloat16)
    else:
        out[1].zero_()
    return out


def unsplit(xs: torch.Tensor) -> torch.Tensor:
    return xs[0].float() + xs[1].float()


def gemm(
    a: torch.Tensor,  # split bf16 [2, M, K]
    w: torch.Tensor,  # split bf16 [2, N, K]
    *,
    bias: Optional[torch.Tensor] = None,
    act: Optional[str] = None,
    post_relu: bool = False,
    res: Optional[torch.Tensor] = None,
    res_split: Optional[torch.Tensor] = None,
    res_row_mod: int = 0,
    out_f32: Optional[torch.Tensor] = None,
    out_split: Optional[torch.Tensor] = None,
    out_col0: int = 0,
    rows_per_group: int = 0,
    group_stride: int = 0,
    group_offset: int = 0,
    npass: int = 3,
    block_n: int = 0,
    want_f32: bool = True,
    want_split: bool = False,
    conv: Optional[tuple] = None,
    stem: Optional[tuple] = None,
    out_kv24: Optional[torch.Tensor] = None,
    splitk: bool = False,
) -> tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """D = A . W^T with the fused epilogue of ``ralf_gemm`` (include/ralf_b200.h).

    Outputs are allocated when not supplied ([M, N] fp32 and/or [2, M, N] split bf16).
    ``conv = (B, H, W, C, KH, KW[, stride])``: ``a`` is the NHWC activation [2, B*H*W, C] and the call is the
    convolution ``ralf_conv_gemm_strided`` (implicit GEMM, padding KH // 2, stride 1 or 2, K = KH*KW*C taken from ``w``).
    ``stem = (B, Ho, Wo)``: ``a`` is the space-to-depth buffer of :func:`stem_s2d` ([2, B*(Ho+3)*(Wo+3), 16]) and the
    call is ``ralf_stem_gemm`` (K = 256)."""
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.dim() == 3 and w.dim() == 3
    M, K = a.shape[1], a.shape[2]
    N = w.shape[1]
    if stem is not None:
        sB, sHo, sWo = stem
        assert a.shape[1] == sB * (sHo + 3) * (sWo + 3) and K == 16 and w.shape[2] == 256 and a.is_contiguous()
        M, K = sB * sHo * sWo, 256
    if conv is not None:
        cB, cH, cW, cC, cKH, cKW = conv[:6]
        cS = conv[6] if len(conv) > 6 else 1
        assert M == cB * cH * cW and K == cC and a.stride(1) == cC and w.shape[2] == cKH * cKW * cC, (a.shape, w.shape)
        K = w.shape[2]
        M = cB * ((cH + 2 * (cKH // 2) - cKH) // cS + 1) * ((cW + 2 * (cKW // 2) - cKW) // cS + 1)
    assert w.shape[2] == K, (a.shape, w.shape)
    assert a.stride(2) == 1 and w.stride(2) == 1
    if out_f32 is None and want_f32:
        out_f32 = torch.empty((M, N), dtype=torch.float32, device=a.device)
    if out_split is None and want_split:
        out_split = torch.empty((2, M, N), dtype=torch.bfloat16, device=a.device)
    out_ld = 0
    if out_f32 is not None:
        out_ld = out_f32.stride(-2)
    if out_split is not None:
        if out_f32 is not None:
            assert out_split.stride(-2) == out_ld
        out_ld = out_split.stride(-2)
    g = GemmArgs()
    g.A, g.a_plane, g.lda = a.data_ptr(), a.stride(0), a.stride(1)
    g.W, g.w_plane, g.ldw = w.data_ptr(), w.stride(0), w.stride(1)
    g.M, g.N, g.K = M, N, K
    g.npass, g.block_n = npass, block_n
    g.bias = _ptr(bias)
    g.act = ACT[act]
    g.post_relu = int(post_relu)
    g.res = _ptr(res)
    g.res_split = _ptr(res_split)
    g.res_plane = res_split.stride(0) if res_split is not None else 0
    g.res_ld = res.stride(-2) if res is not None else (res_split.stride(-2) if res_split is not None else 0)
    g.res_row_mod = res_row_mod
    g.out_f32 = _ptr(out_f32)
    g.out_split = _ptr(out_split)
    g.out_plane = out_split.stride(0) if out_split is not None else 0
    g.out_split_lo = 1 if npass == 3 else 0
    g.out_ld, g.out_col0 = out_ld, out_col0
    g.rows_per_group, g.group_stride, g.group_offset = rows_per_group, group_stride, group_offset
    if splitk:  # weight-gradient shape class: few output tiles, K = all rows -> k slices + deterministic sum
        nbytes = _lib.lib().ralf_gemm_splitk_workspace_bytes(M, N, K)
        if nbytes:
            ws = torch.empty(nbytes, dtype=torch.uint8, device=a.device)
            g.splitk_ws, g.splitk_ws_bytes = ws.data_ptr(), nbytes
            global _LAUNCHES
            _LAUNCHES += 1
    if out_kv24 is not None:  # uint8 rows of the quantised K/V cache (N = 512): 1536 B = 24-bit, 1088 B = 16-bit + scales
        assert out_kv24.dtype == torch.uint8 and out_kv24.shape[-1] in KV_ROW_BYTES.values() and out_kv24.is_contiguous() and N == 512
        g.out_kv24 = out_kv24.data_ptr()
        g.out_kv_fmt = 16 if out_kv24.shape[-1] == KV_ROW_BYTES[16] else 24
        if out_f32 is None and out_split is None:
            g.out_ld = 512  # unused, keeps the alignment checks trivially true
    if stem is not None:
        check(_lib.lib().ralf_stem_gemm(C.byref(g), *stem, _stream()), "ralf_stem_gemm")
    elif conv is not None:
        check(_lib.lib().ralf_conv_gemm_strided(C.byref(g), cB, cH, cW, cC, cKH, cKW, cS, _stream()), "ralf_conv_gemm_strided")
    else:
        check(_lib.lib().ralf_gemm(C.byref(g), _stream()), "ralf_gemm")
    return out_f32, out_split


def gemm_ln(
    x: torch.Tensor,      # fp32 [M <= 128, 256]
    gamma: torch.Tensor,
    beta: torch.Tensor,
    w: torch.Tensor,      # split bf16 [2, N, 256]
    *,
    bias: Optional[torch.Tensor] = None,
    act: Optional[str] = None,
    res: Optional[torch.Tensor] = None,
    out_f32: Optional[torch.Tensor] = None,
    want_f32: bool = True,
    want_split: bool = False,
    eps: float = 1e-5,
) -> tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """LayerNorm(x) . W^T with the LayerNorm fused into the GEMM prologue (ralf_gemm_ln; decode path)."""
    M, K = x.shape
    N = w.shape[1]
    if out_f32 is None and want_f32:
        out_f32 = torch.empty((M, N), dtype=torch.float32, device=x.device)
    out_split = torch.empty((2, M, N), dtype=torch.bfloat16, device=x.device) if want_split else None
    g = GemmArgs()
    g.W, g.w_plane, g.ldw = w.data_ptr(), w.stride(0), w.stride(1)
    g.M, g.N, g.K, g.npass = M, N, K, 3
    g.bias, g.act = _ptr(bias), ACT[act]
    g.res, g.res_ld = _ptr(res), (res.stride(-2) if res is not None else 0)
    g.out_f32, g.out_split = _ptr(out_f32), _ptr(out_split)
    g.out_plane = out_split.stride(0) if out_split is not None else 0
    g.out_split_lo = 1
    g.out_ld = out_f32.stride(-2) if out_f32 is not None else N
    check(_lib.lib().ralf_gemm_ln(x.data_ptr(), x.stride(0), gamma.data_ptr(), beta.data_ptr(), eps, C.byref(g), _stream()),
          "ralf_gemm_ln")
    return out_f32, out_split


def gemm_res_ln(a: torch.Tensor, w: torch.Tensor, x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, *,
                bias: Optional[torch.Tensor] = None, out_f32: Optional[torch.Tensor] = None,
                ln_split: Optional[torch.Tensor] = None, eps: float = 1e-5) -> tuple[torch.Tensor, torch.Tensor]:
    """x_new = a . w^T + bias + x and h = LayerNorm(x_new) in one launch (``ralf_gemm_res_ln``; decode loop).
    a: split bf16 [2, M, K]; w: split bf16 [2, 256, K]; x: fp32 [M, 256] (``out_f32`` defaults to updating it in place).
    Returns (x_new fp32 [M, 256], h split bf16 [2, M, 256])."""
    M, K = a.shape[1], a.shape[2]
    assert w.shape[1] == 256 and w.shape[2] == K and x.shape == (M, 256) and x.dtype == torch.float32 and x.stride(1) == 1
    if out_f32 is None:
        out_f32 = x
    if ln_split is None:
        ln_split = torch.empty((2, M, 256), dtype=torch.bfloat16, device=a.device)
    assert ln_split.stride(1) == 256 and ln_split.stride(2) == 1
    g = GemmArgs()
    g.A, g.a_plane, g.lda = a.data_ptr(), a.stride(0), a.stride(1)
    g.W, g.w_plane, g.ldw = w.data_ptr(), w.stride(0), w.stride(1)
    g.M, g.N, g.K, g.npass = M, 256, K, 3
    g.bias = _ptr(bias)
    g.res, g.res_ld = x.data_ptr(), x.stride(0)
    g.out_f32, g.out_ld = out_f32.data_ptr(), out_f32.stride(0)
    check(_lib.lib().ralf_gemm_res_ln(C.byref(g), gamma.data_ptr(), beta.data_ptr(), eps, ln_split.data_ptr(), ln_split.stride(0),
                                      _stream()), "ralf_gemm_res_ln")
    return out_f32, ln_split


def chain_stage(w: torch.Tensor, *, bias: Optional[torch.Tensor] = None, ln: Optional[tuple] = None,
                in_split: Optional[torch.Tensor] = None, act: Optional[str] = None, add_x: bool = False, to_x: bool = False,
                out_operand: bool = False, out_f32: Optional[torch.Tensor] = None, eps: float = 1e-5) -> "_lib.ChainStage":
    """One stage of ``decode_chain``.  w: split bf16 [2, n_out, k_in].  Input: ``ln=(gamma, beta)`` = LayerNorm of the
    residual row, ``in_split`` = split rows [2, B, 256] from global memory, neither = the previous stage's operand output."""
    st = _lib.ChainStage()
    assert w.dim() == 3 and w.shape[0] == 2 and w.dtype == torch.bfloat16 and w.stride(2) == 1
    st.W, st.w_plane, st.ldw, st.n_out, st.k_in = w.data_ptr(), w.stride(0), w.stride(1), w.shape[1], w.shape[2]
    st.in_mode = 1 if ln is not None else (2 if in_split is not None else 0)
    st.gamma, st.beta, st.eps = (_ptr(ln[0]), _ptr(ln[1]), eps) if ln is not None else (None, None, eps)
    if in_split is not None:
        assert in_split.dtype == torch.bfloat16 and in_split.shape[0] == 2 and in_split.stride(2) == 1
        st.in_split, st.in_plane, st.in_ld = in_split.data_ptr(), in_split.stride(0), in_split.stride(1)
    st.bias = _ptr(bias)
    st.act = ACT[act]
    st.add_x, st.to_x, st.out_operand = int(add_x), int(to_x), int(out_operand)
    if out_f32 is not None:
        assert out_f32.dtype == torch.float32 and out_f32.stride(-1) == 1
        st.out_f32, st.out_ld = out_f32.data_ptr(), out_f32.stride(0)
    st._keep = (w, bias, ln, in_split, out_f32)  # keep the tensors alive until the call
    return st


def decode_chain(x: Optional[torch.Tensor], B: int, stages: list) -> None:
    """ralf_decode_chain: the row-local stages of a decoder-layer step for the B new tokens in one kernel."""
    arr = (_lib.ChainStage * len(stages))(*stages)
    check(_lib.lib().ralf_decode_chain(_ptr(x), x.stride(0) if x is not None else 0, B, arr, len(stages), _stream()),
          "ralf_decode_chain")


def knn_topk(
    gallery: torch.Tensor,
    queries: torch.Tensor,
    k: int,
    *,
    index_base: int = 0,
    gallery_max_norm: float = 0.0,
    exact: bool = False,
    workspace: Optional[torch.Tensor] = None,
    fixup: bool = True,
) -> tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
    """Top-k maximum inner product (ralf_knn_topk).  Returns (idx int64 [q,k], score fp32 [q,k], certified).
    ``fixup`` (default): queries the TF32 bound cannot certify are re-run through the exact scan on the device
    (ralf_knn_fixup_exact, no host sync, capturable) -- the result is the exact top-k unconditionally and
    ``certified`` is 1 (proved by the bound) or 2 (exact scan) for every query.  ``fixup=False`` returns the raw
    phase-1/2 result with certified in {0, 1} (tests of the certificate itself)."""
    assert gallery.is_cuda and queries.is_cuda and gallery.dtype == torch.float32 and queries.dtype == torch.float32
    gallery, queries = gallery.contiguous(), queries.contiguous()
    n, d = gallery.shape
    q = queries.shape[0]
    L = _lib.lib()
    ws_bytes = L.ralf_knn_workspace_bytes(n, d, q, k)
    if workspace is None or workspace.numel() < ws_bytes:
        workspace = torch.empty(ws_bytes, dtype=torch.uint8, device=gallery.device)
    idx = torch.empty((q, k), dtype=torch.int64, device=gallery.device)
    score = torch.empty((q, k), dtype=torch.float32, device=gallery.device)
    if exact:
        check(L.ralf_knn_topk_exact(gallery.data_ptr(), n, d, queries.data_ptr(), q, k, index_base,
                                    idx.data_ptr(), score.data_ptr(), workspace.data_ptr(), ws_bytes, _stream()),
              "ralf_knn_topk_exact")
        return idx, score, None
    cert = torch.zeros((q,), dtype=torch.int32, device=gallery.device)
    check(L.ralf_knn_topk(gallery.data_ptr(), n, d, queries.data_ptr(), q, k, index_base,
                          float(gallery_max_norm), idx.data_ptr(), score.data_ptr(), cert.data_ptr(),
                          workspace.data_ptr(), ws_bytes, _stream()), "ralf_knn_topk")
    if fixup:
        check(L.ralf_knn_fixup_exact(gallery.data_ptr(), n, d, queries.data_ptr(), q, k, index_base, cert.data_ptr(),
                                     idx.data_ptr(), score.data_ptr(), workspace.data_ptr(), ws_bytes, _stream()),
              "ralf_knn_fixup_exact")
    return idx, score, cert


def knn_merge(part_score: torch.Tensor, part_idx: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """Merge [parts, q, k] per-shard results into the global top-k (ralf_knn_merge)."""
    parts, q, k = part_score.shape
    part_score, part_idx = part_score.contiguous(), part_idx.contiguous()
    idx = torch.empty((q, k), dtype=torch.int64, device=part_score.device)
    score = torch.empty((q, k), dtype=torch.float32, device=part_score.device)
    check(_lib.lib().ralf_knn_merge(part_score.data_ptr(), part_idx.data_ptr(), parts, q, k, idx.data_ptr(),
                                    score.data_ptr(), _stream()), "ralf_knn_merge")
    return idx, score


# --------------------------------------------------------------------------------------------------
# non-GEMM kernels (csrc/nn_kernels.cu)
# --------------------------------------------------------------------------------------------------
def _split_out(M: int, D: int, device) -> torch.Tensor:
    return torch.empty((2, M, D), dtype=torch.bfloat16, device=device)


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, *, rows: Optional[int] = None,
              in_ld: Optional[int] = None, want_f32: bool = False, want_split: bool = True, eps: float = 1e-5):
    """Rows of x (fp32, last dim D, row stride in_ld) -> (fp32 [M,D] | None, split [2,M,D] | None)."""
    D = x.shape[-1]
    M = rows if rows is not None else x.numel() // D
    ld = in_ld if in_ld is not None else D
    of = torch.empty((M, D), dtype=torch.float32, device=x.device) if want_f32 else None
    os_ = _split_out(M, D, x.device) if want_split else None
    check(_lib.lib().ralf_layernorm(x.data_ptr(), ld, gamma.data_ptr(), beta.data_ptr(), eps, M, D, _ptr(of),
                                    _ptr(os_), os_.stride(0) if os_ is not None else 0, _stream()), "ralf_layernorm")
    return of, os_


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, B: int, H: int, Tq: int, Tk: int, dh: int, *,
              mask: Optional[torch.Tensor] = None, causal: bool = False, scale: Optional[float] = None,
              dropout: Optional[tuple] = None, lse_out: Optional[torch.Tensor] = None):
    """q [B*Tq, >=H*dh] / k, v [B*Tk, >=H*dh] fp32 views (row stride = stride(0)); returns split [2, B*Tq, H*dh].
    ``dropout`` = (seed int64 device tensor [1], site, p): dropout on the attention probabilities (training);
    ``lse_out`` fp32 [B*H*Tq] (with dropout, p > 0): receives the row log-sum-exp for the backward."""
    out = _split_out(B * Tq, H * dh, q.device)
    assert k.stride(0) == v.stride(0)
    sc = scale if scale is not None else dh ** -0.5
    if dropout is not None:
        seed, site, p = dropout
        check(_lib.lib().ralf_attention_dropout(q.data_ptr(), q.stride(0), k.data_ptr(), v.data_ptr(), k.stride(0),
                                                _ptr(mask), B, H, Tq, Tk, dh, int(causal), sc, out.data_ptr(),
                                                out.stride(0), None, H * dh, seed.data_ptr(), site, p, _ptr(lse_out), _stream()),
              "ralf_attention_dropout")
        return out
    check(_lib.lib().ralf_attention(q.data_ptr(), q.stride(0), k.data_ptr(), v.data_ptr(), k.stride(0), _ptr(mask),
                                    B, H, Tq, Tk, dh, int(causal), sc, out.data_ptr(), out.stride(0), None, H * dh,
                                    _stream()), "ralf_attention")
    return out


def attention_decode(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, kv_bstride: int, Tk: int, B: int, H: int,
                     dh: int, *, mask: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None):
    """One query per (b, h) against cached K/V rows (b*kv_bstride + j); returns split [2, B, H*dh]."""
    if out is None:
        out = _split_out(B, H * dh, q.device)
    check(_lib.lib().ralf_attention_decode(q.data_ptr(), q.stride(0), k.data_ptr(), v.data_ptr(), kv_bstride,
                                           k.stride(0), _ptr(mask), mask.stride(0) if mask is not None else 0, Tk, B,
                                           H, dh, dh ** -0.5, out.data_ptr(), out.stride(0), H * dh, _stream()),
          "ralf_attention_decode")
    return out


KV_ROW_BYTES = {24: 1536, 16: 1088}  # bytes per memory token of the quantised cross-attention K/V cache


def attention_decode_kv24(q: torch.Tensor, kv24: torch.Tensor, kv_bstride: int, Tk: int, B: int, H: int, *,
                          out: Optional[torch.Tensor] = None):
    """Decode-step cross-attention over the quantised K/V cache (uint8 [rows, 1536]: 24-bit format; [rows, 1088]: 16-bit
    per-head-scaled format); returns split [2, B, 256]."""
    assert kv24.dtype == torch.uint8 and kv24.shape[-1] in KV_ROW_BYTES.values() and H == 8
    if out is None:
        out = _split_out(B, H * 32, q.device)
    L = _lib.lib()
    fn, name = ((L.ralf_attention_decode_kv16, "ralf_attention_decode_kv16") if kv24.shape[-1] == KV_ROW_BYTES[16]
                else (L.ralf_attention_decode_kv24, "ralf_attention_decode_kv24"))
    check(fn(q.data_ptr(), q.stride(0), kv24.data_ptr(), kv_bstride, Tk, B, H, 32 ** -0.5, out.data_ptr(), out.stride(0),
             H * 32, _stream()), name)
    return out


def attention_decode_append(qkv: torch.Tensor, kcache: torch.Tensor, vcache: torch.Tensor, pos: int, B: int, H: int,
                            dh: int, *, mask: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None):
    """Self-attention decode step: append K/V (columns [D,3D) of qkv) at `pos`, attend over keys 0..pos."""
    if out is None:
        out = _split_out(B, H * dh, qkv.device)
    check(_lib.lib().ralf_attention_decode_append(qkv.data_ptr(), qkv.stride(0), kcache.data_ptr(), vcache.data_ptr(),
                                                  kcache.shape[1], pos, _ptr(mask),
                                                  mask.stride(0) if mask is not None else 0, B, H, dh, dh ** -0.5,
                                                  out.data_ptr(), out.stride(0), H * dh, _stream()),
          "ralf_attention_decode_append")
    return out


def stem_im2col(img: torch.Tensor, KP: int = 200):
    B, C, H, W = img.shape
    assert C == 4 and img.is_contiguous() and img.dtype == torch.float32
    Ho, Wo = (H + 6 - 7) // 2 + 1, (W + 6 - 7) // 2 + 1
    out = _split_out(B * Ho * Wo, KP, img.device)
    check(_lib.lib().ralf_stem_im2col(img.data_ptr(), B, H, W, KP, out.data_ptr(), out.stride(0), _stream()),
          "ralf_stem_im2col")
    return out, Ho, Wo


def stem_s2d(img: torch.Tensor):
    """fp32 NCHW [B,4,H,W] -> zero-bordered space-to-depth split buffer [2, B*(H/2+3)*(W/2+3), 16] (ralf_stem_s2d)."""
    B, C4, H, W = img.shape
    assert C4 == 4 and img.is_contiguous() and img.dtype == torch.float32 and H % 2 == 0 and W % 2 == 0
    Ho, Wo = H // 2, W // 2
    out = _split_out(B * (Ho + 3) * (Wo + 3), 16, img.device)
    check(_lib.lib().ralf_stem_s2d(img.data_ptr(), B, H, W, out.data_ptr(), out.stride(0), _stream()), "ralf_stem_s2d")
    return out, Ho, Wo


def im2col(x: torch.Tensor, B: int, H: int, W: int, C: int, KH: int, KW: int, stride: int, pad: int):
    """x split [2, B*H*W, C] (NHWC) -> split [2, B*Ho*Wo, KH*KW*C]."""
    Ho, Wo = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1
    out = _split_out(B * Ho * Wo, KH * KW * C, x.device)
    check(_lib.lib().ralf_im2col(x.data_ptr(), x.stride(0), B, H, W, C, KH, KW, stride, pad, out.data_ptr(),
                                 out.stride(0), _stream()), "ralf_im2col")
    return out, Ho, Wo


def maxpool3x3s2(x: torch.Tensor, B: int, H: int, W: int, C: int):
    Ho, Wo = (H + 2 - 3) // 2 + 1, (W + 2 - 3) // 2 + 1
    out = _split_out(B * Ho * Wo, C, x.device)
    check(_lib.lib().ralf_maxpool3x3s2(x.data_ptr(), x.stride(0), B, H, W, C, out.data_ptr(), out.stride(0),
                                       _stream()), "ralf_maxpool3x3s2")
    return out, Ho, Wo


def fpn_merge(c5: torch.Tensor, c4: torch.Tensor, B: int, h5: int, w5: int, h4: int, w4: int, C: int):
    fused = _split_out(B * h4 * w4, 2 * C, c5.device)
    summ = _split_out(B * h4 * w4, C, c5.device)
    check(_lib.lib().ralf_fpn_merge(c5.data_ptr(), c4.data_ptr(), B, h5, w5, h4, w4, C, fused.data_ptr(),
                                    fused.stride(0), 2 * C, summ.data_ptr(), summ.stride(0), _stream()),
          "ralf_fpn_merge")
    return fused, summ


def rows_affine(inp: Optional[torch.Tensor], M: int, D: int, *, in_ld: Optional[int] = None, scale: float = 1.0,
                add: float = 0.0, table: Optional[torch.Tensor] = None, tab_mod: int = 0, rows_per_group: int = 0,
                group_stride: int = 0, group_offset: int = 0, out_f32: Optional[torch.Tensor] = None,
                out_split: Optional[torch.Tensor] = None, out_ld: Optional[int] = None):
    ld = in_ld if in_ld is not None else D
    old = out_ld if out_ld is not None else D
    check(_lib.lib().ralf_rows_affine(_ptr(inp), ld, M, D, scale, add, _ptr(table), tab_mod, rows_per_group,
                                      group_stride, group_offset, _ptr(out_f32), _ptr(out_split),
                                      out_split.stride(0) if out_split is not None else 0, old, _stream()),
          "ralf_rows_affine")


def embed(tok: torch.Tensor, tok_col: int, S: int, emb: torch.Tensor, scale: float, pe: torch.Tensor, pos0: int = 0):
    """tok int64 [B, >=tok_col+S] -> fp32 [B*S, D] = emb[tok]*scale + pe[pos0+s]."""
    B, D = tok.shape[0], emb.shape[1]
    out = torch.empty((B * S, D), dtype=torch.float32, device=emb.device)
    check(_lib.lib().ralf_embed(tok.data_ptr(), tok.stride(0), tok_col, B, S, emb.data_ptr(), D, scale, pe.data_ptr(),
                                pos0, out.data_ptr(), _stream()), "ralf_embed")
    return out


def fid_embed(cx, cy, w, h, label, fc_w, fc_b, emb):
    rows, D = cx.numel(), emb.shape[1]
    out = _split_out(rows, 2 * D, emb.device)
    check(_lib.lib().ralf_fid_embed(cx.data_ptr(), cy.data_ptr(), w.data_ptr(), h.data_ptr(), label.data_ptr(), rows,
                                    D, fc_w.data_ptr(), fc_b.data_ptr(), emb.data_ptr(), out.data_ptr(),
                                    out.stride(0), _stream()), "ralf_fid_embed")
    return out


def argmax_next(logits, allowed, seq, pos, pad_mask, pad_id, emb, scale, pe, x_next):
    B, V = logits.shape
    check(_lib.lib().ralf_argmax_next(logits.data_ptr(), logits.stride(0), B, V, allowed.data_ptr(), seq.data_ptr(),
                                      seq.stride(0), pos, _ptr(pad_mask), pad_mask.stride(0) if pad_mask is not None else 0,
                                      pad_id, _ptr(emb), emb.shape[1] if emb is not None else 0, scale, _ptr(pe),
                                      _ptr(x_next), _stream()), "ralf_argmax_next")


def dropout(seed: torch.Tensor, site: int, p: float, *, x_f32: Optional[torch.Tensor] = None,
            x_split: Optional[torch.Tensor] = None, res: Optional[torch.Tensor] = None,
            out_f32: Optional[torch.Tensor] = None, out_split: Optional[torch.Tensor] = None) -> None:
    """y = (res or 0) + dropout_p(x) elementwise over a contiguous [M, C] tensor (fp32 or split planes); the keep mask
    of element i is a pure function of (seed, site, i) -- see include/ralf_b200.h."""
    src = x_f32 if x_f32 is not None else x_split[0]
    assert src.is_contiguous() and (res is None or res.is_contiguous())
    assert out_f32 is None or out_f32.is_contiguous()
    check(_lib.lib().ralf_dropout(_ptr(x_f32), _ptr(x_split), x_split.stride(0) if x_split is not None else 0, _ptr(res),
                                  src.numel(), seed.data_ptr(), site, p, _ptr(out_f32), _ptr(out_split),
                                  out_split.stride(0) if out_split is not None else 0, _stream()), "ralf_dropout")


def dropout_mask(seed: torch.Tensor, site: int, p: float, total: int) -> torch.Tensor:
    out = torch.empty(total, dtype=torch.uint8, device=seed.device)
    check(_lib.lib().ralf_dropout_mask(seed.data_ptr(), site, p, total, out.data_ptr(), _stream()), "ralf_dropout_mask")
    return out


SAMPLING_MODES = {"deterministic": 0, "random": 1, "top_k": 2, "top_p": 3, "gumbel": 4}


def sample_next(logits, allowed, seq, pos, pad_mask, pad_id, emb, scale, pe, x_next, *, forced=None, step=0,
                mode="deterministic", temperature=1.0, top_k=5, top_p=0.9, uniform=None, noise=None):
    """Step tail with decoding-space restriction (``forced`` int32 [B, S], -1 = free) and helpers/sampling.py's
    stochastic samplers; ``uniform`` fp32 [B] (and ``noise`` fp32 [B, V] for gumbel) are the random numbers."""
    B, V = logits.shape
    check(_lib.lib().ralf_sample_next(logits.data_ptr(), logits.stride(0), B, V, allowed.data_ptr(), _ptr(forced),
                                      forced.stride(0) if forced is not None else 0, step, SAMPLING_MODES[mode],
                                      float(temperature), int(top_k), float(top_p), _ptr(uniform), _ptr(noise),
                                      noise.stride(0) if noise is not None else 0, seq.data_ptr(), seq.stride(0), pos,
                                      _ptr(pad_mask), pad_mask.stride(0) if pad_mask is not None else 0, pad_id, _ptr(emb),
                                      emb.shape[1] if emb is not None else 0, scale, _ptr(pe), _ptr(x_next), _stream()),
          "ralf_sample_next")


def kv_append(qkv, kcache, vcache, pos):
    B, D = qkv.shape[0], qkv.shape[1] // 3
    check(_lib.lib().ralf_kv_append(qkv.data_ptr(), B, D, kcache.data_ptr(), vcache.data_ptr(), kcache.shape[1], pos,
                                    _stream()), "ralf_kv_append")


def ce_label_smooth(logits: torch.Tensor, targets: torch.Tensor, eps: float, ignore_index: int) -> torch.Tensor:
    """Mean label-smoothed CE over logits [..., V] / targets [...] (ralf_ce_label_smooth); 0-dim fp32 tensor."""
    V = logits.shape[-1]
    lg = logits.reshape(-1, V)
    tg = targets.reshape(-1).to(torch.int64).contiguous()
    M = lg.shape[0]
    ws = torch.empty(2 * M, dtype=torch.float32, device=lg.device)
    out = torch.empty(1, dtype=torch.float32, device=lg.device)
    check(_lib.lib().ralf_ce_label_smooth(lg.data_ptr(), lg.stride(0), tg.data_ptr(), M, V, eps, ignore_index,
                                          ws.data_ptr(), out.data_ptr(), _stream()), "ralf_ce_label_smooth")
    return out[0]


def gather_layouts(table: torch.Tensor, idx: torch.Tensor, index_base: int = 0) -> torch.Tensor:
    """table fp32 [N, 6, E]; idx int64 [...] -> fp32 [..., 6, E] (ralf_gather_layouts)."""
    n, six, E = table.shape
    flat = idx.reshape(-1).contiguous()
    out = torch.empty((*idx.shape, six, E), dtype=torch.float32, device=table.device)
    check(_lib.lib().ralf_gather_layouts(table.data_ptr(), flat.data_ptr(), flat.numel(), six * E, n, index_base,
                                         out.data_ptr(), _stream()), "ralf_gather_layouts")
    return out


def fid_embed_packed(packed: torch.Tensor, fc_w, fc_b, emb):
    """packed fp32 [nseq, 6, E] -> (split rows [2, nseq*E, 2D], pad mask uint8 [nseq, E+1])."""
    nseq, _, E = packed.shape
    D = emb.shape[1]
    out = _split_out(nseq * E, 2 * D, emb.device)
    pad = torch.empty((nseq, E + 1), dtype=torch.uint8, device=emb.device)
    check(_lib.lib().ralf_fid_embed_packed(packed.data_ptr(), nseq, E, D, fc_w.data_ptr(), fc_b.data_ptr(),
                                           emb.data_ptr(), emb.shape[0], out.data_ptr(), out.stride(0), pad.data_ptr(),
                                           _stream()), "ralf_fid_embed_packed")
    return out, pad
