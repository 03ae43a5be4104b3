"""Torch-tensor front end of the C ABI (device pointers + current stream; no torch types cross the ABI)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import GemmArgs, check

ACT = {None: 0, "none": 0, "relu": 1, "gelu": 2}


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
) -> tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """D = A . W^T with the fused epilogue of ``ralf_gemm`` (include/ralf_b200.h).

    Outputs are allocated when not supplied ([M, N] fp32 and/or [2, M, N] split bf16)."""
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.dim() == 3 and w.dim() == 3
    M, K = a.shape[1], a.shape[2]
    N = w.shape[1]
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
    check(_lib.lib().ralf_gemm(C.byref(g), _stream()), "ralf_gemm")
    return out_f32, out_split


def knn_topk(
    gallery: torch.Tensor,
    queries: torch.Tensor,
    k: int,
    *,
    index_base: int = 0,
    gallery_max_norm: float = 0.0,
    exact: bool = False,
    workspace: Optional[torch.Tensor] = None,
) -> tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
    """Top-k maximum inner product (ralf_knn_topk).  Returns (idx int64 [q,k], score fp32 [q,k], certified)."""
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
