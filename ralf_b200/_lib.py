"""ctypes binding of the C ABI declared in include/ralf_b200.h.

There is deliberately no fallback: if the CUDA library is missing, importing the product path
raises.  (The library is built in-tree by ``ralf_b200.build`` / ``__graft_entry__.build``.)
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RALF_B200_LIB", os.path.join(HERE, "libralf_b200.so"))  # override: A/B builds

STATUS = {
    0: "RALF_OK", -1: "RALF_ERR_SHAPE", -2: "RALF_ERR_ALIGN", -3: "RALF_ERR_NULL", -4: "RALF_ERR_CUDA",
    -5: "RALF_ERR_DRIVER", -6: "RALF_ERR_ARCH", -7: "RALF_ERR_WORKSPACE",
}


class RalfError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("a_plane", C.c_longlong), ("lda", C.c_int),
        ("W", C.c_void_p), ("w_plane", C.c_longlong), ("ldw", C.c_int),
        ("M", C.c_int), ("N", C.c_int), ("K", C.c_int),
        ("npass", C.c_int), ("block_n", C.c_int),
        ("bias", C.c_void_p), ("act", C.c_int), ("post_relu", C.c_int),
        ("res", C.c_void_p), ("res_split", C.c_void_p), ("res_plane", C.c_longlong),
        ("res_ld", C.c_int), ("res_row_mod", C.c_int),
        ("out_f32", C.c_void_p), ("out_split", C.c_void_p), ("out_plane", C.c_longlong),
        ("out_split_lo", C.c_int), ("out_ld", C.c_int), ("out_col0", C.c_int),
        ("rows_per_group", C.c_int), ("group_stride", C.c_int), ("group_offset", C.c_int),
        ("out_kv24", C.c_void_p), ("out_kv_fmt", C.c_int),
        ("splitk_ws", C.c_void_p), ("splitk_ws_bytes", C.c_size_t),
    ]


class RefreshTask(C.Structure):
    _fields_ = [("src", C.c_void_p), ("ld_in", C.c_longlong), ("R", C.c_int), ("C", C.c_int),
                ("w", C.c_void_p), ("w_plane", C.c_longlong), ("w_ld", C.c_longlong),
                ("wt", C.c_void_p), ("wt_plane", C.c_longlong), ("wt_ld", C.c_longlong),
                ("tile0", C.c_int), ("tiles_x", C.c_int)]


class ChainStage(C.Structure):
    _fields_ = [
        ("W", C.c_void_p), ("w_plane", C.c_longlong), ("ldw", C.c_int), ("n_out", C.c_int), ("k_in", C.c_int),
        ("in_mode", C.c_int), ("gamma", C.c_void_p), ("beta", C.c_void_p), ("eps", C.c_float),
        ("in_split", C.c_void_p), ("in_plane", C.c_longlong), ("in_ld", C.c_int),
        ("bias", C.c_void_p), ("act", C.c_int), ("add_x", C.c_int), ("to_x", C.c_int), ("out_operand", C.c_int),
        ("out_f32", C.c_void_p), ("out_ld", C.c_int),
    ]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RalfError(
                f"{LIB_PATH} is missing: the sm_100a CUDA library has not been built "
                "(run `python -m ralf_b200.build`); there is no CPU fallback."
            )
        L = C.CDLL(LIB_PATH)
        L.ralf_last_cuda_error.restype = C.c_char_p
        L.ralf_knn_workspace_bytes.restype = C.c_size_t
        L.ralf_knn_workspace_bytes.argtypes = [C.c_int] * 4
        L.ralf_knn_topk.argtypes = [
            C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_float,
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
        ]
        L.ralf_knn_topk_exact.argtypes = [
            C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_longlong,
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
        ]
        L.ralf_knn_fixup_exact.argtypes = [
            C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_longlong,
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
        ]
        L.ralf_knn_merge.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_void_p]
        L.ralf_gemm.argtypes = [C.POINTER(GemmArgs), C.c_void_p]
        L.ralf_gemm_splitk_workspace_bytes.restype = C.c_size_t
        L.ralf_gemm_splitk_workspace_bytes.argtypes = [C.c_int] * 3
        L.ralf_conv_gemm.argtypes = [C.POINTER(GemmArgs)] + [C.c_int] * 6 + [C.c_void_p]
        L.ralf_gemm_res_ln.argtypes = [C.POINTER(GemmArgs), C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_longlong, C.c_void_p]
        L.ralf_conv_gemm_strided.argtypes = [C.POINTER(GemmArgs)] + [C.c_int] * 7 + [C.c_void_p]
        L.ralf_stem_gemm.argtypes = [C.POINTER(GemmArgs)] + [C.c_int] * 3 + [C.c_void_p]
        L.ralf_stem_s2d.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p]
        L.ralf_gemm_ln.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.POINTER(GemmArgs), C.c_void_p]
        L.ralf_decode_chain.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(ChainStage), C.c_int, C.c_void_p]
        L.ralf_check_device.argtypes = [C.c_int]
        vp, i, ll, f = C.c_void_p, C.c_int, C.c_longlong, C.c_float
        L.ralf_layernorm.argtypes = [vp, ll, vp, vp, f, i, i, vp, vp, ll, vp]
        L.ralf_attention.argtypes = [vp, i, vp, vp, i, vp, i, i, i, i, i, i, f, vp, ll, vp, i, vp]
        L.ralf_attention_decode.argtypes = [vp, i, vp, vp, ll, i, vp, i, i, i, i, i, f, vp, ll, i, vp]
        L.ralf_attention_decode_kv24.argtypes = [vp, i, vp, ll, i, i, i, f, vp, ll, i, vp]
        L.ralf_attention_decode_kv16.argtypes = [vp, i, vp, ll, i, i, i, f, vp, ll, i, vp]
        L.ralf_attention_decode_append.argtypes = [vp, i, vp, vp, i, i, vp, i, i, i, i, f, vp, ll, i, vp]
        L.ralf_stem_im2col.argtypes = [vp, i, i, i, i, vp, ll, vp]
        L.ralf_im2col.argtypes = [vp, ll, i, i, i, i, i, i, i, i, vp, ll, vp]
        L.ralf_maxpool3x3s2.argtypes = [vp, ll, i, i, i, i, vp, ll, vp]
        L.ralf_fpn_merge.argtypes = [vp, vp, i, i, i, i, i, i, vp, ll, i, vp, ll, vp]
        L.ralf_rows_affine.argtypes = [vp, ll, i, i, f, f, vp, i, i, i, i, vp, vp, ll, i, vp]
        L.ralf_embed.argtypes = [vp, ll, i, i, i, vp, i, f, vp, i, vp, vp]
        L.ralf_fid_embed.argtypes = [vp, vp, vp, vp, vp, i, i, vp, vp, vp, vp, ll, vp]
        L.ralf_argmax_next.argtypes = [vp, i, i, i, vp, vp, i, i, vp, i, ll, vp, i, f, vp, vp, vp]
        L.ralf_kv_append.argtypes = [vp, i, i, vp, vp, i, i, vp]
        L.ralf_sample_next.argtypes = [vp, i, i, i, vp, vp, i, i, i, f, i, f, vp, vp, i, vp, i, i, vp, i, ll, vp, i, f, vp,
                                       vp, vp]
        L.ralf_gather_layouts.argtypes = [vp, vp, i, i, ll, ll, vp, vp]
        L.ralf_fid_embed_packed.argtypes = [vp, i, i, i, vp, vp, vp, i, vp, ll, vp, vp]
        L.ralf_transpose_to_split.argtypes = [vp, vp, ll, ll, i, i, vp, ll, ll, vp]
        L.ralf_to_split.argtypes = [vp, ll, vp, ll, vp]
        L.ralf_refresh_operands.argtypes = [vp, i, i, vp]
        L.ralf_split_and_transpose.argtypes = [vp, ll, i, i, vp, ll, ll, vp, ll, ll, vp]
        L.ralf_colsum.argtypes = [vp, ll, i, i, vp, i, vp, vp]
        L.ralf_colsum_workspace_bytes.restype = C.c_size_t
        L.ralf_colsum_workspace_bytes.argtypes = [i, i]
        L.ralf_layernorm_bwd.argtypes = [vp, ll, vp, vp, f, i, i, vp, vp, vp, vp, vp, vp]
        L.ralf_attention_bwd.argtypes = [vp, i, vp, vp, i, vp, i, i, i, i, i, i, f, vp, ll, vp, i, vp, vp, vp, i, vp, vp, i, vp]
        u32 = C.c_uint
        L.ralf_attention_bwd_dropout.argtypes = [vp, i, vp, vp, i, vp, i, i, i, i, i, i, f, vp, ll, vp, i, vp, vp, vp, i, vp, vp,
                                                 i, vp, u32, f, i, vp]
        L.ralf_attention_dropout.argtypes = [vp, i, vp, vp, i, vp, i, i, i, i, i, i, f, vp, ll, vp, i, vp, u32, f, vp, vp]
        L.ralf_dropout.argtypes = [vp, vp, ll, vp, ll, vp, u32, f, vp, vp, ll, vp]
        L.ralf_dropout_mask.argtypes = [vp, u32, f, ll, vp, vp]
        L.ralf_ce_label_smooth_bwd.argtypes = [vp, i, vp, i, i, f, ll, vp, f, vp, i, vp]
        L.ralf_relu_bwd.argtypes = [vp, vp, ll, vp]
        L.ralf_gelu_fwd.argtypes = [vp, ll, vp, ll, vp]
        L.ralf_gelu_bwd.argtypes = [vp, vp, ll, vp]
        L.ralf_axpy.argtypes = [vp, vp, f, ll, vp]
        L.ralf_rows_gather.argtypes = [vp, ll, i, i, f, i, i, i, vp, i, vp]
        L.ralf_embed_bwd.argtypes = [vp, ll, i, i, i, vp, i, f, vp, i, vp]
        L.ralf_grad_norm.argtypes = [vp, ll, vp, vp, vp]
        L.ralf_adamw_step.argtypes = [vp, vp, vp, vp, ll, vp, f, f, f, f, f, f, i, vp]
        L.ralf_adamw_step_dyn.argtypes = [vp, vp, vp, vp, ll, vp, f, f, f, f, f, f, vp, vp]
        L.ralf_bn_colstats.argtypes = [vp, vp, vp, vp, i, i, i, f, f, vp, vp, vp, vp, vp, vp]
        L.ralf_bn_apply.argtypes = [vp, vp, vp, vp, vp, vp, ll, i, i, i, vp, ll, vp, vp]
        L.ralf_bn_bwd_apply.argtypes = [vp, vp, vp, vp, vp, vp, vp, i, i, vp, vp]
        L.ralf_col2im.argtypes = [vp, i, i, i, i, i, i, i, i, vp, i, vp]
        L.ralf_maxpool3x3s2_bwd.argtypes = [vp, ll, vp, i, i, i, i, vp, vp, vp]
        L.ralf_upsample_nearest_bwd.argtypes = [vp, ll, i, i, i, i, i, i, vp, vp]
        L.ralf_conv_weight_to_gemm.argtypes = [vp, i, i, i, i, vp, ll, vp, ll, i, vp]
        L.ralf_conv_grad_from_gemm.argtypes = [vp, i, i, i, i, vp, vp]
        L.ralf_ce_label_smooth.argtypes = [vp, i, vp, i, i, f, ll, vp, vp, vp]
        _lib = L
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = STATUS.get(rc, str(rc))
        if rc == -4:
            msg += ": " + lib().ralf_last_cuda_error().decode()
        raise RalfError(f"{what} failed: {msg}")
