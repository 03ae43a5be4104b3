"""A minimal reverse-mode tape over the ralf_b200 kernels (training rows a12/a13 of SURVEY.md 8).

There is no torch.autograd on this path: every forward op below launches our kernels through the C ABI and
records a closure that launches the matching backward kernels.  ``Tape.backward()`` replays the closures in
reverse.  Gradients are fp32; GEMM operands (activations, weights and their transposes) are split bf16 so the
backward contractions  dX = dY.W  and  dW = dY^T.X  run on the same tcgen05 GEMM as the forward (bf16x3).

Node = a [M, C] activation held as fp32 (``f32``), split bf16 (``s``), or both, plus its accumulated ``grad``.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from . import _lib, ops
from .ops import _ptr, _stream, check


class Node:
    __slots__ = ("f32", "s", "grad", "M", "Cn", "need_grad")

    def __init__(self, M: int, Cn: int, f32: Optional[torch.Tensor] = None, s: Optional[torch.Tensor] = None,
                 need_grad: bool = True) -> None:
        self.M, self.Cn, self.f32, self.s, self.grad, self.need_grad = M, Cn, f32, s, None, need_grad


class DropoutState:
    """Dropout of one training step: probability, the device-resident step seed and a call-site counter (sites are
    numbered in forward order, so forward and backward of an op agree and a captured graph keeps its numbering)."""

    def __init__(self, p: float, seed: torch.Tensor) -> None:
        self.p, self.seed, self.site = p, seed, 0

    def next_site(self) -> int:
        self.site += 1
        return self.site


class Tape:
    def __init__(self, drop: Optional[DropoutState] = None) -> None:
        self.ops: list[Callable[[], None]] = []
        self.drop = drop if (drop is not None and drop.p > 0.0) else None

    def record(self, fn: Callable[[], None]) -> None:
        self.ops.append(fn)

    def backward(self) -> None:
        for fn in reversed(self.ops):
            fn()
        self.ops.clear()


# ---- thin wrappers over the training kernels ---------------------------------------------------------
def _L():
    return _lib.lib()


def to_split(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 [M, C] -> split [2, M, C]; the row pitch is padded to a multiple of 8 elements when C is not (TMA needs
    16-byte row pitches), in which case a [:, :, :C] view of the padded buffer is returned.  ``out``: a tensor this
    function returned earlier for the same shape -- rewritten in place (operand buffers must keep their address across
    optimiser steps so that a captured training graph keeps reading the live weights)."""
    M, Cn = x.shape
    if Cn % 8 == 0 and x.is_contiguous():
        if out is None:
            out = torch.empty((2, M, Cn), dtype=torch.bfloat16, device=x.device)
        check(_L().ralf_to_split(x.data_ptr(), x.numel(), out.data_ptr(), out.stride(0), _stream()), "ralf_to_split")
        return out
    Cp = (Cn + 7) // 8 * 8
    if out is None:
        out = torch.empty((2, M, Cp), dtype=torch.bfloat16, device=x.device)[:, :, :Cn]
    ops.rows_affine(x, M, Cn, in_ld=x.stride(0), out_split=out, out_ld=Cp)
    return out


def transpose_to_split(x_f32: Optional[torch.Tensor] = None, x_split: Optional[torch.Tensor] = None,
                       out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[R, C] (fp32 or split) -> split [2, C, Rp] with Rp = R rounded up to 8 (row pitch for the tensor map);
    the logical K extent of the result is R (columns >= R are never read)."""
    src = x_f32 if x_f32 is not None else x_split[0]
    R, Cn = src.shape[-2], src.shape[-1]
    Rp = (R + 7) // 8 * 8
    if out is None:  # else: the [2, C, R] view (row stride Rp) returned by an earlier call, rewritten in place
        out = torch.empty((2, Cn, Rp), dtype=torch.bfloat16, device=src.device)[:, :, :R]
    check(_L().ralf_transpose_to_split(_ptr(x_f32), _ptr(x_split), x_split.stride(0) if x_split is not None else 0,
                                       src.stride(-2), R, Cn, out.data_ptr(), out.stride(0), Rp, _stream()),
          "ralf_transpose_to_split")
    return out  # shape [2, C, R], row stride Rp


def split_and_transpose(x: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """fp32 [R, C] (contiguous rows) -> (split [2, R, C], split^T [2, C, R] with row stride R rounded up to 8) in one pass.
    Falls back to the two separate kernels when C is not a multiple of 8 (padded row pitch of ``to_split``)."""
    R, Cn = x.shape
    if Cn % 8 or not x.is_contiguous():
        return to_split(x), transpose_to_split(x_f32=x)
    Rp = (R + 7) // 8 * 8
    s = torch.empty((2, R, Cn), dtype=torch.bfloat16, device=x.device)
    t = torch.empty((2, Cn, Rp), dtype=torch.bfloat16, device=x.device)[:, :, :R]
    check(_L().ralf_split_and_transpose(x.data_ptr(), x.stride(0), R, Cn, s.data_ptr(), s.stride(0), s.stride(1),
                                        t.data_ptr(), t.stride(0), Rp, _stream()), "ralf_split_and_transpose")
    return s, t


def colsum(x: torch.Tensor, out: torch.Tensor, accumulate: bool = False) -> None:
    M, Cn = x.shape
    nb = _L().ralf_colsum_workspace_bytes(M, Cn)
    ws = torch.empty(nb // 4, dtype=torch.float32, device=x.device) if nb else None
    check(_L().ralf_colsum(x.data_ptr(), x.stride(0), M, Cn, out.data_ptr(), int(accumulate), _ptr(ws), _stream()),
          "ralf_colsum")


def axpy(a: torch.Tensor, b: torch.Tensor, alpha: float = 1.0) -> None:
    check(_L().ralf_axpy(a.data_ptr(), b.data_ptr(), alpha, a.numel(), _stream()), "ralf_axpy")


def accumulate(node: Node, g: torch.Tensor) -> None:
    """node.grad += g (takes ownership of g on first use; every grad buffer is consumed exactly once)."""
    if not node.need_grad:
        return
    if node.grad is None:
        node.grad = g
    else:
        axpy(node.grad, g)


# ---- parameters -------------------------------------------------------------------------------------
class ParamStore:
    """Flat fp32 parameter / gradient / Adam-moment buffers with per-parameter views, plus the split-bf16 GEMM
    operands (W and W^T) that are refreshed from the master weights after every optimizer step."""

    def __init__(self, named_params: list[tuple[str, torch.Tensor]], groups: list[list[str]], device) -> None:
        order = [n for g in groups for n in g]
        byname = dict(named_params)
        self.offsets, off = {}, 0
        for n in order:
            self.offsets[n] = (off, byname[n].numel(), tuple(byname[n].shape))
            off += (byname[n].numel() + 3) // 4 * 4  # 16-byte aligned views
        self.total = off
        self.flat_p = torch.zeros(off, dtype=torch.float32, device=device)
        self.flat_g = torch.zeros(off, dtype=torch.float32, device=device)
        self.flat_m = torch.zeros(off, dtype=torch.float32, device=device)
        self.flat_v = torch.zeros(off, dtype=torch.float32, device=device)
        self.group_ranges = []
        for g in groups:
            if g:
                a = self.offsets[g[0]][0]
                last = self.offsets[g[-1]]
                self.group_ranges.append((a, (last[0] + last[1] + 3) // 4 * 4))
            else:
                self.group_ranges.append((0, 0))
        for n in order:
            o, cnt, shape = self.offsets[n]
            self.flat_p[o:o + cnt].copy_(byname[n].detach().reshape(-1))
        self.w: dict[str, torch.Tensor] = {}
        self.wT: dict[str, torch.Tensor] = {}
        self.gemm_weights: dict[str, tuple[str, int, int]] = {}  # name -> (param, row0, rows)

    def p(self, name: str) -> torch.Tensor:
        o, cnt, shape = self.offsets[name]
        return self.flat_p[o:o + cnt].view(shape)

    def g(self, name: str) -> torch.Tensor:
        o, cnt, shape = self.offsets[name]
        return self.flat_g[o:o + cnt].view(shape)

    def register_gemm_weight(self, name: str, param: str, row0: int = 0, rows: Optional[int] = None) -> None:
        shape = self.offsets[param][2]
        self.gemm_weights[name] = (param, row0, rows if rows is not None else shape[0])

    def weight_view(self, name: str, grad: bool = False) -> torch.Tensor:
        param, r0, rows = self.gemm_weights[name]
        t = self.g(param) if grad else self.p(param)
        return t.reshape(t.shape[0], -1)[r0:r0 + rows]

    def refresh_operands(self) -> None:
        """master fp32 weights -> split W [2,N,K] and W^T [2,K,N] for every registered GEMM weight.  The first call
        allocates the operand buffers weight by weight; later calls (one per optimiser step) are ONE multi-tensor launch
        over a device-resident task table (the buffers and the master weights keep their addresses)."""
        if getattr(self, "_refresh_table", None) is not None and self._refresh_names == list(self.gemm_weights):
            tab, n, tiles = self._refresh_table
            check(_L().ralf_refresh_operands(tab.data_ptr(), n, tiles, _stream()), "ralf_refresh_operands")
            return
        for name in self.gemm_weights:
            w = self.weight_view(name)
            self.w[name] = to_split(w, out=self.w.get(name))
            self.wT[name] = transpose_to_split(x_f32=w, out=self.wT.get(name))
        import ctypes as C

        tasks, tile0 = [], 0
        for name in self.gemm_weights:
            w, ws, wt = self.weight_view(name), self.w[name], self.wT[name]
            t = _lib.RefreshTask()
            t.src, t.ld_in, t.R, t.C = w.data_ptr(), w.stride(0), w.shape[0], w.shape[1]
            t.w, t.w_plane, t.w_ld = ws.data_ptr(), ws.stride(0), ws.stride(1)
            t.wt, t.wt_plane, t.wt_ld = wt.data_ptr(), wt.stride(0), wt.stride(1)
            t.tiles_x = (t.C + 31) // 32
            t.tile0 = tile0
            tile0 += t.tiles_x * ((t.R + 31) // 32)
            tasks.append(t)
        if tasks:
            arr = (_lib.RefreshTask * len(tasks))(*tasks)
            host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
            self._refresh_table = (host.to(self.flat_p.device), len(tasks), tile0)
            self._refresh_names = list(self.gemm_weights)


# ---- ops --------------------------------------------------------------------------------------------
def linear(tape: Tape, ps: ParamStore, x: Node, wname: str, bias: Optional[str] = None, *, act: Optional[str] = None,
           res: Optional[Node] = None, res_table: Optional[torch.Tensor] = None, res_row_mod: int = 0,
           want_f32: bool = True, want_split: bool = False, npass: int = 3) -> Node:
    """y = act(x W^T + b) (+ res).  act in (None, "relu").  GELU is a separate op (needs the pre-activation)."""
    assert x.s is not None
    b = ps.p(bias) if bias is not None else None
    f32, s = ops.gemm(x.s, ps.w[wname], bias=b, act=act, res=(res.f32 if res is not None else res_table),
                      res_row_mod=res_row_mod, want_f32=want_f32, want_split=want_split or act == "relu", npass=npass)
    y = Node(x.M, ps.w[wname].shape[1], f32, s)

    def bwd() -> None:
        dy = y.grad
        if dy is None:
            return
        if act == "relu":
            check(_L().ralf_relu_bwd(dy.data_ptr(), y.s.data_ptr(), dy.numel(), _stream()), "ralf_relu_bwd")
        if res is not None:
            accumulate(res, dy)  # residual branch shares dy (read-only from here on)
        if x.need_grad:
            dys, dyT = split_and_transpose(dy)        # [2, M, N] and [2, N, M] in one pass
        else:
            dyT = transpose_to_split(x_f32=dy)        # [2, N, M]
        xT = transpose_to_split(x_split=x.s)          # [2, K, M]
        ops.gemm(dyT, xT, out_f32=ps.weight_view(wname, grad=True), npass=npass, splitk=True)   # dW = dY^T X
        if bias is not None:
            colsum(dy, ps.g(bias))
        if x.need_grad:
            dx, _ = ops.gemm(dys, ps.wT[wname], res=x.grad, npass=npass)          # dX = dY W (+ existing grad)
            x.grad = dx
        y.grad = None

    tape.record(bwd)
    return y


def gelu(tape: Tape, z: Node) -> Node:
    out = torch.empty((2, z.M, z.Cn), dtype=torch.bfloat16, device=z.f32.device)
    check(_L().ralf_gelu_fwd(z.f32.data_ptr(), z.f32.numel(), out.data_ptr(), out.stride(0), _stream()), "ralf_gelu_fwd")
    y = Node(z.M, z.Cn, None, out)

    def bwd() -> None:
        if y.grad is None:
            return
        check(_L().ralf_gelu_bwd(y.grad.data_ptr(), z.f32.data_ptr(), y.grad.numel(), _stream()), "ralf_gelu_bwd")
        accumulate(z, y.grad)
        y.grad = None

    tape.record(bwd)
    return y


def layernorm(tape: Tape, ps: ParamStore, x: Node, name: str, *, want_f32: bool = False, rows: Optional[int] = None,
              in_ld: Optional[int] = None) -> Node:
    """nn.LayerNorm over rows of x.f32 (optionally a strided row subset: rows / in_ld)."""
    M = rows if rows is not None else x.M
    ld = in_ld if in_ld is not None else x.Cn
    gamma, beta = ps.p(name + ".weight"), ps.p(name + ".bias")
    f32, s = ops.layernorm(x.f32, gamma, beta, rows=M, in_ld=ld, want_f32=want_f32, want_split=True)
    y = Node(M, x.Cn, f32, s)
    assert rows is None, "strided-row LayerNorm backward is not needed on the trainable path"

    def bwd() -> None:
        if y.grad is None:
            return
        D = x.Cn
        nblocks = min((M + 7) // 8, 4 * torch.cuda.get_device_properties(x.f32.device).multi_processor_count)
        ws = torch.empty(2 * D * nblocks, dtype=torch.float32, device=x.f32.device)
        dx = torch.empty_like(x.f32)
        check(_L().ralf_layernorm_bwd(x.f32.data_ptr(), ld, y.grad.data_ptr(), gamma.data_ptr(), 1e-5, M, D,
                                      _ptr(x.grad), dx.data_ptr(), ps.g(name + ".weight").data_ptr(),
                                      ps.g(name + ".bias").data_ptr(), ws.data_ptr(), _stream()), "ralf_layernorm_bwd")
        x.grad = dx
        y.grad = None

    tape.record(bwd)
    return y


def dropout_inplace(tape: Tape, node: Node) -> None:
    """nn.Dropout applied to ``node`` in place (fp32 and / or split storage); the backward masks node.grad in place,
    between the consumer that produced it and the producer that reads it."""
    dr = tape.drop
    if dr is None:
        return
    site = dr.next_site()
    if node.f32 is not None:
        ops.dropout(dr.seed, site, dr.p, x_f32=node.f32, out_f32=node.f32, out_split=node.s)
    else:
        ops.dropout(dr.seed, site, dr.p, x_split=node.s, out_split=node.s)

    def bwd() -> None:
        if node.grad is not None:
            ops.dropout(dr.seed, site, dr.p, x_f32=node.grad, out_f32=node.grad)

    tape.record(bwd)


def dropout_add(tape: Tape, z: Node, res: Node) -> Node:
    """y = res + dropout(z)  (the ``x + dropoutN(sublayer(x))`` of nn.TransformerEncoder/DecoderLayer)."""
    dr = tape.drop
    site = dr.next_site()
    y = Node(z.M, z.Cn, torch.empty_like(z.f32), None)
    ops.dropout(dr.seed, site, dr.p, x_f32=z.f32, res=res.f32, out_f32=y.f32)

    def bwd() -> None:
        dy = y.grad
        if dy is None:
            return
        g = torch.empty_like(dy)
        ops.dropout(dr.seed, site, dr.p, x_f32=dy, out_f32=g)
        accumulate(z, g)
        accumulate(res, dy)
        y.grad = None

    tape.record(bwd)
    return y


def linear_res(tape: Tape, ps: ParamStore, x: Node, wname: str, bias: Optional[str], res: Node) -> Node:
    """res + dropout(x W^T + b): fused into the GEMM epilogue when dropout is off."""
    if tape.drop is None:
        return linear(tape, ps, x, wname, bias, res=res)
    return dropout_add(tape, linear(tape, ps, x, wname, bias), res)


def _attention_bwd(q, ldq, k, v, ldk, mask, B, H, Tq, Tk, dh, causal, o_split, dO, dq, lddq, dk, dv, lddk, dropout=None,
                   lse=None):
    """``lse``: the row log-sum-exp saved by the forward kernel (training with dropout): the dQ kernel then sweeps the keys
    once instead of twice."""
    dev = dO.device
    lse_given = lse is not None
    if lse is None:
        lse = torch.empty(B * H * Tq, dtype=torch.float32, device=dev)
    delta = torch.empty(B * H * Tq, dtype=torch.float32, device=dev)
    if dropout is not None:
        seed, site, p = dropout
        check(_L().ralf_attention_bwd_dropout(q.data_ptr(), ldq, k.data_ptr(), v.data_ptr(), ldk, _ptr(mask), B, H, Tq, Tk,
                                              dh, int(causal), dh ** -0.5, o_split.data_ptr(), o_split.stride(0),
                                              dO.data_ptr(), dO.stride(0), lse.data_ptr(), delta.data_ptr(), dq.data_ptr(),
                                              lddq, dk.data_ptr(), dv.data_ptr(), lddk, seed.data_ptr(), site, p,
                                              int(lse_given), _stream()), "ralf_attention_bwd_dropout")
        return
    check(_L().ralf_attention_bwd(q.data_ptr(), ldq, k.data_ptr(), v.data_ptr(), ldk, _ptr(mask), B, H, Tq, Tk, dh,
                                  int(causal), dh ** -0.5, o_split.data_ptr(), o_split.stride(0), dO.data_ptr(),
                                  dO.stride(0), lse.data_ptr(), delta.data_ptr(), dq.data_ptr(), lddq, dk.data_ptr(),
                                  dv.data_ptr(), lddk, _stream()), "ralf_attention_bwd")


def self_attention(tape: Tape, qkv: Node, B: int, T: int, H: int, dh: int, *, mask: Optional[torch.Tensor] = None,
                   causal: bool = False) -> Node:
    """qkv.f32 [B*T, 3*H*dh] (fused projection) -> attention output (split) [B*T, H*dh]."""
    Dm = H * dh
    x = qkv.f32
    dr = tape.drop
    dropout = (dr.seed, dr.next_site(), dr.p) if dr is not None else None  # attention-probability dropout
    lse = torch.empty(B * H * T, dtype=torch.float32, device=x.device) if (dropout is not None and dropout[2] > 0) else None
    out = ops.attention(x[:, :Dm], x[:, Dm:2 * Dm], x[:, 2 * Dm:], B, H, T, T, dh, mask=mask, causal=causal,
                        dropout=dropout, lse_out=lse)
    y = Node(qkv.M, Dm, None, out)

    def bwd() -> None:
        if y.grad is None:
            return
        d = torch.empty_like(x)
        _attention_bwd(x[:, :Dm], x.stride(0), x[:, Dm:2 * Dm], x[:, 2 * Dm:], x.stride(0), mask, B, H, T, T, dh, causal,
                       out, y.grad, d[:, :Dm], d.stride(0), d[:, Dm:2 * Dm], d[:, 2 * Dm:], d.stride(0), dropout=dropout, lse=lse)
        accumulate(qkv, d)
        y.grad = None

    tape.record(bwd)
    return y


def cross_attention(tape: Tape, q: Node, kv: Node, kcol: int, vcol: int, B: int, Tq: int, Tk: int, H: int, dh: int, *,
                    use_dropout: bool = True) -> Node:
    """q.f32 [B*Tq, H*dh]; kv.f32 [B*Tk, ncols] with K at columns [kcol, kcol+H*dh), V at [vcol, ...).
    ``use_dropout=False``: the fusion Attention of the reference is built with dropout=0.0 (:701-703)."""
    Dm = H * dh
    kk, vv = kv.f32[:, kcol:kcol + Dm], kv.f32[:, vcol:vcol + Dm]
    dr = tape.drop if use_dropout else None
    dropout = (dr.seed, dr.next_site(), dr.p) if dr is not None else None
    lse = torch.empty(B * H * Tq, dtype=torch.float32, device=q.f32.device) if (dropout is not None and dropout[2] > 0) else None
    out = ops.attention(q.f32, kk, vv, B, H, Tq, Tk, dh, dropout=dropout, lse_out=lse)
    y = Node(q.M, Dm, None, out)

    def bwd() -> None:
        if y.grad is None:
            return
        assert kv.Cn == 2 * Dm, "K and V column blocks must cover the kv tensor"
        dq = torch.empty_like(q.f32)
        g = torch.empty_like(kv.f32)
        _attention_bwd(q.f32, q.f32.stride(0), kk, vv, kv.f32.stride(0), None, B, H, Tq, Tk, dh, False, out, y.grad, dq,
                       dq.stride(0), g[:, kcol:kcol + Dm], g[:, vcol:vcol + Dm], g.stride(0), dropout=dropout, lse=lse)
        accumulate(kv, g)
        accumulate(q, dq)
        y.grad = None

    tape.record(bwd)
    return y


def embed(tape: Tape, ps: ParamStore, tok: torch.Tensor, S: int, emb_name: str, scale: float, pe: torch.Tensor) -> Node:
    """emb[tok] * scale + pe[s]  (BaseDecoder front end); backward scatters into the embedding gradient."""
    emb = ps.p(emb_name)
    x = ops.embed(tok, 0, S, emb, scale, pe, 0)
    y = Node(x.shape[0], x.shape[1], x, None)

    def bwd() -> None:
        if y.grad is None:
            return
        check(_L().ralf_embed_bwd(tok.data_ptr(), tok.stride(0), 0, tok.shape[0], S, y.grad.data_ptr(), emb.shape[1], scale,
                                  ps.g(emb_name).data_ptr(), emb.shape[0], _stream()), "ralf_embed_bwd")
        y.grad = None

    tape.record(bwd)
    return y


def place_rows(tape: Tape, src: Node, dst_f32: torch.Tensor, dst_node: Node, rpg: int, gs: int, go: int, *,
               scale: float = 1.0, add: float = 0.0, dst_split: Optional[torch.Tensor] = None) -> None:
    """dst[(r/rpg)*gs + go + r%rpg, :] = src[r, :] * scale + add   (concatenation along the token axis)."""
    ops.rows_affine(src.f32, src.M, src.Cn, scale=scale, add=add, rows_per_group=rpg, group_stride=gs, group_offset=go,
                    out_f32=dst_f32, out_split=dst_split)

    def bwd() -> None:
        if dst_node.grad is None or not src.need_grad:
            return
        g = torch.empty_like(src.f32)
        check(_L().ralf_rows_gather(dst_node.grad.data_ptr(), dst_node.grad.stride(0), src.M, src.Cn, scale, rpg, gs, go,
                                    g.data_ptr(), 0, _stream()), "ralf_rows_gather")
        accumulate(src, g)

    tape.record(bwd)


def ce_loss(tape: Tape, logits: Node, targets: torch.Tensor, eps: float, ignore_index: int) -> torch.Tensor:
    """Mean label-smoothed cross entropy; backward seeds logits.grad."""
    V = logits.Cn
    lg = logits.f32
    tg = targets.reshape(-1).to(torch.int64).contiguous()
    M = lg.shape[0]
    ws = torch.empty(2 * M, dtype=torch.float32, device=lg.device)
    out = torch.empty(1, dtype=torch.float32, device=lg.device)
    check(_L().ralf_ce_label_smooth(lg.data_ptr(), lg.stride(0), tg.data_ptr(), M, V, eps, ignore_index, ws.data_ptr(),
                                    out.data_ptr(), _stream()), "ralf_ce_label_smooth")

    def bwd() -> None:
        d = torch.empty_like(lg)
        check(_L().ralf_ce_label_smooth_bwd(lg.data_ptr(), lg.stride(0), tg.data_ptr(), M, V, eps, ignore_index,
                                            ws.data_ptr(), 1.0, d.data_ptr(), d.stride(0), _stream()),
              "ralf_ce_label_smooth_bwd")
        logits.grad = d

    tape.record(bwd)
    return out[0]


# ---- optimizer --------------------------------------------------------------------------------------
def grad_norm(flat_g: torch.Tensor) -> torch.Tensor:
    ws = torch.empty(1024, dtype=torch.float32, device=flat_g.device)
    out = torch.empty(1, dtype=torch.float32, device=flat_g.device)
    check(_L().ralf_grad_norm(flat_g.data_ptr(), flat_g.numel(), ws.data_ptr(), out.data_ptr(), _stream()), "ralf_grad_norm")
    return out


def adamw_step(ps: ParamStore, group_cfg: list[tuple[float, float]], step: int, max_norm: float, norm: torch.Tensor,
               betas=(0.9, 0.999), eps: float = 1e-8, dyn: Optional[torch.Tensor] = None) -> None:
    """torch.optim.AdamW + clip_grad_norm_(max_norm) over the flat buffers, one launch per (lr, wd) group.
    ``dyn``: device fp32 [3] = {lr scale, 1 - beta1^t, 1 - beta2^t} read by the kernel instead of ``step`` (graph replay)."""
    for (a, b), (lr, wd) in zip(ps.group_ranges, group_cfg):
        if b <= a:
            continue
        n = b - a
        if dyn is not None:
            check(_L().ralf_adamw_step_dyn(ps.flat_p[a:b].data_ptr(), ps.flat_g[a:b].data_ptr(), ps.flat_m[a:b].data_ptr(),
                                           ps.flat_v[a:b].data_ptr(), n, norm.data_ptr(), max_norm, lr, betas[0], betas[1],
                                           eps, wd, dyn.data_ptr(), _stream()), "ralf_adamw_step_dyn")
            continue
        check(_L().ralf_adamw_step(ps.flat_p[a:b].data_ptr(), ps.flat_g[a:b].data_ptr(), ps.flat_m[a:b].data_ptr(),
                                   ps.flat_v[a:b].data_ptr(), n, norm.data_ptr(), max_norm, lr, betas[0], betas[1], eps, wd,
                                   step, _stream()), "ralf_adamw_step")
