"""Training forward / backward of the ResNet50-FPN trunk on the tape (image branch).

Reference graph: ``ResnetBackbone.forward`` (image2layout/train/models/common/image.py:90-120) in ``model.train()``
mode, i.e. BatchNorm uses batch statistics and updates its running buffers (momentum 0.1).  Every convolution is
im2col (or nothing for 1x1/stride 1) + the tcgen05 GEMM; BatchNorm, ReLU and the residual add are one fused
elementwise kernel after a deterministic column-statistics reduction.  Backward: BN backward (two column sums
+ apply), dW = dZ^T . Xcol and dXcol = dZ . W on the same GEMM, col2im gather for k > 1 or stride > 1.
Activations are NHWC, split bf16; gradients fp32.
"""
from __future__ import annotations

import torch

from . import autograd as ag
from . import ops
from .autograd import Node, ParamStore, Tape, _L
from .ops import _ptr, _stream, check

BODY = "encoder.extractor.body"
EXT = "encoder.extractor"
LAYERS = [(3, 1, 64), (4, 2, 128), (6, 2, 256), (3, 2, 512)]


class Trunk:
    def __init__(self, ps: ParamStore, model, device) -> None:
        self.ps, self.dev = ps, device
        self.buffers = dict(model.named_buffers())
        self.cw: dict[str, torch.Tensor] = {}    # k > 1 conv weights in GEMM layout, split [2, N, Kp]
        self.cwT: dict[str, torch.Tensor] = {}   # and transposed, split [2, Kp, Np]
        self.kconvs: list[tuple[str, int, int, int]] = []  # (param, N, C, T)
        reg = ps.register_gemm_weight
        self.kconvs.append((BODY + ".conv1.weight", 64, 4, 49))
        for li, (nblk, stride, planes) in enumerate(LAYERS, start=1):
            for bi in range(nblk):
                p = f"{BODY}.layer{li}.{bi}"
                reg(p + ".conv1", p + ".conv1.weight")
                self.kconvs.append((p + ".conv2.weight", planes, planes, 9))
                reg(p + ".conv3", p + ".conv3.weight")
                if bi == 0:
                    reg(p + ".downsample.0", p + ".downsample.0.weight")
        for n in ("fpn_conv11_4", "fpn_conv11_5", "proj"):
            reg(f"{EXT}.{n}", f"{EXT}.{n}.weight")
        self.kconvs.append((EXT + ".fpn_conv33.weight", 256, 256, 9))
        for (pname, N, Cc, T) in self.kconvs:
            Kp = (Cc * T + 7) // 8 * 8
            Np = (N + 7) // 8 * 8
            self.cw[pname] = torch.zeros((2, N, Kp), dtype=torch.bfloat16, device=device)
            self.cwT[pname] = torch.zeros((2, Kp, Np), dtype=torch.bfloat16, device=device)

    def refresh_operands(self) -> None:
        for (pname, N, Cc, T) in self.kconvs:
            w, wT = self.cw[pname], self.cwT[pname]
            check(_L().ralf_conv_weight_to_gemm(self.ps.p(pname).data_ptr(), N, Cc, T, w.shape[2], w.data_ptr(), w.stride(0),
                                                wT.data_ptr(), wT.stride(0), wT.shape[2], _stream()),
                  "ralf_conv_weight_to_gemm")

    # ---- BatchNorm(train) pieces ---------------------------------------------------------------
    def _bn_stats(self, z: torch.Tensor, bn: str):
        M, Cn = z.shape
        mean = torch.empty(Cn, dtype=torch.float32, device=self.dev)
        rstd = torch.empty(Cn, dtype=torch.float32, device=self.dev)
        ws = torch.empty(2 * Cn * ((M + 255) // 256), dtype=torch.float32, device=self.dev)
        rm, rv = self.buffers[bn + ".running_mean"], self.buffers[bn + ".running_var"]
        check(_L().ralf_bn_colstats(z.data_ptr(), None, None, None, 0, M, Cn, 1e-5, 0.1, mean.data_ptr(), rstd.data_ptr(),
                                    rm.data_ptr(), rv.data_ptr(), ws.data_ptr(), _stream()), "ralf_bn_colstats")
        self.buffers[bn + ".num_batches_tracked"].add_(1)
        return mean, rstd

    # ---- conv + BN(train) (+ residual) (+ ReLU) --------------------------------------------------
    def conv_bn(self, tape: Tape, x: Node, geom, conv: str, bn: str, k: int, stride: int, pad: int, relu: bool,
                res: Node | None = None):
        """x.s split NHWC [2, B*H*W, Cin]; returns (Node with split output [2, B*Ho*Wo, Cout], (B, Ho, Wo))."""
        ps = self.ps
        B, H, W = geom
        Cin = x.Cn
        pname = conv + ".weight"
        direct = (k == 1 and stride == 1)
        if direct:
            col, Ho, Wo = x.s, H, W
        else:
            col, Ho, Wo = ops.im2col(x.s, B, H, W, Cin, k, k, stride, pad)
        wfw = ps.w[conv] if k == 1 else self.cw[pname]
        z, _ = ops.gemm(col, wfw)
        M, Cout = z.shape
        mean, rstd = self._bn_stats(z, bn)
        gamma, beta = ps.p(bn + ".weight"), ps.p(bn + ".bias")
        out = torch.empty((2, M, Cout), dtype=torch.bfloat16, device=self.dev)
        check(_L().ralf_bn_apply(z.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                                 _ptr(res.s) if res is not None else None, res.s.stride(0) if res is not None else 0,
                                 int(relu), M, Cout, out.data_ptr(), out.stride(0), None, _stream()), "ralf_bn_apply")
        y = Node(M, Cout, None, out)

        def bwd() -> None:
            dy = y.grad
            if dy is None:
                return
            if relu:
                check(_L().ralf_relu_bwd(dy.data_ptr(), out.data_ptr(), dy.numel(), _stream()), "ralf_relu_bwd")
            if res is not None:
                ag.accumulate(res, dy)
            ws = torch.empty(2 * Cout * ((M + 255) // 256), dtype=torch.float32, device=self.dev)
            check(_L().ralf_bn_colstats(dy.data_ptr(), z.data_ptr(), mean.data_ptr(), rstd.data_ptr(), 1, M, Cout, 1e-5, 0.0,
                                        ps.g(bn + ".bias").data_ptr(), ps.g(bn + ".weight").data_ptr(), None, None,
                                        ws.data_ptr(), _stream()), "ralf_bn_colstats")
            dz = torch.empty_like(z)
            check(_L().ralf_bn_bwd_apply(dy.data_ptr(), z.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                                         ps.g(bn + ".bias").data_ptr(), ps.g(bn + ".weight").data_ptr(), M, Cout,
                                         dz.data_ptr(), _stream()), "ralf_bn_bwd_apply")
            self._conv_bwd(x, geom, col, dz, conv, k, stride, pad, direct)
            y.grad = None

        tape.record(bwd)
        return y, (B, Ho, Wo)

    def _conv_bwd(self, x: Node, geom, col: torch.Tensor, dz: torch.Tensor, conv: str, k: int, stride: int, pad: int,
                  direct: bool, need_dx: bool = True) -> None:
        """dW (into the parameter gradient) and dX (accumulated into x.grad) of z = col . W^T."""
        ps = self.ps
        B, H, W = geom
        pname = conv + ".weight"
        want_dx = need_dx and x.need_grad
        if want_dx:
            dzs, dzT = ag.split_and_transpose(dz)  # dgrad and wgrad operands in one pass over dz
        else:
            dzT = ag.transpose_to_split(x_f32=dz)
        colT = ag.transpose_to_split(x_split=col)
        if k == 1:
            ops.gemm(dzT, colT, out_f32=ps.weight_view(conv, grad=True), splitk=True)
        else:
            N, Kp = self.cw[pname].shape[1], self.cw[pname].shape[2]
            dwg, _ = ops.gemm(dzT, colT, splitk=True)                                # [N, Kcol] in (tap, cin) order
            _, Nn, Cc, T = next(kc for kc in self.kconvs if kc[0] == pname)
            check(_L().ralf_conv_grad_from_gemm(dwg.data_ptr(), Nn, Cc, T, dwg.stride(0), ps.g(pname).data_ptr(), _stream()),
                  "ralf_conv_grad_from_gemm")
        if not want_dx:
            return
        wT = ps.wT[conv] if k == 1 else self.cwT[pname][:, :col.shape[2], :dz.shape[1]]
        if direct:
            dx, _ = ops.gemm(dzs, wT, res=x.grad)
            x.grad = dx
        else:
            dcol, _ = ops.gemm(dzs, wT)
            dx = x.grad if x.grad is not None else torch.empty((x.M, x.Cn), dtype=torch.float32, device=self.dev)
            check(_L().ralf_col2im(dcol.data_ptr(), B, H, W, x.Cn, k, k, stride, pad, dx.data_ptr(),
                                   int(x.grad is not None), _stream()), "ralf_col2im")
            x.grad = dx

    # ---- whole trunk ---------------------------------------------------------------------------
    def forward(self, tape: Tape, image: torch.Tensor, pos2d_fn, mark=None) -> tuple[Node, int, int]:
        """image fp32 [B,4,H,W] -> tokens Node (fp32 [B*h*w, 256], 2-D sine PE added), h, w.
        ``mark(tag)``: called DURING BACKWARD when the gradients of a parameter bucket are complete ("rest": everything
        outside the ResNet body, i.e. when the body's backward is about to start; "layer4" / "layer3" / "layer2": that
        stage's parameters) -- the data-parallel trainer starts the bucket's all-reduce there, under the remaining backward."""
        ps, dev = self.ps, self.dev
        if mark is None:
            mark = lambda tag: None
        B = image.shape[0]
        col0, H, W = ops.stem_im2col(image.contiguous())
        stem_name = BODY + ".conv1"
        z, _ = ops.gemm(col0, self.cw[stem_name + ".weight"])
        M = z.shape[0]
        mean, rstd = self._bn_stats(z, BODY + ".bn1")
        gamma, beta = ps.p(BODY + ".bn1.weight"), ps.p(BODY + ".bn1.bias")
        s0 = torch.empty((2, M, 64), dtype=torch.bfloat16, device=dev)
        check(_L().ralf_bn_apply(z.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), beta.data_ptr(), None, 0, 1,
                                 M, 64, s0.data_ptr(), s0.stride(0), None, _stream()), "ralf_bn_apply")
        stem = Node(M, 64, None, s0)
        img_node = Node(M, 200, None, col0, need_grad=False)

        def stem_bwd() -> None:
            dy = stem.grad
            if dy is None:
                return
            check(_L().ralf_relu_bwd(dy.data_ptr(), s0.data_ptr(), dy.numel(), _stream()), "ralf_relu_bwd")
            ws = torch.empty(2 * 64 * ((M + 255) // 256), dtype=torch.float32, device=dev)
            check(_L().ralf_bn_colstats(dy.data_ptr(), z.data_ptr(), mean.data_ptr(), rstd.data_ptr(), 1, M, 64, 1e-5, 0.0,
                                        ps.g(BODY + ".bn1.bias").data_ptr(), ps.g(BODY + ".bn1.weight").data_ptr(), None, None,
                                        ws.data_ptr(), _stream()), "ralf_bn_colstats")
            dz = torch.empty_like(z)
            check(_L().ralf_bn_bwd_apply(dy.data_ptr(), z.data_ptr(), mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(),
                                         ps.g(BODY + ".bn1.bias").data_ptr(), ps.g(BODY + ".bn1.weight").data_ptr(), M, 64,
                                         dz.data_ptr(), _stream()), "ralf_bn_bwd_apply")
            self._conv_bwd(img_node, (B, 0, 0), col0, dz, stem_name, 7, 2, 3, False, need_dx=False)
            stem.grad = None

        tape.record(stem_bwd)
        # max-pool
        p0, Hp, Wp = ops.maxpool3x3s2(s0, B, H, W, 64)
        pool = Node(p0.shape[1], 64, None, p0)

        def pool_bwd() -> None:
            if pool.grad is None:
                return
            dx = torch.empty((M, 64), dtype=torch.float32, device=dev)
            taps = torch.empty(pool.grad.numel(), dtype=torch.uint8, device=dev)
            check(_L().ralf_maxpool3x3s2_bwd(s0.data_ptr(), s0.stride(0), pool.grad.data_ptr(), B, H, W, 64, dx.data_ptr(),
                                             taps.data_ptr(), _stream()), "ralf_maxpool3x3s2_bwd")
            ag.accumulate(stem, dx)
            pool.grad = None

        tape.record(pool_bwd)
        x, geom = pool, (B, Hp, Wp)
        feats = {}
        for li, (nblk, stride, planes) in enumerate(LAYERS, start=1):
            for bi in range(nblk):
                p = f"{BODY}.layer{li}.{bi}"
                s = stride if bi == 0 else 1
                t1, g1 = self.conv_bn(tape, x, geom, p + ".conv1", p + ".bn1", 1, 1, 0, True)
                t2, g2 = self.conv_bn(tape, t1, g1, p + ".conv2", p + ".bn2", 3, s, 1, True)
                if bi == 0:
                    idt, _ = self.conv_bn(tape, x, geom, p + ".downsample.0", p + ".downsample.1", 1, s, 0, False)
                else:
                    idt = x
                x, geom = self.conv_bn(tape, t2, g2, p + ".conv3", p + ".bn3", 1, 1, 0, True, res=idt)
            feats[li] = (x, geom)
            if li < 4:  # backward reaches this point when stage li+1 is done
                tape.record(lambda tag=f"layer{li + 1}": mark(tag))
        tape.record(lambda: mark("rest"))  # backward: FPN + everything downstream is done, the body's backward starts
        # ---- FPN (image.py:99-111) ----
        (l3, (_, h4, w4)), (l4, (_, h5, w5)) = feats[3], feats[4]
        c4 = ag.linear(tape, ps, l3, EXT + ".fpn_conv11_4", EXT + ".fpn_conv11_4.bias")
        c5 = ag.linear(tape, ps, l4, EXT + ".fpn_conv11_5", EXT + ".fpn_conv11_5.bias")
        fused_s, summ_s = ops.fpn_merge(c5.f32, c4.f32, B, h5, w5, h4, w4, 256)
        M4 = B * h4 * w4
        a33, _, _ = ops.im2col(summ_s, B, h4, w4, 256, 3, 3, 1, 1)
        n33 = EXT + ".fpn_conv33"
        ops.gemm(a33, self.cw[n33 + ".weight"], bias=ps.p(n33 + ".bias"), out_split=fused_s, out_col0=256, want_f32=False)
        fused = Node(M4, 512, None, fused_s)
        summ = Node(M4, 256, None, summ_s)

        def fpn_bwd() -> None:
            g = fused.grad  # [M4, 512] fp32
            if g is None:
                return
            d33 = torch.empty((M4, 256), dtype=torch.float32, device=dev)
            dup = torch.empty((M4, 256), dtype=torch.float32, device=dev)
            check(_L().ralf_rows_gather(g[:, 256:].data_ptr(), g.stride(0), M4, 256, 1.0, 0, 0, 0, d33.data_ptr(), 0, _stream()),
                  "ralf_rows_gather")
            check(_L().ralf_rows_gather(g.data_ptr(), g.stride(0), M4, 256, 1.0, 0, 0, 0, dup.data_ptr(), 0, _stream()),
                  "ralf_rows_gather")
            ag.colsum(d33, ps.g(n33 + ".bias"))
            self._conv_bwd(summ, (B, h4, w4), a33, d33, n33, 3, 1, 1, False)   # -> summ.grad = d(up + c4)
            ag.accumulate(c4, summ.grad)
            ag.axpy(dup, summ.grad)
            d5 = torch.empty((B * h5 * w5, 256), dtype=torch.float32, device=dev)
            check(_L().ralf_upsample_nearest_bwd(dup.data_ptr(), 256, B, h5, w5, h4, w4, 256, d5.data_ptr(), _stream()),
                  "ralf_upsample_nearest_bwd")
            ag.accumulate(c5, d5)
            fused.grad = None

        tape.record(fpn_bwd)
        tokens = ag.linear(tape, ps, fused, EXT + ".proj", EXT + ".proj.bias", res_table=pos2d_fn(h4, w4),
                           res_row_mod=h4 * w4)
        return tokens, h4, w4
