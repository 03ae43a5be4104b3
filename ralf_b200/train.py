"""Training step of the RALF model on the ralf_b200 kernels.

Mirrors ``train()`` of the reference (image2layout/train/train.py:431-459): forward (teacher forced) ->
label-smoothed CE -> backward -> clip_grad_norm_(0.1) -> AdamW (4 parameter groups from
``BaseModel.optim_groups``, base_model.py:207-347: decay / no-decay x {ResNet body lr*0.1, rest}) ->
(data parallel) gradient all-reduce over NCCL BEFORE the clip, i.e. what DDP is meant to do -- the reference's
own wrapper never arms the reducer, we deliberately do the real thing.

Status (round 1): every trainable parameter of the reference trains through the tape in autograd.py -- the
ResNet50-FPN trunk (BatchNorm in training mode, train_conv.py), image encoder, layout adapter, fusion attention,
head, constraint encoder, decoder, loss; FIDNetV3 stays frozen like in the reference
(retrieval_augmented_autoreg.py:150-154).  Dropout (p = 0.1) is applied at the reference's sites: attention
probabilities, the three residual branches and the FFN activation of every image-encoder / constraint-encoder / decoder
layer (nn.TransformerEncoder/DecoderLayer(dropout=0.1), :105,116-126; common/common.py:26-35,216) and the three
1-D positional encodings (positional_encoding.py:67-107); the fusion Attention, head and layout adapter are built with
dropout 0.0 in the reference.  Masks come from a counter-based generator (csrc/common.cuh), not torch's Philox stream,
so runs are reproducible per ``seed`` but not mask-identical to the reference.  One deliberate deviation: the frozen
FIDNetV3 is evaluated without dropout (the reference leaves it in train mode under ``model.train()``, which only adds
noise to a frozen feature extractor) -- reported by ``TrainEngine.limits``.
"""
from __future__ import annotations

import math
from typing import Optional

import torch

from . import autograd as ag
from . import ops
from .autograd import Node, ParamStore, Tape
from .engine import D, NHEAD, NLAYER, Engine, _sine_pe_1d


def _is_decay(name: str, p: torch.Tensor) -> bool:
    """base_model.py:207-247: Linear / Conv / MultiheadAttention weights decay; biases, LayerNorm / BatchNorm /
    Embedding weights do not."""
    if name.endswith("bias") or p.dim() < 2:
        return False
    if "emb" in name.split(".")[-2] if "." in name else False:
        return False
    return True


class _TapeLoss(torch.autograd.Function):
    """Bridge between the kernel tape and torch.autograd (see TrainEngine.loss_with_grad)."""

    @staticmethod
    def forward(ctx, anchor, loss, tape, engine):
        ctx.tape, ctx.engine = tape, engine
        return loss.detach().clone()

    @staticmethod
    def backward(ctx, g):
        ctx.tape.backward()
        ctx.engine._publish_grads(g)
        return None, None, None, None


class TrainEngine:
    limits = ("frozen FIDNetV3 evaluated without dropout",)

    @property
    def seed(self) -> int:
        """Per-rank dropout seed: every rank draws its own masks from one base seed."""
        return (self.base_seed * 0x9E3779B97F4A7C15 + self.rank * 0xD1B54A32D192ED03) & (2 ** 64 - 1)

    def __init__(self, model, *, lr: float = 1e-4, weight_decay: float = 1e-4, body_lr_scale: float = 0.1,
                 max_grad_norm: float = 0.1, world_size: int = 1, process_group=None, train_trunk: bool = True,
                 dropout: float = 0.1, seed: int = 0, rank: int = 0) -> None:
        self.model = model
        self.dropout = float(dropout)
        self.rank = int(rank)
        self.base_seed = int(seed)  # what checkpoints store; the per-rank stream is derived from it
        self.dev = model.device
        self.lr, self.wd, self.body_scale, self.max_norm = lr, weight_decay, body_lr_scale, max_grad_norm
        self.world, self.pg = world_size, process_group
        self.step_count = 0
        self.is_ralf = bool(getattr(model, "IS_RALF", True))  # False: the Autoreg baseline (models/autoreg.py:590-622)
        self.infer = Engine(model.state_dict(), self.dev, is_ralf=self.is_ralf, top_k=model.top_k)  # frozen trunk + FIDNet
        self.train_trunk = train_trunk
        named = [(n, p) for n, p in model.named_parameters()
                 if p.requires_grad and (train_trunk or not n.startswith("encoder.extractor"))]
        # optim_groups(custom_lr={"encoder.extractor.body": lr * 0.1}) -> body decay / body no-decay / decay / no-decay
        groups = [[], [], [], []]
        for n, p in named:
            body = n.startswith("encoder.extractor.body")
            groups[(0 if body else 2) + (0 if _is_decay(n, p) else 1)].append(n)
        groups = [sorted(g) for g in groups]
        self.ps = ParamStore(named, groups, self.dev)
        self.group_cfg = [(lr * body_lr_scale, weight_decay), (lr * body_lr_scale, 0.0), (lr, weight_decay), (lr, 0.0)]
        for n, p in named:  # the module's parameters become views of the flat master buffer
            p.data = self.ps.p(n)
        self._named = named
        self._register_weights()
        self.trunk = None
        if train_trunk:
            from .train_conv import Trunk

            self.trunk = Trunk(self.ps, model, self.dev)
        self.pe = _sine_pe_1d(5000, D).to(self.dev)
        # per-step host scalars go through a ring of pinned slots: the async copy of step t must not see step t+1's values
        self._dyn_host = torch.ones((64, 3), dtype=torch.float32).pin_memory()
        self._dyn = torch.ones(3, dtype=torch.float32, device=self.dev)
        self._seed_host = torch.zeros(64, dtype=torch.int64).pin_memory()
        self._seed_dev = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self._slot_done: list = [None] * 64  # event after the async copies out of ring slot i: waited on before rewriting it
        self._graph = None
        self._buckets = self._grad_buckets()
        self._bucket_done: set = set()
        self._comm_stream = torch.cuda.Stream(device=self.dev) if (world_size > 1 and self.dev.type == "cuda") else None
        self.overlap_comm = True   # False: one blocking all-reduce of the flat buffer after the backward (A/B, bench)
        self.skip_comm = False     # bench only: time the step without its all-reduce
        self.refresh_operands()

    # ------------------------------------------------------------------------------------------
    # data-parallel gradient all-reduce, bucketed in backward-completion order and overlapped with the rest of the
    # backward (decoder / encoders / FPN -> ResNet layer4 -> layer3 -> layer2 -> layer1 + stem); the
    # reference's DDPWrapper was meant to do this through torch's reducer (helpers/distrubuted.py:10-31).
    # ------------------------------------------------------------------------------------------
    def _grad_buckets(self) -> dict:
        """tag -> [(lo, hi)] element ranges of the flat gradient buffer.  Parameters are laid out group by group (body
        decay | body no-decay | rest decay | rest no-decay), names sorted inside a group, so every bucket is at most two
        contiguous ranges."""
        ps = self.ps
        body = "encoder.extractor.body."

        def ranges(pred):
            out = []
            for n, (off, cnt, _) in ps.offsets.items():  # insertion order = buffer order
                if not pred(n):
                    continue
                end = off + (cnt + 3) // 4 * 4
                if out and out[-1][1] == off:
                    out[-1][1] = end
                else:
                    out.append([off, end])
            return [tuple(r) for r in out]

        b = {"rest": ranges(lambda n: not n.startswith(body))}
        for li in (4, 3, 2):
            b[f"layer{li}"] = ranges(lambda n, li=li: n.startswith(f"{body}layer{li}."))
        b["tail"] = ranges(lambda n: n.startswith(body) and not any(n.startswith(f"{body}layer{li}.") for li in (2, 3, 4)))
        return {k: v for k, v in b.items() if v}

    def _allreduce_ranges(self, rs) -> None:
        import torch.distributed as dist

        for lo, hi in rs:
            dist.all_reduce(self.ps.flat_g[lo:hi], op=dist.ReduceOp.SUM, group=self.pg)

    def _mark(self, tag: str) -> None:
        """Backward reached the point where bucket ``tag`` is complete: all-reduce it on the communication stream."""
        if self.world <= 1 or self.skip_comm or not self.overlap_comm or tag in self._bucket_done or tag not in self._buckets:
            return
        self._bucket_done.add(tag)
        cur = torch.cuda.current_stream()
        self._comm_stream.wait_stream(cur)  # the bucket's gradients are produced on the compute stream
        with torch.cuda.stream(self._comm_stream):
            self._allreduce_ranges(self._buckets[tag])

    def _finish_allreduce(self) -> None:
        """After the backward: reduce whatever no marker has sent yet, join the communication stream, average."""
        if self.world <= 1 or self.skip_comm:
            return
        ps = self.ps
        if not self.overlap_comm:
            self._allreduce_ranges([(0, ps.total)])
        else:
            cur = torch.cuda.current_stream()
            rest = [r for tag, rs in self._buckets.items() if tag not in self._bucket_done for r in rs]
            self._comm_stream.wait_stream(cur)
            with torch.cuda.stream(self._comm_stream):
                self._allreduce_ranges(rest)
            cur.wait_stream(self._comm_stream)
            self._bucket_done.clear()
        ps.flat_g.mul_(1.0 / self.world)

    def refresh_operands(self) -> None:
        self.ps.refresh_operands()
        if self.trunk is not None:
            self.trunk.refresh_operands()

    # ------------------------------------------------------------------------------------------
    def _register_weights(self) -> None:
        ps = self.ps
        reg = ps.register_gemm_weight

        def enc(p):
            reg(p + ".qkv", p + ".self_attn.in_proj_weight")
            reg(p + ".o", p + ".self_attn.out_proj.weight")
            reg(p + ".l1", p + ".linear1.weight")
            reg(p + ".l2", p + ".linear2.weight")

        for i in range(NLAYER):
            enc(f"transformer_encoder.layers.{i}")
            enc(f"user_const_encoder.encoder.layers.{i}")
            p = f"decoder.transformer.layers.{i}"
            enc(p)
            reg(p + ".cq", p + ".multihead_attn.in_proj_weight", 0, D)
            reg(p + ".ckv", p + ".multihead_attn.in_proj_weight", D, 2 * D)
            reg(p + ".co", p + ".multihead_attn.out_proj.weight")
        if self.is_ralf:
            for n in ("layout_adapter.net.1", "layout_adapter.net.4", "head.net.1", "head.net.4", "attn.to_q", "attn.to_kv",
                      "attn.to_out.0"):
                reg(n, n + ".weight")
        reg("decoder.head.1", "decoder.head.1.weight")

    # ------------------------------------------------------------------------------------------
    def _enc_layer(self, tape, x: Node, p: str, B: int, T: int, mask=None) -> Node:
        ps = self.ps
        h = ag.layernorm(tape, ps, x, p + ".norm1")
        qkv = ag.linear(tape, ps, h, p + ".qkv", p + ".self_attn.in_proj_bias")
        a = ag.self_attention(tape, qkv, B, T, NHEAD, D // NHEAD, mask=mask)
        x1 = ag.linear_res(tape, ps, a, p + ".o", p + ".self_attn.out_proj.bias", x)
        h = ag.layernorm(tape, ps, x1, p + ".norm2")
        f = ag.linear(tape, ps, h, p + ".l1", p + ".linear1.bias", act="relu", want_f32=False)
        ag.dropout_inplace(tape, f)
        return ag.linear_res(tape, ps, f, p + ".l2", p + ".linear2.bias", x1)

    def _ffn_gelu(self, tape, x: Node, p: str) -> Node:
        """common/attention.py:15-30  LN -> Linear -> GELU -> Linear."""
        ps = self.ps
        h = ag.layernorm(tape, ps, x, p + ".net.0")
        z = ag.linear(tape, ps, h, p + ".net.1", p + ".net.1.bias")
        g = ag.gelu(tape, z)
        return ag.linear(tape, ps, g, p + ".net.4", p + ".net.4.bias")

    def _scalar_add_bwd(self, tape, node: Node, pname: str, row: int) -> None:
        """y = x + task_emb[row] was applied while concatenating; route sum(dy) to the scalar parameter."""
        ps = self.ps

        def bwd() -> None:
            if node.grad is None:
                return
            cs = torch.empty(node.Cn, dtype=torch.float32, device=self.dev)
            ag.colsum(node.grad, cs)
            ag.colsum(cs.view(node.Cn, 1), ps.g(pname).view(-1)[row:row + 1], accumulate=True)

        tape.record(bwd)

    # ------------------------------------------------------------------------------------------
    def forward_loss(self, inputs: dict, targets: dict):
        """Teacher-forced forward + loss through the tape.  Returns (loss 0-dim tensor, tape, logits Node)."""
        ps, dev = self.ps, self.dev
        tape = Tape(ag.DropoutState(self.dropout, self._seed_dev))
        image = inputs["image"].to(dev, torch.float32)
        B = image.shape[0]
        K = self.model.top_k
        # ---- ResNet50-FPN trunk: on the tape (BatchNorm batch statistics), or frozen through the inference kernels ----
        if self.trunk is not None:
            x, h, w = self.trunk.forward(tape, image, self.infer.pos2d, mark=self._mark)
        else:
            tokens, h, w = self.infer.resnet_fpn(image)
            x = Node(B * h * w, D, tokens, None, need_grad=False)
        T = h * w
        for i in range(NLAYER):
            x = self._enc_layer(tape, x, f"transformer_encoder.layers.{i}", B, T)
        if self.is_ralf:
            # ---- retrieved layouts: frozen FIDNet CLS features -> trainable adapter ----
            cls = self._fid_cls(inputs["retrieved"], B)              # fp32 [B*K, 256]
            ref0 = self._ffn_gelu(tape, Node(B * K, D, cls, None, need_grad=False), "layout_adapter")
            ref = Node(B * K, D, torch.empty((B * K, D), dtype=torch.float32, device=dev),
                       torch.empty((2, B * K, D), dtype=torch.bfloat16, device=dev))
            ops.rows_affine(ref0.f32, B * K, D, scale=math.sqrt(D), table=self.pe, tab_mod=K, out_f32=ref.f32, out_split=ref.s)

            def ref_bwd() -> None:  # d(ref0) = sqrt(d) * d(ref)
                if ref.grad is None:
                    return
                g = torch.empty_like(ref0.f32)
                ag.check(ag._L().ralf_rows_gather(ref.grad.data_ptr(), ref.grad.stride(0), B * K, D, math.sqrt(D), 0, 0, 0,
                                                  g.data_ptr(), 0, ag._stream()), "ralf_rows_gather")
                ag.accumulate(ref0, g)

            tape.record(ref_bwd)
            ag.dropout_inplace(tape, ref)  # pos_emb_1d's dropout (positional_encoding.py:107)
            # ---- fusion attention + head over cat[img, ca, ref] ----
            hq = ag.layernorm(tape, ps, x, "attn.norm")
            q = ag.linear(tape, ps, hq, "attn.to_q")
            kv = ag.linear(tape, ps, ref, "attn.to_kv")
            a = ag.cross_attention(tape, q, kv, 0, 512, B, T, K, 8, 64, use_dropout=False)
            ca = ag.linear(tape, ps, a, "attn.to_out.0", "attn.to_out.0.bias")
            Tcat = 2 * T + K
            cat = Node(B * Tcat, D, torch.empty((B * Tcat, D), dtype=torch.float32, device=dev), None)
            ag.place_rows(tape, x, cat.f32, cat, T, Tcat, 0)
            ag.place_rows(tape, ca, cat.f32, cat, T, Tcat, T)
            ag.place_rows(tape, ref, cat.f32, cat, K, Tcat, 2 * T)
            mem_img = self._ffn_gelu(tape, cat, "head")
        else:  # Autoreg baseline: the image tokens are the memory's first part as they are
            mem_img, Tcat = x, T
        # ---- constraint encoder ----
        sc = inputs["seq_layout_const"].to(dev).contiguous()
        Tc = sc.shape[1]
        uc = ag.embed(tape, ps, sc, Tc, "user_const_encoder.emb.weight", math.sqrt(D), self.pe)
        ag.dropout_inplace(tape, uc)
        m = inputs["seq_layout_const_pad_mask"].to(dev).to(torch.uint8).contiguous()
        for i in range(NLAYER):
            uc = self._enc_layer(tape, uc, f"user_const_encoder.encoder.layers.{i}", B, Tc, mask=m)
        # ---- memory = cat[mem_img + task_emb[0], uc + task_emb[1]] ----
        Mlen = Tcat + Tc
        mem = Node(B * Mlen, D, torch.empty((B * Mlen, D), dtype=torch.float32, device=dev),
                   torch.empty((2, B * Mlen, D), dtype=torch.bfloat16, device=dev))
        te = ps.p("task_emb.weight").view(-1)
        tv0 = te[0:1].expand(D).contiguous().view(1, D)
        tv1 = te[1:2].expand(D).contiguous().view(1, D)
        ops.rows_affine(mem_img.f32, B * Tcat, D, table=tv0, tab_mod=1, rows_per_group=Tcat, group_stride=Mlen,
                        group_offset=0, out_f32=mem.f32, out_split=mem.s)
        ops.rows_affine(uc.f32, B * Tc, D, table=tv1, tab_mod=1, rows_per_group=Tc, group_stride=Mlen,
                        group_offset=Tcat, out_f32=mem.f32, out_split=mem.s)

        def mem_bwd() -> None:
            if mem.grad is None:
                return
            for src, rpg, go, row in ((mem_img, Tcat, 0, 0), (uc, Tc, Tcat, 1)):
                g = torch.empty_like(src.f32)
                ag.check(ag._L().ralf_rows_gather(mem.grad.data_ptr(), mem.grad.stride(0), src.M, D, 1.0, rpg, Mlen, go,
                                                  g.data_ptr(), 0, ag._stream()), "ralf_rows_gather")
                cs = torch.empty(D, dtype=torch.float32, device=dev)
                ag.colsum(g, cs)
                ag.colsum(cs.view(D, 1), ps.g("task_emb.weight").view(-1)[row:row + 1], accumulate=True)
                ag.accumulate(src, g)

        tape.record(mem_bwd)
        # ---- decoder (teacher forced, causal + key padding) ----
        seq = inputs["seq"].to(dev).contiguous()
        S = seq.shape[1]
        pm = inputs["tgt_key_padding_mask"].to(dev).to(torch.uint8).contiguous()
        y = ag.embed(tape, ps, seq, S, "decoder.emb.weight", math.sqrt(D), self.pe)
        ag.dropout_inplace(tape, y)
        for i in range(NLAYER):
            p = f"decoder.transformer.layers.{i}"
            hh = ag.layernorm(tape, ps, y, p + ".norm1")
            qkv = ag.linear(tape, ps, hh, p + ".qkv", p + ".self_attn.in_proj_bias")
            a = ag.self_attention(tape, qkv, B, S, NHEAD, 32, mask=pm, causal=True)
            y = ag.linear_res(tape, ps, a, p + ".o", p + ".self_attn.out_proj.bias", y)
            hh = ag.layernorm(tape, ps, y, p + ".norm2")
            qn = ag.linear(tape, ps, hh, p + ".cq", None)
            kvn = ag.linear(tape, ps, mem, p + ".ckv", None)
            # in_proj_bias is one parameter [768]: q part and k/v part are added by bias-only epilogues below
            self._add_bias_slice(tape, qn, p + ".multihead_attn.in_proj_bias", 0, D)
            self._add_bias_slice(tape, kvn, p + ".multihead_attn.in_proj_bias", D, 2 * D)
            a = ag.cross_attention(tape, qn, kvn, 0, D, B, S, Mlen, NHEAD, 32)
            y = ag.linear_res(tape, ps, a, p + ".co", p + ".multihead_attn.out_proj.bias", y)
            hh = ag.layernorm(tape, ps, y, p + ".norm3")
            f = ag.linear(tape, ps, hh, p + ".l1", p + ".linear1.bias", act="relu", want_f32=False)
            ag.dropout_inplace(tape, f)
            y = ag.linear_res(tape, ps, f, p + ".l2", p + ".linear2.bias", y)
        hh = ag.layernorm(tape, ps, y, "decoder.head.0")
        logits = ag.linear(tape, ps, hh, "decoder.head.1", None)
        loss = ag.ce_loss(tape, logits, targets["seq"].to(dev), 0.1, self.model.tokenizer.name_to_id("pad"))
        return loss, tape, logits

    # ------------------------------------------------------------------------------------------
    # drop-in loss for the reference's own loop (train.py:440-454): loss.backward() fills p.grad
    # ------------------------------------------------------------------------------------------
    def loss_with_grad(self, inputs: dict, targets: dict):
        """Forward through the tape and return (loss, logits [B, S, V]) where ``loss`` is a scalar tensor attached to
        torch.autograd by ONE node: its backward replays the tape (our kernels), all-reduces in data-parallel runs and
        points every trainable parameter's ``.grad`` at its slice of the flat gradient buffer, so that the reference's
        ``loss.backward(); clip_grad_norm_(model.parameters(), c); optimizer.step()`` works unchanged.  The parameters
        are views of the flat master buffer, so an external optimiser updates the master weights in place; the GEMM
        operands are refreshed from them at the start of the next call."""
        self.refresh_operands()
        self.ps.flat_g.zero_()
        self.step_count += 1
        self._set_step_seed()
        loss, tape, logits = self.forward_loss(inputs, targets)
        anchor = next(p for _, p in self._named)
        B = inputs["image"].shape[0]
        return _TapeLoss.apply(anchor, loss, tape, self), logits.f32.view(B, -1, logits.Cn)

    def _publish_grads(self, scale: Optional[torch.Tensor]) -> None:
        ps = self.ps
        self._finish_allreduce()
        if scale is not None:
            ps.flat_g.mul_(scale)
        for n, p in self._named:
            p.grad = ps.g(n)

    def _add_bias_slice(self, tape, node: Node, bias_name: str, off: int, n: int) -> None:
        """node.f32 += bias[off:off+n] (row broadcast) with the bias gradient routed to that slice."""
        ps = self.ps
        b = ps.p(bias_name)[off:off + n]
        ops.rows_affine(node.f32, node.M, node.Cn, table=b.view(1, n), tab_mod=1, out_f32=node.f32)

        def bwd() -> None:
            if node.grad is not None:
                ag.colsum(node.grad, ps.g(bias_name)[off:off + n])

        tape.record(bwd)

    def _fid_cls(self, retrieved, B: int) -> torch.Tensor:
        """Frozen FIDNetV3 CLS features [B*K, 256] (fid/model.py:95-103) via the inference kernels."""
        eng = self.infer
        K = eng.top_k
        packed = retrieved if torch.is_tensor(retrieved) else eng.pack_retrieved(retrieved, K, self.dev)
        E = packed.shape[-1]
        N, T = B * K, E + 1
        f = "layout_encoer"
        rows, pad = ops.fid_embed_packed(packed.view(N, 6, E), eng.w[f + ".fc_bbox.w"], eng.w[f + ".fc_bbox.b"],
                                         eng.w[f + ".emb_label"])
        x = torch.empty((N * T, D), dtype=torch.float32, device=self.dev)
        xs = torch.empty((2, N * T, D), dtype=torch.bfloat16, device=self.dev)
        ops.rows_affine(None, N, D, table=eng.w[f + ".token"], tab_mod=1, rows_per_group=1, group_stride=T, group_offset=0,
                        out_f32=x, out_split=xs)
        eng._gemm(rows, f + ".enc_fc_in", act="relu", out_f32=x, out_split=xs, rows_per_group=E, group_stride=T,
                  group_offset=1)
        for i in range(4):
            x, xs = eng._postnorm_layer(x, xs, f"{f}.enc_transformer.core.layers.{i}", N, T, pad, 4)
        cls = torch.empty((N, D), dtype=torch.float32, device=self.dev)
        ops.rows_affine(x, N, D, in_ld=T * D, out_f32=cls)
        return cls

    # ------------------------------------------------------------------------------------------
    def _set_step_scalars(self, lr: Optional[float]) -> None:
        """Per-step scalars of the optimiser go through device memory (a captured step must not bake them in)."""
        self.step_count += 1
        t = self.step_count
        self._wait_slot(t % 64)
        slot = self._dyn_host[t % 64]
        slot[0] = 1.0 if lr is None else lr / self.lr  # scheduler (MultiStepLR) scales every group alike
        slot[1] = 1.0 - 0.9 ** t
        slot[2] = 1.0 - 0.999 ** t
        self._dyn.copy_(slot, non_blocking=True)
        self._set_step_seed()

    def _wait_slot(self, i: int) -> None:
        """The host may run more than 64 steps ahead of the device (train_step_graph never syncs): before a pinned ring
        slot is rewritten, wait until the copies that read it 64 steps ago have executed."""
        ev = self._slot_done[i]
        if ev is not None:
            ev.synchronize()

    def _set_step_seed(self) -> None:
        """Dropout seed of step ``step_count`` (SplitMix64 of base seed + step) -> device memory (graph-replay safe)."""
        z = (self.seed + 0x9E3779B97F4A7C15 * self.step_count) & (2 ** 64 - 1)
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & (2 ** 64 - 1)
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & (2 ** 64 - 1)
        z ^= z >> 31
        i = self.step_count % 64
        self._wait_slot(i)
        slot = self._seed_host[i:i + 1]
        slot[0] = z - 2 ** 64 if z >= 2 ** 63 else z
        self._seed_dev.copy_(slot, non_blocking=True)
        if self._seed_dev.is_cuda:
            ev = torch.cuda.Event()
            ev.record()
            self._slot_done[i] = ev

    def _step_body(self, inputs: dict, targets: dict) -> torch.Tensor:
        ps = self.ps
        ps.flat_g.zero_()
        loss, tape, _ = self.forward_loss(inputs, targets)
        tape.backward()
        self._finish_allreduce()
        norm = ag.grad_norm(ps.flat_g)
        ag.adamw_step(ps, self.group_cfg, 0, self.max_norm, norm, dyn=self._dyn)
        self.refresh_operands()
        self.last_grad_norm = norm
        return loss

    def _weights_changed(self) -> None:
        """The master weights were updated in place: the model's cached inference Engine (split-bf16 GEMM operands and
        folded BatchNorm made from the OLD weights, next to live views of the new biases / norms) is stale -- drop it, so
        the next evaluate() / sample() / forward() in eval mode re-prepares it (generator.engine())."""
        self.model._engine = None

    def train_step(self, inputs: dict, targets: dict, lr: Optional[float] = None) -> torch.Tensor:
        """One optimisation step (train.py:440-454), eager launches.  Returns the loss (device scalar; no host sync)."""
        self._set_step_scalars(lr)
        loss = self._step_body(inputs, targets)
        self._weights_changed()
        return loss

    # ------------------------------------------------------------------------------------------
    # CUDA-graph replay of the whole step (fixed batch shape): ~1.9 k kernel launches become one graph launch
    # ------------------------------------------------------------------------------------------
    def capture(self, inputs: dict, targets: dict) -> None:
        """Capture forward + backward + all-reduce + clip + AdamW for this batch SHAPE.  The example batch is copied into
        static device buffers; `train_step_graph` refills them and replays."""
        dev = self.dev

        def static(v):
            if torch.is_tensor(v):
                return v.to(dev).clone()
            if isinstance(v, dict):
                return {k: static(x) for k, x in v.items()}
            return v  # non-tensor entries (e.g. the reference's retrieved["index"] lists) are not inputs of the kernels

        self._g_in, self._g_tg = static(inputs), static(targets)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm-up (a real step): kernel attributes, tensor-map cache, allocator
            self.train_step(self._g_in, self._g_tg)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._g_loss = self._step_body(self._g_in, self._g_tg)

    def train_step_graph(self, inputs: dict, targets: dict, lr: Optional[float] = None) -> torch.Tensor:
        assert getattr(self, "_graph", None) is not None, "call capture(inputs, targets) first"

        def same_shape(dst, src) -> bool:
            if torch.is_tensor(dst):
                return torch.is_tensor(src) and tuple(src.shape) == tuple(dst.shape)
            if isinstance(dst, dict):
                return isinstance(src, dict) and all(k in src and same_shape(dst[k], src[k]) for k in dst)
            return True

        def fill(dst, src):
            if torch.is_tensor(dst):
                dst.copy_(src, non_blocking=True)
            elif isinstance(dst, dict):
                for k in dst:
                    fill(dst[k], src[k])

        if not (same_shape(self._g_in, inputs) and same_shape(self._g_tg, targets)):
            # a captured graph bakes the batch shape in (seq_layout_const varies per batch for the constrained tasks and
            # under use_multitask): such a batch takes the eager step -- same kernels, same result, only the launches differ
            return self.train_step(inputs, targets, lr)
        fill(self._g_in, inputs)
        fill(self._g_tg, targets)
        self._set_step_scalars(lr)
        self._graph.replay()
        self._weights_changed()
        return self._g_loss
