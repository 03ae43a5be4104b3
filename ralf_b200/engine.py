"""B200 execution engine for the RALF / Autoreg forward and greedy-generation path.

Takes a state dict with the REFERENCE's key names (strict-load contract) and runs
the whole path as hand-written sm_100a kernels behind the C ABI (include/ralf_b200.h):

  image --stem im2col / tcgen05 GEMM (BN folded) / max-pool / 16 bottlenecks / FPN-->  [B*hw, 256]
        --6 pre-LN encoder layers-->  memory_img
  retrieved layouts [B,16,E] --FIDNetV3 (4 post-LN layers, batched over B*16) / adapter / PE--> ref [B,16,256]
  fusion attention(memory_img, ref), head FFN over cat[memory_img, memory_ca, ref], constraint encoder,
  concat  -->  memory [B, M, 256];  decoder cross K/V for all 6 layers in ONE GEMM (N = 3072)
  greedy decode with self-attention KV cache + precomputed cross K/V (the reference recomputes the full
  prefix and the memory projections at every step: retrieval_augmented_autoreg.py:271-297).

Torch is used for device memory, streams and CUDA graphs only; every arithmetic kernel on the path is ours.
Reference citations are relative to image2layout/train/.
"""
from __future__ import annotations

import math
import os
from typing import Optional

import torch

from . import ops

NHEAD = 8
NLAYER = 6
D = 256
# LayerNorm folded into the consuming decode GEMM (ralf_gemm_ln) is OFF by default: measured on B200 the in-kernel row
# normalisation costs more than the separate 5 us LayerNorm launch it removes (round 1, B <= 128: 1698 vs 2162 layouts/s;
# round 2, generalised to the bench batch of 1024: 188.2 vs 168.5 ms per step).
# RALF_FUSE_LN=1 switches it on for A/B runs.
FUSE_LN = os.environ.get("RALF_FUSE_LN", "0") != "0"
# Decoder cross-attention K/V cache of the greedy loop in the 24-bit format (3 bytes per value); RALF_KV24=0 keeps fp32.
KV24 = os.environ.get("RALF_KV24", "1") != "0"
# Format of that cache: 16 (default) = 16-bit integers with one fp32 scale per (memory token, head), 2.125 bytes per value;
# 24 = the round-1 24-bit float format (3 bytes).  The decode loop is bound by this stream: 1088 instead of 1536 bytes per
# memory token and layer.  Accuracy on the reference goldens (profiles/r2_precision_study.json): step logits within 8e-5 of
# scale (24-bit: 2.5e-5; bf16 would be 5e-3, over the 1e-3 bar), token ids identical.
KVFMT = 24 if os.environ.get("RALF_KVFMT", "16") == "24" else 16
# Fused decode-step chains (ralf_decode_chain, csrc/decode_chain.cu): the row-local ops of a decoder-layer step in 3 kernels
# (LN1+in_proj | out_proj+LN2+cross-q | out_proj+LN3+FFN [+ next layer's LN1+in_proj / final LN + LM head]) instead of 9.
# Bit-identical to the per-op launches (RALF_CHAIN_ACC=1) but measured SLOWER on B200 (191.8 vs 184.1 ms per 1024-canvas
# step, profiles/r2_decode_chain.md): a CTA that owns 16 canvases must stream ALL of the layer's weights through its own
# tensor core, and an SS-mode tcgen05.mma at N = 16 costs ~92 cycles whatever N is (operand fetch of the 128-row weight
# tile), so the chain is MMA-issue bound at ~0.6 us per 32 KB weight tile.  Opt-in (RALF_DECODE_CHAIN=1) for A/B runs.
DECODE_CHAIN = os.environ.get("RALF_DECODE_CHAIN", "0") != "0"
# 3x3 convolutions as implicit GEMMs (ralf_conv_gemm_strided); RALF_IMPLICIT_CONV=0 restores im2col + GEMM for A/B runs,
# RALF_STRIDED_CONV=0 only for the stride-2 3x3 / 1x1 convolutions (TMA boxes with element strides).
IMPLICIT_CONV = os.environ.get("RALF_IMPLICIT_CONV", "1") != "0"
STRIDED_CONV = os.environ.get("RALF_STRIDED_CONV", "1") != "0"
# Decode loop: LayerNorm in the epilogue of the residual GEMM in front of it (thread-block cluster of the row's eight
# n-tiles, statistics through distributed shared memory).  Built and measured in round 2 (profiles/resln_bench.py, graph
# replay of a dependent chain at M = 1024): GEMM 6.35 us, GEMM + LayerNorm launch 8.44 us, fused cluster kernel 8.31 us --
# the two cluster barriers + the guarded exit cost what the 2.1 us LayerNorm launch costs; step 147.9-148.5 vs 147.0-148.4
# ms with 1080 launches fewer.  Opt-in (RALF_DECODE_RESLN=1); tokens stay bit-exact on every golden with it on.
DECODE_RESLN = os.environ.get("RALF_DECODE_RESLN", "0") != "0"


def _sine_pe_1d(max_len: int, d_model: int) -> torch.Tensor:
    """models/common/positional_encoding.py:71-81."""
    position = torch.arange(max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(max_len, d_model)
    pe[:, 0::2] = torch.sin(position * div_term)
    pe[:, 1::2] = torch.cos(position * div_term)
    return pe


def _pos_emb_2d(h: int, w: int, d_model: int) -> torch.Tensor:
    """PositionEmbeddingSine(normalize=True) as a [h*w, d] table (positional_encoding.py:182-210)."""
    half = d_model // 2
    y, x = torch.meshgrid(torch.arange(h).float(), torch.arange(w).float(), indexing="ij")
    y = y / (h - 1) * (2 * math.pi)
    x = x / (w - 1) * (2 * math.pi)
    dim_t = torch.arange(half).float()
    dim_t = 10000 ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / half)
    px = x.flatten()[:, None] / dim_t
    py = y.flatten()[:, None] / dim_t
    px = torch.stack((px[:, 0::2].sin(), px[:, 1::2].cos()), dim=2).flatten(1)
    py = torch.stack((py[:, 0::2].sin(), py[:, 1::2].cos()), dim=2).flatten(1)
    return torch.cat((py, px), dim=1).contiguous()


class Engine:
    """Prepared weights + kernel orchestration.  ``is_ralf=False`` gives the Autoreg baseline (no retrieval)."""

    def __init__(self, state_dict: dict, device: torch.device, *, is_ralf: bool = True, top_k: int = 16,
                 npass: int = 3) -> None:
        self.dev = device
        self.is_ralf = is_ralf
        self.top_k = top_k
        self.npass = npass
        self.w: dict[str, torch.Tensor] = {}
        self._pos2d: dict[tuple[int, int], torch.Tensor] = {}
        self._prepare({k: v.detach() for k, v in state_dict.items()})

    # ------------------------------------------------------------------------------------------
    # weight preparation (once per load_state_dict)
    # ------------------------------------------------------------------------------------------
    def _put(self, name: str, t: torch.Tensor) -> None:
        self.w[name] = t.to(self.dev, torch.float32).contiguous()

    def _put_w(self, name: str, w2d: torch.Tensor) -> None:
        """Linear/conv weight [N, K] -> split bf16 [2, N, K] (K padded to a multiple of 8)."""
        w2d = w2d.to(self.dev, torch.float32)
        if w2d.shape[1] % 8:
            w2d = torch.nn.functional.pad(w2d, (0, 8 - w2d.shape[1] % 8))
        self.w[name] = ops.split_bf16(w2d.contiguous())

    def _conv_bn(self, sd, conv: str, bn: Optional[str], name: str, bias_key: Optional[str] = None) -> None:
        w = sd[conv + ".weight"].float()
        cout = w.shape[0]
        if bn is not None:  # eval-mode BatchNorm folded into the convolution
            scale = sd[bn + ".weight"].float() / torch.sqrt(sd[bn + ".running_var"].float() + 1e-5)
            bias = sd[bn + ".bias"].float() - sd[bn + ".running_mean"].float() * scale
            w = w * scale[:, None, None, None]
        else:
            bias = sd[bias_key].float()
        self._put_w(name + ".w", w.permute(0, 2, 3, 1).reshape(cout, -1))  # k = (kh*KW + kw)*Cin + c
        self._put(name + ".b", bias)

    def _stem_s2d_weight(self, sd, conv: str, bn: str) -> None:
        """7x7 / stride 2 stem weight (BN folded) in space-to-depth form [64, 4(kh') * 4(kw') * 2(dy) * 2(dx) * 4(c)]:
        original tap kh -> (kh', dy) = ((kh+1)//2, (kh+1)%2) (input row 2*ho - 3 + kh = 2*(ho + kh' - 2) + dy)."""
        w = sd[conv + ".weight"].float()
        scale = sd[bn + ".weight"].float() / torch.sqrt(sd[bn + ".running_var"].float() + 1e-5)
        w = w * scale[:, None, None, None]
        w2 = torch.zeros(w.shape[0], 4, 4, 2, 2, 4)
        for kh in range(7):
            for kw in range(7):
                w2[:, (kh + 1) // 2, (kw + 1) // 2, (kh + 1) % 2, (kw + 1) % 2, :] = w[:, :, kh, kw]
        self._put_w("stem_s2d.w", w2.reshape(w.shape[0], 256))
        self.w["stem_s2d.b"] = self.w["stem.b"]

    def _lin(self, sd, key: str, name: Optional[str] = None, wscale: float = 1.0, badd: float = 0.0) -> None:
        name = name or key
        self._put_w(name + ".w", sd[key + ".weight"].float() * wscale)
        if (key + ".bias") in sd:
            self._put(name + ".b", sd[key + ".bias"].float() * wscale + badd)

    def _norm(self, sd, key: str) -> None:
        self._put(key + ".g", sd[key + ".weight"])
        self._put(key + ".beta", sd[key + ".bias"])

    def _enc_layer(self, sd, p: str, badd_last: float = 0.0) -> None:
        self._put_w(p + ".qkv.w", sd[p + ".self_attn.in_proj_weight"])
        self._put(p + ".qkv.b", sd[p + ".self_attn.in_proj_bias"])
        self._lin(sd, p + ".self_attn.out_proj", p + ".o")
        self._lin(sd, p + ".linear1")
        self._lin(sd, p + ".linear2", badd=badd_last)
        self._norm(sd, p + ".norm1")
        self._norm(sd, p + ".norm2")

    def _prepare(self, sd: dict) -> None:
        b = "encoder.extractor.body"
        self._conv_bn(sd, b + ".conv1", b + ".bn1", "stem")
        self._stem_s2d_weight(sd, b + ".conv1", b + ".bn1")
        self.blocks = []
        for li, (nblk, stride, planes) in enumerate([(3, 1, 64), (4, 2, 128), (6, 2, 256), (3, 2, 512)], start=1):
            for bi in range(nblk):
                p = f"{b}.layer{li}.{bi}"
                for c in (1, 2, 3):
                    self._conv_bn(sd, f"{p}.conv{c}", f"{p}.bn{c}", f"{p}.c{c}")
                has_ds = (p + ".downsample.0.weight") in sd
                if has_ds:
                    self._conv_bn(sd, p + ".downsample.0", p + ".downsample.1", p + ".ds")
                self.blocks.append((p, li, stride if bi == 0 else 1, planes, has_ds))
        e = "encoder.extractor"
        for n in ("fpn_conv11_4", "fpn_conv11_5", "fpn_conv33", "proj"):
            self._conv_bn(sd, f"{e}.{n}", None, f"{e}.{n}", bias_key=f"{e}.{n}.bias")
        for i in range(NLAYER):
            self._enc_layer(sd, f"transformer_encoder.layers.{i}")
        t0 = float(sd["task_emb.weight"][int(sd["flag_img"][0]), 0])
        t1 = float(sd["task_emb.weight"][int(sd["flag_user_const"][0]), 0])
        self.t_img = t0
        for i in range(NLAYER):  # flag embedding of the constraint branch folded into its last bias
            self._enc_layer(sd, f"user_const_encoder.encoder.layers.{i}", badd_last=t1 if i == NLAYER - 1 else 0.0)
        self._put("user_const_encoder.emb", sd["user_const_encoder.emb.weight"])
        self._put("pe1d", _sine_pe_1d(5000, D))
        if self.is_ralf:
            f = "layout_encoer"
            self._put(f + ".emb_label", sd[f + ".emb_label.weight"])
            self._put(f + ".fc_bbox.w", sd[f + ".fc_bbox.weight"])
            self._put(f + ".fc_bbox.b", sd[f + ".fc_bbox.bias"])
            self._lin(sd, f + ".enc_fc_in")
            self._put(f + ".token", sd[f + ".enc_transformer.token"].reshape(1, D))
            for i in range(4):
                self._enc_layer(sd, f"{f}.enc_transformer.core.layers.{i}")
            self._norm(sd, "layout_adapter.net.0")
            self._lin(sd, "layout_adapter.net.1")
            # pos_emb_1d multiplies by sqrt(d) = 16 (an exact power of two): fold into the last linear
            self._lin(sd, "layout_adapter.net.4", wscale=math.sqrt(D))
            self._norm(sd, "attn.norm")
            self._lin(sd, "attn.to_q")
            self._lin(sd, "attn.to_kv")
            self._lin(sd, "attn.to_out.0")
            self._norm(sd, "head.net.0")
            self._lin(sd, "head.net.1")
            self._lin(sd, "head.net.4", badd=t0)  # + task_emb(flag_img)
        # decoder
        d = "decoder.transformer.layers"
        wkv, bkv = [], []
        for i in range(NLAYER):
            p = f"{d}.{i}"
            self._put_w(p + ".qkv.w", sd[p + ".self_attn.in_proj_weight"])
            self._put(p + ".qkv.b", sd[p + ".self_attn.in_proj_bias"])
            self._lin(sd, p + ".self_attn.out_proj", p + ".o")
            cw, cb = sd[p + ".multihead_attn.in_proj_weight"].float(), sd[p + ".multihead_attn.in_proj_bias"].float()
            self._put_w(p + ".cq.w", cw[:D])
            self._put(p + ".cq.b", cb[:D])
            wkv.append(cw[D:])
            bkv.append(cb[D:])
            self._lin(sd, p + ".multihead_attn.out_proj", p + ".co")
            self._lin(sd, p + ".linear1")
            self._lin(sd, p + ".linear2")
            for n in ("norm1", "norm2", "norm3"):
                self._norm(sd, f"{p}.{n}")
        self._put_w("decoder.ckv.w", torch.cat(wkv, 0))  # [6*512, 256]: layer l -> K cols l*512.., V cols l*512+256..
        self._put("decoder.ckv.b", torch.cat(bkv, 0))
        self._put("decoder.emb", sd["decoder.emb.weight"])
        self._norm(sd, "decoder.head.0")
        self._put_w("decoder.head.1.w", sd["decoder.head.1.weight"])
        self.vocab = sd["decoder.emb.weight"].shape[0]

    # ------------------------------------------------------------------------------------------
    # kernel helpers
    # ------------------------------------------------------------------------------------------
    def _gemm(self, a, name, **kw):
        kw.setdefault("npass", self.npass)
        return ops.gemm(a, self.w[name + ".w"], bias=self.w.get(name + ".b"), **kw)

    def _ln(self, x, name, **kw):
        return ops.layernorm(x, self.w[name + ".g"], self.w[name + ".beta"], **kw)

    def _gemm_ln(self, x, ln_name, name, **kw):
        """LayerNorm fused into the consuming GEMM (M <= 128 rows: the decode path); falls back to two kernels above."""
        if self.npass != 3 or not FUSE_LN:
            _, h = self._ln(x, ln_name)
            return self._gemm(h, name, **kw)
        return ops.gemm_ln(x, self.w[ln_name + ".g"], self.w[ln_name + ".beta"], self.w[name + ".w"],
                           bias=self.w.get(name + ".b"), **kw)

    def pos2d(self, h: int, w: int) -> torch.Tensor:
        if (h, w) not in self._pos2d:
            self._pos2d[(h, w)] = _pos_emb_2d(h, w, D).to(self.dev)
        return self._pos2d[(h, w)]

    # ------------------------------------------------------------------------------------------
    # image branch: ResNet50 + FPN (common/image.py:90-120), NHWC split activations
    # ------------------------------------------------------------------------------------------
    def resnet_fpn(self, img: torch.Tensor) -> tuple[torch.Tensor, int, int]:
        """img fp32 [B,4,H,W] -> tokens fp32 [B*h*w, 256] with the 2-D sine PE already added."""
        B = img.shape[0]
        if IMPLICIT_CONV and img.shape[2] % 2 == 0 and img.shape[3] % 2 == 0 and img.shape[3] <= 256:
            a, H, W = ops.stem_s2d(img.contiguous())  # space-to-depth + implicit GEMM: no im2col rows
            _, x = self._gemm(a, "stem_s2d", act="relu", want_f32=False, want_split=True, stem=(B, H, W))
        else:
            a, H, W = ops.stem_im2col(img.contiguous())
            _, x = self._gemm(a, "stem", act="relu", want_f32=False, want_split=True)
        del a
        x, H, W = ops.maxpool3x3s2(x, B, H, W, 64)
        C = 64
        feats = {}
        for (p, li, stride, planes, has_ds) in self.blocks:
            _, t1 = self._gemm(x, p + ".c1", act="relu", want_f32=False, want_split=True)
            strided_ok = STRIDED_CONV and H % 2 == 0 and W % 2 == 0
            if IMPLICIT_CONV and ((stride == 1 and W <= 128) or (stride == 2 and strided_ok and W <= 256)):
                # implicit GEMM: the 3x3 taps are TMA boxes (element strides 2 for the stride-2 blocks), no im2col
                Ho, Wo = H // stride, W // stride
                _, t2 = self._gemm(t1, p + ".c2", act="relu", want_f32=False, want_split=True,
                                   conv=(B, H, W, planes, 3, 3, stride))
            else:
                a2, Ho, Wo = ops.im2col(t1, B, H, W, planes, 3, 3, stride, 1)
                _, t2 = self._gemm(a2, p + ".c2", act="relu", want_f32=False, want_split=True)
                del a2
            del t1
            if has_ds:
                if stride == 2 and IMPLICIT_CONV and strided_ok and W <= 256:
                    _, idt = self._gemm(x, p + ".ds", want_f32=False, want_split=True, conv=(B, H, W, C, 1, 1, 2))
                else:
                    xs = x if stride == 1 else ops.im2col(x, B, H, W, C, 1, 1, stride, 0)[0]
                    _, idt = self._gemm(xs, p + ".ds", want_f32=False, want_split=True)
            else:
                idt = x
            _, x = self._gemm(t2, p + ".c3", res_split=idt, post_relu=True, want_f32=False, want_split=True)
            H, W, C = Ho, Wo, planes * 4
            feats[li] = (x, H, W)
        e = "encoder.extractor"
        l3, h4, w4 = feats[3]
        l4, h5, w5 = feats[4]
        c4, _ = self._gemm(l3, e + ".fpn_conv11_4")
        c5, _ = self._gemm(l4, e + ".fpn_conv11_5")
        fused, summ = ops.fpn_merge(c5, c4, B, h5, w5, h4, w4, D)
        if IMPLICIT_CONV and w4 <= 128:
            self._gemm(summ, e + ".fpn_conv33", out_split=fused, out_col0=D, want_f32=False, conv=(B, h4, w4, D, 3, 3))
        else:
            a33, _, _ = ops.im2col(summ, B, h4, w4, D, 3, 3, 1, 1)
            self._gemm(a33, e + ".fpn_conv33", out_split=fused, out_col0=D, want_f32=False)
        tokens, _ = self._gemm(fused, e + ".proj", res=self.pos2d(h4, w4), res_row_mod=h4 * w4)
        return tokens, h4, w4

    # ------------------------------------------------------------------------------------------
    # transformer layers
    # ------------------------------------------------------------------------------------------
    def _prenorm_layer(self, x, p, B, T, mask=None, nhead=NHEAD, out=None, out_map=None, want_split=False):
        """nn.TransformerEncoderLayer(norm_first=True, relu) in eval mode.  x fp32 [B*T, 256]."""
        _, h = self._ln(x, p + ".norm1")
        qkv, _ = self._gemm(h, p + ".qkv")
        a = ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], B, nhead, T, T, D // nhead, mask=mask)
        x1, _ = self._gemm(a, p + ".o", res=x)
        _, h = self._ln(x1, p + ".norm2")
        _, f = self._gemm(h, p + ".linear1", act="relu", want_f32=False, want_split=True)
        kw = {}
        if out_map is not None:
            kw = dict(rows_per_group=out_map[0], group_stride=out_map[1], group_offset=out_map[2])
        if isinstance(out, tuple):
            return self._gemm(f, p + ".linear2", res=x1, out_f32=out[0], out_split=out[1], **kw)
        return self._gemm(f, p + ".linear2", res=x1, out_f32=out, want_split=want_split, **kw)

    def _postnorm_layer(self, x, xs, p, B, T, mask, nhead):
        """nn.TransformerEncoderLayer default (post-norm) -- FIDNetV3 (fid/model.py:26-33).
        x fp32 [B*T,256] and its split copy xs."""
        qkv, _ = self._gemm(xs, p + ".qkv")
        a = ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], B, nhead, T, T, D // nhead, mask=mask)
        y, _ = self._gemm(a, p + ".o", res=x)
        x1, x1s = self._ln(y, p + ".norm1", want_f32=True)
        _, f = self._gemm(x1s, p + ".linear1", act="relu", want_f32=False, want_split=True)
        y2, _ = self._gemm(f, p + ".linear2", res=x1)
        return self._ln(y2, p + ".norm2", want_f32=True)

    def encode_image(self, img: torch.Tensor):
        tokens, h, w = self.resnet_fpn(img)
        B, T = img.shape[0], h * w
        x = tokens
        for i in range(NLAYER):
            x, _ = self._prenorm_layer(x, f"transformer_encoder.layers.{i}", B, T)
        return x, T

    # ------------------------------------------------------------------------------------------
    # retrieved-layout branch (retrieval_augmented_autoreg.py:526-584; fid/model.py:95-103)
    # ------------------------------------------------------------------------------------------
    @staticmethod
    def pack_retrieved(retrieved: dict, K: int, dev) -> torch.Tensor:
        """retrieved{label, mask, center_x, center_y, width, height: [B, >=K, E]} -> packed fp32 [B, K, 6, E]
        (API convenience for dict inputs; GpuRetriever.fetch produces the packed form directly on the GPU)."""
        if "packed" in retrieved:
            return retrieved["packed"][:, :K].to(dev, torch.float32).contiguous()
        keys = ["label", "mask", "center_x", "center_y", "width", "height"]
        return torch.stack([retrieved[k][:, :K].to(dev, torch.float32) for k in keys], dim=2).contiguous()

    def retrieved_features(self, retrieved, B: int):
        """-> ref_layouts fp32 [B*K, 256] (already x16 + PE) and its split copy."""
        K = self.top_k
        packed = retrieved if torch.is_tensor(retrieved) else self.pack_retrieved(retrieved, K, self.dev)
        E = packed.shape[-1]
        N = B * K
        f = "layout_encoer"
        rows, pad = ops.fid_embed_packed(packed.view(N, 6, E), self.w[f + ".fc_bbox.w"], self.w[f + ".fc_bbox.b"],
                                         self.w[f + ".emb_label"])
        T = E + 1
        x = torch.empty((N * T, D), dtype=torch.float32, device=self.dev)
        xs = torch.empty((2, N * T, D), dtype=torch.bfloat16, device=self.dev)
        # CLS token rows (TransformerWithToken, fid/model.py:42-43) then relu(enc_fc_in(...)) rows
        ops.rows_affine(None, N, D, table=self.w[f + ".token"], tab_mod=1, rows_per_group=1, group_stride=T,
                        group_offset=0, out_f32=x, out_split=xs)
        self._gemm(rows, f + ".enc_fc_in", act="relu", out_f32=x, out_split=xs, rows_per_group=E, group_stride=T,
                   group_offset=1)
        for i in range(4):
            x, xs = self._postnorm_layer(x, xs, f"{f}.enc_transformer.core.layers.{i}", N, T, pad, 4)
        # layout_adapter FeedForward on the CLS rows (row stride T*256), then x16 + PE1d[k]
        _, h = self._ln(x, "layout_adapter.net.0", rows=N, in_ld=T * D)
        _, g = self._gemm(h, "layout_adapter.net.1", act="gelu", want_f32=False, want_split=True)
        return self._gemm(g, "layout_adapter.net.4", res=self.w["pe1d"], res_row_mod=K, want_split=True)

    # ------------------------------------------------------------------------------------------
    # constraint encoder (common/common.py:238-252)
    # ------------------------------------------------------------------------------------------
    def constraint_encoder(self, seq_const: torch.Tensor, pad_mask: torch.Tensor, mem, mem_s, Mlen: int, off: int):
        B, T = seq_const.shape
        x = ops.embed(seq_const.to(self.dev).contiguous(), 0, T, self.w["user_const_encoder.emb"], math.sqrt(D),
                      self.w["pe1d"], 0)
        m = pad_mask.to(self.dev).to(torch.uint8).contiguous()
        for i in range(NLAYER):
            last = i == NLAYER - 1
            x, _ = self._prenorm_layer(x, f"user_const_encoder.encoder.layers.{i}", B, T, mask=m,
                                       out=(mem, mem_s) if last else None, out_map=(T, Mlen, off) if last else None)

    # ------------------------------------------------------------------------------------------
    # memory (retrieval_augmented_autoreg.py:963-994,1004-1033 / autoreg.py:590-622)
    # ------------------------------------------------------------------------------------------
    def encode(self, image: torch.Tensor, retrieved: Optional[dict], seq_const: torch.Tensor,
               seq_const_pad: torch.Tensor):
        """-> (memory fp32 [B, M, 256], split copy [2, B*M, 256])."""
        B = image.shape[0]
        Tc = seq_const.shape[1]
        x_img, T = self.encode_image(image.to(self.dev, torch.float32))
        if self.is_ralf:
            K = self.top_k
            Tcat = 2 * T + K
            Mlen = Tcat + Tc
            ref, ref_s = self.retrieved_features(retrieved, B)
            cat = torch.empty((B * Tcat, D), dtype=torch.float32, device=self.dev)
            ops.rows_affine(x_img, B * T, D, rows_per_group=T, group_stride=Tcat, group_offset=0, out_f32=cat)
            ops.rows_affine(ref, B * K, D, rows_per_group=K, group_stride=Tcat, group_offset=2 * T, out_f32=cat)
            # fusion Attention (common/attention.py:49-71): LN on x only, 8 heads x 64, no residual
            _, h = self._ln(x_img, "attn.norm")
            q, _ = self._gemm(h, "attn.to_q")
            kv, _ = self._gemm(ref_s, "attn.to_kv")
            a = ops.attention(q, kv[:, :512], kv[:, 512:], B, 8, T, K, 64)
            self._gemm(a, "attn.to_out.0", out_f32=cat, rows_per_group=T, group_stride=Tcat, group_offset=T)
            # head FeedForward over the concatenation, written straight into memory rows [0, Tcat)
            mem = torch.empty((B * Mlen, D), dtype=torch.float32, device=self.dev)
            mem_s = torch.empty((2, B * Mlen, D), dtype=torch.bfloat16, device=self.dev)
            _, h = self._ln(cat, "head.net.0")
            _, g = self._gemm(h, "head.net.1", act="gelu", want_f32=False, want_split=True)
            self._gemm(g, "head.net.4", out_f32=mem, out_split=mem_s, rows_per_group=Tcat, group_stride=Mlen,
                       group_offset=0)
            off = Tcat
        else:
            Mlen = T + Tc
            mem = torch.empty((B * Mlen, D), dtype=torch.float32, device=self.dev)
            mem_s = torch.empty((2, B * Mlen, D), dtype=torch.bfloat16, device=self.dev)
            ops.rows_affine(x_img, B * T, D, add=self.t_img, rows_per_group=T, group_stride=Mlen, group_offset=0,
                            out_f32=mem, out_split=mem_s)
            off = T
        self.constraint_encoder(seq_const, seq_const_pad, mem, mem_s, Mlen, off)
        return mem.view(B, Mlen, D), mem_s

    # ------------------------------------------------------------------------------------------
    # decoder
    # ------------------------------------------------------------------------------------------
    def alloc_cross_kv(self, rows: int, kv24: bool = False) -> list:
        """Decoder cross-attention K/V cache, layer-major.  fp32: 6 x [rows, 512] (K = cols 0..255, V = 256..511), so
        one layer's K/V of a canvas is ONE contiguous 2 KB x M stream for the decode kernel.  ``kv24``: 6 x uint8
        [rows, 1536] rows in the 24-bit format of include/ralf_b200.h (3 bytes per value: the greedy decode loop is bound
        by this stream)."""
        if kv24:
            return [torch.empty((rows, ops.KV_ROW_BYTES[KVFMT]), dtype=torch.uint8, device=self.dev) for _ in range(NLAYER)]
        return [torch.empty((rows, 2 * D), dtype=torch.float32, device=self.dev) for _ in range(NLAYER)]

    def cross_kv(self, mem_s: torch.Tensor, out: Optional[list] = None, row0: int = 0, kv24: bool = False) -> list:
        """K/V of the memory rows for all 6 decoder layers (one GEMM per layer, N = 512); written into rows
        [row0, row0 + rows) of ``out`` when given (micro-batched encode)."""
        rows = mem_s.shape[1]
        if out is None:
            out = self.alloc_cross_kv(rows, kv24)
        kv24 = out[0].dtype == torch.uint8
        w, b = self.w["decoder.ckv.w"], self.w["decoder.ckv.b"]
        for i in range(NLAYER):
            wi, bi = w[:, i * 2 * D:(i + 1) * 2 * D], b[i * 2 * D:(i + 1) * 2 * D]
            if kv24:
                ops.gemm(mem_s, wi, bias=bi, npass=self.npass, want_f32=False, out_kv24=out[i][row0:row0 + rows])
            else:
                ops.gemm(mem_s, wi, bias=bi, npass=self.npass, out_f32=out[i][row0:row0 + rows])
        return out

    def decoder_logits(self, seq: torch.Tensor, pad_mask: torch.Tensor, mem_s: torch.Tensor, B: int, Mlen: int):
        """Teacher-forced BaseDecoder.forward (common/common.py:84-135), causal + key padding masks."""
        S = seq.shape[1]
        kvm = self.cross_kv(mem_s)
        x = ops.embed(seq.to(self.dev).contiguous(), 0, S, self.w["decoder.emb"], math.sqrt(D), self.w["pe1d"], 0)
        m = pad_mask.to(self.dev).to(torch.uint8).contiguous()
        for i in range(NLAYER):
            p = f"decoder.transformer.layers.{i}"
            _, h = self._ln(x, p + ".norm1")
            qkv, _ = self._gemm(h, p + ".qkv")
            a = ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], B, NHEAD, S, S, 32, mask=m, causal=True)
            x, _ = self._gemm(a, p + ".o", res=x)
            _, h = self._ln(x, p + ".norm2")
            q, _ = self._gemm(h, p + ".cq")
            a = ops.attention(q, kvm[i][:, :D], kvm[i][:, D:], B, NHEAD, S, Mlen, 32)
            x, _ = self._gemm(a, p + ".co", res=x)
            _, h = self._ln(x, p + ".norm3")
            _, f = self._gemm(h, p + ".linear1", act="relu", want_f32=False, want_split=True)
            x, _ = self._gemm(f, p + ".linear2", res=x)
        _, h = self._ln(x, "decoder.head.0")
        logits, _ = self._gemm(h, "decoder.head.1")
        return logits.view(B, S, self.vocab)

    def _decode_step(self, x: torch.Tensor, t: int, kc: list, vc: list, kvm: list, pad_mask: torch.Tensor, B: int,
                     Mlen: int) -> torch.Tensor:
        """One KV-cached pass of the 6 decoder layers + head over the token embeddings ``x`` [B, 256] at position ``t``
        (updates ``x`` and rows ``t`` of the self-attention caches in place) -> logits [B, V]."""
        kv24 = kvm[0].dtype == torch.uint8
        if DECODE_CHAIN and self.npass == 3:
            return self._decode_step_chained(x, t, kc, vc, kvm, pad_mask, B, Mlen)
        if DECODE_RESLN and self.npass == 3 and not FUSE_LN:
            return self._decode_step_resln(x, t, kc, vc, kvm, pad_mask, B, Mlen)
        for i in range(NLAYER):
            p = f"decoder.transformer.layers.{i}"
            # every LayerNorm of the step is folded into the GEMM that consumes it (ralf_gemm_ln)
            qkv, _ = self._gemm_ln(x, p + ".norm1", p + ".qkv")
            a = ops.attention_decode_append(qkv, kc[i], vc[i], t, B, NHEAD, 32, mask=pad_mask)
            self._gemm(a, p + ".o", res=x, out_f32=x)
            q, _ = self._gemm_ln(x, p + ".norm2", p + ".cq")
            if kv24:
                a = ops.attention_decode_kv24(q, kvm[i], Mlen, Mlen, B, NHEAD)
            else:
                a = ops.attention_decode(q, kvm[i][:, :D], kvm[i][:, D:], Mlen, Mlen, B, NHEAD, 32)
            self._gemm(a, p + ".co", res=x, out_f32=x)
            _, f = self._gemm_ln(x, p + ".norm3", p + ".linear1", act="relu", want_f32=False, want_split=True)
            self._gemm(f, p + ".linear2", res=x, out_f32=x)
        logits, _ = self._gemm_ln(x, "decoder.head.0", "decoder.head.1")
        return logits

    def _decode_step_resln(self, x: torch.Tensor, t: int, kc: list, vc: list, kvm: list, pad_mask: torch.Tensor, B: int,
                           Mlen: int) -> torch.Tensor:
        """``_decode_step`` with every LayerNorm but the first computed in the epilogue of the residual GEMM in front of
        it (ops.gemm_res_ln: out-projections and linear2 write the new residual row AND its normalised split operand):
        8 launches per layer instead of 11."""
        w = self.w
        kv24 = kvm[0].dtype == torch.uint8
        L = "decoder.transformer.layers."

        def res_ln(a, name, ln_name):
            return ops.gemm_res_ln(a, w[name + ".w"], x, w[ln_name + ".g"], w[ln_name + ".beta"], bias=w.get(name + ".b"))[1]

        _, h = self._ln(x, L + "0.norm1")
        for i in range(NLAYER):
            p = L + str(i)
            qkv, _ = self._gemm(h, p + ".qkv")
            a = ops.attention_decode_append(qkv, kc[i], vc[i], t, B, NHEAD, 32, mask=pad_mask)
            h = res_ln(a, p + ".o", p + ".norm2")
            q, _ = self._gemm(h, p + ".cq")
            if kv24:
                a = ops.attention_decode_kv24(q, kvm[i], Mlen, Mlen, B, NHEAD)
            else:
                a = ops.attention_decode(q, kvm[i][:, :D], kvm[i][:, D:], Mlen, Mlen, B, NHEAD, 32)
            h = res_ln(a, p + ".co", p + ".norm3")
            _, f = self._gemm(h, p + ".linear1", act="relu", want_f32=False, want_split=True)
            h = res_ln(f, p + ".linear2", (L + str(i + 1) + ".norm1") if i + 1 < NLAYER else "decoder.head.0")
        logits, _ = self._gemm(h, "decoder.head.1")
        return logits

    def _decode_step_chained(self, x: torch.Tensor, t: int, kc: list, vc: list, kvm: list, pad_mask: torch.Tensor, B: int,
                             Mlen: int) -> torch.Tensor:
        """``_decode_step`` with the row-local ops of every layer fused (ops.decode_chain): per layer
        [LN1 + in_proj] -> self-attention (KV append) -> [out_proj + x, LN2, cross-q] -> cross-attention over the memory
        cache -> [out_proj + x, LN3, linear1, ReLU, linear2 + x, and the NEXT layer's LN1 + in_proj or the final LN + LM
        head]: 4 launches per layer instead of 11, the residual row and the FFN hidden layer never leave the SM."""
        w, dev = self.w, self.dev
        kv24 = kvm[0].dtype == torch.uint8
        L = "decoder.transformer.layers."
        cs = ops.chain_stage

        def ln(name):
            return (w[name + ".g"], w[name + ".beta"])

        qkv = torch.empty((B, 3 * D), dtype=torch.float32, device=dev)
        q = torch.empty((B, D), dtype=torch.float32, device=dev)
        logits = torch.empty((B, self.vocab), dtype=torch.float32, device=dev)
        ops.decode_chain(x, B, [cs(w[L + "0.qkv.w"], bias=w[L + "0.qkv.b"], ln=ln(L + "0.norm1"), out_f32=qkv)])
        for i in range(NLAYER):
            p = L + str(i)
            a = ops.attention_decode_append(qkv, kc[i], vc[i], t, B, NHEAD, 32, mask=pad_mask)
            ops.decode_chain(x, B, [cs(w[p + ".o.w"], bias=w[p + ".o.b"], in_split=a, add_x=True, to_x=True, out_f32=x),
                                    cs(w[p + ".cq.w"], bias=w[p + ".cq.b"], ln=ln(p + ".norm2"), out_f32=q)])
            if kv24:
                a = ops.attention_decode_kv24(q, kvm[i], Mlen, Mlen, B, NHEAD)
            else:
                a = ops.attention_decode(q, kvm[i][:, :D], kvm[i][:, D:], Mlen, Mlen, B, NHEAD, 32)
            stages = [cs(w[p + ".co.w"], bias=w[p + ".co.b"], in_split=a, add_x=True, to_x=True),
                      cs(w[p + ".linear1.w"], bias=w[p + ".linear1.b"], ln=ln(p + ".norm3"), act="relu", out_operand=True),
                      cs(w[p + ".linear2.w"], bias=w[p + ".linear2.b"], add_x=True, to_x=True, out_f32=x)]
            if i + 1 < NLAYER:
                n = L + str(i + 1)
                stages.append(cs(w[n + ".qkv.w"], bias=w[n + ".qkv.b"], ln=ln(n + ".norm1"), out_f32=qkv))
            else:
                stages.append(cs(w["decoder.head.1.w"], ln=ln("decoder.head.0"), out_f32=logits))
            ops.decode_chain(x, B, stages)
        return logits

    def generate(self, mem_s: Optional[torch.Tensor], B: int, Mlen: int, token_mask: torch.Tensor, bos_id: int,
                 pad_id: int, steps: int, return_logits: bool = False, kv: Optional[list] = None, step_hook=None,
                 forced: Optional[torch.Tensor] = None, sampling: Optional[dict] = None,
                 uniform: Optional[torch.Tensor] = None, rng: Optional[torch.Generator] = None):
        """Autoregressive decode (retrieval_augmented_autoreg.py:244-300) with KV caches.
        token_mask: uint8 [steps, V] (tokenizer.token_mask).  Returns seq int64 [B, steps] (BOS dropped).
        ``kv``: precomputed cross-attention cache (cross_kv) of all B canvases; else built from ``mem_s``.
        ``forced``: int32 [B, steps] decoding-space restriction table (ralf_b200.task.forced_token_table), -1 = free.
        ``sampling``: {"name": deterministic|random|top_k|top_p|gumbel, "temperature", "top_k", "top_p"}
        (helpers/sampling.py:18-68); the uniforms come from ``uniform`` fp32 [steps, B] or are drawn with ``rng``."""
        dev = self.dev
        kvm = kv if kv is not None else self.cross_kv(mem_s, kv24=KV24 and self.npass == 3)
        seq = torch.full((B, steps + 1), pad_id, dtype=torch.int64, device=dev)
        seq[:, 0] = bos_id
        pad_mask = torch.zeros((B, steps + 1), dtype=torch.uint8, device=dev)
        kc = [torch.empty((B, steps, D), dtype=torch.float32, device=dev) for _ in range(NLAYER)]
        vc = [torch.empty((B, steps, D), dtype=torch.float32, device=dev) for _ in range(NLAYER)]
        x = ops.embed(seq, 0, 1, self.w["decoder.emb"], math.sqrt(D), self.w["pe1d"], 0)
        tm = token_mask.to(dev).to(torch.uint8).contiguous()
        all_logits = []
        mode = (sampling or {}).get("name") or "deterministic"
        if mode not in ops.SAMPLING_MODES:
            raise NotImplementedError(f"sampling {mode!r}")
        plain = forced is None and mode == "deterministic"
        noise = None
        if not plain:
            if forced is not None:
                forced = forced.to(dev, torch.int32).contiguous()
                assert forced.shape == (B, steps), f"{forced.shape=}"
            if mode != "deterministic" and uniform is None:
                uniform = torch.rand((steps, B), dtype=torch.float32, device=dev, generator=rng)
            if mode == "gumbel":
                noise = torch.empty((B, self.vocab), dtype=torch.float32, device=dev)
        for t in range(steps):
            if step_hook is not None:  # profiling aid (profiles/launch_slice.py): called before every decode step
                step_hook(t)
            logits = self._decode_step(x, t, kc, vc, kvm, pad_mask, B, Mlen)
            if return_logits:
                all_logits.append(logits)
            if plain:
                ops.argmax_next(logits, tm[t], seq, t + 1, pad_mask, pad_id, self.w["decoder.emb"], math.sqrt(D),
                                self.w["pe1d"], x)
                continue
            if noise is not None:
                noise.uniform_(generator=rng)
            ops.sample_next(logits, tm[t], seq, t + 1, pad_mask, pad_id, self.w["decoder.emb"], math.sqrt(D),
                            self.w["pe1d"], x, forced=forced, step=t, mode=mode,
                            temperature=(sampling or {}).get("temperature", 1.0), top_k=(sampling or {}).get("top_k", 5),
                            top_p=(sampling or {}).get("top_p", 0.9), uniform=uniform[t] if uniform is not None else None,
                            noise=noise)
        out = seq[:, 1:]
        return (out, torch.stack(all_logits, 1)) if return_logits else out

    def generate_graphed(self, mem_s: torch.Tensor, B: int, Mlen: int, token_mask: torch.Tensor, bos_id: int, pad_id: int,
                         steps: int) -> torch.Tensor:
        """Plain greedy ``generate`` replayed from a CUDA graph cached per (B, Mlen, steps).

        Through the model-class API (``model.sample`` called batch after batch by inference.py) the decode loop is ~4 k
        launches issued from Python, i.e. host-bound for the batch sizes the reference's loaders use; the loop only
        depends on the memory K/V cache, so it is captured once per shape over a static cache that ``cross_kv`` then fills
        in place.  Default of ``model.sample()`` for plain greedy decoding (``RALF_SAMPLE_GRAPH=0`` switches it off); at most
        4 shapes are kept."""
        cache = self.__dict__.setdefault("_gen_graphs", {})
        key = (B, Mlen, steps, bos_id, pad_id)
        entry = cache.get(key)
        if entry is None:
            if len(cache) >= 4:
                cache.pop(next(iter(cache)))
            kv = self.alloc_cross_kv(B * Mlen, kv24=KV24 and self.npass == 3)
            self.cross_kv(mem_s, out=kv)
            tm = token_mask.to(self.dev).to(torch.uint8).contiguous()
            side = torch.cuda.Stream(device=self.dev)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):  # warm-up outside the capture (allocator, kernel attributes)
                self.generate(None, B, Mlen, tm, bos_id, pad_id, steps, kv=kv)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self.generate(None, B, Mlen, tm, bos_id, pad_id, steps, kv=kv)
            entry = cache[key] = (kv, graph, out, tm)
        else:
            self.cross_kv(mem_s, out=entry[0])
        entry[1].replay()
        return entry[2].clone()


class DecodeSession:
    """KV-cached next-token logits for ONE canvas with rewind, for samplers that backtrack on the host
    (``ralf_b200.relation.sample_with_backtracking``; reference: retrieval_augmented_autoreg.py:365-383, which re-runs the
    whole decoder over the prefix at every step).  ``logits_of(prefix)`` keeps the self-attention K/V rows of the longest
    prefix it has in common with the previous call and only runs the decoder for the tokens after it, so a rewind to
    position p costs nothing but the re-decode of what follows p."""

    def __init__(self, engine: "Engine", kvm: list, Mlen: int, steps: int, pad_id: int) -> None:
        self.eng, self.kvm, self.Mlen, self.steps, self.pad_id = engine, kvm, Mlen, steps, pad_id
        dev = engine.dev
        self.seq = torch.full((1, steps + 1), pad_id, dtype=torch.int64, device=dev)
        self.pad_mask = torch.zeros((1, steps + 1), dtype=torch.uint8, device=dev)
        self.kc = [torch.empty((1, steps, D), dtype=torch.float32, device=dev) for _ in range(NLAYER)]
        self.vc = [torch.empty((1, steps, D), dtype=torch.float32, device=dev) for _ in range(NLAYER)]
        self.cached: list = []       # tokens whose K/V rows are in the cache
        self.last_logits: Optional[torch.Tensor] = None

    def logits_of(self, prefix: list) -> torch.Tensor:
        """prefix = [<bos>, t1, ..., t_s] (1 <= len <= steps) -> host fp32 [V] logits of token s + 1."""
        n = len(prefix)
        assert 1 <= n <= self.steps, f"prefix of {n} tokens, cache holds {self.steps}"
        keep = 0
        while keep < min(n, len(self.cached)) and self.cached[keep] == prefix[keep]:
            keep += 1
        if keep == n and keep == len(self.cached) and self.last_logits is not None:
            return self.last_logits
        keep = min(keep, n - 1)  # the last token is always re-run: its logits are what the caller wants
        eng = self.eng
        new = torch.tensor(prefix[keep:], dtype=torch.int64)
        self.seq[0, keep:n] = new.to(eng.dev)
        self.pad_mask[0, keep:n] = (new == self.pad_id).to(torch.uint8).to(eng.dev)
        logits = None
        for t in range(keep, n):
            x = ops.embed(self.seq, t, 1, eng.w["decoder.emb"], math.sqrt(D), eng.w["pe1d"], t)
            logits = eng._decode_step(x, t, self.kc, self.vc, self.kvm, self.pad_mask, 1, self.Mlen)
        self.cached = list(prefix)
        self.last_logits = logits[0].float().cpu()
        return self.last_logits
