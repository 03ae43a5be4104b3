"""TEST INFRASTRUCTURE ONLY -- CPU (torch fp32) restatement of the RALF / Autoreg forward and greedy
generation path, written against the reference's state-dict schema.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  The product path (ralf_b200/engine.py + csrc/) never does.

Pinned against the UNMODIFIED reference classes run in the build container: tests/golden/*.npz hold the
reference's own outputs (memory, logits, greedy token ids) for seeded synthetic weights and inputs;
tests/test_oracle_golden.py checks this restatement against them (script: tests/golden/make_golden.py).

Every function cites the reference code it follows (paths relative to image2layout/train/).
It deliberately keeps the reference's algorithmic shape: no KV cache, the decoder is re-run over the
whole prefix at every greedy step (models/retrieval_augmented_autoreg.py:271-297).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import Tensor

NHEAD = 8
NUM_LAYERS = 6


# ---------------------------------------------------------------------------------------------------
# building blocks
# ---------------------------------------------------------------------------------------------------
def _lin(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def _ln(sd, p, x):
    return F.layer_norm(x, (x.shape[-1],), sd[p + ".weight"], sd[p + ".bias"], 1e-5)


def _mha(sd, p, xq, xkv, nhead, attn_mask=None, key_padding_mask=None):
    """nn.MultiheadAttention (batch_first), packed in_proj [Wq; Wk; Wv]; masks are additive -inf."""
    d = xq.shape[-1]
    w, b = sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"]
    q = F.linear(xq, w[:d], b[:d])
    k = F.linear(xkv, w[d:2 * d], b[d:2 * d])
    v = F.linear(xkv, w[2 * d:], b[2 * d:])
    B, Tq, _ = q.shape
    Tk = k.shape[1]
    dh = d // nhead
    q = q.view(B, Tq, nhead, dh).transpose(1, 2)
    k = k.view(B, Tk, nhead, dh).transpose(1, 2)
    v = v.view(B, Tk, nhead, dh).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(dh)
    if attn_mask is not None:
        s = s + attn_mask
    if key_padding_mask is not None:
        s = s.masked_fill(key_padding_mask[:, None, None, :], float("-inf"))
    o = torch.softmax(s, dim=-1) @ v
    o = o.transpose(1, 2).reshape(B, Tq, d)
    return _lin(sd, p + ".out_proj", o)


def _enc_layer_prenorm(sd, p, x, key_padding_mask=None):
    """nn.TransformerEncoderLayer(norm_first=True, activation=relu), eval mode."""
    h = _ln(sd, p + ".norm1", x)
    x = x + _mha(sd, p + ".self_attn", h, h, NHEAD, key_padding_mask=key_padding_mask)
    h = _ln(sd, p + ".norm2", x)
    return x + _lin(sd, p + ".linear2", torch.relu(_lin(sd, p + ".linear1", h)))


def _enc_layer_postnorm(sd, p, x, nhead, key_padding_mask=None):
    """nn.TransformerEncoderLayer default (post-norm), used by FIDNetV3 (fid/model.py:26-33)."""
    x = _ln(sd, p + ".norm1", x + _mha(sd, p + ".self_attn", x, x, nhead, key_padding_mask=key_padding_mask))
    return _ln(sd, p + ".norm2", x + _lin(sd, p + ".linear2", torch.relu(_lin(sd, p + ".linear1", x))))


def _dec_layer_prenorm(sd, p, x, memory, tgt_mask, tgt_key_padding_mask):
    """nn.TransformerDecoderLayer(norm_first=True): x += SA(LN1 x); x += CA(LN2 x, mem); x += FF(LN3 x)."""
    h = _ln(sd, p + ".norm1", x)
    x = x + _mha(sd, p + ".self_attn", h, h, NHEAD, attn_mask=tgt_mask, key_padding_mask=tgt_key_padding_mask)
    h = _ln(sd, p + ".norm2", x)
    x = x + _mha(sd, p + ".multihead_attn", h, memory, NHEAD)
    h = _ln(sd, p + ".norm3", x)
    return x + _lin(sd, p + ".linear2", torch.relu(_lin(sd, p + ".linear1", h)))


def _feed_forward(sd, p, x):
    """common/attention.py:15-30  LN -> Linear -> GELU -> Linear (dropouts are identity in eval)."""
    return _lin(sd, p + ".net.4", F.gelu(_lin(sd, p + ".net.1", _ln(sd, p + ".net.0", x))))


def _pe1d(sd, p, x):
    """common/positional_encoding.py:94-107 (scale_input=True, batch_first)."""
    return x * math.sqrt(x.shape[-1]) + sd[p + ".pe"][:, : x.shape[1]]


# ---------------------------------------------------------------------------------------------------
# image encoder
# ---------------------------------------------------------------------------------------------------
BN_TRAIN = False  # tests of the training step set this: BatchNorm2d batch statistics (model.train())


def _bn(sd, p, x):
    if BN_TRAIN:
        return F.batch_norm(x, sd[p + ".running_mean"].clone(), sd[p + ".running_var"].clone(), sd[p + ".weight"],
                            sd[p + ".bias"], True, 0.1, 1e-5)
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, 1e-5)


def _bottleneck(sd, p, x, stride):
    """timm / torchvision ResNet-v1.5 Bottleneck: 1x1 -> 3x3 (stride) -> 1x1, relu(out + identity)."""
    idt = x
    out = torch.relu(_bn(sd, p + ".bn1", F.conv2d(x, sd[p + ".conv1.weight"])))
    out = torch.relu(_bn(sd, p + ".bn2", F.conv2d(out, sd[p + ".conv2.weight"], stride=stride, padding=1)))
    out = _bn(sd, p + ".bn3", F.conv2d(out, sd[p + ".conv3.weight"]))
    if (p + ".downsample.0.weight") in sd:
        idt = _bn(sd, p + ".downsample.1", F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride))
    return torch.relu(out + idt)


def resnet_fpn(sd, img, prefix="encoder.extractor"):
    """ResnetBackbone.forward (models/common/image.py:90-120), head == "transformer"."""
    b = prefix + ".body"
    x = torch.relu(_bn(sd, b + ".bn1", F.conv2d(img, sd[b + ".conv1.weight"], stride=2, padding=3)))
    x = F.max_pool2d(x, 3, 2, 1)
    feats = {}
    for li, (nblk, stride) in enumerate([(3, 1), (4, 2), (6, 2), (3, 2)], start=1):
        for bi in range(nblk):
            x = _bottleneck(sd, f"{b}.layer{li}.{bi}", x, stride if bi == 0 else 1)
        feats[li] = x
    f4 = F.conv2d(feats[3], sd[prefix + ".fpn_conv11_4.weight"], sd[prefix + ".fpn_conv11_4.bias"])
    f5 = F.conv2d(feats[4], sd[prefix + ".fpn_conv11_5.weight"], sd[prefix + ".fpn_conv11_5.bias"])
    f5up = F.interpolate(f5, size=f4.shape[2:], mode="nearest")
    fused = torch.cat(
        [f5up, F.conv2d(f5up + f4, sd[prefix + ".fpn_conv33.weight"], sd[prefix + ".fpn_conv33.bias"], padding=1)],
        dim=1)
    return F.conv2d(fused, sd[prefix + ".proj.weight"], sd[prefix + ".proj.bias"])


def pos_emb_2d(h: int, w: int, d_model: int = 256) -> Tensor:
    """PositionEmbeddingSine(normalize=True) table [h*w, d] (common/positional_encoding.py:182-210)."""
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
    return torch.cat((py, px), dim=1)


def encode_image(sd, image):
    """encoder -> pos_emb_2d -> transformer_encoder (retrieval_augmented_autoreg.py:967-971)."""
    f = resnet_fpn(sd, image)
    B, C, h, w = f.shape
    x = f.flatten(2).transpose(1, 2) + pos_emb_2d(h, w, C)[None].to(f.device)
    for i in range(NUM_LAYERS):
        x = _enc_layer_prenorm(sd, f"transformer_encoder.layers.{i}", x)
    return x


# ---------------------------------------------------------------------------------------------------
# retrieved-layout branch, fusion, constraint encoder
# ---------------------------------------------------------------------------------------------------
def fidnet_features(sd, lay, p="layout_encoer"):
    """FIDNetV3.extract_features (fid/model.py:95-103) for layouts [N, E] -> CLS feature [N, 256]."""
    bbox = torch.stack([lay[k] for k in ["center_x", "center_y", "width", "height"]], dim=-1).float()
    h = torch.cat([_lin(sd, p + ".fc_bbox", bbox), sd[p + ".emb_label.weight"][lay["label"].long()]], dim=-1)
    x = torch.relu(_lin(sd, p + ".enc_fc_in", h))
    N = x.shape[0]
    x = torch.cat([sd[p + ".enc_transformer.token"].reshape(1, 1, -1).expand(N, 1, -1), x], dim=1)
    pad = torch.cat([torch.zeros(N, 1, dtype=torch.bool, device=x.device), ~lay["mask"].bool()], dim=1)
    for i in range(4):
        x = _enc_layer_postnorm(sd, f"{p}.enc_transformer.core.layers.{i}", x, 4, key_padding_mask=pad)
    return x[:, 0]


def retrieved_features(sd, retrieved, top_k):
    """extract_retrieved_features (retrieval_augmented_autoreg.py:526-584), use_reference_image=False."""
    refs = []
    for k in range(top_k):
        lay = {key: retrieved[key][:, k] for key in ["center_x", "center_y", "width", "height", "label", "mask"]}
        refs.append(_feed_forward(sd, "layout_adapter", fidnet_features(sd, lay)))
    return _pe1d(sd, "pos_emb_1d", torch.stack(refs, dim=1))


def fusion_attention(sd, x, ctx, p="attn", heads=8, dim_head=64):
    """Attention.forward (common/attention.py:49-71): LN on x only, no residual."""
    x = _ln(sd, p + ".norm", x)
    q = F.linear(x, sd[p + ".to_q.weight"])
    k, v = F.linear(ctx, sd[p + ".to_kv.weight"]).chunk(2, dim=-1)
    B, n, _ = q.shape
    sp = lambda t: t.view(B, t.shape[1], heads, dim_head).transpose(1, 2)
    q, k, v = sp(q), sp(k), sp(v)
    a = torch.softmax((q @ k.transpose(-1, -2)) * dim_head ** -0.5, dim=-1)
    o = (a @ v).transpose(1, 2).reshape(B, n, heads * dim_head)
    return _lin(sd, p + ".to_out.0", o)


def constraint_encoder(sd, seq, pad_mask, p="user_const_encoder"):
    """UserConstraintTransformerEncoder.forward (common/common.py:238-252), task_token=None."""
    h = _pe1d(sd, p + ".pos_emb", sd[p + ".emb.weight"][seq])
    for i in range(NUM_LAYERS):
        h = _enc_layer_prenorm(sd, f"{p}.encoder.layers.{i}", h, key_padding_mask=pad_mask)
    return h


def uncond_constraint(tok_n_total: int, B: int):
    """UnconditionalPreprocessor (layoutformerpp/task_preprocessor.py:354-384): [bos, uncondition,
    end_of_task, eos]; ids follow the preprocessor vocabulary = tokenizer vocab (minus nothing) + special
    + task tokens; resolved by the caller from the golden fixture / reference instance."""
    raise NotImplementedError


def encode_ralf_memory(sd, image, retrieved, seq_const, seq_const_pad, top_k=16):
    """ConcateAuxilaryTaskConcateCrossAttnRetrievalAugmentedAutoreg._encode_into_memory
    (retrieval_augmented_autoreg.py:963-994 + 1004-1033)."""
    memory = encode_image(sd, image)
    ref = retrieved_features(sd, retrieved, top_k)
    memory_ca = fusion_attention(sd, memory, ref)
    mem = _feed_forward(sd, "head", torch.cat([memory, memory_ca, ref], dim=1))
    uc = constraint_encoder(sd, seq_const, seq_const_pad)
    t = sd["task_emb.weight"]
    return torch.cat([mem + t[sd["flag_img"]], uc + t[sd["flag_user_const"]]], dim=1)


def encode_autoreg_memory(sd, image, seq_const, seq_const_pad):
    """ConcateAuxilaryTaskAutoreg._encode_into_memory (models/autoreg.py:590-622)."""
    memory = encode_image(sd, image)
    uc = constraint_encoder(sd, seq_const, seq_const_pad)
    t = sd["task_emb.weight"]
    return torch.cat([memory + t[sd["flag_img"]], uc + t[sd["flag_user_const"]]], dim=1)


# ---------------------------------------------------------------------------------------------------
# decoder
# ---------------------------------------------------------------------------------------------------
def decoder_logits(sd, tgt, memory, tgt_key_padding_mask, p="decoder"):
    """BaseDecoder.forward with is_causal=True (common/common.py:84-135)."""
    h = _pe1d(sd, p + ".pos_emb", sd[p + ".emb.weight"][tgt])
    S = h.shape[1]
    causal = torch.triu(torch.full((S, S), float("-inf"), device=h.device, dtype=h.dtype), diagonal=1)
    for i in range(NUM_LAYERS):
        h = _dec_layer_prenorm(sd, f"{p}.transformer.layers.{i}", h, memory, causal, tgt_key_padding_mask)
    return F.linear(_ln(sd, p + ".head.0", h), sd[p + ".head.1.weight"])


def greedy_sample(sd, memory, token_mask, bos_id, pad_id, max_token_length, return_logits=False):
    """BaseRetrievalAugmentedAutoreg.sample greedy loop, cond_type uncond
    (retrieval_augmented_autoreg.py:244-300; helpers/sampling.py:24-25).  Returns seq without BOS."""
    B = memory.shape[0]
    inp = torch.full((B, 1), bos_id, dtype=torch.long, device=memory.device)
    step_logits = []
    for i in range(max_token_length):
        logits = decoder_logits(sd, inp, memory, inp == pad_id)[:, i].clone()
        logits[:, ~token_mask[i]] = float("-inf")
        if return_logits:
            step_logits.append(logits)
        inp = torch.cat([inp, torch.argmax(logits, dim=1, keepdim=True)], dim=1)
    return (inp[:, 1:], torch.stack(step_logits, 1)) if return_logits else inp[:, 1:]


# ---------------------------------------------------------------------------------------------------
# constrained tasks: decoding-space restriction, literal per-sample restatement
# ---------------------------------------------------------------------------------------------------
def restrict_logits(cond_type, sampling_idx, cond_seq, logits, pad_id, eos_id):
    """DECODE_SPACE_RESTRICTION[cond_type] (layoutformerpp/decoding_space_restriction.py:5-106): c / cwh keep only the
    given token (or <eos> from the first <pad> on); refinement does the same at label slots only; others: identity."""
    if cond_type in (None, "none", "uncond", "partial"):
        return logits
    if cond_type == "refinement" and (sampling_idx - 1) % 5 != 0:
        return logits
    for b in range(cond_seq.shape[0]):
        row = cond_seq[b]
        pads = (row == pad_id).nonzero()
        first_pad = int(pads[0]) if len(pads) else float("inf")
        given = int(row[sampling_idx])
        if sampling_idx < first_pad:
            if given == pad_id or given == -1:
                continue
            keep = given
        else:
            keep = eos_id
        v = logits[b, keep].clone()
        logits[b] = float("-inf")
        logits[b, keep] = v
    return logits


def constrained_greedy_sample(sd, memory, token_mask, cond_type, cond_seq, bos_id, pad_id, eos_id, max_token_length):
    """BaseRetrievalAugmentedAutoreg.sample for the constrained tasks (retrieval_augmented_autoreg.py:244-300), greedy,
    no KV cache.  ``cond_seq`` is cond.seq AFTER the task preprocessor ran (it rewrites <eos> to <pad> in place)."""
    B = memory.shape[0]
    inp = torch.full((B, 1), bos_id, dtype=torch.long)
    start = 0
    if cond_type == "partial":
        inp = torch.cat([inp, cond_seq[:, 1:6]], dim=1)
        start = 5
    for i in range(start, max_token_length):
        logits = decoder_logits(sd, inp, memory, inp == pad_id)[:, i].clone()
        logits[:, ~token_mask[i]] = float("-inf")
        logits = restrict_logits(cond_type, i + 1, cond_seq, logits, pad_id, eos_id)
        inp = torch.cat([inp, torch.argmax(logits, dim=1, keepdim=True)], dim=1)
    return inp[:, 1:]


# ---------------------------------------------------------------------------------------------------
# stochastic sampling (helpers/sampling.py:10-68)
# ---------------------------------------------------------------------------------------------------
def filtered_logits(logits, name, temperature=1.0, top_k=5, top_p=0.9):
    """logits / temperature with helpers/sampling.py's filter applied (-inf = dropped), name in random / top_k / top_p.
    top_k keeps every logit >= the k-th largest (ties included, :10-15); top_p sorts descending, drops every entry whose
    inclusive cumulative probability exceeds top_p except the first (:35-50)."""
    x = logits / temperature
    if name == "top_k":
        kth = torch.topk(x, top_k, dim=1).values[:, -1:]
        x = torch.where(x < kth, torch.full_like(x, float("-inf")), x)
    elif name == "top_p":
        s, idx = torch.sort(x, descending=True, dim=1)
        cum = torch.cumsum(torch.softmax(s, dim=1), dim=1)
        drop = cum > top_p
        drop[:, 0] = False
        s = torch.where(drop, torch.full_like(s, float("-inf")), s)
        x = torch.empty_like(s).scatter_(1, idx, s)
    elif name != "random":
        raise NotImplementedError(name)
    return x


def filtered_probs(logits, name, temperature=1.0, top_k=5, top_p=0.9):
    """The probabilities helpers/sampling.py hands to torch.multinomial (:60)."""
    return torch.softmax(filtered_logits(logits, name, temperature, top_k, top_p), dim=1)


def inverse_cdf_draw(x, u):
    """The draw ralf_b200 defines in place of torch.multinomial's generator-specific one (same distribution): walk the
    kept tokens of the filtered logits ``x`` in (logit desc, index asc) order and return the first whose cumulative
    softmax mass exceeds u * (kept mass).  float64 here.  Also returns, per row, the set of tokens that are acceptable
    when fp32 rounding moves the boundary: every token whose cumulative interval comes within 1e-5 of the target."""
    x = x.double()
    picks, accept = [], []
    for b in range(x.shape[0]):
        order = sorted((c for c in range(x.shape[1]) if x[b, c] > float("-inf")), key=lambda c: (-float(x[b, c]), c))
        m = float(x[b, order[0]])
        p = [math.exp(float(x[b, c]) - m) for c in order]
        tot = sum(p)
        target = float(u[b]) * tot
        acc, pick, ok = 0.0, order[-1], set()
        found = False
        for c, pc in zip(order, p):
            lo, acc = acc, acc + pc
            if lo - 1e-5 * tot <= target <= acc + 1e-5 * tot:
                ok.add(c)
            if not found and acc > target:
                pick, found = c, True
        picks.append(pick)
        accept.append(ok | {pick})
    return torch.tensor(picks), accept
