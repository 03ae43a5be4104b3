"""CPU error-budget study (no GPU): which reduced-precision choices keep the greedy token ids of the reference goldens
bit-exact?  Emulated inside the CPU oracle (oracle/ralf_oracle.py) on the golden inputs of tests/golden/:

  kv formats  -- the decode loop's cross-attention K/V cache: fp32 | 24-bit float (current: bf16-sized top half + one
                 mantissa byte, round to nearest) | int16 with one fp32 scale per (row, head) | bf16
  gemm passes -- "bf16x2" GEMMs in the ResNet trunk: activations rounded to bf16 (the x_lo . w_hi pass dropped)

For each variant: max |logit - fp32 oracle| / max |logit| over the greedy loop, and the number of token ids that differ
from the reference's golden sequence.

    python profiles/precision_study.py > profiles/r2_precision_study.json
"""
import json
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from oracle import ralf_oracle as O  # noqa: E402
from tests import helpers  # noqa: E402


def q24(x):
    u = x.contiguous().view(torch.int32)
    return ((u + 0x80) & ~0xFF).view(torch.float32)


def q_int16_head(x, nhead=8):
    sh = x.shape
    y = x.reshape(*sh[:-1], nhead, sh[-1] // nhead)
    amax = y.abs().amax(dim=-1, keepdim=True).clamp_min(1e-30)
    q = torch.round(y * (32767.0 / amax)).clamp(-32767, 32767)
    return (q * (amax / 32767.0)).reshape(sh)


KV = {"fp32": lambda x: x, "float24": q24, "int16_per_head_scale": q_int16_head, "bf16": lambda x: x.bfloat16().float()}


def patched_mha(kvq):
    orig = O._mha

    def mha(sd, p, xq, xkv, nhead, attn_mask=None, key_padding_mask=None):
        if not p.endswith("multihead_attn"):
            return orig(sd, p, xq, xkv, nhead, attn_mask, key_padding_mask)
        d = xq.shape[-1]
        w, b = sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"]
        q = F.linear(xq, w[:d], b[:d])
        k = kvq(F.linear(xkv, w[d:2 * d], b[d:2 * d]))
        v = kvq(F.linear(xkv, w[2 * d:], b[2 * d:]))
        B, Tq, _ = q.shape
        Tk, dh = k.shape[1], d // nhead
        q = q.view(B, Tq, nhead, dh).transpose(1, 2)
        k = k.view(B, Tk, nhead, dh).transpose(1, 2)
        v = v.view(B, Tk, nhead, dh).transpose(1, 2)
        o = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(dh), dim=-1) @ v
        return O._lin(sd, p + ".out_proj", o.transpose(1, 2).reshape(B, Tq, d))

    return mha


def run(name, schema, is_ralf):
    z, meta = helpers.load_golden(name)
    sd = helpers.synth_weights(schema, meta["seed"])
    batch = helpers.synth_batch(meta)
    labels = 3 if "pku" in name else 4
    tok = helpers.make_tokenizer(max_seq_length=meta.get("E", 10)) if labels == 4 else None
    if tok is None:
        from ralf_b200.tokenizer import LayoutSequenceTokenizer

        tok = LayoutSequenceTokenizer(["text", "logo", "underlay"], meta.get("E", 10))
    sc, sp = torch.from_numpy(z["seq_layout_const"]), torch.from_numpy(z["seq_layout_const_pad_mask"]).bool()
    ids = meta["special"]
    out = {}
    with torch.no_grad():
        def encode(conv_round=False):
            if conv_round:
                orig = F.conv2d
                F.conv2d = lambda x, *a, **k: orig(x.bfloat16().float(), *a, **k)
            try:
                img = helpers.image4(batch)
                if is_ralf:
                    return O.encode_ralf_memory(sd, img, {k: v.float() for k, v in batch["retrieved"].items()}, sc, sp)
                return O.encode_autoreg_memory(sd, img, sc, sp)
            finally:
                if conv_round:
                    F.conv2d = orig

        mem = encode()
        ref_seq, ref_lg = O.greedy_sample(sd, mem, tok.token_mask, ids["bos"], ids["pad"], tok.max_token_length, return_logits=True)
        fin = torch.isfinite(ref_lg)
        scale = ref_lg[fin].abs().max().item()
        assert (ref_seq.numpy() == z["gen_seq"]).all(), "oracle must reproduce the golden"
        for kname, kq in KV.items():
            orig = O._mha
            O._mha = patched_mha(kq)
            try:
                seq, lg = O.greedy_sample(sd, mem, tok.token_mask, ids["bos"], ids["pad"], tok.max_token_length, return_logits=True)
            finally:
                O._mha = orig
            same = (seq == ref_seq)
            # logits are only comparable up to the first differing token of a canvas
            ok = same.cumprod(dim=1).bool()
            m = fin & ok[:, :, None]
            out["kv_" + kname] = {"tokens_differing_from_golden": int((~same).sum()),
                                  "max_logit_err_of_scale": float(((lg - ref_lg)[m]).abs().max().item() / scale)}
        mem2 = encode(conv_round=True)
        seq, lg = O.greedy_sample(sd, mem2, tok.token_mask, ids["bos"], ids["pad"], tok.max_token_length, return_logits=True)
        same = (seq == ref_seq)
        ok = same.cumprod(dim=1).bool()
        m = fin & ok[:, :, None]
        out["resnet_bf16x2 (conv inputs rounded to bf16)"] = {
            "tokens_differing_from_golden": int((~same).sum()),
            "memory_err_of_scale": float((mem2 - mem).abs().max().item() / mem.abs().max().item()),
            "max_logit_err_of_scale": float(((lg - ref_lg)[m]).abs().max().item() / scale)}
    return out


if __name__ == "__main__":
    torch.set_num_threads(8)
    res = {}
    for name, schema, is_ralf in [("ralf_cgl_256", "ralf_cgl", True), ("ralf_cgl_350x240", "ralf_cgl", True),
                                  ("autoreg_cgl_350x240", "autoreg_cgl", False)]:
        res[name] = run(name, schema, is_ralf)
        print(name, json.dumps(res[name]), file=sys.stderr)
    print(json.dumps(res, indent=1))
