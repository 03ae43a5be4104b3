"""Micro-benchmark of ralf_decode_chain (csrc/decode_chain.cu) at the bench shape (B = 1024 rows): the three chains of a
decoder-layer step timed alone with CUDA events, (a) back to back (weights L2-resident) and (b) with a 256 MiB L2 flush
between launches (what the decode loop sees: the K/V streams of a token evict the weights), next to the per-op launches
they replace.  MMA issue order / prefetch are selected by RALF_CHAIN_ACC / RALF_CHAIN_PAIR / RALF_CHAIN_PREFETCH (read
once per process).

    python profiles/chain_bench.py [B]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    dev = torch.device("cuda:0")
    from ralf_b200 import ops

    g = torch.Generator(device=dev).manual_seed(0)

    def w(n, k):
        return ops.split_bf16(torch.randn(n, k, device=dev, generator=g) * 0.05)

    def v(n):
        return torch.randn(n, device=dev, generator=g)

    W = {"qkv": w(768, 256), "o": w(256, 256), "cq": w(256, 256), "co": w(256, 256), "l1": w(1024, 256), "l2": w(256, 1024),
         "head": w(519, 256)}
    bias = {k: v(t.shape[1]) for k, t in W.items()}
    ln = [(v(256), v(256)) for _ in range(4)]
    x = torch.randn(B, 256, device=dev, generator=g)
    a = ops.split_bf16(torch.randn(B, 256, device=dev, generator=g))
    qkv = torch.empty(B, 768, device=dev)
    q = torch.empty(B, 256, device=dev)
    logits = torch.empty(B, 519, device=dev)
    cs = ops.chain_stage
    chains = {
        "K1 ln+qkv (24 tiles)": lambda: ops.decode_chain(x, B, [cs(W["qkv"], bias=bias["qkv"], ln=ln[0], out_f32=qkv)]),
        "K2 o+ln+cq (16 tiles)": lambda: ops.decode_chain(x, B, [
            cs(W["o"], bias=bias["o"], in_split=a, add_x=True, to_x=True, out_f32=x),
            cs(W["cq"], bias=bias["cq"], ln=ln[1], out_f32=q)]),
        "K3 co+ln+ffn+ln+qkv (96 tiles)": lambda: ops.decode_chain(x, B, [
            cs(W["co"], bias=bias["co"], in_split=a, add_x=True, to_x=True),
            cs(W["l1"], bias=bias["l1"], ln=ln[2], act="relu", out_operand=True),
            cs(W["l2"], bias=bias["l2"], add_x=True, to_x=True, out_f32=x),
            cs(W["qkv"], bias=bias["qkv"], ln=ln[3], out_f32=qkv)]),
        "K3h co+ln+ffn+ln+head (92 tiles)": lambda: ops.decode_chain(x, B, [
            cs(W["co"], bias=bias["co"], in_split=a, add_x=True, to_x=True),
            cs(W["l1"], bias=bias["l1"], ln=ln[2], act="relu", out_operand=True),
            cs(W["l2"], bias=bias["l2"], add_x=True, to_x=True, out_f32=x),
            cs(W["head"], ln=ln[3], out_f32=logits)]),
    }

    def per_op():  # the launches K2 + K3 replace (one decoder layer minus the two attentions)
        ops.gemm(a, W["o"], bias=bias["o"], res=x, out_f32=x)
        _, h = ops.layernorm(x, *ln[1])
        ops.gemm(h, W["cq"], bias=bias["cq"], out_f32=q)
        ops.gemm(a, W["co"], bias=bias["co"], res=x, out_f32=x)
        _, h = ops.layernorm(x, *ln[2])
        _, f = ops.gemm(h, W["l1"], bias=bias["l1"], act="relu", want_f32=False, want_split=True)
        ops.gemm(f, W["l2"], bias=bias["l2"], res=x, out_f32=x)
        _, h = ops.layernorm(x, *ln[3])
        ops.gemm(h, W["qkv"], bias=bias["qkv"], out_f32=qkv)

    chains["per-op launches of K2 + K3 (9 kernels)"] = per_op
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    out = {"B": B, "switches": {k: v_ for k, v_ in os.environ.items() if k.startswith("RALF_")}, "us": {}}
    for name, fn in chains.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            fn()
        e1.record()
        torch.cuda.synchronize()
        warm = e0.elapsed_time(e1) / 50 * 1e3
        cold = []
        for _ in range(10):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            cold.append(e0.elapsed_time(e1) * 1e3)
        out["us"][name] = {"l2_warm": round(warm, 1), "l2_flushed_median": round(sorted(cold)[5], 1)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
