"""GEMM micro-benchmark over the shapes the RALF step actually launches (micro-batch of 128 canvases, 256x256).
Prints time, tensor rate (3 passes counted) and HBM-side bytes/s per shape.  Diagnostic; CUDA events, L2 flushed.

    python profiles/gemm_bench.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ralf_b200 import ops  # noqa: E402

SHAPES = [  # name, M, N, K, residual ("split" | "f32" | None), out ("split" | "f32"), act
    ("l1.c3  ", 524288, 256, 64, "split", "split", None),
    ("l1.ds  ", 524288, 256, 64, None, "split", None),
    ("l1.c1  ", 524288, 64, 256, None, "split", "relu"),
    ("l1.c2  ", 524288, 64, 576, None, "split", "relu"),
    ("l2.c3  ", 131072, 512, 128, "split", "split", None),
    ("l3.c3  ", 32768, 1024, 256, "split", "split", None),
    ("l3.c2  ", 32768, 256, 2304, None, "split", "relu"),
    ("l4.c2  ", 8192, 512, 4608, None, "split", "relu"),
    ("l2.c1  ", 131072, 128, 512, None, "split", "relu"),
    ("l3.c1  ", 32768, 256, 1024, None, "split", "relu"),
    ("l3.ds  ", 32768, 1024, 512, None, "split", None),
    ("l4.c3  ", 8192, 2048, 512, "split", "split", None),
    ("l4.ds  ", 8192, 2048, 1024, None, "split", None),
    ("enc.qkv", 32768, 768, 256, None, "f32", None),
    ("enc.o  ", 32768, 256, 256, "f32", "f32", None),
    ("enc.l1 ", 32768, 1024, 256, None, "split", "relu"),
    ("enc.l2 ", 32768, 256, 1024, "f32", "f32", None),
    ("ckv    ", 68096, 512, 256, None, "f32", None),
    ("dec.qkv", 1024, 768, 256, None, "f32", None),
    ("dec.o  ", 1024, 256, 256, "f32", "f32", None),
    ("dec.l1 ", 1024, 1024, 256, None, "split", "relu"),
]


def main():
    dev = torch.device("cuda:0")
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    rows = []
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    for name, M, N, K, res, out, act in SHAPES:
        if only and only not in name:
            continue
        a = torch.randn(2, M, K, device=dev).to(torch.bfloat16)
        w = torch.randn(2, N, K, device=dev).to(torch.bfloat16)
        bias = torch.randn(N, device=dev)
        kw = dict(bias=bias, act=act, block_n=int(os.environ.get("RALF_BENCH_BN", "0")))
        byt = 2 * M * K * 2 + 2 * N * K * 2
        if res == "split":
            kw["res_split"] = torch.randn(2, M, N, device=dev).to(torch.bfloat16)
            kw["post_relu"] = True
            byt += 2 * M * N * 2
        elif res == "f32":
            kw["res"] = torch.randn(M, N, device=dev)
            byt += M * N * 4
        if out == "split":
            kw["out_split"] = torch.empty(2, M, N, dtype=torch.bfloat16, device=dev)
            kw["want_f32"] = False
        else:
            kw["out_f32"] = torch.empty(M, N, device=dev)
        byt += M * N * 4
        for _ in range(2):
            ops.gemm(a, w, **kw)
        ts = []
        for _ in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.gemm(a, w, **kw)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        row = {"shape": name.strip(), "M": M, "N": N, "K": K, "us": round(ms * 1e3, 1),
               "tflops_3pass": round(3 * 2.0 * M * N * K / ms / 1e9, 1), "gbps": round(byt / ms / 1e6, 1)}
        rows.append(row)
        print(f"{name} M={M:7d} N={N:5d} K={K:5d}  {ms * 1e3:8.1f} us  {row['tflops_3pass']:7.1f} TF/s(x3)  {row['gbps']:7.1f} GB/s",
              flush=True)
        del a, w, kw
    print(json.dumps(rows))


if __name__ == "__main__":
    main()
