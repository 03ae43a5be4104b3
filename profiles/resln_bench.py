"""Decode-loop micro-benchmark: residual GEMM + LayerNorm as two launches vs ralf_gemm_res_ln (cluster kernel), replayed from a
CUDA graph of 64 dependent iterations (M = 1024 canvases).  Diagnostic; CUDA events."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ralf_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
M = 1024
for K in (256, 1024):
    a = ops.split_bf16(torch.randn(M, K, device=dev))
    w = ops.split_bf16(torch.randn(256, K, device=dev) / K ** 0.5)
    bias = torch.randn(256, device=dev)
    gamma, beta = torch.randn(256, device=dev), torch.randn(256, device=dev)
    x = torch.randn(M, 256, device=dev)
    h = torch.empty(2, M, 256, dtype=torch.bfloat16, device=dev)

    def two():
        ops.gemm(a, w, bias=bias, res=x, out_f32=x)
        ops.layernorm(x, gamma, beta)

    def one():
        ops.gemm_res_ln(a, w, x, gamma, beta, bias=bias, ln_split=h)

    def gemm_only():
        ops.gemm(a, w, bias=bias, res=x, out_f32=x)

    for name, fn in (("gemm + layernorm", two), ("gemm_res_ln", one), ("gemm only", gemm_only)):
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for _ in range(3):
                fn()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=s):
                for _ in range(64):
                    fn()
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                g.replay()
            e1.record()
        torch.cuda.synchronize()
        print(f"K={K} {name}: {e0.elapsed_time(e1) / 5 / 64 * 1e3:.2f} us per iteration")
        x.normal_()
