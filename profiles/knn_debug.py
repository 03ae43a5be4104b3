"""Profiling aid: time the k-NN scan (N = 1M x 512, Q = 128) in RALF_KNN_DEBUG modes 0/1/2 (set by the caller's env).
Mode 0 = product; 1 = epilogue reads TMEM but skips the filter; 2 = epilogue only returns the accumulator."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from ralf_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
out = {"mode": os.environ.get("RALF_KNN_DEBUG", "0")}
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for q in (1, 32, 128):
    G = torch.nn.functional.normalize(torch.randn(1_000_000, 512, device=dev, generator=g), dim=1)
    Q = torch.nn.functional.normalize(torch.randn(q, 512, device=dev, generator=g), dim=1)
    for _ in range(3):
        ops.knn_topk(G, Q, 16)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.knn_topk(G, Q, 16); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    out[f"q{q}_ms_median"] = round(ts[len(ts) // 2], 4)
    out[f"q{q}_gbps"] = round(2.048e9 / (ts[len(ts) // 2] * 1e-3) / 1e9, 1)
print(json.dumps(out))
