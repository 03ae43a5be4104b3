"""Encoder self-attention micro-benchmark (B = 128 canvases, T = 256 tokens, 8 heads x 32): tcgen05 kernel vs the
CUDA-core kernel (RALF_ATTN_TC=0).  Diagnostic; CUDA events."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ralf_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
B, T, H, dh = 128, 256, 8, 32
qkv = torch.randn(B * T, 3 * H * dh, device=dev)
D = H * dh
for _ in range(3):
    ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], B, H, T, T, dh)
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ops.attention(qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], B, H, T, T, dh)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
fl = 2 * 2.0 * B * H * T * T * dh
print(f"attention B={B} T={T}: {ts[5] * 1e3:.1f} us  ({fl / ts[5] / 1e9:.1f} TFLOP/s algorithmic, tc={os.environ.get('RALF_ATTN_TC', '1')})")
