"""BASELINE configs[3]: k-NN retrieval sweep -- gallery 10k -> 1M x 512-d fp32, top-16, Q in {1, 32, 128}.
Reports achieved GB/s (algorithmic bytes N*d*4 + Q*d*4 + Q*k*12 per call / CUDA-event time, L2 flushed between calls)
against MEASURED_PEAKS.json:hbm_gbs, the exact CUDA-core kernel for comparison, and a CPU baseline (numpy fp32 G @ q^T +
argpartition on the host cores, the restatement of FAISS IndexFlat IP).  With WORLD_SIZE > 1 the gallery is row-sharded
and the per-rank time is the max over ranks (all-gather + merge included).

    python profiles/knn_sweep.py            # 1 GPU
    torchrun --nproc-per-node N ... profiles/knn_sweep.py
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    from ralf_b200 import ops
    from ralf_b200.retrieval import GpuRetriever, shard_bounds

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs", 6650.0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    rows = []
    for n in (10_000, 100_000, 1_000_000):
        lo, hi = shard_bounds(n, world, rank)
        g = torch.Generator(device=dev).manual_seed(5 + rank)
        G = torch.nn.functional.normalize(torch.randn(hi - lo, 512, device=dev, generator=g), dim=1)
        retr = GpuRetriever(G, None, device=dev, rank=rank, world_size=world, index_base=lo)
        for q in (1, 32, 128):
            Q = torch.nn.functional.normalize(torch.randn(q, 512, device=dev, generator=g), dim=1)
            if world > 1:
                import torch.distributed as dist

                dist.broadcast(Q, 0)
            for _ in range(3):
                retr.search(Q, 16)
            times = []
            for _ in range(10):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                idx, score = retr.search(Q, 16)
                e1.record()
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1))
            ms = float(np.median(times))
            if world > 1:
                import torch.distributed as dist

                t = torch.tensor([ms], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            row = {"n": n, "q": q, "n_gpus": world, "ms": round(ms, 4),
                   "algorithmic_mb_per_gpu": round(((hi - lo) * 512 * 4 + q * 512 * 4 + q * 16 * 12) / 1e6, 2),
                   "certified": bool(retr.last_certified.all().item())}
            row["gbps_per_gpu"] = round(row["algorithmic_mb_per_gpu"] / ms, 1)
            row["frac_of_measured_hbm_peak"] = round(row["gbps_per_gpu"] / peak, 4)
            if world == 1:
                ts = []
                for _ in range(3):
                    flush.zero_()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    ei, es, _ = ops.knn_topk(G, Q, 16, exact=True)
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                row["exact_cuda_core_kernel_ms"] = round(float(np.median(ts)), 4)
                row["matches_exact_kernel"] = bool(torch.equal(ei, idx) and torch.equal(es, score))
                if q in (1, 32) or n <= 100_000:  # CPU restatement (bounded): numpy sgemm + argpartition
                    Gh, Qh = G.cpu().numpy(), Q.cpu().numpy()
                    t0 = time.time()
                    s = Gh @ Qh.T
                    np.argpartition(-s, 16, axis=0)
                    row["cpu_numpy_ms"] = round((time.time() - t0) * 1e3, 2)
            rows.append(row)
            if rank == 0:
                print(json.dumps(row), flush=True)
    if rank == 0:
        print(json.dumps({"peak_gbs": peak, "rows": rows}))
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()
