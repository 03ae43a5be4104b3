"""Training-step throughput (BASELINE configs[1]: RALF CGL k=16, d_model 256, 6-layer decoder, batch 32 per GPU).
Diagnostic companion of bench.py (whose contract metric is inference layouts/s): samples/s of
TrainEngine.train_step on synthetic 256x256 canvases, CUDA events, eager launches.

    python profiles/train_bench.py [batch] [steps]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:  # data parallel: per-GPU batch B, gradient all-reduce (SUM / world) over NCCL before the clip
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    from oracle import synth  # data generator only
    from ralf_b200 import generator as G
    from ralf_b200 import ops
    from ralf_b200.tokenizer import LayoutSequenceTokenizer
    from ralf_b200.train import TrainEngine

    tok = LayoutSequenceTokenizer(["logo", "text", "underlay", "embellishment"], 10)
    model = G.RALF(features=None, tokenizer=tok, dataset_name="cgl", max_seq_length=10, top_k=16, auxilary_task="uncond")
    model.load_state_dict(bench.synth_weights_for(model), strict=True)
    model.to(dev)
    batch = synth.synth_batch(B, 256, 256, 10, 16, 4, seed=3 + rank)
    inputs, targets = model.preprocess(batch)
    inputs = {k: (v.to(dev) if torch.is_tensor(v) else {kk: vv.to(dev) for kk, vv in v.items()}) for k, v in inputs.items()}
    targets = {k: v.to(dev) for k, v in targets.items()}
    te = TrainEngine(model, world_size=world)
    use_graph = os.environ.get("RALF_TRAIN_GRAPH", "1") != "0"
    losses = []
    if use_graph:  # one captured CUDA graph per step instead of ~1.9 k eager launches
        te.capture(inputs, targets)
        step = te.train_step_graph
    else:
        step = te.train_step
    for _ in range(2):
        losses.append(float(step(inputs, targets)))
    torch.cuda.synchronize()
    n0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step(inputs, targets)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    losses.append(float(loss))
    if world > 1:
        import torch.distributed as dist

        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        # replicas must stay bit-identical: same init, same averaged gradients
        chk = te.ps.flat_p[:4096].clone()
        ref = chk.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(chk, ref), "data-parallel replicas diverged"
        dist.destroy_process_group()
        if rank != 0:
            return
    print(json.dumps({"n_gpus": world, "global_batch": B * world, "what": "train_step (fwd + bwd + clip + AdamW), bf16x3 tensor-core GEMMs, fp32 master weights" + (", CUDA-graph replay" if use_graph else ", eager launches"),
                      "batch": B, "canvas": "256x256x4", "ms_per_step": round(ms, 2), "samples_per_s": round(B * world / ms * 1e3, 1),
                      "kernels_per_step": (ops.launch_count() - n0) // steps, "loss_trace": [round(x, 4) for x in losses],
                      "limits": list(te.limits), "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 2)}))


if __name__ == "__main__":
    main()
