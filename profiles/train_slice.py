"""ncu launch list of ONE training step (BASELINE configs[1]: batch 32, 256x256, dropout on, eager launches -- the same
kernels the captured graph replays): two warm-up steps, then one step between cudaProfilerStart / Stop.

    ncu --profile-from-start off --kernel-name-base demangled --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/train_slice.csv python profiles/train_slice.py
    python profiles/train_slice.py --summarize gpurun_out/train_slice.csv
"""
import collections
import csv
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def summarize(path):
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"])
        unit = row["Metric Unit"]
        v = v / 1000 if unit == "ns" else (v * 1000 if unit == "ms" else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"training step (cold-cache, serialised under ncu): {tot / 1000:.2f} ms over {sum(v[0] for v in agg.values())} kernels")
    print("| kernel | total ms | share | launches | avg us |\n|---|---:|---:|---:|---:|")
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{name}` | {us / 1000:.3f} | {100 * us / tot:.1f}% | {n} | {us / n:.1f} |")


def main():
    import torch

    import bench
    from oracle import synth
    from ralf_b200 import generator as G
    from ralf_b200.tokenizer import LayoutSequenceTokenizer
    from ralf_b200.train import TrainEngine

    dev = torch.device("cuda:0")
    tok = LayoutSequenceTokenizer(["logo", "text", "underlay", "embellishment"], 10)
    model = G.RALF(features=None, tokenizer=tok, dataset_name="cgl", max_seq_length=10, top_k=16, auxilary_task="uncond")
    model.load_state_dict(bench.synth_weights_for(model), strict=True)
    model.to(dev)
    batch = synth.synth_batch(32, 256, 256, 10, 16, 4, seed=3)
    inputs, targets = model.preprocess(batch)
    inputs = {k: (v.to(dev) if torch.is_tensor(v) else {kk: vv.to(dev) for kk, vv in v.items()}) for k, v in inputs.items()}
    targets = {k: v.to(dev) for k, v in targets.items()}
    te = TrainEngine(model)
    for _ in range(2):
        te.train_step(inputs, targets)
    torch.cuda.synchronize()
    cudart = torch.cuda.cudart()
    cudart.cudaProfilerStart()
    te.train_step(inputs, targets)
    torch.cuda.synchronize()
    cudart.cudaProfilerStop()
    print("train slice done")


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--summarize":
        summarize(sys.argv[2])
    else:
        main()
