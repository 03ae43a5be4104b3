"""Per-kernel time of one training step (batch 32, 256x256) with torch.profiler (CUPTI): cheap alternative to an ncu
launch list for the ~1.9 k kernels of a step.  Prints the top kernels by total device time."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    dev = torch.device("cuda:0")
    from oracle import synth
    from ralf_b200 import generator as G
    from ralf_b200.tokenizer import LayoutSequenceTokenizer
    from ralf_b200.train import TrainEngine

    tok = LayoutSequenceTokenizer(["logo", "text", "underlay", "embellishment"], 10)
    model = G.RALF(features=None, tokenizer=tok, dataset_name="cgl", max_seq_length=10, top_k=16, auxilary_task="uncond")
    model.load_state_dict(bench.synth_weights_for(model), strict=True)
    model.to(dev)
    batch = synth.synth_batch(B, 256, 256, 10, 16, 4, seed=3)
    inputs, targets = model.preprocess(batch)
    inputs = {k: (v.to(dev) if torch.is_tensor(v) else {kk: vv.to(dev) for kk, vv in v.items()}) for k, v in inputs.items()}
    targets = {k: v.to(dev) for k, v in targets.items()}
    te = TrainEngine(model)
    for _ in range(2):
        te.train_step(inputs, targets)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        te.train_step(inputs, targets)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=70))


if __name__ == "__main__":
    main()
