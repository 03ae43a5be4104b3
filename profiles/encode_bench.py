"""Encoder-only timing at the reference's real canvas size (350 x 240 -> 22 x 15 = 330 image tokens) and at the synthetic
256 x 256 bench size: Engine.encode of 128 canvases, CUDA events.  Used for the RALF_ATTN_TC_BIG A/B (tcgen05 attention for
256 < Tk <= 480; without it those shapes take the CUDA-core kernel).

    python profiles/encode_bench.py [H W]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    H = int(sys.argv[1]) if len(sys.argv) > 2 else 350
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 240
    dev = torch.device("cuda:0")
    from oracle import synth  # data generator only
    from ralf_b200 import generator as G
    from ralf_b200.tokenizer import LayoutSequenceTokenizer

    tok = LayoutSequenceTokenizer(["logo", "text", "underlay", "embellishment"], 10)
    model = G.RALF(features=None, tokenizer=tok, dataset_name="cgl", max_seq_length=10, top_k=16, auxilary_task="uncond")
    model.load_state_dict(bench.synth_weights_for(model), strict=True)
    model.eval().to(dev)
    eng = model.engine()
    B = 128
    b = synth.synth_batch(B, H, W, 10, 16, 4, seed=1)
    img = torch.cat([b["image"], b["saliency"]], 1).to(dev)
    retrieved = {k: v.to(dev) for k, v in b["retrieved"].items()}
    const = model.preprocessor(G.ConditionalInputs(image=img))
    cs, cp = const["seq"].to(dev), const["pad_mask"].to(dev)
    for _ in range(2):
        eng.encode(img, retrieved, cs, cp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        eng.encode(img, retrieved, cs, cp)
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"canvas": f"{H}x{W}", "canvases": B, "encode_ms": round(e0.elapsed_time(e1) / 5, 3),
                      "switches": {k: v for k, v in os.environ.items() if k.startswith("RALF_")}}))


if __name__ == "__main__":
    main()
