"""ncu launch-list slice of one bench step (1024 canvases): profiles ONE encoder micro-batch (128 canvases: retrieval
pass + fetch + ResNet/FPN + encoders + cross-K/V GEMMs) and TWO decode steps (t = 30, 31) between
cudaProfilerStart/Stop, so the per-kernel list costs ~1 GPU-minute instead of ~15.  Scale: encoder slice x 8,
decode slice x 30 for the step's shares.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/slice.csv python profiles/launch_slice.py
"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    B, MB = 1024, 128
    args = types.SimpleNamespace(elems=12, gallery=1_000_000, hw=256, batch=B, precision="bf16x3")
    dev = torch.device("cuda:0")
    from ralf_b200 import generator as G

    retr, tok = bench.synth_world(args, 0, 1, dev)
    model = G.RALF(features=None, tokenizer=tok, dataset_name="cgl", max_seq_length=12, top_k=16, auxilary_task="uncond")
    model.load_state_dict(bench.synth_weights_for(model), strict=True)
    model.eval().to(dev)
    eng = model.engine()
    img = torch.rand(B, 4, 256, 256, device=dev)
    qry = torch.nn.functional.normalize(torch.randn(B, 512, device=dev), dim=1)
    const = model.preprocessor(G.ConditionalInputs(image=img))
    cs, cp = const["seq"].to(dev).contiguous(), const["pad_mask"].to(dev).to(torch.uint8).contiguous()
    tm = tok.token_mask.to(dev).to(torch.uint8)
    ids = model.special_token_ids
    cudart = torch.cuda.cudart()

    def encode_mb(b0, kv, profile):
        if profile:
            torch.cuda.synchronize()
            cudart.cudaProfilerStart()
        idx, _ = retr.search_local(qry[b0:b0 + MB], 16)
        packed = retr.fetch(idx)["packed"]
        mem, mem_s = eng.encode(img[b0:b0 + MB], packed, cs[b0:b0 + MB], cp[b0:b0 + MB])
        Mlen = mem.shape[1]
        if kv is None:
            kv = eng.alloc_cross_kv(B * Mlen, kv24=True)
        eng.cross_kv(mem_s, out=kv, row0=b0 * Mlen)
        if profile:
            torch.cuda.synchronize()
            cudart.cudaProfilerStop()
        return kv, Mlen

    kv, Mlen = None, 0
    for rep in range(2):  # first pass warms everything up un-profiled
        for b0 in range(0, B, MB):
            kv, Mlen = encode_mb(b0, kv, profile=(rep == 1 and b0 == MB))

    def hook(t):
        if t == 30:
            torch.cuda.synchronize()
            cudart.cudaProfilerStart()
        if t == 32:
            torch.cuda.synchronize()
            cudart.cudaProfilerStop()

    eng.generate(None, B, Mlen, tm, ids["bos"], ids["pad"], 60, kv=kv)            # warm
    eng.generate(None, B, Mlen, tm, ids["bos"], ids["pad"], 60, kv=kv, step_hook=hook)
    torch.cuda.synchronize()
    print("slice done")


if __name__ == "__main__":
    main()
