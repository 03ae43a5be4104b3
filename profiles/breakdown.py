"""Phase breakdown of one bench step (eager launches, CUDA events per phase).  Diagnostic only -- numbers
include per-launch host overhead; the bench value comes from bench.py (CUDA graphs)."""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    args = types.SimpleNamespace(elems=12, gallery=1_000_000, hw=256, batch=B, precision="bf16x3")
    dev = torch.device("cuda:0")
    from ralf_b200 import generator as G
    from ralf_b200 import ops

    retr, tok = bench.synth_world(args, 0, 1, dev)
    model = G.RALF(features=None, tokenizer=tok, dataset_name="cgl", max_seq_length=12, top_k=16, auxilary_task="uncond")
    model.load_state_dict(bench.synth_weights_for(model), strict=True)
    model.eval().to(dev)
    eng = model.engine()
    img = torch.rand(B, 4, 256, 256, device=dev)
    qry = torch.nn.functional.normalize(torch.randn(B, 512, device=dev), dim=1)
    const = model.preprocessor(G.ConditionalInputs(image=img))
    tm = tok.token_mask.to(dev).to(torch.uint8)
    ids = model.special_token_ids
    res = {}

    def phase(name, fn, reps=3):
        fn()
        torch.cuda.synchronize()
        n0 = ops.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        res[name] = {"ms": round(e0.elapsed_time(e1) / reps, 3), "kernels": (ops.launch_count() - n0) // reps}
        return out

    idx, _ = phase("knn_search", lambda: retr.search_local(qry, 16))
    retrieved = phase("fetch", lambda: retr.fetch(idx))
    tokens, h, w = phase("resnet_fpn", lambda: eng.resnet_fpn(img))
    phase("encode_image(resnet+6 enc layers)", lambda: eng.encode_image(img))
    phase("retrieved_features(fidnet)", lambda: eng.retrieved_features(retrieved["packed"], B))
    mem, mem_s = phase("encode(all)", lambda: eng.encode(img, retrieved["packed"], const["seq"], const["pad_mask"]))
    phase("cross_kv", lambda: eng.cross_kv(mem_s))
    phase("generate(60 steps)", lambda: eng.generate(mem_s, B, mem.shape[1], tm, ids["bos"], ids["pad"], 60), reps=2)
    print(json.dumps({"batch": B, "phases": res}, indent=1))


if __name__ == "__main__":
    main()
