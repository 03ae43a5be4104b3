"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference classes
(/root/reference, CPU, fp32) on seeded synthetic weights and inputs.  Build-container only.

    python tests/golden/make_golden.py

Outputs (committed):
  schema_ralf_cgl.json / schema_autoreg_cgl.json   state-dict key -> shape/dtype of the reference classes
  ralf_cgl_256.npz      RALF (shipped class), B=2, 256x256 canvases, k=16, E=10  (BASELINE configs 2/3/5 shape)
  ralf_cgl_350x240.npz  RALF, B=1, real canvas size 350x240
  autoreg_cgl_350x240.npz  Autoreg baseline, B=1 (BASELINE config 1)
Each npz: tokenizer outputs (seq, mask, token_mask), constraint sequence, encoder memory, teacher-forced
logits, nll loss, greedy token ids + per-step masked logits, decoded layout.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_bootstrap as rb  # noqa: E402
from oracle import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def schema_of(model):
    return {k: {"shape": list(v.shape), "dtype": str(v.dtype).replace("torch.", "")} for k, v in model.state_dict().items()}


def run(model, tok, name, B, H, W, seed, is_ralf):
    torch.manual_seed(0)
    sd = synth.synth_state_dict(schema_of(model), seed=seed)
    model.load_state_dict(sd, strict=True)
    model.eval()
    batch = synth.synth_batch(B, H, W, 10, 16, tok.N_label, seed=seed)
    if not is_ralf:
        batch.pop("retrieved")
    out = {}
    import copy

    with torch.no_grad():
        inputs, targets = model.preprocess(copy.deepcopy(batch))
        out["seq_in"] = inputs["seq"].numpy()
        out["tgt_key_padding_mask"] = inputs["tgt_key_padding_mask"].numpy()
        out["targets"] = targets["seq"].numpy()
        out["seq_layout_const"] = inputs["seq_layout_const"].numpy()
        out["seq_layout_const_pad_mask"] = inputs["seq_layout_const_pad_mask"].numpy()
        out["token_mask"] = tok.token_mask.numpy()
        enc = tok.encode({k: batch[k] for k in ["label", "mask", "center_x", "center_y", "width", "height"]})
        out["tok_seq"], out["tok_mask"] = enc["seq"].numpy(), enc["mask"].numpy()
        mem = model._encode_into_memory(copy.deepcopy(inputs))["memory"]
        out["memory"] = mem.numpy()
        outputs, losses = model.train_loss(copy.deepcopy(inputs), targets)
        out["logits"] = outputs["logits"].numpy()
        out["nll_loss"] = losses["nll_loss"].numpy()
        # greedy generation through the reference's own sample()
        from image2layout.train.helpers.task import get_condition

        cond, _ = get_condition(copy.deepcopy(batch), "uncond", tok)
        res = model.sample(cond=cond, sampling_cfg=rb.DictConfig(name="deterministic"), cond_type="uncond",
                           return_violation=False)
        for k in ["label", "mask", "center_x", "center_y", "width", "height"]:
            out["gen_" + k] = res[k].numpy()
        # token ids + per-step masked logits, replaying the loop of retrieval_augmented_autoreg.py:271-297
        enc_in, _ = model._create_encoder_inputs(cond)
        if is_ralf:
            enc_in["retrieved"] = {k: v.type_as(cond.image) for k, v in enc_in["retrieved"].items() if torch.is_tensor(v)}
        memory = model._encode_into_memory(enc_in)
        ids = model.special_token_ids
        inp = torch.full((B, 1), ids["bos"])
        steps = []
        for i in range(tok.max_token_length):
            lg = model.decoder(tgt=inp, tgt_key_padding_mask=(inp == ids["pad"]), is_causal=True, **memory)[:, i].clone()
            lg[:, ~tok.token_mask[i]] = -float("inf")
            steps.append(lg)
            inp = torch.cat([inp, lg.argmax(dim=1, keepdim=True)], dim=1)
        out["gen_seq"] = inp[:, 1:].numpy()
        out["gen_step_logits"] = torch.stack(steps, 1).numpy()
        dec = tok.decode(inp[:, 1:])
        assert all(torch.equal(dec[k], res[k]) for k in ["label", "mask"]), "sample() and replay disagree"
    out["meta"] = np.array(json.dumps({"B": B, "H": H, "W": W, "seed": seed, "E": 10, "K": 16, "dataset": "cgl",
                                       "special": {k: int(v) for k, v in ids.items()}}))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, {k: getattr(v, "shape", None) for k, v in out.items()})


def main():
    rb.bootstrap("/tmp/ralf_ref_work")
    torch.backends.mha.set_fastpath_enabled(False)
    torch.set_num_threads(8)
    ralf, tok, _ = rb.make_ralf("cgl")
    with open(os.path.join(OUT, "schema_ralf_cgl.json"), "w") as f:
        json.dump(schema_of(ralf), f)
    run(ralf, tok, "ralf_cgl_256", B=2, H=256, W=256, seed=1, is_ralf=True)
    run(ralf, tok, "ralf_cgl_350x240", B=1, H=350, W=240, seed=2, is_ralf=True)
    ar, tok2, _ = rb.make_autoreg("cgl")
    with open(os.path.join(OUT, "schema_autoreg_cgl.json"), "w") as f:
        json.dump(schema_of(ar), f)
    run(ar, tok2, "autoreg_cgl_350x240", B=1, H=350, W=240, seed=3, is_ralf=False)


if __name__ == "__main__":
    main()
