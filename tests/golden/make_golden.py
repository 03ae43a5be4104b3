"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference classes
(/root/reference, CPU, fp32) on seeded synthetic weights and inputs.  Build-container only.

    python tests/golden/make_golden.py

Outputs (committed):
  schema_ralf_cgl.json / schema_autoreg_cgl.json   state-dict key -> shape/dtype of the reference classes
  ralf_cgl_256.npz      RALF (shipped class), B=2, 256x256 canvases, k=16, E=10  (BASELINE configs 2/3/5 shape)
  ralf_cgl_350x240.npz  RALF, B=1, real canvas size 350x240
  autoreg_cgl_350x240.npz  Autoreg baseline, B=1 (BASELINE config 1)
  tasks_cgl_256.npz     constrained tasks c / cwh / partial / refinement through the reference's get_condition,
                        task preprocessors, DECODE_SPACE_RESTRICTION and greedy sample() (same weights/batch as ralf_cgl_256)
  optim_groups_ralf_cgl.json   BaseModel.optim_groups as train.py calls it -> [(lr, weight_decay, [parameter names])]
  schema_ralf_pku.json / tokenizer_pku.npz   PKU (3 labels) variants of the state-dict schema and tokenizer outputs
  sampling_filters.npz  helpers/sampling.py on random logits: the post-filter probabilities handed to torch.multinomial
  relation_cgl_128.npz + relation_table_reference_pickle.pt   cond_type="relation" (Gen-R): see run_relation
  coarse_saliency.npz   the retrieval feature of retrieval_backbone="saliency" (models/retrieval/image.py:35-44)
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


def run_b32(model, tok, name="ralf_cgl_b32_128", B=32, H=128, W=128, seed=17, full=4):
    """SURVEY.md 8c asks for goldens at B in {1, 2, 32}: a 32-canvas batch through the UNMODIFIED reference (preprocess ->
    memory -> teacher-forced logits -> sample()).  To keep the fixture small, memory / logits are stored in full for the
    first `full` canvases only; for every canvas the greedy token ids, the decoded layout, and per (canvas, position)
    summaries of the logits (max, log-sum-exp, arg-max) and per-row norms of the memory are stored."""
    import copy

    torch.manual_seed(0)
    sd = synth.synth_state_dict(schema_of(model), seed=seed)
    model.load_state_dict(sd, strict=True)
    model.eval()
    batch = synth.synth_batch(B, H, W, 10, 16, tok.N_label, seed=seed)
    out = {}
    with torch.no_grad():
        inputs, targets = model.preprocess(copy.deepcopy(batch))
        out["seq_in"], out["targets"] = inputs["seq"].numpy(), targets["seq"].numpy()
        out["tgt_key_padding_mask"] = inputs["tgt_key_padding_mask"].numpy()
        out["seq_layout_const"] = inputs["seq_layout_const"].numpy()
        out["seq_layout_const_pad_mask"] = inputs["seq_layout_const_pad_mask"].numpy()
        mem = model._encode_into_memory(copy.deepcopy(inputs))["memory"]
        out["memory_head"] = mem[:full].numpy()
        out["memory_row_norm"] = mem.norm(dim=-1).numpy()
        outputs, losses = model.train_loss(copy.deepcopy(inputs), targets)
        lg = outputs["logits"]
        out["logits_head"] = lg[:full].numpy()
        out["logits_max"], out["logits_lse"] = lg.max(-1).values.numpy(), torch.logsumexp(lg, -1).numpy()
        out["logits_argmax"] = lg.argmax(-1).numpy()
        out["nll_loss"] = losses["nll_loss"].numpy()
        from image2layout.train.helpers.task import get_condition

        cond, _ = get_condition(copy.deepcopy(batch), "uncond", tok)
        res = model.sample(cond=cond, sampling_cfg=rb.DictConfig(name="deterministic"), cond_type="uncond",
                           return_violation=False)
        for k in ["label", "mask", "center_x", "center_y", "width", "height"]:
            out["gen_" + k] = res[k].numpy()
        # raw greedy token ids, replaying the loop of retrieval_augmented_autoreg.py:271-297 (as run() does)
        enc_in, _ = model._create_encoder_inputs(cond)
        enc_in["retrieved"] = {k: v.type_as(cond.image) for k, v in enc_in["retrieved"].items() if torch.is_tensor(v)}
        memory = model._encode_into_memory(enc_in)
        ids = model.special_token_ids
        inp = torch.full((B, 1), ids["bos"])
        for i in range(tok.max_token_length):
            lgi = model.decoder(tgt=inp, tgt_key_padding_mask=(inp == ids["pad"]), is_causal=True, **memory)[:, i].clone()
            lgi[:, ~tok.token_mask[i]] = -float("inf")
            inp = torch.cat([inp, lgi.argmax(dim=1, keepdim=True)], dim=1)
        out["gen_seq"] = inp[:, 1:].numpy()
        dec = tok.decode(inp[:, 1:])
        assert all(torch.equal(dec[k], res[k]) for k in ["label", "mask"]), "sample() and replay disagree"
    out["meta"] = np.array(json.dumps({"B": B, "H": H, "W": W, "seed": seed, "E": 10, "K": 16, "dataset": "cgl", "full": full,
                                       "special": {k: int(v) for k, v in ids.items()}}))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, {k: getattr(v, "shape", None) for k, v in out.items()})


TASKS = ["c", "cwh", "partial", "refinement"]


def run_tasks(tok, name, B, H, W, seed):
    """Constrained tasks on the shipped RALF class built with use_multitask=True, which makes sample()
    pick the task preprocessor from cond.task (retrieval_augmented_autoreg.py:745-750); same weights as `run`."""
    import copy

    from image2layout.train.helpers.task import get_condition
    from image2layout.train.models.layoutformerpp.decoding_space_restriction import DECODE_SPACE_RESTRICTION
    from image2layout.train.models.retrieval_augmented_autoreg import (
        ConcateAuxilaryTaskConcateCrossAttnRetrievalAugmentedAutoreg as RALF,
    )

    _, features = rb.make_tokenizer("cgl", 10)
    model = RALF(features=features, tokenizer=tok, dataset_name="cgl", max_seq_length=10, db_dataset=None,
                 retrieval_backbone="dreamsim", random_retrieval=False, top_k=16, saliency_k="None",
                 auxilary_task="uncond", use_multitask=True)
    torch.manual_seed(0)
    model.load_state_dict(synth.synth_state_dict(schema_of(model), seed=seed), strict=True)
    model.eval()
    batch = synth.synth_batch(B, H, W, 10, 16, tok.N_label, seed=seed)
    ids = model.special_token_ids
    out = {}
    for ti, task in enumerate(TASKS):
        rng_seed = 100 + ti
        with torch.no_grad():
            # (1) the reference's own sample(): decoded layout + violation
            torch.manual_seed(rng_seed)
            cond, _ = get_condition(copy.deepcopy(batch), task, tok)
            res, vio = model.sample(cond=cond, sampling_cfg=rb.DictConfig(name="deterministic"), cond_type=task,
                                    return_violation=True)
            # (2) replay with the same host RNG stream, recording every intermediate of the boundary
            torch.manual_seed(rng_seed)
            b2 = copy.deepcopy(batch)
            cond, b2 = get_condition(b2, task, tok)
            out[f"{task}_cond_seq"], out[f"{task}_cond_mask"] = cond.seq.clone().numpy(), cond.mask.clone().numpy()
            if task == "refinement":
                for k in ["center_x", "center_y", "width", "height"]:
                    out[f"{task}_noisy_{k}"] = b2[k].numpy()
            model.set_task_preprocessor(task)
            enc_in, const = model._create_encoder_inputs(cond)
            out[f"{task}_const_seq"], out[f"{task}_const_pad_mask"] = const["seq"].numpy(), const["pad_mask"].numpy()
            out[f"{task}_cond_seq_after"] = cond.seq.clone().numpy()  # parse_seq_into_vars rewrites <eos> in place
            enc_in["retrieved"] = {k: v.type_as(cond.image) for k, v in enc_in["retrieved"].items() if torch.is_tensor(v)}
            memory = model._encode_into_memory(enc_in)
            inp = torch.full((B, 1), ids["bos"])
            start = 0
            if task == "partial":
                inp = torch.cat([inp, cond.seq[:, 1:6]], dim=1)
                start = 5
            for i in range(start, tok.max_token_length):
                lg = model.decoder(tgt=inp, tgt_key_padding_mask=(inp == ids["pad"]), is_causal=True, **memory)[:, i].clone()
                lg[:, ~tok.token_mask[i]] = -float("inf")
                lg = DECODE_SPACE_RESTRICTION[task](i + 1, cond.seq, lg, pad_id=ids["pad"], eos_id=ids["eos"],
                                                    max_length=tok.max_token_length)
                inp = torch.cat([inp, lg.argmax(dim=1, keepdim=True)], dim=1)
            out[f"{task}_gen_seq"] = inp[:, 1:].numpy()
            dec = tok.decode(inp[:, 1:])
            for k in ["label", "mask", "center_x", "center_y", "width", "height"]:
                assert torch.equal(dec[k], res[k]), f"{task}: sample() and replay disagree on {k}"
                out[f"{task}_gen_{k}"] = res[k].numpy()
            out[f"{task}_violation"] = np.array([vio["total"], vio["viorated"]])
    out["meta"] = np.array(json.dumps({"B": B, "H": H, "W": W, "seed": seed, "tasks": TASKS,
                                       "rng_seed": {t: 100 + i for i, t in enumerate(TASKS)}}))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, {k: getattr(v, "shape", None) for k, v in out.items()})


def run_sampling_filters():
    """helpers/sampling.py:18-68 on random masked logits: capture what the reference hands to torch.multinomial."""
    from image2layout.train.helpers import sampling as S

    g = torch.Generator().manual_seed(5)
    N, V = 24, 519
    logits = torch.randn((N, V), generator=g) * 3.0
    logits[:, 4:] += torch.where(torch.rand((N, V - 4), generator=g) < 0.5, -float("inf"), 0.0)  # masked vocabulary
    logits[3] = -float("inf")
    logits[3, 7] = 0.25  # a forced token: single finite entry
    logits[5, 10:14] = 2.0  # ties
    cfgs = [dict(name="random", temperature=1.0), dict(name="random", temperature=0.7),
            dict(name="top_k", top_k=5, temperature=1.0), dict(name="top_k", top_k=1, temperature=1.3),
            dict(name="top_k", top_k=40, temperature=0.5), dict(name="top_p", top_p=0.9, temperature=1.0),
            dict(name="top_p", top_p=0.5, temperature=2.0), dict(name="top_p", top_p=1.0, temperature=1.0)]
    out = {"logits": logits.numpy(), "cfgs": np.array(json.dumps(cfgs))}
    captured = []
    real = torch.multinomial

    def spy(probs, num_samples, *a, **k):
        captured.append(probs.clone())
        return real(probs, num_samples, *a, **k)

    torch.multinomial = spy
    try:
        for i, c in enumerate(cfgs):
            S.sample(logits.clone(), rb.DictConfig(**c))
            out[f"probs_{i}"] = captured[-1].numpy()
    finally:
        torch.multinomial = real
    np.savez_compressed(os.path.join(OUT, "sampling_filters.npz"), **out)
    print("sampling_filters", len(cfgs))


def run_optim_groups(model):
    """BaseModel.optim_groups as train.py:217-223 calls it -> [(lr, weight_decay, [parameter names])]."""
    names = {id(p): n for n, p in model.named_parameters()}
    groups = model.optim_groups(base_lr=1e-4, weight_decay=1e-4, custom_lr={"encoder.extractor.body": 1e-5})
    out = [{"lr": g["lr"], "weight_decay": g["weight_decay"], "params": [names[id(p)] for p in g["params"]]} for g in groups]
    with open(os.path.join(OUT, "optim_groups_ralf_cgl.json"), "w") as f:
        json.dump(out, f)
    print("optim_groups", [(g["lr"], g["weight_decay"], len(g["params"])) for g in out])


def run_tokenizer_edge_cases():
    """Reference LayoutSequenceTokenizer on empty / full layouts, out-of-range and bin-edge geometry, and its decode of
    arbitrary (also invalid) token sequences."""
    tok, _ = rb.make_tokenizer("cgl", 10)
    g = torch.Generator().manual_seed(0)
    B, E = 64, 10
    n = torch.randint(0, E + 1, (B,), generator=g)
    n[0], n[1], n[2] = 0, E, 1
    mask = torch.arange(E)[None] < n[:, None]
    lay = {"label": torch.randint(0, 4, (B, E), generator=g) * mask, "mask": mask}
    for k in ["center_x", "center_y", "width", "height"]:
        v = torch.rand((B, E), generator=g) * 1.4 - 0.2
        v[3, :], v[4, :], v[5, :] = 0.0, 1.0, torch.arange(E) / 128.0
        lay[k] = v * mask
    enc = tok.encode({k: v.clone() for k, v in lay.items()})
    seqs = torch.randint(0, tok.N_total, (B, 50), generator=g)
    seqs[0, :], seqs[1, :], seqs[2, 7] = tok.name_to_id("eos"), tok.name_to_id("pad"), tok.name_to_id("bos")
    dec = tok.decode(seqs.clone())
    keys = ["label", "mask", "center_x", "center_y", "width", "height"]
    np.savez_compressed(os.path.join(OUT, "tokenizer_edge_cases.npz"), **{f"in_{k}": v.numpy() for k, v in lay.items()},
                        seq=enc["seq"].numpy(), mask=enc["mask"].numpy(), dec_in=seqs.numpy(),
                        **{f"dec_{k}": dec[k].numpy() for k in keys})
    print("tokenizer edge cases", enc["seq"].shape)


def edge_case_batch(seed, B=6):
    """Random layouts; every fifth seed forces a single-element layout, a full layout and a layout with one label."""
    batch = synth.synth_batch(B, 8, 8, 10, 1, 4, seed=1000 + seed)
    if seed % 5 == 0:
        batch["mask"][0] = torch.arange(10) < 1
        batch["mask"][1] = True
        batch["label"][2] = 1
        for k in ["label", "center_x", "center_y", "width", "height"]:
            batch[k] = batch[k] * batch["mask"]
    return batch


def run_task_edge_cases():
    """get_condition + task preprocessors of the reference on the batches of `edge_case_batch` (seeds 0, 5, 17)."""
    import copy

    from image2layout.train.helpers.task import get_condition
    from image2layout.train.models.layoutformerpp.task_preprocessor import PREPROCESSOR

    tok, _ = rb.make_tokenizer("cgl", 10)
    out = {}
    for seed in (0, 5, 17):
        batch = edge_case_batch(seed)
        for task in TASKS:
            torch.manual_seed(seed)
            cond, _ = get_condition(copy.deepcopy(batch), task, tok)
            out[f"{seed}_{task}_cond_seq"], out[f"{seed}_{task}_cond_mask"] = cond.seq.clone().numpy(), cond.mask.numpy()
            const = PREPROCESSOR[task](tokenizer=tok, global_task_embedding=False)(cond)
            out[f"{seed}_{task}_const_seq"] = const["seq"].numpy()
    np.savez_compressed(os.path.join(OUT, "tasks_edge_cases.npz"), **out)
    print("task edge cases", len(out))


def corrupted_output(tok, batch, cond_seq, task, seed):
    """A generated sequence that reproduces the condition except for ~15 % corrupted label / geometry tokens."""
    pad, eos = tok.name_to_id("pad"), tok.name_to_id("eos")
    keys = ["label", "mask", "center_x", "center_y", "width", "height"]
    gt = cond_seq[:, 1:].clone() if task == "refinement" else tok.encode({k: batch[k] for k in keys})["seq"][:, 1:]
    out = gt.clone()
    out[out == pad] = eos
    g = torch.Generator().manual_seed(seed)
    flip = torch.rand(out.shape, generator=g) < 0.15
    valid = (gt != pad) & (gt != eos)
    noise = torch.randint(0, 4, out.shape, generator=g)
    lab = (torch.arange(out.shape[1])[None] % 5 == 0).expand_as(out)
    out = torch.where(flip & valid & lab, noise, out)
    return torch.where(flip & valid & ~lab, torch.clamp(out + 1, max=515), out)


def run_violation_cases():
    """calculate_violation of the reference (violate.py:24-139) on corrupted outputs, tasks c / cwh / refinement."""
    import copy

    from image2layout.train.helpers.task import get_condition
    from image2layout.train.models.layoutformerpp.task_preprocessor import PREPROCESSOR
    from image2layout.train.models.layoutformerpp.violate import calculate_violation

    tok, _ = rb.make_tokenizer("cgl", 10)
    out = {}
    for seed in (1, 3, 7):
        batch = synth.synth_batch(5, 8, 8, 10, 1, 4, seed=2000 + seed)
        for task in ("c", "cwh", "refinement"):
            torch.manual_seed(seed)
            cond, _ = get_condition(copy.deepcopy(batch), task, tok)
            PREPROCESSOR[task](tokenizer=tok, global_task_embedding=False)(cond)
            seq = corrupted_output(tok, batch, cond.seq, task, seed)
            vio = calculate_violation(task, cond, seq.clone(), tok.decode(seq.clone()), tok, [])
            out[f"{seed}_{task}_seq"] = seq.numpy()
            out[f"{seed}_{task}_violation"] = np.array([vio["total"], vio["viorated"]])
    np.savez_compressed(os.path.join(OUT, "violation_cases.npz"), **out)
    print("violation cases", {k: v.tolist() for k, v in out.items() if k.endswith("violation")})


def collate_examples(trial):
    """Ragged dataset rows (0, 1, 3 or 10 elements) as `dataset[i]` yields them (lists + image tensors)."""
    import random

    rnd = random.Random(trial)
    exs = []
    for b in range(rnd.randint(1, 6)):
        n = rnd.choice([0, 1, 3, 10])
        g = torch.Generator().manual_seed(trial * 100 + b)
        exs.append({"id": str(trial * 10 + b), "label": [rnd.randint(0, 3) for _ in range(n)],
                    **{k: [rnd.random() for _ in range(n)] for k in ["center_x", "center_y", "width", "height"]},
                    "image": torch.rand((3, 4, 4), generator=g), "saliency": torch.rand((1, 4, 4), generator=g)})
    return exs


def run_collate_cases():
    """The reference's collate_fn (data.py:42-117) on ragged rows.  data.py cannot be imported here (hydra config store), so
    the function object is built from its source in the reference tree with the module globals it uses."""
    import ast
    import copy
    from typing import Optional

    from torch.utils.data import default_collate

    from image2layout.train.global_variables import DUMMY_LAYOUT, RETRIEVED_KEYS

    src = open(os.path.join(rb.REFERENCE_ROOT, "image2layout/train/data.py")).read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "collate_fn"][0]
    glb = {"torch": torch, "Tensor": torch.Tensor, "default_collate": default_collate, "Optional": Optional,
           "compute_validity": None, "RETRIEVED_KEYS": RETRIEVED_KEYS, "DUMMY_LAYOUT": DUMMY_LAYOUT}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "reference_collate_fn", "exec"), glb)
    out = {}
    for trial in (0, 1, 2, 3):
        ref = glb["collate_fn"](copy.deepcopy(collate_examples(trial)), max_seq_length=10)
        for k in ["label", "mask", "center_x", "center_y", "width", "height", "image", "saliency"]:
            out[f"{trial}_{k}"] = ref[k].numpy()
        out[f"{trial}_id"] = np.array(list(ref["id"]))
    np.savez_compressed(os.path.join(OUT, "collate_cases.npz"), **out)
    print("collate cases", len(out))


def run_pku_contract():
    """BASELINE configs[2] (PKU, 3 labels): state-dict schema of the reference class and the tokenizer's outputs."""
    ralf, tok, _ = rb.make_ralf("pku")
    with open(os.path.join(OUT, "schema_ralf_pku.json"), "w") as f:
        json.dump(schema_of(ralf), f)
    b = synth.synth_batch(3, 8, 8, 10, 1, tok.N_label, seed=21)
    enc = tok.encode({k: b[k] for k in ["label", "mask", "center_x", "center_y", "width", "height"]})
    np.savez_compressed(os.path.join(OUT, "tokenizer_pku.npz"), seq=enc["seq"].numpy(), mask=enc["mask"].numpy(),
                        token_mask=tok.token_mask.numpy(),
                        special=np.array([tok.name_to_id("pad"), tok.name_to_id("bos"), tok.name_to_id("eos")]))
    print("pku contract", tok.N_total)
    # a small PKU forward golden: memory + greedy token ids of the reference class on 2 canvases of 128x128
    import copy

    from image2layout.train.helpers.task import get_condition

    sd = synth.synth_state_dict(schema_of(ralf), seed=4)
    ralf.load_state_dict(sd, strict=True)
    ralf.eval()
    B, H, W = 2, 128, 128
    batch = synth.synth_batch(B, H, W, 10, 16, tok.N_label, seed=6)
    with torch.no_grad():
        cond, _ = get_condition(copy.deepcopy(batch), "uncond", tok)
        enc_in, _ = ralf._create_encoder_inputs(cond)
        enc_in["retrieved"] = {k: v.type_as(cond.image) for k, v in enc_in["retrieved"].items() if torch.is_tensor(v)}
        mem = ralf._encode_into_memory(enc_in)["memory"]
        ids = ralf.special_token_ids
        inp = torch.full((B, 1), ids["bos"])
        for i in range(tok.max_token_length):
            lg = ralf.decoder(tgt=inp, tgt_key_padding_mask=(inp == ids["pad"]), is_causal=True, memory=mem)[:, i].clone()
            lg[:, ~tok.token_mask[i]] = -float("inf")
            inp = torch.cat([inp, lg.argmax(dim=1, keepdim=True)], dim=1)
    np.savez_compressed(os.path.join(OUT, "ralf_pku_128.npz"), memory=mem.numpy(), gen_seq=inp[:, 1:].numpy(),
                        seq_layout_const=enc_in["seq_layout_const"].numpy(),
                        seq_layout_const_pad_mask=enc_in["seq_layout_const_pad_mask"].numpy(),
                        meta=np.array(json.dumps({"B": B, "H": H, "W": W, "seed": 6, "weights_seed": 4, "E": 10, "K": 16,
                                                  "dataset": "pku", "special": {k: int(v) for k, v in ids.items()}})))


# ---- relation (Gen-R) ------------------------------------------------------------------------------------------------
REL_SEEDS = {"deterministic": 400, "random": 401}


def _reference_table_builder():
    """describe_relationships / generate_unique_labels of preprocess/precompute_relationship.py:31-131, executed
    unmodified.  The module itself cannot be imported here (hydra / dataset plumbing at import time), so the two function
    definitions are compiled from its source; the drawing calls (cv2 / seaborn) only paint a scratch canvas."""
    import ast

    import cv2
    from image2layout.train.helpers import relationships as R
    from image2layout.train.helpers.util import convert_xywh_to_ltrb

    path = os.path.join(rb.REFERENCE_ROOT, "image2layout/preprocess/precompute_relationship.py")
    tree = ast.parse(open(path).read())
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("generate_unique_labels", "describe_relationships")]
    sns = type("sns", (), {"color_palette": staticmethod(lambda name, n: [(0.1, 0.2, 0.3)] * n)})
    from typing import Any

    ns = {"cv2": cv2, "np": np, "sns": sns, "torch": torch, "Tensor": torch.Tensor, "Any": Any, "PAD_ELEMENT": "pad",
          "RelElement": R.RelElement, "convert_xywh_to_ltrb": convert_xywh_to_ltrb,
          "detect_loc_relation_between_element_and_canvas": R.detect_loc_relation_between_element_and_canvas,
          "detect_loc_relation_between_elements": R.detect_loc_relation_between_elements,
          "detect_size_relation": R.detect_size_relation,
          # the script's own name lists are alphabetical; the fixture uses the tokenizer's order so that names round-trip
          "ELEMENTS": {"cgl": rb.LABELS["cgl"], "pku": rb.LABELS["pku"]}}
    exec(compile(ast.Module(keep, []), path, "exec"), ns)
    return ns["describe_relationships"]


def relation_batch(B, H, W, seed):
    """Small canvases (2-5 elements), elements sorted by label like the reference's sort_label transform; canvas 0 carries a
    single label so that the label shuffle of the constraint sequence cannot disagree with the decoding order."""
    batch = synth.synth_batch(B, H, W, 10, 16, 4, seed=seed)
    g = torch.Generator().manual_seed(seed)
    n = torch.randint(2, 6, (B,), generator=g)
    mask = torch.arange(10)[None] < n[:, None]
    label = torch.where(mask, batch["label"], torch.full_like(batch["label"], 99))
    label[0] = torch.where(mask[0], torch.ones_like(label[0]), label[0])
    order = torch.argsort(label, dim=1, stable=True)
    batch["mask"] = mask
    batch["label"] = torch.where(mask, torch.gather(label, 1, order), torch.zeros_like(label))
    for k in ["center_x", "center_y", "width", "height"]:
        v = torch.gather(batch[k], 1, order)
        if k in ("width", "height"):
            v = v * 0.5 + 0.05
        batch[k] = torch.where(mask, v, torch.zeros_like(v))
    return batch


def run_relation(name="relation_cgl_128", B=3, H=128, W=128, seed=5):
    """cond_type="relation": relationship table, compute_relation, RelationshipPreprocessor, prepare(), every per-step
    relation mask + back index the backtracking sampler asked for, final tokens, decoded layouts, violation counts --
    all from the UNMODIFIED reference (retrieval_augmented_autoreg.py:335-507 and the modules it calls), under recorded
    seeds of `random` and torch."""
    import copy
    import random

    from image2layout.train.helpers import relationships as R
    from image2layout.train.helpers.task import get_condition
    from image2layout.train.models.layoutformerpp import relation_restriction as RR
    from image2layout.train.models.retrieval_augmented_autoreg import (
        ConcateAuxilaryTaskConcateCrossAttnRetrievalAugmentedAutoreg as RALF,
    )

    import seaborn

    if not hasattr(seaborn, "color_palette"):  # the stub module: violate.py only uses it to paint a scratch canvas
        seaborn.color_palette = lambda name, n: [(0.1, 0.2, 0.3)] * n
    tok, features = rb.make_tokenizer("cgl", 10)
    batch = relation_batch(B, H, W, seed)
    out = {k: batch[k].numpy() for k in ["label", "mask", "center_x", "center_y", "width", "height"]}
    # (a) the table, written where the reference constructor looks for it
    table = _reference_table_builder()(batch, "cgl")
    os.makedirs("cache", exist_ok=True)
    torch.save(table, "cache/pku_cgl_relationships_dic_using_canvas_sort_label_lexico.pt")
    torch.save(table, os.path.join(OUT, "relation_table_reference_pickle.pt"))
    random.seed(300)
    # torch >= 2.6 loads with weights_only=True unless told otherwise; the reference (torch 2.0) expects plain pickle
    torch.serialization.add_safe_globals([R.RelElement, R.RelLoc, R.RelSize])
    model = RALF(features=features, tokenizer=tok, dataset_name="cgl", max_seq_length=10, db_dataset=None,
                 retrieval_backbone="dreamsim", random_retrieval=False, top_k=16, saliency_k="None",
                 auxilary_task="relation")
    model.load_state_dict(synth.synth_state_dict(schema_of(model), seed=seed), strict=True)
    model.eval()
    pre = model.preprocessor
    for key, rows in table.items():
        out[f"table_{key}"] = np.array([[pre.name_to_id(x) for x in r] for r in rows], dtype=np.int64).reshape(-1, 5)
    for key, rows in pre.table.items():  # after the constructor's shuffle
        out[f"table_shuffled_{key}"] = np.array([[pre.name_to_id(x) for x in r] for r in rows], dtype=np.int64).reshape(-1, 5)

    def encode_constraints(cons):
        rows = []
        for e, mine in enumerate(cons):
            for kind, tgt in mine:
                rows.append([e, -1, int(tgt)] if kind == "canvas" else [e, int(kind), int(tgt)])
        return np.array(rows, dtype=np.int64).reshape(-1, 3)

    ids = model.special_token_ids
    for mode, rng_seed in REL_SEEDS.items():
        cfg = rb.DictConfig(name=mode, temperature=1.0, top_k=5, top_p=0.9)
        random.seed(rng_seed)
        torch.manual_seed(rng_seed)
        cond, _ = get_condition(copy.deepcopy(batch), "relation", tok)
        if mode == "deterministic":
            out["cond_seq"], out["cond_mask"] = cond.seq.clone().numpy(), cond.mask.clone().numpy()
            out["edge_indexes"], out["edge_attributes"] = cond.edge_indexes.numpy(), cond.edge_attributes.numpy()
        log = {"calls": [], "const": None, "memory": None, "prepared": []}
        enc0, mem0, call0, prep0 = model._create_encoder_inputs, model._encode_into_memory, \
            RR.TransformerSortByDictRelationConstraint.__call__, RR.TransformerSortByDictRelationConstraint.prepare

        def enc_spy(c):
            r = enc0(c)
            log["const"] = (r[1]["seq"].clone(), r[1]["pad_mask"].clone())
            return r

        def mem_spy(x):
            r = mem0(x)
            log["memory"] = r["memory"].clone()
            return r

        def prep_spy(self, seq):
            r = prep0(self, seq)
            log["prepared"].append(copy.deepcopy(r))
            return r

        def call_spy(self, token_ids, rel_constraints):
            m, back = call0(self, token_ids, rel_constraints)
            log["calls"].append((len(log["prepared"]) - 1, token_ids[0].tolist(), m.clone(), -1 if back is None else int(back)))
            return m, back

        model._create_encoder_inputs, model._encode_into_memory = enc_spy, mem_spy
        RR.TransformerSortByDictRelationConstraint.__call__ = call_spy
        RR.TransformerSortByDictRelationConstraint.prepare = prep_spy
        try:
            with torch.no_grad():
                res, vio = model.sample(cond=cond, sampling_cfg=cfg, cond_type="relation", return_violation=True,
                                        use_backtrack=True)
        finally:
            model._create_encoder_inputs, model._encode_into_memory = enc0, mem0
            RR.TransformerSortByDictRelationConstraint.__call__ = call0
            RR.TransformerSortByDictRelationConstraint.prepare = prep0
        p = f"bt_{mode}_"
        out[p + "const_seq"], out[p + "const_pad_mask"] = log["const"][0].numpy(), log["const"][1].numpy()
        out[p + "cond_seq_after"] = cond.seq.clone().numpy()
        if mode == "deterministic":
            out["memory"] = log["memory"].numpy()  # the constraint sequence differs per mode (shuffles), so does the memory
        else:
            out[p + "memory"] = log["memory"].numpy()
        for b, cons in enumerate(log["prepared"]):
            out[p + f"prepared_{b}"] = encode_constraints(cons)
        calls = log["calls"]
        out[p + "call_sample"] = np.array([c[0] for c in calls], dtype=np.int64)
        out[p + "call_len"] = np.array([len(c[1]) for c in calls], dtype=np.int64)
        pref = np.full((len(calls), tok.max_token_length + 2), -1, dtype=np.int64)
        for i, c in enumerate(calls):
            pref[i, :len(c[1])] = c[1]
        out[p + "call_prefix"] = pref
        out[p + "call_mask"] = np.packbits(torch.stack([c[2] for c in calls]).numpy(), axis=1)
        out[p + "call_back"] = np.array([c[3] for c in calls], dtype=np.int64)
        for k in ["label", "mask", "center_x", "center_y", "width", "height"]:
            out[p + f"gen_{k}"] = res[k].numpy()
        out[p + "violation"] = np.array([vio["total"], vio["viorated"]])
        print(name, mode, "constraint calls:", len(calls), "violation:", vio)
    # (b') every relation kind: all table rows as constraints (relation size 100 %), random geometry prefixes walked forward
    random.seed(420)
    torch.manual_seed(420)
    cond, _ = get_condition(copy.deepcopy(batch), "relation", tok)
    pre.set_relation_size(100)
    const = pre(cond)
    pre.set_relation_size(10)
    out["sweep_const_seq"] = const["seq"].numpy()
    fn = RR.TransformerSortByDictRelationConstraint(pre)
    g = torch.Generator().manual_seed(421)
    sweep_pref, sweep_mask, sweep_back, sweep_sample = [], [], [], []
    for b in range(B):
        cons = fn.prepare(const["seq"][b])
        out[f"sweep_prepared_{b}"] = encode_constraints(cons)
        types = fn.type_constraint_token_id.tolist()
        for trial in range(12):
            fn.prepare(const["seq"][b])  # fresh decode state
            toks = [ids["bos"]]
            scale = [128, 64, 24][trial % 3]  # large, medium and small boxes
            for e, lab in enumerate(types):
                w, h = torch.randint(0, scale, (2,), generator=g).tolist()
                cx, cy = torch.randint(0, 128, (2,), generator=g).tolist()
                elem = [lab, fn.width_start_idx + w, fn.height_start_idx + h, fn.center_x_start_idx + cx,
                        fn.center_y_start_idx + cy]
                for t in elem:
                    m, back = fn(torch.tensor([toks]), cons)
                    sweep_pref.append(list(toks))
                    sweep_mask.append(m.clone())
                    sweep_back.append(-1 if back is None else int(back))
                    sweep_sample.append(b)
                    toks.append(t)
            m, back = fn(torch.tensor([toks]), cons)  # all elements complete
            sweep_pref.append(list(toks))
            sweep_mask.append(m.clone())
            sweep_back.append(-1 if back is None else int(back))
            sweep_sample.append(b)
    # crafted constraint lists: every relation kind against every earlier element, several per element
    kinds = [R.RelSize.SMALLER, R.RelSize.EQUAL, R.RelSize.LARGER, R.RelSize.UNKNOWN, R.RelLoc.LEFT, R.RelLoc.TOP,
             R.RelLoc.RIGHT, R.RelLoc.BOTTOM, R.RelLoc.CENTER, R.RelLoc.UNKNOWN]
    canvas_kinds = [R.RelLoc.TOP, R.RelLoc.CENTER, R.RelLoc.BOTTOM]
    n_craft = 40
    for trial in range(n_craft):
        b = trial % B
        fn.prepare(const["seq"][b])
        types = fn.type_constraint_token_id.tolist()
        cons = [[] for _ in types]
        for e in range(len(types)):
            for _ in range(int(torch.randint(0, 4, (1,), generator=g))):
                if e == 0 or int(torch.randint(0, 4, (1,), generator=g)) == 0:
                    cons[e].append(("canvas", canvas_kinds[int(torch.randint(0, 3, (1,), generator=g))]))
                else:
                    cons[e].append((kinds[int(torch.randint(0, len(kinds), (1,), generator=g))],
                                    int(torch.randint(0, e, (1,), generator=g))))
        out[f"craft_prepared_{trial}"] = encode_constraints(cons)
        toks = [ids["bos"]]
        scale = [128, 48, 16][trial % 3]
        for e, lab in enumerate(types):
            w, h = torch.randint(0, scale, (2,), generator=g).tolist()
            cx, cy = torch.randint(0, 128, (2,), generator=g).tolist()
            for t in [lab, fn.width_start_idx + w, fn.height_start_idx + h, fn.center_x_start_idx + cx,
                      fn.center_y_start_idx + cy]:
                m, back = fn(torch.tensor([toks]), cons)
                sweep_pref.append(list(toks))
                sweep_mask.append(m.clone())
                sweep_back.append(-1 if back is None else int(back))
                sweep_sample.append(B + trial)  # B + trial: crafted constraint list `trial`
                toks.append(t)
    out["craft_count"] = np.array(n_craft)
    # live cross-check (build container only, nothing stored): the product's RelationConstraint.mask against the reference
    # on many more crafted constraint lists / prefixes than the fixture carries
    from ralf_b200 import relation as PR
    from ralf_b200.tokenizer import LayoutSequenceTokenizer as HostTok

    host_tok = HostTok(rb.LABELS["cgl"], 10)
    random.seed(300)
    mine_pre = PR.RelationPreprocessor(host_tok, {k: [[x if isinstance(x, str) else getattr(PR, type(x).__name__)(int(x))
                                                         for x in r] for r in rows] for k, rows in table.items()})
    mine = PR.RelationConstraint(mine_pre)
    to_mine = lambda cons: [[(PR.CANVAS, PR.RelLoc(int(t))) if k == "canvas" else
                             ((PR.RelSize if int(k) < 4 else PR.RelLoc)(int(k)), int(t)) for k, t in c] for c in cons]
    live = 0
    for trial in range(600):
        b = trial % B
        fn.prepare(const["seq"][b])
        mine.types = fn.type_constraint_token_id.clone()
        types = fn.type_constraint_token_id.tolist()
        cons = [[] for _ in types]
        for e in range(len(types)):
            for _ in range(int(torch.randint(0, 5, (1,), generator=g))):
                if e == 0 or int(torch.randint(0, 4, (1,), generator=g)) == 0:
                    cons[e].append(("canvas", canvas_kinds[int(torch.randint(0, 3, (1,), generator=g))]))
                else:
                    cons[e].append((kinds[int(torch.randint(0, len(kinds), (1,), generator=g))],
                                    int(torch.randint(0, e, (1,), generator=g))))
        mc = to_mine(cons)
        toks = [ids["bos"]]
        scale = [128, 64, 32, 8][trial % 4]
        for e, lab in enumerate(types):
            w, h = torch.randint(0, scale, (2,), generator=g).tolist()
            cx, cy = torch.randint(0, 128, (2,), generator=g).tolist()
            for t in [lab, fn.width_start_idx + w, fn.height_start_idx + h, fn.center_x_start_idx + cx,
                      fn.center_y_start_idx + cy]:
                m, back = fn(torch.tensor([toks]), cons)
                m2, back2 = mine.mask(toks, mc)
                assert torch.equal(m, m2) and back == back2, (trial, toks, cons)
                live += 1
                toks.append(t)
    out["live_checked_calls"] = np.array(live)
    print(name, "live cross-check of RelationConstraint.mask against the reference:", live, "calls, no mismatch")
    pref = np.full((len(sweep_pref), tok.max_token_length + 2), -1, dtype=np.int64)
    for i, c in enumerate(sweep_pref):
        pref[i, :len(c)] = c
    out["sweep_prefix"], out["sweep_len"] = pref, np.array([len(c) for c in sweep_pref], dtype=np.int64)
    out["sweep_mask"] = np.packbits(torch.stack(sweep_mask).numpy(), axis=1)
    out["sweep_back"], out["sweep_sample"] = np.array(sweep_back, dtype=np.int64), np.array(sweep_sample, dtype=np.int64)
    print(name, "sweep calls:", len(sweep_pref), "relation kinds:",
          sorted({int(r[1]) for b in range(B) for r in out[f"sweep_prepared_{b}"]}))
    # (b'') helpers/sampling.py:18-68 on single rows, as the backtracking loop calls it (temperature override included)
    from image2layout.train.helpers.sampling import sample as ref_sample

    gl = torch.Generator().manual_seed(430)
    rows = torch.randn(24, tok.N_total, generator=gl) * 2.0
    rows[:, ::3] = -float("inf")
    out["draw_logits"] = rows.numpy()
    for mode in ["deterministic", "random", "top_k", "top_p", "gumbel"]:
        cfg = rb.DictConfig(name=mode, temperature=0.8, top_k=5, top_p=0.9)
        torch.manual_seed(431)
        out[f"draw_{mode}"] = np.array([int(ref_sample(rows[i:i + 1].clone(), cfg, temperature=1.5 if i % 2 else None))
                                        for i in range(rows.size(0))], dtype=np.int64)
    # (b-live) cross-check of the whole backtracking sampler on further seeds (nothing stored): product host code
    # (get_condition -> RelationPreprocessor -> prepare -> sample_with_backtracking, oracle decoder on the reference's memory)
    # against the reference's sample_relation under the same `random` / torch seeds
    from oracle import ralf_oracle as O
    from ralf_b200 import task as PT

    sd_live = synth.synth_state_dict(schema_of(model), seed=seed)
    pairs = 0
    for live_seed in range(500, 504):
        for mode in ("deterministic", "random"):
            cfg = rb.DictConfig(name=mode, temperature=1.0, top_k=5, top_p=0.9)
            random.seed(live_seed)
            torch.manual_seed(live_seed)
            cond, _ = get_condition(copy.deepcopy(batch), "relation", tok)
            grabbed = {}
            mem0 = model._encode_into_memory

            def mem_spy2(x):
                r = mem0(x)
                grabbed["memory"] = r["memory"].clone()
                return r

            model._encode_into_memory = mem_spy2
            try:
                with torch.no_grad():
                    res, vio = model.sample(cond=cond, sampling_cfg=cfg, cond_type="relation", return_violation=True,
                                            use_backtrack=True)
            finally:
                model._encode_into_memory = mem0
            random.seed(live_seed)
            torch.manual_seed(live_seed)
            cond2, _ = PT.get_condition(copy.deepcopy(batch), "relation", host_tok)
            const2 = mine_pre(cond2)
            forced = PT.forced_token_table("relation", cond2.seq, ids["pad"], ids["eos"], tok.max_token_length)
            rows2, prepared2 = [], []
            for b in range(B):
                cons2 = mine.prepare(const2["seq"][b])
                mem_b = grabbed["memory"][b:b + 1]

                def logits_of(prefix, mem_b=mem_b):
                    tgt = torch.tensor([prefix])
                    with torch.no_grad():
                        return O.decoder_logits(sd_live, tgt, mem_b, tgt == ids["pad"])[0, -1]

                rows2.append(PR.sample_with_backtracking(logits_of, mine, cons2, forced[b], bos_id=ids["bos"], eos_id=ids["eos"],
                                                         max_token_length=tok.max_token_length, sampling_cfg=dict(cfg)))
                prepared2.append(cons2)
            out2 = host_tok.decode(PR.pad_like_reference(rows2, tok.max_token_length))
            for k in ["label", "mask", "center_x", "center_y", "width", "height"]:
                assert torch.equal(out2[k], res[k]), (live_seed, mode, k)
            assert PR.violation_count(out2, prepared2) == vio, (live_seed, mode)
            pairs += 1
    out["live_checked_sampler_runs"] = np.array(pairs)
    print(name, "live cross-check of the backtracking sampler against the reference:", pairs, "seed x mode runs, no mismatch")
    # (c) without backtracking: batched decode under the label restriction only (:218-325)
    random.seed(410)
    torch.manual_seed(410)
    cond, _ = get_condition(copy.deepcopy(batch), "relation", tok)
    with torch.no_grad():
        res, vio = model.sample(cond=cond, sampling_cfg=rb.DictConfig(name="deterministic", temperature=1.0),
                                cond_type="relation", return_violation=True, use_backtrack=False)
    for k in ["label", "mask", "center_x", "center_y", "width", "height"]:
        out[f"nobt_gen_{k}"] = res[k].numpy()
    out["nobt_violation"] = np.array([vio["total"], vio["viorated"]])
    out["meta"] = np.array(json.dumps({"B": B, "H": H, "W": W, "seed": seed, "ctor_seed": 300, "rng_seed": REL_SEEDS,
                                       "nobt_seed": 410, "special": {k: int(v) for k, v in ids.items()},
                                       "label_names": rb.LABELS["cgl"]}))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, {k: getattr(v, "shape", None) for k, v in out.items() if not k.startswith("table")})


def run_coarse_saliency():
    """retrieval_backbone="saliency": the query / gallery feature is the 16x16 nearest-neighbour thumbnail of the saliency
    map mapped to [-1, 1] (models/retrieval/image.py:35-44).  The module needs dreamsim at import time, so the function
    definition alone is compiled from its source, unmodified."""
    import ast

    import torch.nn.functional as F
    from einops import rearrange

    path = os.path.join(rb.REFERENCE_ROOT, "image2layout/train/models/retrieval/image.py")
    tree = ast.parse(open(path).read())
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "coarse_saliency"]
    ns = {"F": F, "torch": torch, "rearrange": rearrange, "np": np, "Tensor": torch.Tensor}
    exec(compile(ast.Module(keep, []), path, "exec"), ns)
    g = torch.Generator().manual_seed(12)
    # some values outside [0, 1] (the clamp matters); rounded to fp16 first so the fixture stores the exact inputs compactly
    sal = (torch.rand(2, 1, 350, 240, generator=g) * 1.4 - 0.2).half().float()
    out = np.stack([ns["coarse_saliency"](sal[i]) for i in range(sal.size(0))])
    np.savez_compressed(os.path.join(OUT, "coarse_saliency.npz"), saliency=sal.numpy().astype(np.float16), feature=out)
    print("coarse_saliency", out.shape)


def run_max_length():
    """Maximum size the reference can be built with: max_seq_length = 11 (one RelElement letter A..K per element,
    task_preprocessor.py:64-67) -> 55 tokens, every canvas and every exemplar FULL (11 elements).  Needs a FIDNet checkpoint
    with max_bbox = 11 (fid/model.py:131-147 loads it strictly), written into a scratch dir of its own."""
    import copy

    from image2layout.train.fid.model import FIDNetV3
    from image2layout.train.helpers.task import get_condition

    work = "/tmp/ralf_ref_work11"
    rb.bootstrap(work)
    torch.manual_seed(0)
    torch.save({"state_dict": FIDNetV3(num_label=4, max_bbox=11).state_dict()},
               os.path.join(work, "tmp/fidnet/cgl/model_best.pth.tar"))
    ralf, tok, _ = rb.make_ralf("cgl", 11)
    with open(os.path.join(OUT, "schema_ralf_cgl_e11.json"), "w") as f:
        json.dump(schema_of(ralf), f)
    sd = synth.synth_state_dict(schema_of(ralf), seed=8)
    ralf.load_state_dict(sd, strict=True)
    ralf.eval()
    B, H, W, E = 2, 128, 128, 11
    batch = synth.synth_batch(B, H, W, E, 16, tok.N_label, seed=14)
    g = torch.Generator().manual_seed(15)
    for tgt in (batch, batch["retrieved"]):  # fill every slot
        tgt["mask"] = torch.ones_like(tgt["mask"])
        tgt["label"] = torch.randint(0, tok.N_label, tgt["label"].shape, generator=g)
        for k in ["center_x", "center_y", "width", "height"]:
            tgt[k] = torch.rand(tgt[k].shape, generator=g)
    out = {f"batch_{k}": batch[k].numpy() for k in ["label", "mask", "center_x", "center_y", "width", "height"]}
    out.update({f"retrieved_{k}": batch["retrieved"][k].numpy() for k in ["label", "mask", "center_x", "center_y", "width", "height"]})
    with torch.no_grad():
        inputs, targets = ralf.preprocess(copy.deepcopy(batch))
        out["seq_in"], out["targets"] = inputs["seq"].numpy(), targets["seq"].numpy()
        logits = ralf.train_loss(copy.deepcopy(inputs), targets)[0]["logits"]
        out["logits"] = logits.numpy()
        cond, _ = get_condition(copy.deepcopy(batch), "uncond", tok)
        enc_in, _ = ralf._create_encoder_inputs(cond)
        enc_in["retrieved"] = {k: v.type_as(cond.image) for k, v in enc_in["retrieved"].items() if torch.is_tensor(v)}
        mem = ralf._encode_into_memory(enc_in)["memory"]
        ids = ralf.special_token_ids
        inp = torch.full((B, 1), ids["bos"])
        for i in range(tok.max_token_length):
            lg = ralf.decoder(tgt=inp, tgt_key_padding_mask=(inp == ids["pad"]), is_causal=True, memory=mem)[:, i].clone()
            lg[:, ~tok.token_mask[i]] = -float("inf")
            inp = torch.cat([inp, lg.argmax(dim=1, keepdim=True)], dim=1)
    np.savez_compressed(os.path.join(OUT, "ralf_cgl_e11_128.npz"), memory=mem.numpy(), gen_seq=inp[:, 1:].numpy(),
                        seq_layout_const=enc_in["seq_layout_const"].numpy(),
                        seq_layout_const_pad_mask=enc_in["seq_layout_const_pad_mask"].numpy(), **out,
                        meta=np.array(json.dumps({"B": B, "H": H, "W": W, "seed": 14, "weights_seed": 8, "E": E, "K": 16,
                                                  "dataset": "cgl", "special": {k: int(v) for k, v in ids.items()}})))
    print("ralf_cgl_e11_128", mem.shape, inp.shape)
    rb.bootstrap("/tmp/ralf_ref_work")  # back to the scratch dir whose FIDNet checkpoint has max_bbox = 10


def run_init_stats():
    """Freshly constructed reference models (before any weights are loaded over them): per state-dict key the mean / std /
    min / max / numel of the initial values, for both classes.  Pins the initial distributions a from-scratch training run
    starts from (xavier for the transformers, N(0, 0.02) for the decoder embedding / head, PyTorch defaults elsewhere;
    ResNet50 + FIDNet come from the checkpoint files the constructors read)."""
    out = {}
    for name, make in [("ralf_cgl", lambda: rb.make_ralf("cgl")), ("autoreg_cgl", lambda: rb.make_autoreg("cgl"))]:
        torch.manual_seed(1234)
        model = make()[0]
        stats = {}
        for k, v in model.state_dict().items():
            if not v.is_floating_point() or k.endswith(".pe"):
                continue
            v = v.double()
            stats[k] = [float(v.mean()), float(v.std()) if v.numel() > 1 else 0.0, float(v.min()), float(v.max()), v.numel()]
        out[name] = stats
    with open(os.path.join(OUT, "init_stats.json"), "w") as f:
        json.dump(out, f)
    print("init_stats", {k: len(v) for k, v in out.items()})


def run_reference_gradients(ralf, tok):
    """Backward of the reference's own train_loss on the golden batch (eval mode: dropout off, BatchNorm on its running
    statistics -- the deterministic setting): per-parameter gradient L2 norms, the small gradient tensors in full.  Pins the
    oracle's backward (torch.autograd over oracle/ralf_oracle.py, which the GPU gradient tests compare against) to the
    reference itself."""
    import copy

    sd = synth.synth_state_dict(schema_of(ralf), seed=1)
    ralf.load_state_dict(sd, strict=True)
    ralf.eval()
    batch = synth.synth_batch(2, 256, 256, 10, 16, tok.N_label, seed=1)  # the ralf_cgl_256 golden batch
    inputs, targets = ralf.preprocess(copy.deepcopy(batch))
    ralf.zero_grad()
    _, losses = ralf.train_loss(inputs, targets)
    losses["nll_loss"].backward()
    out = {"nll_loss": losses["nll_loss"].detach().numpy()}
    names, norms = [], []
    for n, p in ralf.named_parameters():
        if p.grad is None:
            continue
        names.append(n)
        norms.append(float(p.grad.double().norm()))
        if p.numel() <= 1024:
            out["grad__" + n] = p.grad.numpy()
    out["names"], out["norms"] = np.array(names), np.array(norms)
    np.savez_compressed(os.path.join(OUT, "grads_ralf_cgl_256.npz"), **out)
    print("grads_ralf_cgl_256", len(names), "parameters with gradients, total norm", float(np.sqrt((np.array(norms) ** 2).sum())))


def main():
    rb.bootstrap("/tmp/ralf_ref_work")
    torch.backends.mha.set_fastpath_enabled(False)
    torch.set_num_threads(8)
    ralf, tok, _ = rb.make_ralf("cgl")
    with open(os.path.join(OUT, "schema_ralf_cgl.json"), "w") as f:
        json.dump(schema_of(ralf), f)
    if "--relation-only" in sys.argv:
        run_relation()
        return
    if "--grads-only" in sys.argv:
        run_reference_gradients(ralf, tok)
        return
    if "--init-stats-only" in sys.argv:
        run_init_stats()
        return
    if "--max-length-only" in sys.argv:
        run_max_length()
        return
    if "--b32-only" in sys.argv:
        run_b32(ralf, tok)
        return
    if "--saliency-only" in sys.argv:
        run_coarse_saliency()
        return
    if "--tasks-only" in sys.argv:
        run_tasks(tok, "tasks_cgl_256", B=2, H=256, W=256, seed=1)
        run_sampling_filters()
        run_optim_groups(ralf)
        run_pku_contract()
        run_tokenizer_edge_cases()
        run_task_edge_cases()
        run_violation_cases()
        run_collate_cases()
        run_relation()
        run_coarse_saliency()
        run_max_length()
        run_init_stats()
        run_reference_gradients(rb.make_ralf("cgl")[0], tok)
        return
    run(ralf, tok, "ralf_cgl_256", B=2, H=256, W=256, seed=1, is_ralf=True)
    run_tasks(tok, "tasks_cgl_256", B=2, H=256, W=256, seed=1)
    run_sampling_filters()
    run(ralf, tok, "ralf_cgl_350x240", B=1, H=350, W=240, seed=2, is_ralf=True)
    ar, tok2, _ = rb.make_autoreg("cgl")
    with open(os.path.join(OUT, "schema_autoreg_cgl.json"), "w") as f:
        json.dump(schema_of(ar), f)
    run(ar, tok2, "autoreg_cgl_350x240", B=1, H=350, W=240, seed=3, is_ralf=False)
    run_b32(rb.make_ralf("cgl")[0], tok)


if __name__ == "__main__":
    main()
