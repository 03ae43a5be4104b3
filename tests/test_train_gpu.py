"""GPU: training-step parity.  Gradients from the kernel tape (ralf_b200/autograd.py)
vs torch.autograd on the CPU oracle for the same seeded weights/batch; optimizer vs torch.optim.AdamW.
Dropout is off on both sides (round-1 limit); BatchNorm uses batch statistics when the trunk trains."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests import helpers

pytestmark = pytest.mark.gpu
GRAD_RTOL = 2e-3


def _model(dev, seed):
    from ralf_b200 import generator as G

    m = G.RALF(features=None, tokenizer=helpers.make_tokenizer(), dataset_name="cgl", max_seq_length=10, top_k=16,
               auxilary_task="uncond")
    m.load_state_dict(helpers.synth_weights("ralf_cgl", seed), strict=True)
    return m.to(dev)


def test_adamw_clip_matches_torch(cuda_device):
    from ralf_b200 import autograd as ag

    g = torch.Generator().manual_seed(0)
    params = [("a.weight", torch.randn(300, 40, generator=g)), ("b.bias", torch.randn(77, generator=g))]
    ps = ag.ParamStore([(n, p.to(cuda_device)) for n, p in params], [["a.weight"], ["b.bias"]], cuda_device)
    ref = [torch.nn.Parameter(p.clone().to(cuda_device)) for _, p in params]
    opt = torch.optim.AdamW([{"params": [ref[0]], "weight_decay": 1e-2, "lr": 1e-3},
                             {"params": [ref[1]], "weight_decay": 0.0, "lr": 1e-3}], betas=(0.9, 0.999), eps=1e-8)
    for step in range(1, 4):
        grads = [torch.randn(p.shape, generator=g).to(cuda_device) * 3 for _, p in params]
        for (n, _), gr, r in zip(params, grads, ref):
            ps.g(n).copy_(gr)
            r.grad = gr.clone()
        norm = ag.grad_norm(ps.flat_g)
        tn = torch.nn.utils.clip_grad_norm_(ref, 0.5)
        assert abs(float(norm) - float(tn)) <= 1e-5 * float(tn)
        ag.adamw_step(ps, [(1e-3, 1e-2), (1e-3, 0.0)], step, 0.5, norm)
        opt.step()
        for (n, _), r in zip(params, ref):
            assert (ps.p(n) - r.data).abs().max().item() <= 1e-6


def _oracle_loss_and_grads(sd, batch, inputs, targets, pad_id, train_trunk):
    """torch.autograd over the CPU oracle; model.train() semantics for BatchNorm when the trunk trains, dropout off."""
    from oracle import ralf_oracle as O

    sd = {k: v.clone() for k, v in sd.items()}
    leaves = {}
    frozen = ("layout_encoer",) if train_trunk else ("encoder.extractor", "layout_encoer")
    for k, v in sd.items():
        if v.dtype == torch.float32 and not k.startswith(frozen) and not k.endswith(".pe") and "running_" not in k:
            v.requires_grad_(True)
            leaves[k] = v
    retrieved = {k: v.float() for k, v in batch["retrieved"].items()}
    O.BN_TRAIN = train_trunk
    try:
        mem = O.encode_ralf_memory(sd, inputs["image"], retrieved, inputs["seq_layout_const"],
                                   inputs["seq_layout_const_pad_mask"])
    finally:
        O.BN_TRAIN = False
    logits = O.decoder_logits(sd, inputs["seq"], mem, inputs["tgt_key_padding_mask"])
    loss = F.cross_entropy(logits.permute(0, 2, 1), targets["seq"], label_smoothing=0.1, ignore_index=pad_id)
    loss.backward()
    return float(loss), {k: v.grad for k, v in leaves.items() if v.grad is not None}


@pytest.mark.parametrize("train_trunk", [False, True])
def test_training_gradients_match_oracle(cuda_device, train_trunk):
    from oracle import synth
    from ralf_b200.train import TrainEngine

    torch.set_num_threads(8)
    model = _model(cuda_device, seed=21)
    sd = helpers.synth_weights("ralf_cgl", 21)
    batch = synth.synth_batch(2, 128, 128, 10, 16, 4, seed=9)
    inputs, targets = model.preprocess(batch)
    pad = model.tokenizer.name_to_id("pad")
    ref_loss, ref_grads = _oracle_loss_and_grads(sd, batch, inputs, targets, pad, train_trunk)
    te = TrainEngine(model, train_trunk=train_trunk)
    te.ps.flat_g.zero_()
    loss, tape, _ = te.forward_loss(inputs, targets)
    tape.backward()
    torch.cuda.synchronize()
    assert abs(float(loss) - ref_loss) <= 1e-4 * abs(ref_loss), (float(loss), ref_loss)
    worst = []
    for name, gref in ref_grads.items():
        assert name in te.ps.offsets, name
        g = te.ps.g(name).cpu()
        scale = gref.abs().max().item()
        err = (g - gref).abs().max().item()
        worst.append((err / (scale + 1e-12), name, scale))
    worst.sort(reverse=True)
    print("worst grads:", worst[:5])
    import json, os
    os.makedirs(os.path.join(helpers.ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(helpers.ROOT, "gpurun_out", f"train_grad_errors_trunk{int(train_trunk)}.json"), "w") as f:
        json.dump([[round(e, 6), n, s_] for e, n, s_ in worst], f, indent=0)
    bad = [w for w in worst if w[0] > GRAD_RTOL and w[2] > 1e-9]
    assert not bad, bad[:10]
    assert set(ref_grads) == set(te.ps.offsets), set(te.ps.offsets) ^ set(ref_grads)


def test_train_steps_reduce_loss_and_update_state_dict(cuda_device):
    from oracle import synth
    from ralf_b200.train import TrainEngine

    model = _model(cuda_device, seed=22)
    batch = synth.synth_batch(4, 128, 128, 10, 16, 4, seed=10)
    inputs, targets = model.preprocess(batch)
    te = TrainEngine(model, lr=1e-3, max_grad_norm=0.0)
    before = model.state_dict()["decoder.head.1.weight"].clone()
    losses = [float(te.train_step(inputs, targets)) for _ in range(6)]
    assert losses[-1] < losses[0], losses
    assert not torch.equal(before, model.state_dict()["decoder.head.1.weight"])
    assert float(te.last_grad_norm) > 0
