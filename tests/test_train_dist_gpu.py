"""GPU, world_size 2: the data-parallel training step -- bucketed gradient all-reduce
overlapped with the trunk backward, clip on the global norm, AdamW -- against a single-process restatement of the same
two micro-batches (local gradients of each rank computed one after the other, averaged by hand, one optimiser step).

Two processes are spawned.  With two or more GPUs visible each rank takes its own GPU and the backend is NCCL (what the
product uses; `bench.py`'s training leg covers N = 2 / 4 / 8 on the driver's scaling run); on a one-GPU box both ranks
share cuda:0 and the exchange goes through gloo's CUDA-tensor all-reduce -- NCCL refuses two ranks on one device -- so the
protocol (bucket ranges, marker order, stream join, averaging, replicas staying identical) is still exercised on hardware.
A sum of two operands is commutative, so every comparison is bit-exact."""
import os
import socket
import tempfile

import pytest
import torch

from tests import helpers

pytestmark = pytest.mark.gpu
STEPS = 2


def _model(dev, seed):
    from ralf_b200 import generator as G

    m = G.RALF(features=None, tokenizer=helpers.make_tokenizer(), dataset_name="cgl", max_seq_length=10, top_k=16,
               auxilary_task="uncond")
    m.load_state_dict(helpers.synth_weights("ralf_cgl", seed), strict=True)
    return m.to(dev)


def _batch(model, rank):
    from oracle import synth

    return model.preprocess(synth.synth_batch(2, 128, 128, 10, 16, 4, seed=40 + rank))


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist

    from ralf_b200.train import TrainEngine

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    multi = torch.cuda.device_count() >= world
    dev = torch.device("cuda", rank if multi else 0)
    torch.cuda.set_device(dev)
    if multi:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    res = {}
    for name, overlap in (("overlap", True), ("blocking", False)):
        model = _model(dev, seed=31)
        inputs, targets = _batch(model, rank)
        te = TrainEngine(model, lr=1e-3, world_size=world, rank=rank, seed=3)
        te.overlap_comm = overlap
        assert set(te._buckets) == {"rest", "layer4", "layer3", "layer2", "tail"}
        covered = sorted(r for rs in te._buckets.values() for r in rs)
        assert covered[0][0] == 0 and covered[-1][1] == te.ps.total
        assert all(a[1] == b[0] for a, b in zip(covered, covered[1:])), "buckets must tile the flat gradient buffer"
        losses = [float(te.train_step(inputs, targets)) for _ in range(STEPS)]
        torch.cuda.synchronize()
        res[name] = {"losses": losses, "p": te.ps.flat_p.cpu(), "g": te.ps.flat_g.cpu(), "norm": float(te.last_grad_norm)}
    assert torch.equal(res["overlap"]["p"], res["blocking"]["p"]), "bucketed / overlapped all-reduce changed the result"
    res["backend"] = "nccl" if multi else "gloo"
    torch.save(res, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_step_equals_hand_averaged_gradients(cuda_device):
    import torch.multiprocessing as mp

    from ralf_b200 import autograd as ag
    from ralf_b200.train import TrainEngine

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(2, port, d), nprocs=2, join=True)
        got = [torch.load(os.path.join(d, f"rank{r}.pt"), weights_only=False) for r in range(2)]
    assert torch.equal(got[0]["overlap"]["p"], got[1]["overlap"]["p"]), "replicas diverged"
    assert torch.equal(got[0]["overlap"]["g"], got[1]["overlap"]["g"])
    # single-process restatement: rank r's engine (same per-rank dropout stream) computes its local gradient
    tes, data = [], []
    for r in range(2):
        m = _model(cuda_device, seed=31)
        data.append(_batch(m, r))
        tes.append(TrainEngine(m, lr=1e-3, world_size=1, rank=r, seed=3))
    master = tes[0]
    for step in range(STEPS):
        grads, losses = [], []
        for r, te in enumerate(tes):
            if r:  # replicas share the weights
                te.ps.flat_p.copy_(master.ps.flat_p)
                te.refresh_operands()
            te._set_step_scalars(None)
            te.ps.flat_g.zero_()
            loss, tape, _ = te.forward_loss(*data[r])
            tape.backward()
            grads.append(te.ps.flat_g.clone())
            losses.append(float(loss))
        master.ps.flat_g.copy_((grads[0] + grads[1]) * 0.5)
        norm = ag.grad_norm(master.ps.flat_g)
        ag.adamw_step(master.ps, master.group_cfg, 0, master.max_norm, norm, dyn=master._dyn)
        master.refresh_operands()
        for r in range(2):
            assert got[r]["overlap"]["losses"][step] == losses[r], (step, r, got[r]["overlap"]["losses"], losses)
    torch.cuda.synchronize()
    assert torch.equal(got[0]["overlap"]["g"], master.ps.flat_g.cpu())
    assert torch.equal(got[0]["overlap"]["p"], master.ps.flat_p.cpu())
    assert got[0]["overlap"]["norm"] == float(norm)
