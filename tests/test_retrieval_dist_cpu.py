"""CPU, world_size 2, gloo: the sharded-retrieval exchange (ralf_b200.retrieval.exchange_and_merge + shard_bounds)
reproduces the unsharded top-k.  Per-shard search and the merge rule are CPU stand-ins built on the C oracle
(test infrastructure); the protocol code is the product's."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import oracle_knn


def _merge_cpu(all_s, all_i):
    """Same rule as ralf_knn_merge: (score desc, index asc), missing = -1."""
    W, Q, k = all_s.shape
    out_i = torch.full((Q, k), -1, dtype=torch.int64)
    out_s = torch.full((Q, k), float("-inf"))
    for q in range(Q):
        cand = [(-float(all_s[w, q, e]), int(all_i[w, q, e])) for w in range(W) for e in range(k) if all_i[w, q, e] >= 0]
        cand.sort()
        for r, (ns, i) in enumerate(cand[:k]):
            out_i[q, r], out_s[q, r] = i, -ns
    return out_i, out_s


def _worker(rank, world, port, G, Q, k, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ralf_b200.retrieval import exchange_and_merge, shard_bounds

    lo, hi = shard_bounds(G.shape[0], world, rank)
    li, ls = oracle_knn.topk(G[lo:hi], Q, k, index_base=lo, threads=2)
    idx, score = exchange_and_merge(torch.from_numpy(li), torch.from_numpy(ls), world, None, _merge_cpu)
    if rank == 0:
        ret["idx"], ret["score"] = idx.numpy(), score.numpy()
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_exchange_matches_unsharded():
    rng = np.random.default_rng(4)
    G = rng.standard_normal((3001, 64)).astype(np.float32)
    Q = rng.standard_normal((5, 64)).astype(np.float32)
    k = 16
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, G, Q, k, ret), nprocs=2, join=True)
    oi, os_ = oracle_knn.topk(G, Q, k)
    np.testing.assert_array_equal(ret["idx"], oi)
    np.testing.assert_array_equal(ret["score"].view(np.uint32), os_.view(np.uint32))


def test_shard_bounds_cover_exactly():
    from ralf_b200.retrieval import shard_bounds

    for n in (1, 7, 1000, 1_000_003):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))


class _StubModel(torch.nn.Module):
    """Host-side stand-in with the surface checkpoint.evaluate() uses (preprocess / train_loss / device / eval)."""

    def __init__(self):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(1))

    @property
    def device(self):
        return self.w.device

    def preprocess(self, batch):
        return {"x": batch["x"]}, {"y": batch["x"]}

    def train_loss(self, inputs, targets, test=False):
        assert test and not torch.is_grad_enabled()
        return {}, {"nll_loss": inputs["x"].float().mean()}


def _eval_worker(rank, world, port, values, ret):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ralf_b200.checkpoint import evaluate

    batches = [{"x": torch.tensor([v])} for v in values]
    out = evaluate(_StubModel(), batches, rank=rank, world_size=world)
    ret[rank] = out["nll_loss"]
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_evaluate_returns_global_mean_on_every_rank():
    """the validation set is sharded over ranks (the reference evaluates all of it on every rank) and the
    mean is all-reduced; world_size 2 over gloo, an odd number of batches so the shards are uneven."""
    values = [1.0, 2.0, 4.0, 8.0, 16.0]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_eval_worker, args=(2, port, values, ret), nprocs=2, join=True)
    want = sum(values) / len(values)
    assert abs(ret[0] - want) < 1e-6 and abs(ret[1] - want) < 1e-6
