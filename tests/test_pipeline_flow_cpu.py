"""CPU: control flow of the batch-level drivers (ralf_b200/pipeline.py) with the CUDA runtime objects replaced by
recording fakes.  No kernel runs and a "graph" replays nothing, so this checks what can be checked without a GPU: that
LayoutPipeline / OverlappedPipeline construct, capture and step without errors for 1 and several decode ways, that the
overlapped driver alternates its two slots, orders its streams through the events it is meant to (encode(i+1) waits for
decode(i-1); decode(i) waits for encode(i)), keeps a separate graph pool for the decode graphs and refuses a third batch
in flight.  Equality of results against the sequential pipeline is the GPU test's job (tests/test_pipeline_gpu.py)."""
import contextlib

import pytest
import torch

from tests import helpers

LOG = []


class FakeStream:
    n = 0

    def __init__(self, device=None):
        FakeStream.n += 1
        self.name = f"s{FakeStream.n}"

    def wait_stream(self, other):
        LOG.append(("wait_stream", self.name, other.name))

    def wait_event(self, ev):
        LOG.append(("wait_event", self.name, ev.name, ev.recorded_on))

    cuda_stream = 0


class FakeEvent:
    n = 0

    def __init__(self, enable_timing=False):
        FakeEvent.n += 1
        self.name, self.recorded_on = f"e{FakeEvent.n}", None

    def record(self, stream=None):
        self.recorded_on = (stream or CUR[-1]).name
        LOG.append(("record", self.name, self.recorded_on))

    def synchronize(self):
        LOG.append(("sync_event", self.name))

    def elapsed_time(self, other):
        return 1.0


class FakeGraph:
    n = 0

    def __init__(self):
        FakeGraph.n += 1
        self.name, self.pool_id = f"g{FakeGraph.n}", None

    def pool(self):
        return ("pool-of", self.name)

    def replay(self):
        LOG.append(("replay", self.name, CUR[-1].name))


MAIN = FakeStream()
MAIN.name = "main"
CUR = [MAIN]


@contextlib.contextmanager
def fake_stream_ctx(s):
    CUR.append(s)
    try:
        yield
    finally:
        CUR.pop()


@contextlib.contextmanager
def fake_graph_ctx(g, pool=None):
    g.pool_id = pool if pool is not None else g.pool()
    LOG.append(("capture", g.name, g.pool_id))
    yield


class FakeEngine:
    npass, dev = 3, torch.device("cpu")

    def encode(self, img, packed, seq, pad):
        b = img.shape[0]
        mem = torch.zeros(b, 7, 256)
        return mem, torch.zeros(2, b * 7, 256)

    def alloc_cross_kv(self, rows, kv24=False):
        return [torch.zeros(rows, 1536 if kv24 else 512, dtype=torch.uint8 if kv24 else torch.float32) for _ in range(6)]

    def cross_kv(self, mem_s, out=None, row0=0, kv24=False):
        LOG.append(("cross_kv", row0))
        return out

    def generate(self, mem_s, B, Mlen, token_mask, bos, pad, steps, kv=None, **kw):
        assert kv[0].shape[0] == B * Mlen
        LOG.append(("generate", B, CUR[-1].name))
        return torch.zeros(B, steps, dtype=torch.int64)


class FakeRetriever:
    world, rank, pg = 1, 0, None

    def search_local(self, q, k):
        return torch.zeros(q.shape[0], k, dtype=torch.int64), torch.zeros(q.shape[0], k)

    def fetch(self, idx):
        return {"packed": torch.zeros(idx.shape[0], idx.shape[1], 6, 10)}


@pytest.fixture
def fakes(monkeypatch):
    from ralf_b200 import generator as G
    from ralf_b200 import ops

    LOG.clear()
    monkeypatch.setattr(torch.cuda, "Stream", FakeStream)
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "CUDAGraph", FakeGraph)
    monkeypatch.setattr(torch.cuda, "graph", fake_graph_ctx)
    monkeypatch.setattr(torch.cuda, "stream", fake_stream_ctx)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: CUR[-1])
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "graph_pool_handle", lambda: ("pool", "decode"))
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    monkeypatch.setattr(ops, "launch_count", lambda: 0)
    model = G.RALF(features=None, tokenizer=helpers.make_tokenizer(), dataset_name="cgl", max_seq_length=10, pretrained=False)
    monkeypatch.setattr(model, "engine", lambda: FakeEngine())
    return model


@pytest.mark.parametrize("ways", [1, 2, 4])
def test_sequential_pipeline_flow(fakes, ways):
    from ralf_b200.pipeline import LayoutPipeline

    B = 6
    pipe = LayoutPipeline(fakes, FakeRetriever(), B, 32, 32, micro_batch=4, decode_ways=ways)
    captured = [e for e in LOG if e[0] == "capture"]
    assert len(captured) == 2 + 2 + 1  # search, fetch, two encoder micro-batches, decode
    LOG.clear()
    out = pipe.generate_layouts(torch.zeros(B, 4, 32, 32), torch.zeros(B, 512))
    assert out["seq"].shape == (B, 50) and out["retrieved_idx"].shape == (B, 16)
    assert [e[1] for e in LOG if e[0] == "replay"] == [g.name for g in [pipe.g_search, pipe.g_fetch, *pipe.g_enc, pipe.g_dec]]
    eager = LayoutPipeline(fakes, FakeRetriever(), B, 32, 32, micro_batch=4, decode_ways=ways, use_graph=False)
    LOG.clear()
    eager.generate_layouts(torch.zeros(B, 4, 32, 32), torch.zeros(B, 512))
    gens = [e for e in LOG if e[0] == "generate"]
    assert sum(e[1] for e in gens) == B and len(gens) == min(ways, 3 if ways == 4 else ways)  # 6 canvases: groups of 2 -> 3
    if ways > 1:
        assert len({e[2] for e in gens}) == len(gens) and "main" not in {e[2] for e in gens}  # one side stream per group
        forks = [e for e in LOG if e[0] == "wait_stream" and e[2] == "main"]
        joins = [e for e in LOG if e[0] == "wait_stream" and e[1] == "main"]
        assert {e[1] for e in forks} == {e[2] for e in joins} == {e[2] for e in gens}  # only forked branches are joined


def test_overlapped_pipeline_flow(fakes):
    from ralf_b200.pipeline import OverlappedPipeline

    B = 4
    pipe = OverlappedPipeline(fakes, FakeRetriever(), B, 32, 32, micro_batch=2)
    caps = {e[1]: e[2] for e in LOG if e[0] == "capture"}
    dec_pools = {caps[s.dec.name] for s in pipe.slots}
    enc_pools = {caps[g.name] for s in pipe.slots for g in s.enc}
    assert dec_pools == {("pool", "decode")} and dec_pools.isdisjoint(enc_pools)  # decode graphs never share scratch with encode
    assert len(pipe.slots) == 2 and pipe.slots[0].kv is not pipe.slots[1].kv
    img, q = torch.zeros(B, 4, 32, 32), torch.zeros(B, 512)
    LOG.clear()
    s0 = pipe.submit_host(img, q)
    s1 = pipe.submit_host(img, q)
    assert (s0, s1) == (0, 1)
    with pytest.raises(AssertionError):
        pipe.submit_host(img, q)  # a third batch before the first is collected
    r0 = pipe.collect(s0)
    assert r0["seq"].shape == (B, 50) and not pipe.slots[0].busy
    s2 = pipe.submit_host(img, q)
    assert s2 == 0
    pipe.collect(s1), pipe.collect(s2)
    pipe.drain()
    dec = pipe.dec_stream.name
    replays = [e for e in LOG if e[0] == "replay"]
    for slot in pipe.slots:
        assert all(e[2] == dec for e in replays if e[1] == slot.dec.name)           # decode graphs on the decode stream
        assert all(e[2] == "main" for e in replays if e[1] in {g.name for g in slot.enc})
    # per submit: main waits for the slot's previous decode BEFORE its encoder graphs; the decode stream waits for this encode
    first_enc = [i for i, e in enumerate(LOG) if e[0] == "replay" and e[1] == pipe.slots[0].enc[0].name]
    waits = [i for i, e in enumerate(LOG) if e[:3] == ("wait_event", "main", pipe.slots[0].dec_done.name)]
    assert len(first_enc) == 2 and waits[0] < first_enc[0] and any(first_enc[0] < w < first_enc[1] for w in waits)
    enc_waits = [e for e in LOG if e[0] == "wait_event" and e[1] == dec]
    assert len(enc_waits) == 3 and all(e[3] == "main" for e in enc_waits)              # enc_done recorded on the main stream
    # the blocking interface of the parent still works (submit + wait)
    out = pipe.generate_layouts(img, q)
    assert out["seq"].shape == (B, 50)


def test_knn_ways_flow_equals_single_stream(fakes, monkeypatch):
    """GpuRetriever.knn_ways > 1: passes of 128 queries round-robin over the side streams, all of them forked from and
    joined into the caller's stream, results concatenated in query order = the single-stream result (oracle as kernel)."""
    import numpy as np

    from ralf_b200 import ops
    from ralf_b200.retrieval import GpuRetriever
    from tests import oracle_knn

    def oracle_knn_topk(gallery, queries, k, *, index_base=0, gallery_max_norm=0.0, exact=False, workspace=None):
        LOG.append(("knn", queries.shape[0], CUR[-1].name))
        i, s = oracle_knn.topk(gallery.numpy(), queries.numpy(), k)
        return torch.from_numpy(i) + index_base, torch.from_numpy(s), torch.ones(queries.shape[0], dtype=torch.int32)

    monkeypatch.setattr(ops, "knn_topk", oracle_knn_topk)
    rng = np.random.default_rng(1)
    retr = GpuRetriever(torch.from_numpy(rng.standard_normal((3000, 64)).astype(np.float32)), device="cpu", index_base=100)
    q = torch.from_numpy(rng.standard_normal((300, 64)).astype(np.float32))
    want_i, want_s = retr.search_local(q, 16)
    retr.knn_ways = 2
    LOG.clear()
    got_i, got_s = retr.search_local(q, 16)
    assert torch.equal(got_i, want_i) and torch.equal(got_s, want_s) and int(retr.last_certified.sum()) == 300
    calls = [e for e in LOG if e[0] == "knn"]
    assert [e[1] for e in calls] == [128, 128, 44] and calls[0][2] == calls[2][2] != calls[1][2] and "main" not in {e[2] for e in calls}
    forks = {e[1] for e in LOG if e[0] == "wait_stream" and e[2] == "main"}
    joins = {e[2] for e in LOG if e[0] == "wait_stream" and e[1] == "main"}
    assert forks == joins == {s.name for s in retr._way_streams}
    retr.search_local(q[:100], 16)  # a single pass stays on the caller's stream
    assert LOG[-1] == ("knn", 100, "main")


def test_generate_graphed_cache_flow(fakes):
    """Engine.generate_graphed: one capture per (B, Mlen, steps) shape, later calls only refill the static K/V cache and
    replay; at most 4 shapes are kept."""
    from ralf_b200.engine import Engine

    class Self(FakeEngine):
        generate_graphed = Engine.generate_graphed

    eng = Self()
    tm = torch.ones(50, 519, dtype=torch.uint8)
    for _ in range(3):
        out = eng.generate_graphed(torch.zeros(2, 3 * 7, 256), 3, 7, tm, 517, 516, 50)
        assert out.shape == (3, 50)
    assert len([e for e in LOG if e[0] == "capture"]) == 1 and len([e for e in LOG if e[0] == "replay"]) == 3
    assert len([e for e in LOG if e[0] == "cross_kv"]) == 3  # filled in place before every replay
    for B in (1, 2, 4, 5):
        eng.generate_graphed(torch.zeros(2, B * 7, 256), B, 7, tm, 517, 516, 50)
    assert len(eng._gen_graphs) == 4 and (3, 7, 50, 517, 516) not in eng._gen_graphs  # the oldest shape was dropped
