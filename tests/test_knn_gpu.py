"""GPU parity of the k-NN kernels against the C oracle: bit-exact indices and scores."""
import numpy as np
import pytest
import torch

from tests import oracle_knn

pytestmark = pytest.mark.gpu


def _data(n, d, q, seed, normalize=True):
    rng = np.random.default_rng(seed)
    G = rng.standard_normal((n, d)).astype(np.float32)
    Q = rng.standard_normal((q, d)).astype(np.float32)
    if normalize:
        G /= np.linalg.norm(G, axis=1, keepdims=True)
        Q /= np.linalg.norm(Q, axis=1, keepdims=True)
    return G, Q


@pytest.mark.parametrize("n,d,q,k", [
    (10000, 512, 1, 16), (10000, 512, 32, 16), (7734, 512, 5, 33), (48544, 512, 128, 16), (3000, 256, 200, 16),
    (1000, 100, 3, 8), (257, 512, 2, 16), (20, 512, 2, 16), (5, 64, 1, 16),
])
def test_knn_topk_matches_oracle(cuda_device, n, d, q, k):
    from ralf_b200 import ops

    G, Q = _data(n, d, q, seed=n + q)
    oi, os_ = oracle_knn.topk(G, Q, k, index_base=1000)
    gi, gs, cert = ops.knn_topk(torch.from_numpy(G).to(cuda_device), torch.from_numpy(Q).to(cuda_device), k,
                                index_base=1000, gallery_max_norm=1.0001)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(gi.cpu().numpy(), oi)
    np.testing.assert_array_equal(gs.cpu().numpy().view(np.uint32), os_.view(np.uint32))
    assert cert.cpu().numpy().all()


def test_knn_exact_kernel_matches_oracle(cuda_device):
    from ralf_b200 import ops

    G, Q = _data(20000, 512, 3, seed=3, normalize=False)
    oi, os_ = oracle_knn.topk(G, Q, 16)
    gi, gs, _ = ops.knn_topk(torch.from_numpy(G).to(cuda_device), torch.from_numpy(Q).to(cuda_device), 16, exact=True)
    np.testing.assert_array_equal(gi.cpu().numpy(), oi)
    np.testing.assert_array_equal(gs.cpu().numpy().view(np.uint32), os_.view(np.uint32))


def test_knn_ties_and_duplicates(cuda_device):
    """Exact duplicates in the gallery: ties must resolve to the lower index (oracle rule)."""
    from ralf_b200 import ops

    G, Q = _data(4096, 512, 4, seed=11)
    G[100:140] = G[7]          # 41 identical rows
    Q[0] = G[7]                # query hits the duplicate cluster
    oi, os_ = oracle_knn.topk(G, Q, 16)
    gi, gs, cert = ops.knn_topk(torch.from_numpy(G).to(cuda_device), torch.from_numpy(Q).to(cuda_device), 16,
                                gallery_max_norm=1.0001, fixup=False)
    np.testing.assert_array_equal(gi.cpu().numpy(), oi)
    # 41 equal scores straddle the candidate cut, so the TF32 bound cannot certify query 0 ...
    assert cert.cpu().numpy()[0] == 0 and cert.cpu().numpy()[1:].all()
    # ... and the exact kernel is the certified fallback
    ei, es, _ = ops.knn_topk(torch.from_numpy(G).to(cuda_device), torch.from_numpy(Q).to(cuda_device), 16, exact=True)
    np.testing.assert_array_equal(ei.cpu().numpy(), oi)


def _near_duplicate_cluster(n, q, seed, width=200, eps=2e-4):
    """A gallery with `width` near-copies of one row (perturbations far below the TF32 resolution 2^-11, far above
    fp32's) and queries aimed at it: TF32 scores of the cluster are indistinguishable, so which 32 of them reach the
    candidate list is arbitrary, while the exact ranking inside the cluster is well defined."""
    rng = np.random.default_rng(seed)
    G, Q = _data(n, 512, q, seed)
    rows = rng.choice(n, width, replace=False)
    G[rows] = G[rows[0]] + eps * rng.standard_normal((width, 512)).astype(np.float32) / np.sqrt(512)
    Q[0] = G[rows[0]]
    if q > 2:
        Q[q // 2] = G[rows[1]] / np.linalg.norm(G[rows[1]])
    return G, Q, rows


def test_knn_uncertified_queries_are_fixed_on_the_device(cuda_device):
    """ralf_knn_fixup_exact: near-duplicate cluster wider than the candidate list.  Without the fix-up phase 1/2 returns a
    wrong top-16 for the cluster queries (and says so: certified == 0); with it (the default, what LayoutPipeline and
    GpuRetriever use) the result equals the oracle bit for bit, certified == 2 for exactly those queries, no host sync
    (the call is captured into a CUDA graph and replayed)."""
    from ralf_b200 import ops

    G, Q, _ = _near_duplicate_cluster(20000, 6, seed=13)
    oi, os_ = oracle_knn.topk(G, Q, 16)
    Gd, Qd = torch.from_numpy(G).to(cuda_device), torch.from_numpy(Q).to(cuda_device)
    ri, rs, rc = ops.knn_topk(Gd, Qd, 16, gallery_max_norm=1.01, fixup=False)
    rc = rc.cpu().numpy()
    assert rc[0] == 0 and rc[3] == 0, "the cluster queries must be flagged"
    bad = [j for j in range(6) if not np.array_equal(ri[j].cpu().numpy(), oi[j])]
    assert set(bad) <= {0, 3} and all(rc[j] == 0 for j in bad), "a wrong row that is not flagged would be a certificate bug"
    gi, gs, gc = ops.knn_topk(Gd, Qd, 16, gallery_max_norm=1.01)
    np.testing.assert_array_equal(gi.cpu().numpy(), oi)
    np.testing.assert_array_equal(gs.cpu().numpy().view(np.uint32), os_.view(np.uint32))
    gc = gc.cpu().numpy()
    assert gc[0] == 2 and gc[3] == 2 and (gc[[1, 2, 4, 5]] == 1).all()
    # capturable: no host round trip anywhere in the call
    ws = torch.empty(ops._lib.lib().ralf_knn_workspace_bytes(20000, 512, 6, 16), dtype=torch.uint8, device=cuda_device)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ops.knn_topk(Gd, Qd, 16, gallery_max_norm=1.01, workspace=ws)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        ci, cs, cc = ops.knn_topk(Gd, Qd, 16, gallery_max_norm=1.01, workspace=ws)
    ci.zero_()
    graph.replay()
    torch.cuda.synchronize()
    np.testing.assert_array_equal(ci.cpu().numpy(), oi)
    np.testing.assert_array_equal(cs.cpu().numpy().view(np.uint32), os_.view(np.uint32))


def test_knn_fixup_with_more_than_1024_queries(cuda_device):
    """The uncertified-query list is built in chunks of 1024 (knn_collect_uncertified_kernel) and the exact scan / re-rank walk
    it with grid-stride loops: 1300 queries (11 query tiles in one launch per phase), cluster queries in the first, the
    tenth and the last tile."""
    from ralf_b200 import ops

    G, Q, rows = _near_duplicate_cluster(30000, 1300, seed=17)
    for pos, r in ((5, rows[2]), (1100, rows[3]), (1299, rows[4])):
        Q[pos] = G[r] / np.linalg.norm(G[r])
    oi, os_ = oracle_knn.topk(G, Q, 16)
    gi, gs, cert = ops.knn_topk(torch.from_numpy(G).to(cuda_device), torch.from_numpy(Q).to(cuda_device), 16,
                                gallery_max_norm=1.01)
    np.testing.assert_array_equal(gi.cpu().numpy(), oi)
    np.testing.assert_array_equal(gs.cpu().numpy().view(np.uint32), os_.view(np.uint32))
    cert = cert.cpu().numpy()
    assert (cert > 0).all() and cert[5] == 2 and cert[1100] == 2 and cert[1299] == 2 and (cert == 2).sum() <= 8


@pytest.mark.parametrize("n,q", [(100_000, 1), (100_000, 32), (100_000, 128), (1_000_000, 1), (1_000_000, 32),
                                 (1_000_000, 128)])
def test_knn_matches_oracle_at_baseline_sizes(cuda_device, n, q):
    """BASELINE configs[3] sizes (gallery 100 k / 1 M x 512, Q in {1, 32, 128}) against the C oracle: indices and score
    bits.  Plus the closest available stand-in for FAISS (not installable here, parity unpinned): torch.topk over the
    fp32 matmul G @ q on the host -- a different summation order, so positions may only differ where the oracle's
    neighbouring scores are closer than the summation-order error (first-divergence-margin rule)."""
    from ralf_b200 import ops

    g = torch.Generator().manual_seed(n + q)
    G = torch.randn(n, 512, generator=g)
    G /= G.norm(dim=1, keepdim=True)
    Q = torch.randn(q, 512, generator=g)
    Q /= Q.norm(dim=1, keepdim=True)
    oi, os_ = oracle_knn.topk(G.numpy(), Q.numpy(), 16)
    gi, gs, cert = ops.knn_topk(G.to(cuda_device), Q.to(cuda_device), 16, gallery_max_norm=1.0001)
    np.testing.assert_array_equal(gi.cpu().numpy(), oi)
    np.testing.assert_array_equal(gs.cpu().numpy().view(np.uint32), os_.view(np.uint32))
    assert (cert > 0).all()
    ts, ti = torch.topk(Q @ G.T, 17, dim=1)  # fp32 BLAS: what IndexFlat's sgemm path computes, another summation order
    ti, ts = ti.numpy(), ts.numpy()
    err = 512 * 2.0 ** -24  # |sum-order error| of a 512-term fp32 dot of unit vectors (loose bound)
    for j in range(q):
        for r in range(16):
            if ti[j, r] != oi[j, r]:
                gap = min(abs(os_[j, r] - os_[j, r - 1]) if r else 1.0, abs(os_[j, r] - os_[j, r + 1]) if r < 15 else
                          abs(os_[j, 15] - ts[j, 16]))
                assert gap <= 2 * err, f"query {j} rank {r}: BLAS ranking differs with a score gap of {gap}"
        np.testing.assert_allclose(ts[j, :16], os_[j], rtol=0, atol=2 * err)


def test_knn_sharded_merge_equals_unsharded(cuda_device):
    from ralf_b200 import ops

    G, Q = _data(30000, 512, 16, seed=5)
    oi, os_ = oracle_knn.topk(G, Q, 16)
    Gd, Qd = torch.from_numpy(G).to(cuda_device), torch.from_numpy(Q).to(cuda_device)
    parts_s, parts_i = [], []
    bounds = [0, 7000, 15000, 22001, 30000]
    for a, b in zip(bounds[:-1], bounds[1:]):
        i, s, _ = ops.knn_topk(Gd[a:b], Qd, 16, index_base=a)
        parts_s.append(s)
        parts_i.append(i)
    mi, ms = ops.knn_merge(torch.stack(parts_s), torch.stack(parts_i))
    np.testing.assert_array_equal(mi.cpu().numpy(), oi)
    np.testing.assert_array_equal(ms.cpu().numpy().view(np.uint32), os_.view(np.uint32))


def test_knn_large_roundtrip_property(cuda_device):
    """Full-size property (1M x 512): every gallery row queried against the gallery finds itself first."""
    from ralf_b200 import ops

    n, d = 1_000_000, 512
    g = torch.Generator(device=cuda_device).manual_seed(0)
    G = torch.randn(n, d, device=cuda_device, generator=g)
    G = G / G.norm(dim=1, keepdim=True)
    rows = torch.arange(0, n, n // 128, device=cuda_device)[:128]
    idx, score, cert = ops.knn_topk(G, G[rows].contiguous(), 16, gallery_max_norm=1.0001)
    assert torch.equal(idx[:, 0], rows)
    assert (score[:, :-1] >= score[:, 1:]).all()
    assert cert.all()
    ei, es, _ = ops.knn_topk(G, G[rows[:2]].contiguous(), 16, exact=True)
    assert torch.equal(ei, idx[:2]) and torch.equal(es, score[:2])


def test_knn_adversarial_row_order(cuda_device):
    """Gallery sorted by the score of query 0 (ascending: every later row beats all earlier ones, so the running
    threshold never protects the candidate lists and they overflow / compact constantly), descending for query 1
    after a flip, plus a query batch wider than one 128-query pass."""
    from ralf_b200 import ops

    G, Q = _data(60000, 512, 130, seed=21)
    G = G[np.argsort(G @ Q[0], kind="stable")]
    for Gx in (G, np.ascontiguousarray(G[::-1])):
        oi, os_ = oracle_knn.topk(Gx, Q, 16)
        gi, gs, cert = ops.knn_topk(torch.from_numpy(Gx).to(cuda_device), torch.from_numpy(Q).to(cuda_device), 16,
                                    gallery_max_norm=1.0001)
        np.testing.assert_array_equal(gi.cpu().numpy(), oi)
        np.testing.assert_array_equal(gs.cpu().numpy().view(np.uint32), os_.view(np.uint32))
        assert cert.cpu().numpy().all()


def test_flat_ip_index_as_hf_custom_index(cuda_device, monkeypatch):
    """FlatIPIndex behind HF datasets' FaissIndex (``custom_index=``), the reference's own route to FAISS
    (retrieval/retriever.py:79-84,200-202): neighbours and scores bit-exact against the C oracle, host arrays in/out."""
    import sys
    import types

    import datasets as ds

    from ralf_b200.retrieval import FlatIPIndex

    monkeypatch.setitem(sys.modules, "faiss", sys.modules.get("faiss") or types.ModuleType("faiss"))
    monkeypatch.setattr(ds.search, "_has_faiss", True)
    G, Q = _data(6000, 512, 4, seed=77, normalize=False)
    db = ds.Dataset.from_dict({"id": [str(i) for i in range(len(G))]})
    index = FlatIPIndex(512, device=cuda_device)
    db.add_faiss_index_from_external_arrays(G, index_name="search_feat", custom_index=index)
    assert index.ntotal == len(G)
    oi, os_ = oracle_knn.topk(G, Q, 17)
    for j in range(len(Q)):
        scores, examples = db.get_nearest_examples("search_feat", Q[j], k=17)
        assert examples["id"] == [str(i) for i in oi[j]]
        np.testing.assert_array_equal(np.asarray(scores, dtype=np.float32).view(np.uint32), os_[j].view(np.uint32))
    s, i = index.search(Q, 33)  # batched, the k of the top_k32 cache tables
    oi, os_ = oracle_knn.topk(G, Q, 33)
    np.testing.assert_array_equal(i, oi)
    np.testing.assert_array_equal(s.view(np.uint32), os_.view(np.uint32))
    np.testing.assert_array_equal(index.reconstruct(123), G[123])


def test_knn_passes_on_parallel_streams_equal_single_stream(cuda_device):
    """GpuRetriever.knn_ways > 1 (passes of 128 queries round-robin on parallel streams) returns what the single-stream
    call returns, eagerly and as parallel branches of a captured graph."""
    from ralf_b200.retrieval import GpuRetriever

    G, Q = _data(30000, 512, 700, seed=3)
    retr = GpuRetriever(torch.from_numpy(G), device=cuda_device)
    q = torch.from_numpy(Q).to(cuda_device)
    want_i, want_s = retr.search_local(q, 16)
    for ways in (2, 3):
        retr.knn_ways = ways
        got_i, got_s = retr.search_local(q, 16)
        torch.cuda.synchronize()
        assert torch.equal(got_i, want_i) and torch.equal(got_s, want_s) and int(retr.last_certified.sum()) == 700
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        retr.search_local(q, 16)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = retr.search_local(q, 16)
    for _ in range(2):
        graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out[0], want_i) and torch.equal(out[1], want_s)
