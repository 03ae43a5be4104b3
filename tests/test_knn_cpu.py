"""CPU: the C k-NN oracle (canonical fp32 dot, (score desc, idx asc)) against a float64 ranking."""
import os

import numpy as np
import pytest

from tests import oracle_knn


def test_canonical_dot_close_to_fp64():
    rng = np.random.default_rng(0)
    for d in (1, 31, 32, 33, 100, 512, 1000):
        a, b = rng.standard_normal(d).astype(np.float32), rng.standard_normal(d).astype(np.float32)
        ref = float(a.astype(np.float64) @ b.astype(np.float64))
        assert abs(oracle_knn.canonical_dot(a, b) - ref) <= 1e-5 * max(1.0, abs(ref))


def test_oracle_topk_matches_fp64_ranking_where_gaps_are_resolvable():
    rng = np.random.default_rng(1)
    G = rng.standard_normal((5000, 512)).astype(np.float32)
    Q = rng.standard_normal((6, 512)).astype(np.float32)
    idx, score, allsc = oracle_knn.topk(G, Q, 16, return_all=True)
    s64 = Q.astype(np.float64) @ G.astype(np.float64).T
    ref = np.argsort(-s64, axis=1, kind="stable")[:, :16]
    for q in range(6):
        gaps = np.abs(np.diff(s64[q, ref[q]]))
        if gaps.min() > 1e-4:
            np.testing.assert_array_equal(idx[q], ref[q])
        assert (np.diff(score[q]) <= 0).all()
        np.testing.assert_array_equal(score[q], allsc[q, idx[q]])


def test_oracle_ties_and_short_gallery():
    G = np.ones((5, 8), np.float32)
    Q = np.ones((1, 8), np.float32)
    idx, score = oracle_knn.topk(G, Q, 8, index_base=10)
    assert idx[0].tolist() == [10, 11, 12, 13, 14, -1, -1, -1]
    assert np.isinf(score[0, 5:]).all() and (score[0, :5] == 8).all()


def test_thread_count_does_not_change_results():
    rng = np.random.default_rng(2)
    G = rng.standard_normal((4000, 96)).astype(np.float32)
    Q = rng.standard_normal((3, 96)).astype(np.float32)
    a = oracle_knn.topk(G, Q, 16, threads=1)
    b = oracle_knn.topk(G, Q, 16, threads=7)
    np.testing.assert_array_equal(a[0], b[0])
    np.testing.assert_array_equal(a[1].view(np.uint32), b[1].view(np.uint32))


def test_flat_ip_index_through_hf_datasets_custom_index(monkeypatch):
    """The reference reaches FAISS only through HF datasets (retrieval/retriever.py:79-84,200-202).  FlatIPIndex plugs into
    that boundary as ``custom_index=``: add_faiss_index_from_external_arrays -> get_nearest_examples return the oracle's
    neighbours.  Host logic only: the kernel call is replaced by the C oracle here (GPU: tests/test_knn_gpu.py)."""
    import sys
    import types

    import datasets as ds
    import torch

    from ralf_b200 import ops
    from ralf_b200.retrieval import FlatIPIndex

    def oracle_knn_topk(gallery, queries, k, *, index_base=0, gallery_max_norm=0.0, exact=False, workspace=None):
        i, s = oracle_knn.topk(gallery.numpy(), queries.numpy(), k)
        return torch.from_numpy(i) + index_base, torch.from_numpy(s), torch.ones(queries.shape[0], dtype=torch.int32)

    monkeypatch.setattr(ops, "knn_topk", oracle_knn_topk)
    monkeypatch.setitem(sys.modules, "faiss", sys.modules.get("faiss") or types.ModuleType("faiss"))  # not installed here
    monkeypatch.setattr(ds.search, "_has_faiss", True)
    rng = np.random.default_rng(5)
    n, d = 2500, 64
    vectors = rng.standard_normal((n, d)).astype(np.float32)
    db = ds.Dataset.from_dict({"id": [str(i) for i in range(n)], "label": [[i % 4] for i in range(n)]})
    index = FlatIPIndex(d, device="cpu")
    db.add_faiss_index_from_external_arrays(vectors, index_name="search_feat", custom_index=index)
    assert index.ntotal == n  # HF adds in batches of 1000: three add() calls
    q = rng.standard_normal((3, d)).astype(np.float32)
    oi, os_ = oracle_knn.topk(vectors, q, 17)
    for j in range(3):
        scores, examples = db.get_nearest_examples("search_feat", q[j], k=17)  # retriever.py:200-202 (top_k + 1)
        assert examples["id"] == [str(i) for i in oi[j]]
        np.testing.assert_array_equal(np.asarray(scores, dtype=np.float32).view(np.uint32), os_[j].view(np.uint32))
    # FAISS conventions at the edges: k beyond ntotal pads with -1 / -inf, empty index, shape errors
    small = FlatIPIndex(d, device="cpu")
    small.add(vectors[:5])
    s, i = small.search(q, 8)
    assert (i[:, 5:] == -1).all() and np.isneginf(s[:, 5:]).all() and sorted(i[0, :5].tolist()) == [0, 1, 2, 3, 4]
    small.reset()
    assert small.ntotal == 0 and (small.search(q, 2)[1] == -1).all()
    with pytest.raises(ValueError):
        small.add(np.zeros((2, d + 4), np.float32))
    with pytest.raises(ValueError):
        small.search(q, 64)


def test_coarse_saliency_matches_reference():
    """retrieval_backbone="saliency" features vs the reference function (fixture: tests/golden/make_golden.py)."""
    import os

    import torch

    from ralf_b200.retrieval import coarse_saliency
    from tests import helpers

    z = np.load(os.path.join(helpers.GOLDEN, "coarse_saliency.npz"))
    sal = torch.from_numpy(z["saliency"].astype(np.float32))
    np.testing.assert_array_equal(coarse_saliency(sal).numpy(), z["feature"])
    np.testing.assert_array_equal(coarse_saliency(sal[:, 0]).numpy(), z["feature"])
    # the interpolation rule for other canvas sizes (the synthetic 256 x 256 canvases of the bench)
    x = torch.rand(3, 1, 256, 256)
    want = 2 * torch.nn.functional.interpolate(x, size=(16, 16)).clamp(0, 1).flatten(1) - 1
    assert torch.equal(coarse_saliency(x), want)


def test_retriever_class_sample_and_cache_tables(monkeypatch, tmp_path):
    """Drop-in for the reference's Retriever (retrieval/retriever.py:24-229), host logic with the kernels replaced by the
    oracle: saliency features of the database, sample() = the nearest canvas's layout, preprocess_retrieval_cache() = the
    reference's table format and file names (self dropped on the train split)."""
    import torch

    from ralf_b200 import data as D
    from ralf_b200 import ops
    from ralf_b200.generator import ConditionalInputs
    from ralf_b200.retrieval import Retriever, coarse_saliency

    def oracle_knn_topk(gallery, queries, k, *, index_base=0, gallery_max_norm=0.0, exact=False, workspace=None):
        i, s = oracle_knn.topk(gallery.numpy(), queries.numpy(), k)
        return torch.from_numpy(i) + index_base, torch.from_numpy(s), torch.ones(queries.shape[0], dtype=torch.int32)

    monkeypatch.setattr(ops, "knn_topk", oracle_knn_topk)
    monkeypatch.setattr(ops, "gather_layouts", lambda packed, idx: packed[idx])
    g = torch.Generator().manual_seed(4)
    n, E = 40, 10
    db = []
    for i in range(n):
        m = int(torch.randint(1, 8, (1,), generator=g))
        db.append({"id": str(1000 + i), "saliency": torch.rand(1, 64, 48, generator=g), "image": torch.rand(3, 64, 48, generator=g),
                   "label": torch.randint(0, 3, (m,), generator=g).tolist(),
                   **{k: torch.rand(m, generator=g).tolist() for k in ["center_x", "center_y", "width", "height"]}})
    r = Retriever(features=None, db_dataset=db, max_seq_length=E, top_k=1, dataset_name="pku", retrieval_backbone="saliency",
                  device="cpu")
    feats = coarse_saliency(torch.stack([x["saliency"] for x in db]))
    assert torch.equal(r.retr.emb, feats) and r.table_paired_id_idx[1003] == 3  # "pku" ids are ints in the tables
    # sample(): every database canvas retrieves itself
    rows = [5, 17, 33]
    image = torch.stack([torch.cat([db[i]["image"], db[i]["saliency"]]) for i in rows])
    out, vio = r.sample(ConditionalInputs(image=image))
    assert vio == {"total": 1, "viorated": 0} and set(out) == set(Retriever.output_keys)
    for b, i in enumerate(rows):
        m = len(db[i]["label"])
        assert out["mask"][b].tolist() == [True] * m + [False] * (E - m)
        assert out["label"][b][:m].tolist() == db[i]["label"]
        np.testing.assert_allclose(out["width"][b][:m].numpy(), np.array(db[i]["width"], dtype=np.float32))
    # cache tables: reference file names, self dropped on train, kept (top_k + 1 entries) elsewhere
    table = r.preprocess_retrieval_cache("train", db, top_k=8, root=str(tmp_path), save_scores=True)
    sims = feats.numpy() @ feats.numpy().T
    for i in (0, 9, 39):
        order = np.argsort(-sims[i], kind="stable")
        assert order[0] == i and table[1000 + i] == order[1:9].tolist()
    path = D.cache_table_path("pku", "train", "saliency", 8, root=str(tmp_path))
    assert D.load_cache_table(path, top_k=4)[1005] == table[1005][:4]
    assert os.path.exists(path.replace("indexes", "scores")) and os.path.exists(D.paired_table_path("pku", "saliency", str(tmp_path)))
    val = r.preprocess_retrieval_cache("val", db[:5], top_k=8, root=str(tmp_path))
    assert len(val[1002]) == 9 and val[1002][0] == 2
    with pytest.raises(NotImplementedError):
        Retriever(features=None, db_dataset=db, max_seq_length=E, retrieval_backbone="dreamsim", device="cpu")
