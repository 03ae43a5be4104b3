"""CPU: the C k-NN oracle (canonical fp32 dot, (score desc, idx asc)) against a float64 ranking."""
import numpy as np

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
