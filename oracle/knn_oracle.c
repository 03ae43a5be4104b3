/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the k-NN retrieval row of the hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * call this.  The product path (ralf_b200/csrc/knn.cu) never links or executes it.
 *
 * What it restates: the reference searches one query at a time with
 *   faiss.IndexFlat(d, faiss.METRIC_INNER_PRODUCT)           (HF datasets FaissIndex, built at
 *   image2layout/train/models/retrieval/retriever.py:79-84, searched at :193-213 with
 *   k = top_k + 1), i.e. an exact maximum-inner-product scan over the fp32 gallery, scores in
 *   descending order.  The arithmetic lives in faiss-cpu ^1.7.4 (pyproject.toml:30), which is NOT
 *   vendored under /root/reference and not installed here, and the reference ships no golden
 *   vectors for it (the published tables under data_splits/retrieval/ come from embeddings that
 *   are not shipped).  PARITY UNPINNED: this file restates IndexFlat's published algorithm
 *   (score = <q, g> in fp32, keep the k largest) and fixes the two things FAISS leaves open --
 *   the fp32 summation order (canonical order below) and ties (lower index first).
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* Canonical fp32 inner product: 32 lane-strided fmaf partial sums (lane l takes elements
 * l, l+32, l+64, ... in ascending order), then the xor butterfly 16, 8, 4, 2, 1.
 * Matches canonical_dot_warp() in ralf_b200/csrc/knn.cu bit for bit. */
float knn_canonical_dot(const float* a, const float* b, int d) {
  float p[32];
  for (int l = 0; l < 32; ++l) p[l] = 0.0f;
  int base = 0;
  for (; base + 32 <= d; base += 32)
    for (int l = 0; l < 32; ++l) p[l] = fmaf(a[base + l], b[base + l], p[l]);
  for (int l = 0; base + l < d; ++l) p[l] = fmaf(a[base + l], b[base + l], p[l]);
  for (int off = 16; off >= 1; off >>= 1) {
    float t[32];
    for (int l = 0; l < 32; ++l) t[l] = p[l] + p[l ^ off];
    memcpy(p, t, sizeof(p));
  }
  return p[0];
}

static int better(float s1, int64_t i1, float s2, int64_t i2) {
  return (s1 > s2) || (s1 == s2 && i1 < i2);
}

/* top-k of gallery rows [r0, r1) for one query; local row ids; returns count */
static int scan_rows(const float* gallery, int r0, int r1, int d, const float* qv, int k, int64_t* bi,
                     float* bs, float* scores_all) {
  int m = 0;
  for (int r = r0; r < r1; ++r) {
    const float s = knn_canonical_dot(qv, gallery + (size_t)r * d, d);
    if (scores_all) scores_all[r] = s;
    if (m < k || better(s, r, bs[m - 1], bi[m - 1])) {
      int j = m < k ? m : k - 1;
      while (j > 0 && better(s, r, bs[j - 1], bi[j - 1])) {
        bs[j] = bs[j - 1];
        bi[j] = bi[j - 1];
        --j;
      }
      bs[j] = s;
      bi[j] = r;
      if (m < k) ++m;
    }
  }
  return m;
}

typedef struct {
  const float* gallery;
  const float* queries;
  int n, d, q, k, slices;
  int64_t* part_idx; /* [q, slices, k] */
  float* part_score;
  int* part_cnt; /* [q, slices] */
  float* scores_all;
  int next; /* work counter */
  pthread_mutex_t mu;
} knn_job;

static void* knn_worker(void* arg) {
  knn_job* J = (knn_job*)arg;
  for (;;) {
    pthread_mutex_lock(&J->mu);
    const int w = J->next++;
    pthread_mutex_unlock(&J->mu);
    if (w >= J->q * J->slices) break;
    const int qi = w / J->slices, sl = w % J->slices;
    const int r0 = (int)(((int64_t)sl * J->n) / J->slices), r1 = (int)(((int64_t)(sl + 1) * J->n) / J->slices);
    J->part_cnt[w] = scan_rows(J->gallery, r0, r1, J->d, J->queries + (size_t)qi * J->d, J->k,
                               J->part_idx + (size_t)w * J->k, J->part_score + (size_t)w * J->k,
                               J->scores_all ? J->scores_all + (size_t)qi * J->n : NULL);
  }
  return NULL;
}

/* Top-k for every query; (score desc, index asc); missing -> idx -1, score -inf.
 * scores_all (optional, [q, n]) receives every canonical score.  nthreads >= 1 host threads. */
void knn_oracle_topk(const float* gallery, int n, int d, const float* queries, int q, int k,
                     int64_t index_base, int64_t* out_idx, float* out_score, float* scores_all,
                     int nthreads) {
  if (nthreads < 1) nthreads = 1;
  knn_job J;
  memset(&J, 0, sizeof(J));
  J.gallery = gallery; J.queries = queries; J.n = n; J.d = d; J.q = q; J.k = k;
  J.slices = q >= nthreads ? 1 : (nthreads + q - 1) / q;
  if (J.slices > n) J.slices = n > 0 ? n : 1;
  J.part_idx = (int64_t*)malloc(sizeof(int64_t) * (size_t)q * J.slices * k);
  J.part_score = (float*)malloc(sizeof(float) * (size_t)q * J.slices * k);
  J.part_cnt = (int*)calloc((size_t)q * J.slices, sizeof(int));
  J.scores_all = scores_all;
  pthread_mutex_init(&J.mu, NULL);
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * nthreads);
  for (int t = 1; t < nthreads; ++t) pthread_create(&th[t], NULL, knn_worker, &J);
  knn_worker(&J);
  for (int t = 1; t < nthreads; ++t) pthread_join(th[t], NULL);
  for (int qi = 0; qi < q; ++qi) { /* merge the slices of each query */
    int64_t* bi = out_idx + (size_t)qi * k;
    float* bs = out_score + (size_t)qi * k;
    int m = 0;
    for (int sl = 0; sl < J.slices; ++sl) {
      const int w = qi * J.slices + sl;
      for (int e = 0; e < J.part_cnt[w]; ++e) {
        const float s = J.part_score[(size_t)w * k + e];
        const int64_t r = J.part_idx[(size_t)w * k + e];
        if (m < k || better(s, r, bs[m - 1], bi[m - 1])) {
          int j = m < k ? m : k - 1;
          while (j > 0 && better(s, r, bs[j - 1], bi[j - 1])) { bs[j] = bs[j - 1]; bi[j] = bi[j - 1]; --j; }
          bs[j] = s; bi[j] = r;
          if (m < k) ++m;
        }
      }
    }
    for (int j = 0; j < m; ++j) bi[j] += index_base;
    for (int j = m; j < k; ++j) { bs[j] = -INFINITY; bi[j] = -1; }
  }
  free(th); free(J.part_idx); free(J.part_score); free(J.part_cnt);
  pthread_mutex_destroy(&J.mu);
}
