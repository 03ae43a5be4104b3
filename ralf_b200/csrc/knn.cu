// K1: exact maximum-inner-product top-k over an fp32 embedding gallery.
//
// Replaces the FAISS IndexFlat(METRIC_INNER_PRODUCT) scan the reference runs one query at a time
// on the CPU (image2layout/train/models/retrieval/retriever.py:79-84,193-213).
//
// Phase 1 (knn_scan_kernel): the gallery [n, d] fp32 is streamed ONCE from HBM by TMA
// (SWIZZLE_128B boxes of 256 rows x 32 floats) and multiplied with a 128-query tile by
// tcgen05.mma kind::tf32 (M = 128 queries, N = 256 gallery rows, fp32 accumulators, two TMEM
// buffers of 256 columns so the epilogue of tile t overlaps the MMAs of tile t+1).  Epilogue
// thread q owns query q: it reads its accumulator row from TMEM and keeps a running top-C
// candidate list (C >= 2k) in shared memory -- one fp32 compare per score in the steady state,
// append on success; when any list of a warp nears capacity the warp sorts the over-full lists
// cooperatively (bitonic network in registers, one list at a time) and tightens their thresholds.
// TF32 scores only nominate candidates; they never decide the result.
// Phase 2 (knn_rerank_kernel): per query, merge the per-CTA candidate lists, re-score the best
// C candidates with the canonical fp32 dot product (bit-identical to oracle/knn_oracle.c),
// order by (score desc, index asc), emit top-k, and certify exactness from the TF32 error bound.
//
// Algorithmic HBM bytes per call: n*d*4 (gallery once) + q*d*4 + q*k*12.
#include <float.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "ralf_internal.h"

namespace ralf {

// Order-preserving key: higher score first, then LOWER index first.
__device__ __forceinline__ uint64_t knn_key(float s, uint32_t idx) {
  uint32_t u = __float_as_uint(s);
  u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return (static_cast<uint64_t>(u) << 32) | static_cast<uint64_t>(0xFFFFFFFFu - idx);
}
__device__ __forceinline__ float knn_key_score(uint64_t key) {
  uint32_t u = static_cast<uint32_t>(key >> 32);
  u = (u & 0x80000000u) ? (u ^ 0x80000000u) : ~u;
  return __uint_as_float(u);
}
__device__ __forceinline__ uint32_t knn_key_index(uint64_t key) {
  return 0xFFFFFFFFu - static_cast<uint32_t>(key & 0xFFFFFFFFull);
}

// Canonical fp32 inner product (see oracle/knn_oracle.c: knn_canonical_dot):
//   lane l accumulates  acc_l = fmaf(a[i], b[i], acc_l)  for i = l, l+32, l+64, ... (ascending),
//   then the 32 partials are combined by the xor butterfly 16, 8, 4, 2, 1 (p_l += p_{l^off}).
// Every lane returns the same value.  Executed by one full warp.
__device__ __forceinline__ float canonical_dot_warp(const float* __restrict__ a,
                                                    const float* __restrict__ b, int d) {
  float acc = 0.f;
  for (int i = lane_id(); i < d; i += 32) acc = fmaf(a[i], b[i], acc);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) acc = acc + __shfl_xor_sync(0xffffffffu, acc, off);
  return acc;
}

template <int C>
struct KnnCfg {
  static constexpr int NS = C / 16;            // sort slots per lane: bitonic network over 32*NS keys
  static constexpr int CAP = (C <= 32) ? 64 : 96;  // candidate-list capacity per query (C kept + appends)
  static constexpr int STRIDE = CAP + 1;       // u64 per query list (+1: spreads the appends over banks)
  static constexpr int A_BYTES = 128 * 128;    // 128 queries x 32 fp32
  static constexpr int B_BYTES = 256 * 128;    // 256 gallery rows x 32 fp32
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (C <= 32) ? 3 : 2;
  static constexpr int CAND_BYTES = ((128 * STRIDE * 8 + 1023) / 1024) * 1024;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + CAND_BYTES + 1024 + 256;
  static_assert(32 * NS >= CAP, "sort network must cover the list");
};

// Bitonic sort (descending) of 32*NS u64 keys held NS per lane; key index i = lane + 32*slot.
// After the sort x[s] of lane l is the key of rank l + 32*s.
template <int NS>
__device__ __forceinline__ void bitonic_desc(uint64_t (&x)[NS], const int lane) {
#pragma unroll
  for (int k = 2; k <= 32 * NS; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 32) {  // partner lives in another slot of the same lane
        const int js = j >> 5;
        uint64_t y[NS];
#pragma unroll
        for (int sl = 0; sl < NS; ++sl) {
          const uint64_t o = x[sl ^ js];
          const int i = lane + 32 * sl;
          const bool keep_max = (((i & k) == 0) == ((i & j) == 0));
          y[sl] = keep_max ? (x[sl] > o ? x[sl] : o) : (x[sl] < o ? x[sl] : o);
        }
#pragma unroll
        for (int sl = 0; sl < NS; ++sl) x[sl] = y[sl];
      } else {
#pragma unroll
        for (int sl = 0; sl < NS; ++sl) {
          const uint64_t o = __shfl_xor_sync(0xffffffffu, x[sl], j);
          const int i = lane + 32 * sl;
          const bool keep_max = (((i & k) == 0) == ((i & j) == 0));
          x[sl] = keep_max ? (x[sl] > o ? x[sl] : o) : (x[sl] < o ? x[sl] : o);
        }
      }
    }
  }
}

// One key per lane: bitonic sort (descending) of 32 keys across the warp.
__device__ __forceinline__ uint64_t warp_sort32_desc(uint64_t x, const int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const uint64_t o = __shfl_xor_sync(0xffffffffu, x, j);
      const bool keep_max = (((lane & k) == 0) == ((lane & j) == 0));
      x = keep_max ? (x > o ? x : o) : (x < o ? x : o);
    }
  }
  return x;
}
// a, b sorted descending (one key per lane each) -> the 32 largest of the 64, sorted descending.
__device__ __forceinline__ uint64_t warp_merge_top32(uint64_t a, uint64_t b, const int lane) {
  const uint64_t rb = __shfl_sync(0xffffffffu, b, 31 - lane);
  uint64_t c = a > rb ? a : rb;  // bitonic sequence holding the top half
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) {
    const uint64_t o = __shfl_xor_sync(0xffffffffu, c, j);
    c = ((lane & j) == 0) ? (c > o ? c : o) : (c < o ? c : o);
  }
  return c;
}

// Warp-cooperative compaction round.  Each lane owns one query list (cnt entries in shared memory at
// lists + lane*STRIDE).  For every list longer than C the whole warp reduces it to its best C entries and
// publishes the new threshold (score of rank C-1) to the owning lane; shorter lists are left alone (their
// threshold stays, any new score still passes).
//   C == 32: the first 32 entries stay sorted between rounds (`sorted` flag per list), so a round sorts only
//            the <= 32 appended keys (15 compare-exchange steps) and merges (1 + 5 steps);
//   C == 64: full bitonic sort of the 128-slot list.
template <int C>
__device__ __forceinline__ void knn_compact_round(uint64_t* lists, const int lane, int& cnt, float& thr, int& sorted,
                                                  const int min_len) {
  using Cfg = KnnCfg<C>;
  __syncwarp();  // the owners' appends (plain shared-memory stores) must be visible to the lanes that sort their lists
                 // (compute-sanitizer racecheck flagged the read below against the append in knn_scan_kernel)
#pragma unroll 1
  for (int ql = 0; ql < 32; ++ql) {
    const int n = __shfl_sync(0xffffffffu, cnt, ql);
    if (n <= min_len) continue;  // warp-uniform: only lists that are about to overflow (or, at the end, exceed C)
    uint64_t* col = lists + ql * Cfg::STRIDE;
    uint64_t kth;
    if constexpr (C == 32) {
      const int was_sorted = __shfl_sync(0xffffffffu, sorted, ql);
      uint64_t a = col[lane];
      uint64_t b = (32 + lane < n) ? col[32 + lane] : 0ull;
      if (!was_sorted) a = warp_sort32_desc(a, lane);
      b = warp_sort32_desc(b, lane);
      a = warp_merge_top32(a, b, lane);
      col[lane] = a;
      kth = __shfl_sync(0xffffffffu, a, 31);
    } else {
      constexpr int NS = Cfg::NS;
      uint64_t x[NS];
#pragma unroll
      for (int sl = 0; sl < NS; ++sl) {
        const int e = lane + 32 * sl;
        x[sl] = (e < n && e < Cfg::CAP) ? col[e] : 0ull;
      }
      bitonic_desc<NS>(x, lane);
#pragma unroll
      for (int sl = 0; sl < C / 32; ++sl) col[lane + 32 * sl] = x[sl];
      kth = __shfl_sync(0xffffffffu, x[(C - 1) >> 5], (C - 1) & 31);
    }
    if (lane == ql) {
      cnt = C;
      thr = knn_key_score(kth);
      sorted = 1;
    }
    __syncwarp();
  }
}

// Algorithm outline of the tensor-core path (all launches on the caller's stream):
//   1. knn_scan_kernel<C, true>   "pre-pass": <= 128 gallery tiles (strided over the whole gallery) are scored and
//      only the maximum per 32-row group and query is kept (8 values per tile).
//   2. knn_threshold_kernel<C>    per query: the C-th largest group maximum.  At least C distinct rows score >= that
//      value, so it is a valid lower bound of the final C-th best TF32 score; the scan starts from it.
//   3. knn_scan_kernel<C, false>  the full gallery scan.  Steady state per 8 scores: 8 compares + one warp vote; the
//      (rare) passing scores are appended to the query's candidate list in shared memory.
//   4. knn_rerank_kernel<C>       merge the per-CTA lists, exact canonical re-score, order, certificate.
constexpr int KNN_SAMPLE_TILES = 128;  // pre-pass tiles (32 K rows)
constexpr int KNN_GROUPS = 8;          // group maxima per tile (32 columns each)

__device__ __forceinline__ float knn_next_below(float s) {
  if (isinf(s) || isnan(s)) return s;
  uint32_t u = __float_as_uint(s);
  if (s > 0.f) u -= 1u;
  else if (s < 0.f) u += 1u;
  else u = 0x80000001u;
  return __uint_as_float(u);
}

template <int C, bool PRE>
__global__ void __launch_bounds__(192, 1)
knn_scan_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmG,
                const int n, const int d, const int nq, uint64_t* __restrict__ cand,
                const float* __restrict__ thr_init, float* __restrict__ tmax, const int debug_mode) {
  // debug_mode (RALF_KNN_DEBUG, profiling only; results are garbage): 2 = the epilogue only hands the accumulator back.
  using Cfg = KnnCfg<C>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base_u32 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base_u32 - smem_u32(smem_raw));
  uint64_t* cand_s = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::CAND_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;  // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;      // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qtile = blockIdx.y;
  const int total_tiles = (n + 255) / 256;
  // scan: CTA x owns a contiguous range of tiles; pre-pass: CTA x scores ONE tile, strided over the gallery
  const int t_begin = static_cast<int>((static_cast<long long>(blockIdx.x) * total_tiles) / gridDim.x);
  const int t_end = PRE ? t_begin + 1
                        : static_cast<int>((static_cast<long long>(blockIdx.x + 1) * total_tiles) / gridDim.x);
  const int nkb = (d + 31) / 32;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmG);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int t = t_begin; t < t_end; ++t) {
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          mbar_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
          uint8_t* st = smem + s * Cfg::STAGE_BYTES;
          tma_load_3d(&tmQ, &full_bar[s], st, kb * 32, qtile * 128, 0);
          tma_load_3d(&tmG, &full_bar[s], st + Cfg::A_BYTES, kb * 32, t * 256, 0);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(2, 128, 256);  // tf32 x tf32 -> fp32
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int t = t_begin; t < t_end; ++t, ++it) {
        const int buf = it & 1;
        mbar_wait(&tempty_bar[buf], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + buf * 256;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_addr = base_u32 + s * Cfg::STAGE_BYTES;
          const uint64_t da = make_sw128_kmajor_desc(a_addr);
          const uint64_t db = make_sw128_kmajor_desc(a_addr + Cfg::A_BYTES);
#pragma unroll
          for (int k = 0; k < 4; ++k)  // UMMA_K = 8 tf32 = 32 bytes -> descriptor += 2
            mma_tf32_ss(tacc, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
          tc_commit(&empty_bar[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        tc_commit(&tfull_bar[buf]);
      }
    }
  } else {
    const int quad = warp & 3;
    const int ql = quad * 32 + lane;  // query lane within the tile == TMEM lane
    const bool active = (qtile * 128 + ql) < nq;
    uint64_t* lists = cand_s + static_cast<size_t>(quad) * 32 * Cfg::STRIDE;  // this warp's 32 query lists
    uint64_t* mine = lists + lane * Cfg::STRIDE;
    int cnt = 0, sorted = 0;
    // inactive lanes (query >= nq) never pass the filter
    float thr = !active ? INFINITY : (thr_init ? thr_init[qtile * 128 + ql] : -INFINITY);
    int it = 0;
    for (int t = t_begin; t < t_end; ++t, ++it) {
      const int buf = it & 1;
      mbar_wait(&tfull_bar[buf], (it >> 1) & 1);
      tc_fence_after();
      const int g0 = t * 256;
      const int ncols = min(256, n - g0);
      const uint32_t tacc = tmem_base + buf * 256 + (static_cast<uint32_t>(quad * 32) << 16);
      if constexpr (PRE) {
        float* out = tmax + (static_cast<size_t>(qtile) * gridDim.x + blockIdx.x) * KNN_GROUPS * 128 + ql;
#pragma unroll 1
        for (int c0 = 0; c0 < 256; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32(tacc + c0, v);
          tmem_ld_wait();
          float m = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; ++j) m = fmaxf(m, (c0 + j < ncols) ? __uint_as_float(v[j]) : -INFINITY);
          out[(c0 >> 5) * 128] = m;
        }
      } else {
#pragma unroll 1
        for (int c0 = 0; c0 < ncols; c0 += 64) {
          if (debug_mode == 2) break;
          uint32_t v[64];
          tmem_ld_32x32(tacc + c0, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
          tmem_ld_32x32(tacc + c0 + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            // steady state: 8 compares + one vote.  Columns >= ncols hold zero-row products: they may trigger the slow
            // path spuriously, where the `j < lim` test rejects them.
            bool hit = false;
#pragma unroll
            for (int j = 0; j < 8; ++j) hit |= __uint_as_float(v[g * 8 + j]) > thr;
            if (__any_sync(0xffffffffu, hit)) {
              // at most 8 appends before the next capacity check
              if (__any_sync(0xffffffffu, cnt > Cfg::CAP - 8))
                knn_compact_round<C>(lists, lane, cnt, thr, sorted, Cfg::CAP - 8);
              const int lim = ncols - c0 - g * 8;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float s = __uint_as_float(v[g * 8 + j]);
                if (s > thr && j < lim) {
                  mine[cnt] = knn_key(s, static_cast<uint32_t>(g0 + c0 + g * 8 + j));
                  ++cnt;
                }
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[buf]);
    }
    if constexpr (!PRE) {
      __syncwarp();
      knn_compact_round<C>(lists, lane, cnt, thr, sorted, C);  // trims every list to <= C entries
      // candidate lists: [qtile][cta][query lane][C]; unused slots are 0 (= empty)
      uint64_t* out_w = cand + ((static_cast<size_t>(qtile) * gridDim.x + blockIdx.x) * 128 + quad * 32) * C;
#pragma unroll 1
      for (int q = 0; q < 32; ++q) {
        const int n_q = __shfl_sync(0xffffffffu, cnt, q);
        const uint64_t* col = lists + q * Cfg::STRIDE;
        for (int e = lane; e < C; e += 32) out_w[static_cast<size_t>(q) * C + e] = (e < n_q) ? col[e] : 0ull;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// Block-wide (256 threads) selection of the C largest keys out of `parts` lists of C keys (0 = empty slot).
// Every warp folds its share of the lists into a running sorted top-C held in registers (warp bitonic sort + merge),
// then warp 0 merges the eight partial results.  sel[0..C) ends sorted descending; scratch = 8*C keys.
template <int C, typename LoadFn>
__device__ __forceinline__ void block_top_c(LoadFn load, const int parts, uint64_t* sel, uint64_t* scratch) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if constexpr (C == 32) {
    uint64_t r = 0ull;
    for (int p0 = warp; p0 < parts; p0 += 64) {  // 8 lists requested at once: one memory round trip per 8 merges
      uint64_t pre[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) pre[u] = (p0 + 8 * u < parts) ? load(p0 + 8 * u, lane) : 0ull;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        uint64_t b = pre[u];
        if (!__any_sync(0xffffffffu, b != 0ull)) continue;
        b = warp_sort32_desc(b, lane);
        r = warp_merge_top32(r, b, lane);
      }
    }
    scratch[warp * 32 + lane] = r;
    __syncthreads();
    if (warp == 0) {
      r = scratch[lane];
      for (int w = 1; w < 8; ++w) r = warp_merge_top32(r, scratch[w * 32 + lane], lane);
      sel[lane] = r;
    }
  } else {
    uint64_t x[4] = {0ull, 0ull, 0ull, 0ull};
    for (int p = warp; p < parts; p += 8) {
      x[2] = load(p, lane);
      x[3] = load(p, 32 + lane);
      if (!__any_sync(0xffffffffu, (x[2] | x[3]) != 0ull)) continue;
      bitonic_desc<4>(x, lane);
    }
    scratch[warp * 64 + lane] = x[0];
    scratch[warp * 64 + 32 + lane] = x[1];
    __syncthreads();
    if (warp == 0) {
      x[0] = scratch[lane];
      x[1] = scratch[32 + lane];
      for (int w = 1; w < 8; ++w) {
        x[2] = scratch[w * 64 + lane];
        x[3] = scratch[w * 64 + 32 + lane];
        bitonic_desc<4>(x, lane);
      }
      sel[lane] = x[0];
      sel[32 + lane] = x[1];
    }
  }
  __syncthreads();
}

// Step 2: starting threshold per query = just below the C-th largest of the pre-pass group maxima
// (tmax: [qtile][sample tile][group][128 query lanes]); -inf when fewer than C groups exist.
template <int C>
__global__ void __launch_bounds__(256)
knn_threshold_kernel(const float* __restrict__ tmax, const int sample_tiles, float* __restrict__ thr_out) {
  __shared__ uint64_t sel[C];
  __shared__ uint64_t scratch[8 * C];
  const int q = blockIdx.x;
  const int qtile = q >> 7, ql = q & 127;
  const int ns = sample_tiles * KNN_GROUPS;
  const float* src = tmax + static_cast<size_t>(qtile) * ns * 128 + ql;
  auto load = [&](int p, int e) -> uint64_t {
    const int i = p * C + e;
    return (i < ns) ? knn_key(src[static_cast<size_t>(i) * 128], static_cast<uint32_t>(i)) : 0ull;
  };
  block_top_c<C>(load, (ns + C - 1) / C, sel, scratch);
  if (threadIdx.x == 0) thr_out[q] = (sel[C - 1] != 0ull) ? knn_next_below(knn_key_score(sel[C - 1])) : -INFINITY;
}

// Step 4.  One CTA (256 threads) per query.  `parts` candidate lists of C keys each (0 = empty slot).
template <int C>
__device__ __forceinline__ void
knn_rerank_body(const int q, const uint64_t* __restrict__ cand, const int parts, const float* __restrict__ gallery,
                const float* __restrict__ queries, const int n, const int d, const int k,
                const long long index_base, const float gmax_norm, const float* __restrict__ thr_init,
                long long* __restrict__ out_idx, float* __restrict__ out_score, int* __restrict__ certified) {
  __shared__ uint64_t sel[C];
  __shared__ uint64_t scratch[8 * C];
  __shared__ float exact[C];
  __shared__ uint64_t fin[C];
  __shared__ int frank[C];
  __shared__ float qnorm2_s[8];
  const int qtile = q >> 7, ql = q & 127;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint64_t* src = cand + (static_cast<size_t>(qtile) * parts * 128 + ql) * C;
  auto load = [&](int p, int e) -> uint64_t { return src[static_cast<size_t>(p) * 128 * C + e]; };
  block_top_c<C>(load, parts, sel, scratch);  // top-C by approximate (TF32) key
  // exact canonical re-score of the selected candidates: warp w owns candidates w*C/8 .. +C/8, four at a time so
  // their gallery rows stream concurrently (each accumulator chain keeps the canonical order).
  const float* qv = queries + static_cast<size_t>(q) * d;
  constexpr int PER_WARP = C / 8;
#pragma unroll 1
  for (int c0 = warp * PER_WARP; c0 < (warp + 1) * PER_WARP; c0 += 4) {
    const float* g[4];
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint64_t kv = sel[c0 + u];
      g[u] = gallery + static_cast<size_t>(kv != 0ull ? knn_key_index(kv) : 0u) * d;
    }
    for (int i = lane; i < d; i += 32) {
      const float a = qv[i];
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] = fmaf(a, g[u][i], acc[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float s = acc[u];
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) s = s + __shfl_xor_sync(0xffffffffu, s, off);
      if (lane == 0) exact[c0 + u] = (sel[c0 + u] != 0ull) ? s : -INFINITY;
    }
  }
  {
    float a = 0.f;
    for (int i = tid; i < d; i += 256) a = fmaf(qv[i], qv[i], a);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
    if (lane == 0) qnorm2_s[warp] = a;
  }
  __syncthreads();
  // final order: exact score descending, index ascending
  if (tid < C) {
    const uint64_t kv = sel[tid];
    const uint64_t mine = (kv != 0) ? knn_key(exact[tid], knn_key_index(kv)) : 0ull;
    int rank = 0;
    for (int j = 0; j < C; ++j) {
      const uint64_t kj = sel[j];
      const uint64_t other = (kj != 0) ? knn_key(exact[j], knn_key_index(kj)) : 0ull;
      rank += (other > mine) || (other == mine && j < tid);
    }
    if (rank < k) {
      out_idx[static_cast<size_t>(q) * k + rank] =
          (kv != 0) ? static_cast<long long>(knn_key_index(kv)) + index_base : -1ll;
      out_score[static_cast<size_t>(q) * k + rank] = (kv != 0) ? exact[tid] : -INFINITY;
    }
    fin[tid] = mine;
    frank[tid] = rank;
  }
  __syncthreads();
  if (tid == 0 && certified) {
    // Every row that is NOT a candidate has an approximate score <= a_c: the C-th candidate's when C candidates
    // exist, else the scan's starting threshold (-inf without a pre-pass: then every row was a candidate).
    const float a_c = (sel[C - 1] != 0ull) ? knn_key_score(sel[C - 1]) : (thr_init ? thr_init[q] : -INFINITY);
    int ok = 1;
    if (a_c != -INFINITY) {
      if (gmax_norm > 0.f) {
        float qn2 = 0.f;
        for (int w = 0; w < 8; ++w) qn2 += qnorm2_s[w];
        const float err = (1.953125e-3f + static_cast<float>(d) * 1.2e-7f) * sqrtf(qn2) * gmax_norm;
        float s_k = -INFINITY;
        for (int j = 0; j < C; ++j)
          if (frank[j] == k - 1 && fin[j] != 0ull) s_k = knn_key_score(fin[j]);
        ok = (s_k > a_c + err) ? 1 : 0;
      } else {
        ok = 0;  // no norm bound supplied: cannot certify
      }
    }
    certified[q] = ok;
  }
}

template <int C>
__global__ void __launch_bounds__(256)
knn_rerank_kernel(const uint64_t* __restrict__ cand, const int parts, const float* __restrict__ gallery,
                  const float* __restrict__ queries, const int n, const int d, const int k,
                  const long long index_base, const float gmax_norm, const float* __restrict__ thr_init,
                  long long* __restrict__ out_idx, float* __restrict__ out_score, int* __restrict__ certified) {
  knn_rerank_body<C>(blockIdx.x, cand, parts, gallery, queries, n, d, k, index_base, gmax_norm, thr_init, out_idx,
                     out_score, certified);
}

// Exact fix-up of the queries phase 1 could not certify (ralf_knn_fixup_exact), entirely on the device:
//   knn_collect_uncertified_kernel  compacts {q : certified[q] == 0} into qlist / qcount;
//   knn_exact_scan_kernel           (qlist form) scans the gallery with the canonical dot for the listed queries only;
//   knn_rerank_listed_kernel        orders their candidates, overwrites their result rows and marks them certified = 2.
// With nothing to fix (the usual case) the three launches exit after one load each.
__global__ void __launch_bounds__(1024)
knn_collect_uncertified_kernel(const int* __restrict__ certified, const int q, int* __restrict__ qlist,
                               int* __restrict__ qcount) {
  __shared__ int warp_tot[32];
  __shared__ int base_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  for (int q0 = 0; q0 < q; q0 += 1024) {  // ascending order: the list (and with it every later launch) is deterministic
    const int i = q0 + threadIdx.x;
    const bool bad = i < q && certified[i] == 0;
    const unsigned m = __ballot_sync(0xffffffffu, bad);
    if (lane == 0) warp_tot[warp] = __popc(m);
    __syncthreads();
    int off = base_s;
    for (int w = 0; w < warp; ++w) off += warp_tot[w];
    if (bad) qlist[off + __popc(m & ((1u << lane) - 1u))] = i;
    __syncthreads();
    if (threadIdx.x == 0) {
      int t = 0;
      for (int w = 0; w < 32; ++w) t += warp_tot[w];
      base_s += t;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *qcount = base_s;
}

template <int C>
__global__ void __launch_bounds__(256)
knn_rerank_listed_kernel(const int* __restrict__ qlist, const int* __restrict__ qcount,
                         const uint64_t* __restrict__ cand, const int parts, const float* __restrict__ gallery,
                         const float* __restrict__ queries, const int n, const int d, const int k,
                         const long long index_base, long long* __restrict__ out_idx, float* __restrict__ out_score,
                         int* __restrict__ certified) {
  const int cnt = *qcount;
  for (int qi = blockIdx.x; qi < cnt; qi += gridDim.x) {  // CTA-uniform trip count
    const int q = qlist[qi];
    knn_rerank_body<C>(q, cand, parts, gallery, queries, n, d, k, index_base, 0.f, nullptr, out_idx, out_score,
                       nullptr);
    __syncthreads();
    if (threadIdx.x == 0) certified[q] = 2;  // exact by construction
  }
}

// Exact CUDA-core scan (fallback + independent check).  grid = (slices, q); block = 256 (8 warps);
// each warp scores rows with the canonical dot and keeps a sorted top-C list (lane 0 inserts).
template <int C>
__global__ void __launch_bounds__(256)
knn_exact_scan_kernel(const float* __restrict__ gallery, const float* __restrict__ queries, const int n,
                      const int d, uint64_t* __restrict__ cand, const int* __restrict__ qlist,
                      const int* __restrict__ qcount) {
  __shared__ uint64_t lists[8][C];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cnt = qlist ? *qcount : gridDim.y;
  for (int qi = blockIdx.y; qi < cnt; qi += gridDim.y) {
  const int q = qlist ? qlist[qi] : qi;
  const float* qv = queries + static_cast<size_t>(q) * d;
  const long long r0 = (static_cast<long long>(blockIdx.x) * n) / gridDim.x;
  const long long r1 = (static_cast<long long>(blockIdx.x + 1) * n) / gridDim.x;
  uint64_t* lst = lists[warp];
  for (int e = lane; e < C; e += 32) lst[e] = 0;
  __syncwarp();
  int m = 0;
  float thr = -INFINITY;
  for (long long r = r0 + warp; r < r1; r += 8) {
    const float s = canonical_dot_warp(qv, gallery + static_cast<size_t>(r) * d, d);
    if (s > thr || m < C) {
      if (lane == 0) {
        const uint64_t key = knn_key(s, static_cast<uint32_t>(r));
        int j = m;
        while (j > 0 && lst[j - 1] < key) {
          if (j < C) lst[j] = lst[j - 1];
          --j;
        }
        if (j < C) lst[j] = key;
      }
      if (m < C) ++m;
      __syncwarp();
      thr = (m == C) ? knn_key_score(lst[C - 1]) : -INFINITY;
    }
  }
  __syncwarp();
  // layout expected by knn_rerank_kernel with parts = slices * 8: [qtile][part][lane][C]
  const int qtile = q >> 7, ql = q & 127;
  const int parts = gridDim.x * 8;
  const int part = blockIdx.x * 8 + warp;
  uint64_t* out = cand + ((static_cast<size_t>(qtile) * parts + part) * 128 + ql) * C;
  for (int e = lane; e < C; e += 32) out[e] = lst[e];
  __syncwarp();
  }
}

__global__ void knn_merge_kernel(const float* __restrict__ ps, const long long* __restrict__ pi, const int parts,
                                 const int nq, const int k, long long* __restrict__ out_idx,
                                 float* __restrict__ out_score) {
  const int q = blockIdx.x;
  const int total = parts * k;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int p = i / k, e = i - p * k;
    const size_t src = (static_cast<size_t>(p) * nq + q) * k + e;
    const float s = ps[src];
    const long long id = pi[src];
    if (id < 0) continue;
    int rank = 0;
    for (int j = 0; j < total; ++j) {
      const int pj = j / k, ej = j - pj * k;
      const size_t sj = (static_cast<size_t>(pj) * nq + q) * k + ej;
      const float s2 = ps[sj];
      const long long id2 = pi[sj];
      if (id2 < 0) continue;
      rank += (s2 > s) || (s2 == s && id2 < id);
    }
    if (rank < k) {
      out_idx[static_cast<size_t>(q) * k + rank] = id;
      out_score[static_cast<size_t>(q) * k + rank] = s;
    }
  }
}
__global__ void knn_fill_kernel(long long* idx, float* score, size_t count) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < count) { idx[i] = -1; score[i] = -INFINITY; }
}

static int knn_c_for_k(int k) { return k <= 24 ? 32 : (k <= 48 ? 64 : 0); }
static int knn_grid_x(int n) {
  const int tiles = (n + 255) / 256;
  const int sms = num_sms();
  return tiles < sms ? tiles : sms;
}
static int knn_exact_slices(int C) { return C <= 32 ? 64 : 32; }

}  // namespace ralf

using namespace ralf;

// Pre-pass sample: 1/16 of the gallery tiles, at least 16 (128 group maxima per query >= C for both list sizes... C = 64
// needs 8 tiles) and at most KNN_SAMPLE_TILES; every tile when the gallery has fewer than 16.
static int knn_sample_tiles(int n) {
  const int tiles = (n + 255) / 256;
  int ns = tiles / 16;
  if (ns < 16) ns = 16;
  if (ns > KNN_SAMPLE_TILES) ns = KNN_SAMPLE_TILES;
  return tiles < ns ? tiles : ns;
}
static size_t knn_cand_bytes(int n, int q, int C) {
  const size_t qtiles = (q + 127) / 128;
  size_t parts = knn_grid_x(n);
  if (parts < static_cast<size_t>(knn_exact_slices(C)) * 8) parts = static_cast<size_t>(knn_exact_slices(C)) * 8;
  return qtiles * parts * 128 * C * sizeof(uint64_t);
}

// workspace = [candidate lists][pre-pass group maxima: qtiles x 128 tiles x 8 groups x 128 lanes f32][thresholds]
//             [fix-up query list: q + 1 ints (count last)]
static size_t knn_front_bytes(int n, int q, int C) {
  const size_t qtiles = (q + 127) / 128;
  return knn_cand_bytes(n, q, C) + qtiles * KNN_SAMPLE_TILES * KNN_GROUPS * 128 * sizeof(float) +
         qtiles * 128 * sizeof(float);
}
extern "C" size_t ralf_knn_workspace_bytes(int n, int d, int q, int k) {
  (void)d;
  const int C = knn_c_for_k(k);
  if (C == 0 || n <= 0 || q <= 0) return 0;
  return knn_front_bytes(n, q, C) + (static_cast<size_t>(q) + 1) * sizeof(int);
}

template <int C>
static int knn_topk_impl(const float* gallery, int n, int d, const float* queries, int q, int k,
                         long long index_base, float gmax, long long* out_idx, float* out_score, int* certified,
                         void* workspace, cudaStream_t st) {
  using Cfg = KnnCfg<C>;
  CUtensorMap tg;
  int rc = make_kmajor_tmap(&tg, gallery, 4, d, n, 1, d, 0, 256);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(knn_scan_kernel<C, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(knn_scan_kernel<C, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return set_cuda_error(e);
    attr_set = true;
  }
  // All query tiles (128 queries each) go out in ONE launch per phase (blockIdx.y = query tile): a query tile still
  // streams the gallery once and stays in the HBM-bound regime the kernel is built for (beyond ~280 queries per gallery
  // pass the TF32 MMAs, not HBM, set the time -- SURVEY.md 8d), but the pre-pass / threshold / re-rank launches and the
  // scan's launch gaps and tails are paid once per call instead of once per 128 queries (round 1: 4 launches per tile;
  // at an 8-way shard the fixed ~55 us per tile outweighed the 39 us of HBM time).  CTAs are scheduled x
  // fastest, so query tile y+1 starts in the SMs tile y's tail leaves idle.
  uint64_t* cand = reinterpret_cast<uint64_t*>(workspace);
  float* tmax = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + knn_cand_bytes(n, q, C));
  const int qtiles = (q + 127) / 128;
  float* thr = tmax + static_cast<size_t>(qtiles) * KNN_SAMPLE_TILES * KNN_GROUPS * 128;
  static const int debug_mode = getenv("RALF_KNN_DEBUG") ? atoi(getenv("RALF_KNN_DEBUG")) : 0;
  const int ns = knn_sample_tiles(n);
  const int gx = knn_grid_x(n);
  CUtensorMap tq;
  rc = make_kmajor_tmap(&tq, queries, 4, d, q, 1, d, 0, 128);
  if (rc) return rc;
  knn_scan_kernel<C, true><<<dim3(ns, qtiles), 192, Cfg::SMEM_BYTES, st>>>(tq, tg, n, d, q, nullptr, nullptr, tmax,
                                                                          debug_mode);
  knn_threshold_kernel<C><<<qtiles * 128, 256, 0, st>>>(tmax, ns, thr);
  knn_scan_kernel<C, false><<<dim3(gx, qtiles), 192, Cfg::SMEM_BYTES, st>>>(tq, tg, n, d, q, cand, thr, nullptr,
                                                                           debug_mode);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e);
  knn_rerank_kernel<C><<<q, 256, 0, st>>>(cand, gx, gallery, queries, n, d, k, index_base, gmax, thr, out_idx, out_score,
                                          certified);
  e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e);
  return 0;
}

extern "C" int ralf_knn_topk(const float* gallery, int n, int d, const float* queries, int q, int k,
                             long long index_base, float gallery_max_norm, long long* out_idx,
                             float* out_score, int* certified, void* workspace, size_t workspace_bytes,
                             void* stream) {
  if (!gallery || !queries || !out_idx || !out_score) return RALF_ERR_NULL;
  if (n <= 0 || q <= 0 || d <= 0 || k <= 0) return RALF_ERR_SHAPE;
  const int C = knn_c_for_k(k);
  if (C == 0) return RALF_ERR_SHAPE;
  if (d % 4) return RALF_ERR_ALIGN;  // TMA needs 16-byte row pitch
  if (!workspace || workspace_bytes < ralf_knn_workspace_bytes(n, d, q, k)) return RALF_ERR_WORKSPACE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (C == 32)
    return knn_topk_impl<32>(gallery, n, d, queries, q, k, index_base, gallery_max_norm, out_idx, out_score,
                             certified, workspace, st);
  return knn_topk_impl<64>(gallery, n, d, queries, q, k, index_base, gallery_max_norm, out_idx, out_score,
                           certified, workspace, st);
}

template <int C>
static int knn_exact_impl(const float* gallery, int n, int d, const float* queries, int q, int k,
                          long long index_base, long long* out_idx, float* out_score, uint64_t* cand,
                          cudaStream_t st) {
  const int parts = knn_exact_slices(C) * 8;
  dim3 grid(knn_exact_slices(C), q);
  knn_exact_scan_kernel<C><<<grid, 256, 0, st>>>(gallery, queries, n, d, cand, nullptr, nullptr);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e);
  knn_rerank_kernel<C><<<q, 256, 0, st>>>(cand, parts, gallery, queries, n, d, k, index_base, 0.f, nullptr, out_idx,
                                          out_score, nullptr);
  return set_cuda_error(cudaGetLastError());
}

extern "C" int ralf_knn_topk_exact(const float* gallery, int n, int d, const float* queries, int q, int k,
                                   long long index_base, long long* out_idx, float* out_score, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  if (!gallery || !queries || !out_idx || !out_score) return RALF_ERR_NULL;
  if (n <= 0 || q <= 0 || d <= 0 || k <= 0) return RALF_ERR_SHAPE;
  const int C = knn_c_for_k(k);
  if (C == 0) return RALF_ERR_SHAPE;
  if (!workspace || workspace_bytes < ralf_knn_workspace_bytes(n, d, q, k)) return RALF_ERR_WORKSPACE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint64_t* cand = reinterpret_cast<uint64_t*>(workspace);
  if (C == 32) return knn_exact_impl<32>(gallery, n, d, queries, q, k, index_base, out_idx, out_score, cand, st);
  return knn_exact_impl<64>(gallery, n, d, queries, q, k, index_base, out_idx, out_score, cand, st);
}

template <int C>
static int knn_fixup_impl(const float* gallery, int n, int d, const float* queries, int q, int k,
                          long long index_base, int* certified, long long* out_idx, float* out_score,
                          void* workspace, cudaStream_t st) {
  uint64_t* cand = reinterpret_cast<uint64_t*>(workspace);
  int* qlist = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(workspace) + knn_front_bytes(n, q, C));
  int* qcount = qlist + q;
  knn_collect_uncertified_kernel<<<1, 1024, 0, st>>>(certified, q, qlist, qcount);
  const int slices = knn_exact_slices(C);
  const int gy = q < 32 ? q : 32;
  knn_exact_scan_kernel<C><<<dim3(slices, gy), 256, 0, st>>>(gallery, queries, n, d, cand, qlist, qcount);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_cuda_error(e);
  knn_rerank_listed_kernel<C><<<q < 128 ? q : 128, 256, 0, st>>>(qlist, qcount, cand, slices * 8, gallery, queries, n, d,
                                                                k, index_base, out_idx, out_score, certified);
  return set_cuda_error(cudaGetLastError());
}

extern "C" int ralf_knn_fixup_exact(const float* gallery, int n, int d, const float* queries, int q, int k,
                                    long long index_base, int* certified, long long* out_idx, float* out_score,
                                    void* workspace, size_t workspace_bytes, void* stream) {
  if (!gallery || !queries || !out_idx || !out_score || !certified) return RALF_ERR_NULL;
  if (n <= 0 || q <= 0 || d <= 0 || k <= 0) return RALF_ERR_SHAPE;
  const int C = knn_c_for_k(k);
  if (C == 0) return RALF_ERR_SHAPE;
  if (!workspace || workspace_bytes < ralf_knn_workspace_bytes(n, d, q, k)) return RALF_ERR_WORKSPACE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (C == 32)
    return knn_fixup_impl<32>(gallery, n, d, queries, q, k, index_base, certified, out_idx, out_score, workspace, st);
  return knn_fixup_impl<64>(gallery, n, d, queries, q, k, index_base, certified, out_idx, out_score, workspace, st);
}

extern "C" int ralf_knn_merge(const float* part_score, const long long* part_idx, int parts, int q, int k,
                              long long* out_idx, float* out_score, void* stream) {
  if (!part_score || !part_idx || !out_idx || !out_score) return RALF_ERR_NULL;
  if (parts <= 0 || q <= 0 || k <= 0) return RALF_ERR_SHAPE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t count = static_cast<size_t>(q) * k;
  knn_fill_kernel<<<static_cast<unsigned>((count + 255) / 256), 256, 0, st>>>(out_idx, out_score, count);
  knn_merge_kernel<<<q, 128, 0, st>>>(part_score, part_idx, parts, q, k, out_idx, out_score);
  return set_cuda_error(cudaGetLastError());
}
