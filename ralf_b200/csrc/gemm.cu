// tcgen05 GEMM for every dense contraction on the RALF hot path (linear layers, 1x1 convs,
// im2col'd 3x3/7x7 convs, K/V projections, LM head):
//
//     D[M,N] = A[M,K] . W[N,K]^T   (+ bias[N]) (activation) (+ residual)   fp32 accumulate in TMEM
//
// Both operands are K-major bf16 ("split" format: plane 0 = hi = bf16(x), plane 1 = lo =
// bf16(x - hi)).  NPASS = 3 issues hi*lo + lo*hi + hi*hi per k-step, which reproduces an fp32
// product to ~2^-17 relative -- that is what lets the logits / token ids match the fp32
// reference (common/common.py:84-135, nn.Linear / nn.Conv2d in fp32).  NPASS = 1 is plain bf16.
//
// Persistent CTAs (grid = min(#tiles, #SMs)) walk the 128 x BN output tiles round-robin (n fastest, so CTAs that
// run together share the A tile in L2).  Two TMEM accumulator buffers: the epilogue of tile i overlaps the
// TMA/MMA main loop of tile i+1.  Warp 0 = TMA producer (SWIZZLE_128B boxes of 64 bf16 along
// K), warp 1 = TMEM allocator + single-thread tcgen05.mma issuer, warps 2-5 = epilogue
// (tcgen05.ld 32x32b, one accumulator row per thread, fused bias / ReLU / GELU / residual /
// bf16 split, direct vectorised global stores).  smem ring of STAGES stages, full/empty
// mbarriers, tcgen05.commit releases stages and signals the epilogue.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "ralf_internal.h"

namespace ralf {

struct GemmEpi {
  const float* bias;               // [N] or null
  const float* res;                // fp32 residual or null
  const __nv_bfloat16* res_split;  // split residual (hi plane; lo at + res_plane) or null
  long long res_plane;
  int res_ld;
  int res_row_mod;  // > 0: residual row = row % res_row_mod (row-broadcast table)
  float* out_f32;   // or null
  __nv_bfloat16* out_split;  // hi plane; lo plane at + out_plane (or null)
  long long out_plane;
  int out_ld;
  int out_col0;
  int rows_per_group, group_stride, group_offset;  // out_row = (r/rpg)*gs + go + r%rpg
  int act;                                         // 0 none, 1 relu, 2 gelu (erf)
  int post_relu;                                   // relu after the residual add
  int vec_ok;                                      // all strides/offsets allow 16-byte accesses
  int split_lo;                                    // write the lo plane too
  uint8_t* out_kv24;                               // 24-bit K/V cache rows (see ralf_b200.h) or null
  long long kv24_ld;                               // bytes per cache row (1536, or 1088 for the 16-bit format)
  int kv_fmt;                                      // 24: 24-bit float; 16: 16-bit integers + fp32 scale per (row, head)
};

// Implicit-GEMM convolution (stride 1 or 2, padding KH / 2): the A operand is never materialised.  k-block kb maps to
// filter tap (kh, kw) = kb / cblocks and input channels 64 * (kb % cblocks) ...; an M tile is a (Wo x BH x NB) brick
// of output positions, loaded per tap as ONE 5-D TMA box over the NHWC activation shifted by (kh - pad, kw - pad) --
// out-of-image coordinates are zero-filled by TMA, which is exactly the convolution's zero padding.
struct ConvGeom {
  int enabled;
  int B, Ho, Wo;      // output positions = input positions (stride 1)
  int BH, NB;         // rows / images per M tile: rows_box = Wo * BH * NB <= 128
  int hblocks;        // ceil(Ho / BH)
  int cblocks;        // C / 64
  int KW, pad;
  int stride;         // 1 or 2: input position = stride * output position + tap - pad (the tensor map's element strides)
};

__device__ __forceinline__ void tma_load_5d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}

__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == 1) return fmaxf(x, 0.f);
  if (act == 2) return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f));
  return x;
}

template <int BN, int NPASS>
struct GemmCfg {
  static constexpr int P = (NPASS == 3) ? 2 : 1;
  static constexpr int A_BYTES = 128 * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = P * (A_BYTES + B_BYTES);
  static constexpr int STAGES_RAW = (192 * 1024) / STAGE_BYTES;
  static constexpr int MAX_STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  // The ring depth is a launch-time choice (min(MAX_STAGES, k-blocks)): short-K GEMMs then take little
  // shared memory and several CTAs share an SM, hiding each other's prologue / epilogue.
  static constexpr int smem_bytes(int stages) {
    return stages * STAGE_BYTES + 1024 /*align*/ + 512 /*barriers*/ + 32768 /*epilogue staging: 8 warps x 4 KB*/;
  }
  static constexpr uint32_t TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;  // two accumulator buffers
  static constexpr uint32_t TMEM_COLS_FOLD = 4 * BN < 32 ? 32 : 4 * BN;  // folded: each buffer is 2 * BN columns
};

// ---- warp-level staging: row-per-thread registers <-> coalesced global memory ------------------------------------
// The TMEM accumulator hands every epilogue thread ONE ROW of the tile (32 fp32 columns per chunk).  Storing / loading
// global memory in that shape makes each warp instruction touch 32 different rows (32 L1 wavefronts per instruction).
// Instead every warp owns a 4 KB shared-memory buffer: rows are written 16 bytes at a time with an XOR swizzle
// (conflict free) and then moved with the lanes running ALONG the rows, so a warp instruction covers full 64 / 128-byte
// row segments.  NV = 16-byte vectors per row chunk: 8 for fp32 (32 columns = 128 B), 4 for one bf16 plane (64 B).
template <int NV>
__device__ __forceinline__ int epi_swz(int row) {
  return NV == 8 ? (row & 7) : (NV == 4 ? ((row >> 1) & 3) : ((row >> 2) & 1));
}

// coalesced global -> registers: pass p moves rows p*(32/NV) .. +32/NV; lane = (row within pass, vector).
// my_base = global address of this lane's own row chunk (0 when the row is out of range).
template <int NV>
__device__ __forceinline__ void epi_load_coalesced(const unsigned long long my_base, const int lane, uint4 (&r)[NV]) {
#pragma unroll
  for (int p = 0; p < NV; ++p) {
    const int rr = p * (32 / NV) + lane / NV;
    const unsigned long long base = __shfl_sync(0xffffffffu, my_base, rr);
    r[p] = make_uint4(0u, 0u, 0u, 0u);
    if (base) r[p] = *reinterpret_cast<const uint4*>(base + static_cast<unsigned long long>(lane % NV) * 16ull);
  }
}
// registers (coalesced order) -> shared -> this thread's own row
template <int NV>
__device__ __forceinline__ void epi_stage_to_row(uint4* buf, const int lane, const uint4 (&r)[NV], uint4 (&mine)[NV]) {
#pragma unroll
  for (int p = 0; p < NV; ++p) {
    const int rr = p * (32 / NV) + lane / NV;
    buf[rr * NV + ((lane % NV) ^ epi_swz<NV>(rr))] = r[p];
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < NV; ++i) mine[i] = buf[lane * NV + (i ^ epi_swz<NV>(lane))];
  __syncwarp();
}
// this thread's own row -> shared -> coalesced global store
template <int NV>
__device__ __forceinline__ void epi_row_to_global(uint4* buf, const int lane, const uint4 (&mine)[NV],
                                                  const unsigned long long my_base) {
#pragma unroll
  for (int i = 0; i < NV; ++i) buf[lane * NV + (i ^ epi_swz<NV>(lane))] = mine[i];
  __syncwarp();
#pragma unroll
  for (int p = 0; p < NV; ++p) {
    const int rr = p * (32 / NV) + lane / NV;
    const unsigned long long base = __shfl_sync(0xffffffffu, my_base, rr);
    const uint4 v = buf[rr * NV + ((lane % NV) ^ epi_swz<NV>(rr))];
    if (base) *reinterpret_cast<uint4*>(base + static_cast<unsigned long long>(lane % NV) * 16ull) = v;
  }
  __syncwarp();
}

// two bf16 planes (hi, lo) at once: plane p lives in buf[p*128 ..], one barrier pair for both
__device__ __forceinline__ void epi_stage_to_row2(uint4* buf, const int lane, const uint4 (&rh)[4], const uint4 (&rl)[4],
                                                  uint4 (&mh)[4], uint4 (&ml)[4]) {
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int rr = p * 8 + lane / 4;
    const int o = rr * 4 + ((lane % 4) ^ epi_swz<4>(rr));
    buf[o] = rh[p];
    buf[128 + o] = rl[p];
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int o = lane * 4 + (i ^ epi_swz<4>(lane));
    mh[i] = buf[o];
    ml[i] = buf[128 + o];
  }
  __syncwarp();
}
__device__ __forceinline__ void epi_row_to_global2(uint4* buf, const int lane, const uint4 (&mh)[4], const uint4 (&ml)[4],
                                                   const unsigned long long base_hi, const unsigned long long base_lo,
                                                   const bool write_lo) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int o = lane * 4 + (i ^ epi_swz<4>(lane));
    buf[o] = mh[i];
    buf[128 + o] = ml[i];
  }
  __syncwarp();
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const int rr = p * 8 + lane / 4;
    const unsigned long long bh = __shfl_sync(0xffffffffu, base_hi, rr);
    const unsigned long long bl = __shfl_sync(0xffffffffu, base_lo, rr);
    const int o = rr * 4 + ((lane % 4) ^ epi_swz<4>(rr));
    const unsigned long long off = static_cast<unsigned long long>(lane % 4) * 16ull;
    if (bh) {
      *reinterpret_cast<uint4*>(bh + off) = buf[o];
      if (write_lo) *reinterpret_cast<uint4*>(bl + off) = buf[128 + o];
    }
  }
  __syncwarp();
}

// Epilogue of columns [c_begin, c_end) of one 128 x BN accumulator tile: each thread owns one row (TMEM lane) and
// walks the columns in chunks of 32.  stg = this warp's 4 KB staging buffer (256 uint4).
template <int BN>
__device__ __forceinline__ void gemm_epilogue_tile(const GemmEpi& ep, const uint32_t tacc, const int quad, const int lane,
                                                   const int m0, const int n0, const int M, const int N, uint4* stg,
                                                   const int c_begin = 0, const int c_end = BN,
                                                   const long long f32_off = 0, const int fold_cols = 0) {
    const int r = m0 + quad * 32 + lane;
    const bool row_ok = r < M;
    const long long out_row =
        (long long)(r / ep.rows_per_group) * ep.group_stride + ep.group_offset + r % ep.rows_per_group;
    const long long res_row = ep.res_row_mod > 0 ? (r % ep.res_row_mod) : r;
    auto gaddr = [&](const void* p) { return row_ok ? reinterpret_cast<unsigned long long>(p) : 0ull; };
#pragma unroll 1
    for (int c = c_begin; c < c_end; c += 32) {
      if (n0 + c >= N) break;
      const int nbase = n0 + c;
      const bool full = ep.vec_ok && (nbase + 32 <= N);  // warp-uniform
      uint32_t v[32];
      tmem_ld_32x32(tacc + (static_cast<uint32_t>(quad * 32) << 16) + c, v);
      float x[32];
      if (fold_cols) {  // folded bf16x3: x_hi.w_hi in column c, (x_hi.w_lo + x_lo.w_hi) in column fold_cols + c
        uint32_t v2[32];
        tmem_ld_32x32(tacc + (static_cast<uint32_t>(quad * 32) << 16) + fold_cols + c, v2);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]) + __uint_as_float(v2[j]);
      } else {
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
      }
      if (full) {
        if (ep.bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(ep.bias + nbase + j));
            x[j] += b.x; x[j + 1] += b.y; x[j + 2] += b.z; x[j + 3] += b.w;
          }
        }
        if (ep.act == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
        } else if (ep.act == 2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = 0.5f * x[j] * (1.f + erff(x[j] * 0.70710678118654752440f));
        }
        if (ep.res) {
          uint4 rq[8], mine[8];
          epi_load_coalesced<8>(gaddr(ep.res + res_row * ep.res_ld + nbase), lane, rq);
          epi_stage_to_row<8>(stg, lane, rq, mine);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            x[4 * j] += __uint_as_float(mine[j].x); x[4 * j + 1] += __uint_as_float(mine[j].y);
            x[4 * j + 2] += __uint_as_float(mine[j].z); x[4 * j + 3] += __uint_as_float(mine[j].w);
          }
        }
        if (ep.res_split) {
          uint4 rqh[4], rql[4], mh[4], ml[4];
          const __nv_bfloat16* rh = ep.res_split + res_row * ep.res_ld + nbase;
          epi_load_coalesced<4>(gaddr(rh), lane, rqh);
          epi_load_coalesced<4>(gaddr(rh + ep.res_plane), lane, rql);
          epi_stage_to_row2(stg, lane, rqh, rql, mh, ml);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint32_t hw[4] = {mh[j].x, mh[j].y, mh[j].z, mh[j].w};
            const uint32_t lw[4] = {ml[j].x, ml[j].y, ml[j].z, ml[j].w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              x[8 * j + 2 * q] += __uint_as_float(hw[q] << 16) + __uint_as_float(lw[q] << 16);
              x[8 * j + 2 * q + 1] +=
                  __uint_as_float(hw[q] & 0xffff0000u) + __uint_as_float(lw[q] & 0xffff0000u);
            }
          }
        }
        if (ep.post_relu) {  // ResNet bottleneck tail: relu(bn3(conv3) + identity)
#pragma unroll
          for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
        }
        if (ep.out_f32) {
          uint4 mine[8];
#pragma unroll
          for (int j = 0; j < 8; ++j)
            mine[j] = make_uint4(__float_as_uint(x[4 * j]), __float_as_uint(x[4 * j + 1]), __float_as_uint(x[4 * j + 2]),
                                 __float_as_uint(x[4 * j + 3]));
          epi_row_to_global<8>(stg, lane, mine, gaddr(ep.out_f32 + f32_off + out_row * ep.out_ld + ep.out_col0 + nbase));
        }
        if (ep.out_kv24 && ep.kv_fmt == 16) {
          // this 32-column chunk is one head of K or V: offset-binary 16-bit integers + one fp32 scale (ralf_b200.h)
          float amax = 0.f;
#pragma unroll
          for (int j = 0; j < 32; ++j) amax = fmaxf(amax, fabsf(x[j]));
          const float inv = amax > 0.f ? 32767.f / amax : 0.f;
          uint4 mq[4];
          uint32_t qw[16];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const uint32_t q0 = static_cast<uint32_t>(__float2int_rn(x[j] * inv) + 32768);
            const uint32_t q1 = static_cast<uint32_t>(__float2int_rn(x[j + 1] * inv) + 32768);
            qw[j >> 1] = (q0 & 0xffffu) | (q1 << 16);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) mq[j] = make_uint4(qw[4 * j], qw[4 * j + 1], qw[4 * j + 2], qw[4 * j + 3]);
          const int part = nbase >> 8, cc = nbase & 255;  // K (0) or V (1) half of the 512 columns
          uint8_t* row = ep.out_kv24 + out_row * ep.kv24_ld;
          epi_row_to_global<4>(stg, lane, mq, gaddr(row + part * 512 + cc * 2));
          if (row_ok) *reinterpret_cast<float*>(row + 1024 + (cc >> 5) * 8 + part * 4) = amax * (1.f / 32767.f);
        } else if (ep.out_kv24) {
          // fp32 rounded to 24 bits: bf16-sized top half (2 B) + one extra mantissa byte -- 3 bytes per value
          uint4 mh[4], ml[2];
          uint32_t hw[16], lb[8];
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const uint32_t u0 = __float_as_uint(x[j]) + 0x80u, u1 = __float_as_uint(x[j + 1]) + 0x80u;
            hw[j >> 1] = (u0 >> 16) | (u1 & 0xffff0000u);
          }
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            uint32_t w = 0;
#pragma unroll
            for (int e = 0; e < 4; ++e) w |= (((__float_as_uint(x[j + e]) + 0x80u) >> 8) & 0xffu) << (8 * e);
            lb[j >> 2] = w;
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) mh[j] = make_uint4(hw[4 * j], hw[4 * j + 1], hw[4 * j + 2], hw[4 * j + 3]);
          ml[0] = make_uint4(lb[0], lb[1], lb[2], lb[3]);
          ml[1] = make_uint4(lb[4], lb[5], lb[6], lb[7]);
          const int part = nbase >> 8, cc = nbase & 255;  // K (0) or V (1) half of the 512 columns
          uint8_t* row = ep.out_kv24 + out_row * ep.kv24_ld;
          epi_row_to_global<4>(stg, lane, mh, gaddr(row + part * 512 + cc * 2));
          epi_row_to_global<2>(stg, lane, ml, gaddr(row + 1024 + part * 256 + cc));
        }
        if (ep.out_split) {
          uint4 mh[4], ml[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              split2_bf16(x[8 * j + 2 * q], x[8 * j + 2 * q + 1], hw[q], lw[q]);
            }
            mh[j] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            ml[j] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
          }
          __nv_bfloat16* oh = ep.out_split + out_row * ep.out_ld + ep.out_col0 + nbase;
          epi_row_to_global2(stg, lane, mh, ml, gaddr(oh), gaddr(oh + ep.out_plane), ep.split_lo != 0);
        }
      } else if (row_ok) {
        // ragged / unaligned tail: scalar, but with STATIC register indices (a dynamically indexed x[] would push the
        // whole accumulator chunk into local memory for the vector path above as well)
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int n = nbase + j;
          if (n < N) {
            float y = x[j];
            if (ep.bias) y += __ldg(ep.bias + n);
            if (ep.act) y = apply_act(y, ep.act);
            if (ep.res) y += ep.res[res_row * ep.res_ld + n];
            if (ep.res_split) {
              y += __bfloat162float(ep.res_split[res_row * ep.res_ld + n]) +
                   __bfloat162float(ep.res_split[ep.res_plane + res_row * ep.res_ld + n]);
            }
            if (ep.post_relu) y = fmaxf(y, 0.f);
            if (ep.out_f32) ep.out_f32[f32_off + out_row * ep.out_ld + ep.out_col0 + n] = y;
            if (ep.out_split) {
              __nv_bfloat16 h, l;
              split_bf16(y, h, l);
              ep.out_split[out_row * ep.out_ld + ep.out_col0 + n] = h;
              if (ep.split_lo) ep.out_split[ep.out_plane + out_row * ep.out_ld + ep.out_col0 + n] = l;
            }
          }
        }
      }
    }
}

// FOLD (NPASS = 3, BN <= 128): two MMAs per k-step instead of three.  The W stage holds its hi rows directly followed by
// its lo rows, so ONE descriptor with N = 2*BN multiplies x_hi with [w_hi ; w_lo]: columns [0, BN) of the accumulator
// collect x_hi.w_hi, columns [BN, 2BN) collect x_hi.w_lo; the second MMA (N = BN) adds x_lo.w_hi into columns [BN, 2BN).
// The epilogue adds the two halves.  Why: an SS-mode tcgen05.mma at M = 128 costs ~92 cycles however narrow N is
// (profiles/r2_decode_chain.md), so at BN <= 128 (<= 64 cycles of math per MMA) the kernel is bound by the NUMBER of MMAs.
// (The round-1 MINB = 2 variant -- two CTAs per SM at 96 registers -- was measured 4 % slower in round 2 and is retired.)
template <int BN, int NPASS, int FOLD>
__global__ void __launch_bounds__(320, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const GemmEpi ep, const int M, const int N, const int K, const int STAGES, const ConvGeom cg,
                 const int splits, const long long split_stride) {
  using Cfg = GemmCfg<BN, NPASS>;
  constexpr int P = Cfg::P;
  pdl_trigger();  // the next kernel of the stream may become resident and run its own prologue while this one works
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base_u32 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base_u32 - smem_u32(smem_raw));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;  // [2] accumulator ready
  uint64_t* tempty_bar = tfull_bar + 2;      // [2] accumulator drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  uint4* stage_all = reinterpret_cast<uint4*>(smem + STAGES * Cfg::STAGE_BYTES + 512);  // 4 x 4 KB, 16-byte aligned

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_n = (N + BN - 1) / BN;
  const int tiles_m = cg.enabled ? ((cg.B + cg.NB - 1) / cg.NB) * cg.hblocks : (M + 127) / 128;
  // split-K (weight-gradient GEMMs: tiny M x N, K = all rows of the batch): tile = (mn tile, k slice); slice s writes
  // its raw fp32 partial to out_f32 + s * split_stride, a second kernel sums the slices (deterministic).
  const int num_tiles = tiles_m * tiles_n * splits;
  const int nkb = (K + 63) / 64;
  const int nkb_s = (nkb + splits - 1) / splits;
  const uint32_t a_tx = cg.enabled ? static_cast<uint32_t>(cg.Wo * cg.BH * cg.NB) * 128u : Cfg::A_BYTES;
  const uint32_t stage_tx = P * (a_tx + Cfg::B_BYTES);
  // first output row and row limit of M tile mt
  auto tile_rows = [&](const int mt, int& m0, int& m_end) {
    if (cg.enabled) {
      const int g = mt / cg.hblocks, hb = mt - g * cg.hblocks;
      const int b0 = g * cg.NB, h0 = hb * cg.BH;
      m0 = (b0 * cg.Ho + h0) * cg.Wo;
      const int valid = cg.NB > 1 ? min(cg.NB, cg.B - b0) * cg.Ho * cg.Wo : min(cg.BH, cg.Ho - h0) * cg.Wo;
      m_end = m0 + valid;
    } else {
      m0 = mt * 128;
      m_end = M;
    }
  };
  // one pipeline stage: A (both planes) + W (both planes) of k-block kb for tile (mt, n0)
  auto load_stage = [&](const int mt, const int n0, const int kb, const int s) {
    mbar_expect_tx(&full_bar[s], stage_tx);
    uint8_t* st = smem + s * Cfg::STAGE_BYTES;
    if (cg.enabled) {
      const int g = mt / cg.hblocks, hb = mt - g * cg.hblocks;
      const int tap = kb / cg.cblocks, cb = kb - tap * cg.cblocks;
      const int kh = tap / cg.KW, kw = tap - kh * cg.KW;
#pragma unroll
      for (int p = 0; p < P; ++p)
        tma_load_5d(&tmA, &full_bar[s], st + p * Cfg::A_BYTES, cb * 64, kw - cg.pad,
                    hb * cg.BH * cg.stride + kh - cg.pad, g * cg.NB, p);
    } else {
#pragma unroll
      for (int p = 0; p < P; ++p) tma_load_3d(&tmA, &full_bar[s], st + p * Cfg::A_BYTES, kb * 64, mt * 128, p);
    }
#pragma unroll
    for (int p = 0; p < P; ++p)
      tma_load_3d(&tmB, &full_bar[s], st + P * Cfg::A_BYTES + p * Cfg::B_BYTES, kb * 64, n0, p);
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], BN >= 64 ? 256 : 128);
    }
    fence_mbar_init();
    pdl_wait();  // PDL: the A operand is the previous kernel's output (everything above touched no global memory)
    // Prologue prefetch: the first ring of TMA loads needs nothing but the barriers this thread just initialised,
    // so it is issued BEFORE the TMEM allocation / block sync below (hides ~0.5 us on latency-bound decode GEMMs).
    {
      const int mn = blockIdx.x / splits, kb0 = (blockIdx.x % splits) * nkb_s;
      const int mt = mn / tiles_n, n0 = (mn % tiles_n) * BN;
      const int nk = min(nkb, kb0 + nkb_s) - kb0;
      const int npre = nk < STAGES ? nk : STAGES;
      for (int kb = 0; kb < npre; ++kb) load_stage(mt, n0, kb0 + kb, kb);
    }
  }
  constexpr uint32_t TCOLS = FOLD ? Cfg::TMEM_COLS_FOLD : Cfg::TMEM_COLS;
  constexpr uint32_t ACC_COLS = FOLD ? 2 * BN : BN;  // columns per accumulator buffer
  if (warp == 1) tmem_alloc<TCOLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();  // every thread: epilogue operands (residual) and outputs are ordered after the previous kernel too
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mn = tile / splits, kb0 = (tile % splits) * nkb_s;
        const int mt = mn / tiles_n, n0 = (mn % tiles_n) * BN;
        const int kb1 = min(nkb, kb0 + nkb_s);
        const int npre = (kb1 - kb0) < STAGES ? (kb1 - kb0) : STAGES;
        for (int kb = kb0; kb < kb1; ++kb) {
          if (tile == static_cast<int>(blockIdx.x) && kb - kb0 < npre) {  // already issued in the prologue
            if (++s == STAGES) { s = 0; ph ^= 1; }
            continue;
          }
          mbar_wait(&empty_bar[s], ph ^ 1);
          load_stage(mt, n0, kb, s);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(1, 128, BN);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(&tempty_bar[buf], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + buf * ACC_COLS;
        const int kb0 = (tile % splits) * nkb_s, kb1 = min(nkb, kb0 + nkb_s);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_hi = base_u32 + s * Cfg::STAGE_BYTES;
          const uint32_t b_hi = a_hi + P * Cfg::A_BYTES;
          const uint64_t da_hi = make_sw128_kmajor_desc(a_hi);
          const uint64_t db_hi = make_sw128_kmajor_desc(b_hi);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t koff = static_cast<uint64_t>(2 * k);  // 32 bytes >> 4 per UMMA_K = 16
            if (NPASS == 3 && FOLD) {
              constexpr uint32_t idesc2 = make_idesc(1, 128, 2 * BN);
              const uint64_t da_lo = make_sw128_kmajor_desc(a_hi + Cfg::A_BYTES);
              mma_bf16_ss(tacc, da_hi + koff, db_hi + koff, idesc2, ((kb - kb0) | k) != 0);  // x_hi . [w_hi ; w_lo]
              mma_bf16_ss(tacc + BN, da_lo + koff, db_hi + koff, idesc, 1);                  // + x_lo . w_hi
            } else if (NPASS == 3) {
              const uint64_t da_lo = make_sw128_kmajor_desc(a_hi + Cfg::A_BYTES);
              const uint64_t db_lo = make_sw128_kmajor_desc(b_hi + Cfg::B_BYTES);
              mma_bf16_ss(tacc, da_hi + koff, db_lo + koff, idesc, ((kb - kb0) | k) != 0);
              mma_bf16_ss(tacc, da_lo + koff, db_hi + koff, idesc, 1);
              mma_bf16_ss(tacc, da_hi + koff, db_hi + koff, idesc, 1);
            } else {
              mma_bf16_ss(tacc, da_hi + koff, db_hi + koff, idesc, ((kb - kb0) | k) != 0);
            }
          }
          tc_commit(&empty_bar[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        tc_commit(&tfull_bar[buf]);
      }
    }
  } else {
    // ------------------------------- epilogue ---------------------------------------------
    // Eight epilogue warps: warp w may only touch TMEM lanes 32*(w%4)..; warps 2-5 take the first half of the tile's
    // columns, warps 6-9 the second half (BN = 32: one chunk, the second group idles).
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    constexpr int CH = BN >= 64 ? BN / 2 : BN;
    int it = 0;
    if (BN >= 64 || half == 0)
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
    const int buf = it & 1;
    const int mn = tile / splits;
    const int n0 = (mn % tiles_n) * BN;
    int m0, m_end;
    tile_rows(mn / tiles_n, m0, m_end);
    mbar_wait(&tfull_bar[buf], (it >> 1) & 1);
    tc_fence_after();
    const uint32_t tacc = tmem_base + buf * ACC_COLS;
    gemm_epilogue_tile<BN>(ep, tacc, quad, lane, m0, n0, m_end, N, stage_all + (warp - 2) * 256, half * CH, half * CH + CH,
                           static_cast<long long>(tile % splits) * split_stride, FOLD ? BN : 0);
    tc_fence_before();
    mbar_arrive(&tempty_bar[buf]);
    }  // tile loop
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TCOLS>(tmem_base);
  }
}


// ------------------------------------------------------------------------------------------------
// Short-K bottleneck-tail GEMMs (ResNet conv3 / downsample: K <= 256, N a multiple of 128, split bf16 output, optional
// split bf16 residual + ReLU): the same producer / MMA warps as gemm_bf16_kernel<128, 3, 1>, but the EPILOGUE moves its
// global traffic with TMA.  These GEMMs are bound by the residual read + output write (K = 64: 1.2 GB per 128 canvases
// against 17 GFLOP x 3), and the register-staged epilogue above keeps only one 4 KB chunk of loads in flight per warp,
// then one chunk of stores: 0.45-0.5 of the HBM peak.  Here every epilogue warp owns a ring of NBUF 4 KB chunk buffers
// ([2 planes][32 rows][32 columns] bf16, SWIZZLE_64B = the XOR pattern of epi_swz<4>, so the row-per-thread accesses
// are conflict free): the residual of chunk g + NBUF - 1 is fetched by cp.async.bulk.tensor while chunk g is computed
// IN PLACE in its buffer and chunk g - 1 drains through a bulk tensor store.  Arithmetic and its order are those of
// gemm_epilogue_tile, so the results are bit-identical to the register-staged path.
// F32 = 1: the same pipeline for fp32-output GEMMs with an optional fp32 residual (transformer out-projections / qkv):
// a chunk is [32 rows][32 columns] fp32 = 128-byte rows, SWIZZLE_128B (piece j of row r at j ^ (r & 7)), 2-D tensor maps.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <int BN_>
struct TepiCfg {
  static constexpr int BN = BN_;                                   // 128 (64 was measured for the 64-channel layers: no gain)
  static constexpr int STAGE_BYTES = 2 * (128 * 128 + BN * 128);  // 64 / 48 KB: A hi/lo + W hi/lo of one k-block
  static constexpr int CHUNK_BYTES = 4096;                         // [2][32][32] bf16
  static constexpr int smem_bytes(int stages, int nbuf) {
    return stages * STAGE_BYTES + 8 * nbuf * CHUNK_BYTES + 1024 /*align*/ + 1024 /*barriers*/;
  }
};

template <int BN, int NBUF, int F32>
__global__ void __launch_bounds__(320, 1)
gemm_bf16_tepi_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmO,
                      const float* __restrict__ bias, const int act, const int has_res, const int post_relu, const int M,
                      const int N, const int K, const int STAGES, const ConvGeom cg) {
  using TCfg = TepiCfg<BN>;
  constexpr int CPT = BN / 64;  // 32-column chunks per tile and epilogue warp
  constexpr int A_BYTES = 128 * 128, B_BYTES = BN * 128;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base_u32 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base_u32 - smem_u32(smem_raw));
  uint8_t* chunks = smem + STAGES * TCfg::STAGE_BYTES;  // 8 warps x NBUF x 4 KB, 1024-byte aligned
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(chunks + 8 * NBUF * TCfg::CHUNK_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* res_bar = tempty_bar + 2;  // [8 warps][NBUF]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 8 * NBUF);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int tiles_n = N / BN;
  const int num_tiles = ((M + 127) / 128) * tiles_n;
  const int nkb = (K + 63) / 64;
  auto load_stage = [&](const int mt, const int n0, const int kb, const int s) {
    mbar_expect_tx(&full_bar[s], TCfg::STAGE_BYTES);
    uint8_t* st = smem + s * TCfg::STAGE_BYTES;
    if (cg.enabled) {  // implicit convolution whose M tiles are all full: tile mt = output rows [128 mt, 128 mt + 128)
      const int g = mt / cg.hblocks, hb = mt - g * cg.hblocks;
      const int tap = kb / cg.cblocks, cb = kb - tap * cg.cblocks;
      const int kh = tap / cg.KW, kw = tap - kh * cg.KW;
#pragma unroll
      for (int p = 0; p < 2; ++p)
        tma_load_5d(&tmA, &full_bar[s], st + p * A_BYTES, cb * 64, kw - cg.pad, hb * cg.BH * cg.stride + kh - cg.pad,
                    g * cg.NB, p);
    } else {
#pragma unroll
      for (int p = 0; p < 2; ++p) tma_load_3d(&tmA, &full_bar[s], st + p * A_BYTES, kb * 64, mt * 128, p);
    }
#pragma unroll
    for (int p = 0; p < 2; ++p) tma_load_3d(&tmB, &full_bar[s], st + 2 * A_BYTES + p * B_BYTES, kb * 64, n0, p);
  };

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmR);
    tma_prefetch_desc(&tmO);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], 256);
    }
    for (int i = 0; i < 8 * NBUF; ++i) mbar_init(&res_bar[i], 1);
    fence_mbar_init();
    pdl_wait();
    const int mt = blockIdx.x / tiles_n, n0 = (blockIdx.x % tiles_n) * BN;
    const int npre = nkb < STAGES ? nkb : STAGES;
    for (int kb = 0; kb < npre; ++kb) load_stage(mt, n0, kb, kb);
  }
  constexpr uint32_t TCOLS = 4 * BN;   // two folded accumulator buffers of 2 * BN columns
  constexpr uint32_t ACC_COLS = 2 * BN;
  if (warp == 1) tmem_alloc<TCOLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const int npre = nkb < STAGES ? nkb : STAGES;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mt = tile / tiles_n, n0 = (tile % tiles_n) * BN;
        for (int kb = 0; kb < nkb; ++kb) {
          if (tile == static_cast<int>(blockIdx.x) && kb < npre) {  // issued in the prologue
            if (++s == STAGES) { s = 0; ph ^= 1; }
            continue;
          }
          mbar_wait(&empty_bar[s], ph ^ 1);
          load_stage(mt, n0, kb, s);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(1, 128, BN);
      constexpr uint32_t idesc2 = make_idesc(1, 128, 2 * BN);
      int s = 0;
      uint32_t ph = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        mbar_wait(&tempty_bar[buf], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tacc = tmem_base + buf * ACC_COLS;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_hi = base_u32 + s * TCfg::STAGE_BYTES;
          const uint32_t b_hi = a_hi + 2 * A_BYTES;
          const uint64_t da_hi = make_sw128_kmajor_desc(a_hi);
          const uint64_t da_lo = make_sw128_kmajor_desc(a_hi + A_BYTES);
          const uint64_t db_hi = make_sw128_kmajor_desc(b_hi);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t koff = static_cast<uint64_t>(2 * k);
            mma_bf16_ss(tacc, da_hi + koff, db_hi + koff, idesc2, (kb | k) != 0);  // x_hi . [w_hi ; w_lo]
            mma_bf16_ss(tacc + BN, da_lo + koff, db_hi + koff, idesc, 1);          // + x_lo . w_hi
          }
          tc_commit(&empty_bar[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        tc_commit(&tfull_bar[buf]);
      }
    }
  } else {
    // Epilogue warp ew: TMEM lanes / tile rows 32 * (ew % 4) .., columns (BN / 2) * (ew / 4) .. + BN / 2 of every tile,
    // as CPT chunks of 32 columns; chunk g of this warp lives in buffer g % NBUF.  Tile coordinates advance by the
    // grid stride without divisions (they would cost more than the chunk's arithmetic).
    const int ew = warp - 2;
    const int quad = warp & 3;
    const int half = ew >> 2;
    uint8_t* cb = chunks + ew * NBUF * TCfg::CHUNK_BYTES;
    uint64_t* rb = res_bar + ew * NBUF;
    const int my_tiles = static_cast<int>(blockIdx.x) < num_tiles ? (num_tiles - 1 - static_cast<int>(blockIdx.x)) / static_cast<int>(gridDim.x) + 1 : 0;
    const int total = CPT * my_tiles;
    const int step_m = static_cast<int>(gridDim.x) / tiles_n, step_n = static_cast<int>(gridDim.x) % tiles_n;
    struct Pos { int mt, nt, c; };
    auto advance = [&](Pos& p) {  // next chunk of this warp
      if (++p.c == CPT) {
        p.c = 0;
        p.mt += step_m;
        p.nt += step_n;
        if (p.nt >= tiles_n) { p.nt -= tiles_n; ++p.mt; }
      }
    };
    Pos cur{static_cast<int>(blockIdx.x) / tiles_n, static_cast<int>(blockIdx.x) % tiles_n, 0};
    Pos pre = cur;  // position of the next residual chunk to request (lane 0)
    int gpre = 0;
    auto issue_res = [&]() {  // lane 0 only: request chunk gpre
      if (gpre < total) {
        mbar_expect_tx(&rb[gpre % NBUF], TCfg::CHUNK_BYTES);
        if constexpr (F32)
          tma_load_2d(&tmR, &rb[gpre % NBUF], cb + (gpre % NBUF) * TCfg::CHUNK_BYTES, pre.nt * BN + half * (BN / 2) + pre.c * 32,
                      pre.mt * 128 + quad * 32);
        else
          tma_load_3d(&tmR, &rb[gpre % NBUF], cb + (gpre % NBUF) * TCfg::CHUNK_BYTES, pre.nt * BN + half * (BN / 2) + pre.c * 32,
                      pre.mt * 128 + quad * 32, 0);
        advance(pre);
      }
      ++gpre;
    };
    if (has_res && lane == 0) {
      for (int g = 0; g < NBUF - 1; ++g) issue_res();
    }
    // SWIZZLE_64B (split bf16 chunk, 64-byte rows): 16-byte piece j of row r sits at piece j ^ ((r >> 1) & 3);
    // SWIZZLE_128B (fp32 chunk, 128-byte rows): at piece j ^ (r & 7)
    const int sw = F32 ? (lane & 7) : ((lane >> 1) & 3);
    for (int g = 0; g < total; ++g, advance(cur)) {
      const int it = g / CPT, c = g % CPT, buf = it & 1;
      const int col = cur.nt * BN + half * (BN / 2) + c * 32, row = cur.mt * 128 + quad * 32;
      if (c == 0) {
        mbar_wait(&tfull_bar[buf], (it >> 1) & 1);
        tc_fence_after();
      }
      const uint32_t tacc = tmem_base + buf * ACC_COLS + (static_cast<uint32_t>(quad * 32) << 16) + half * (BN / 2) + c * 32;
      uint32_t v[32], v2[32];
      tmem_ld_32x32(tacc, v);
      tmem_ld_32x32(tacc + BN, v2);
      tmem_ld_wait();
      if (c == CPT - 1) {  // the accumulator buffer is drained: hand it back to the MMA warp before the memory work
        tc_fence_before();
        mbar_arrive(&tempty_bar[buf]);
      }
      float x[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]) + __uint_as_float(v2[j]);
      if (bias) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(bias + col + j));
          x[j] += b.x; x[j + 1] += b.y; x[j + 2] += b.z; x[j + 3] += b.w;
        }
      }
      if (act == 1) {
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
      }
      uint4* bq = reinterpret_cast<uint4*>(cb + (g % NBUF) * TCfg::CHUNK_BYTES);
      if (has_res && F32) {
        mbar_wait(&rb[g % NBUF], (g / NBUF) & 1);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint4 t = bq[lane * 8 + (j ^ sw)];
          x[4 * j] += __uint_as_float(t.x); x[4 * j + 1] += __uint_as_float(t.y);
          x[4 * j + 2] += __uint_as_float(t.z); x[4 * j + 3] += __uint_as_float(t.w);
        }
      } else if (has_res) {
        mbar_wait(&rb[g % NBUF], (g / NBUF) & 1);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const uint4 mh = bq[lane * 4 + (j ^ sw)], ml = bq[128 + lane * 4 + (j ^ sw)];
          const uint32_t hw[4] = {mh.x, mh.y, mh.z, mh.w};
          const uint32_t lw[4] = {ml.x, ml.y, ml.z, ml.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            x[8 * j + 2 * q] += __uint_as_float(hw[q] << 16) + __uint_as_float(lw[q] << 16);
            x[8 * j + 2 * q + 1] += __uint_as_float(hw[q] & 0xffff0000u) + __uint_as_float(lw[q] & 0xffff0000u);
          }
        }
      } else {
        if (lane == 0) bulk_wait_read<NBUF - 1>();  // the store that last used this buffer (chunk g - NBUF) has read it
        __syncwarp();
      }
      if (post_relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
      }
      if constexpr (F32) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          bq[lane * 8 + (j ^ sw)] = make_uint4(__float_as_uint(x[4 * j]), __float_as_uint(x[4 * j + 1]), __float_as_uint(x[4 * j + 2]),
                                               __float_as_uint(x[4 * j + 3]));
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t hw[4], lw[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            split2_bf16(x[8 * j + 2 * q], x[8 * j + 2 * q + 1], hw[q], lw[q]);
          }
          bq[lane * 4 + (j ^ sw)] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
          bq[128 + lane * 4 + (j ^ sw)] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
        }
      }
      fence_proxy_async_smem();  // generic-proxy writes -> visible to the bulk store
      __syncwarp();
      if (lane == 0) {
        if constexpr (F32) tma_store_2d(&tmO, bq, col, row);
        else tma_store_3d(&tmO, bq, col, row, 0);
        bulk_commit();
        if (has_res) {
          bulk_wait_read<1>();  // chunk g - 1's store has read its buffer: refill it with the residual of g + NBUF - 1
          issue_res();
        }
      }
      __syncwarp();
    }
    if (lane == 0) bulk_wait_all();  // shared memory must outlive the stores
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TCOLS>(tmem_base);
  }
}


// ------------------------------------------------------------------------------------------------
// Decode-loop residual GEMM with the FOLLOWING LayerNorm in its epilogue (round 2): x_new = A . W^T + bias + x (N = 256 =
// d_model) and h = LayerNorm(x_new) as split operand rows for the next GEMM.  A full row spans the eight 32-column
// n-tiles, which keep running on eight different SMs (a single CTA per row block would serialise the MMAs, see
// decode_chain.cu); the eight CTAs form a THREAD-BLOCK CLUSTER and exchange their per-row partial sums through
// distributed shared memory: sum -> cluster barrier -> mean; centred sum of squares -> cluster barrier -> rstd (the
// two-pass form of layernorm_kernel; partials are added in rank order, so every CTA computes the same statistics).
// One tile per CTA, grid = (8, M tiles), cluster = (8, 1, 1).  Replaces 18 of the 19 LayerNorm launches of a decode step.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ float ld_dsmem_f32(const float* local_ptr, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local_ptr)), "r"(rank));
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
  return v;
}

constexpr int RLN_STAGE_BYTES = 2 * (128 * 128 + 32 * 128);  // 40 KB: A hi/lo (128 rows) + W hi/lo (32 rows) of one k-block
constexpr int RLN_STAGES = 4;
constexpr int RLN_SMEM = RLN_STAGES * RLN_STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/ + 2 * 128 * 4 /*row partials*/;

__global__ void __launch_bounds__(192, 1)
gemm_resln_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  const float* __restrict__ bias, const float* res, const int res_ld, float* out_f32, const int out_ld,
                  const float* __restrict__ gamma, const float* __restrict__ beta, const float eps,
                  __nv_bfloat16* __restrict__ ln_split, const long long ln_plane, const int M, const int K) {
  constexpr int BN = 32, A_BYTES = 128 * 128, B_BYTES = BN * 128;
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base_u32 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base_u32 - smem_u32(smem_raw));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + RLN_STAGES * RLN_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + RLN_STAGES;
  uint64_t* tfull_bar = empty_bar + RLN_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);
  float* part1 = reinterpret_cast<float*>(smem + RLN_STAGES * RLN_STAGE_BYTES + 256);  // [128] row sums of this n-tile
  float* part2 = part1 + 128;                                                            // [128] centred sums of squares
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * 128;
  const int nkb = (K + 63) / 64;

  auto load_stage = [&](const int kb, const int s) {
    mbar_expect_tx(&full_bar[s], RLN_STAGE_BYTES);
    uint8_t* st = smem + s * RLN_STAGE_BYTES;
#pragma unroll
    for (int p = 0; p < 2; ++p) tma_load_3d(&tmA, &full_bar[s], st + p * A_BYTES, kb * 64, m0, p);
#pragma unroll
    for (int p = 0; p < 2; ++p) tma_load_3d(&tmB, &full_bar[s], st + 2 * A_BYTES + p * B_BYTES, kb * 64, n0, p);
  };
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int s = 0; s < RLN_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tfull_bar, 1);
    fence_mbar_init();
    pdl_wait();
    const int npre = nkb < RLN_STAGES ? nkb : RLN_STAGES;
    for (int kb = 0; kb < npre; ++kb) load_stage(kb, kb);
  }
  if (warp == 1) tmem_alloc<64>(tmem_slot);  // one folded accumulator: 2 * BN columns
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_wait();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        if (kb >= RLN_STAGES) {  // the first ring was issued in the prologue
          mbar_wait(&empty_bar[s], ph ^ 1);
          load_stage(kb, s);
        }
        if (++s == RLN_STAGES) { s = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(1, 128, BN);
      constexpr uint32_t idesc2 = make_idesc(1, 128, 2 * BN);
      int s = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_hi = base_u32 + s * RLN_STAGE_BYTES;
        const uint64_t da_hi = make_sw128_kmajor_desc(a_hi);
        const uint64_t da_lo = make_sw128_kmajor_desc(a_hi + A_BYTES);
        const uint64_t db_hi = make_sw128_kmajor_desc(a_hi + 2 * A_BYTES);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t koff = static_cast<uint64_t>(2 * k);
          mma_bf16_ss(tmem_base, da_hi + koff, db_hi + koff, idesc2, (kb | k) != 0);  // x_hi . [w_hi ; w_lo]
          mma_bf16_ss(tmem_base + BN, da_lo + koff, db_hi + koff, idesc, 1);          // + x_lo . w_hi
        }
        tc_commit(&empty_bar[s]);
        if (++s == RLN_STAGES) { s = 0; ph ^= 1; }
      }
      tc_commit(tfull_bar);
    }
    __syncwarp();
  }
  // ---- epilogue: warps 2-5, thread = row (TMEM lane quadrant warp % 4); the other warps only join the cluster barriers ----
  const bool epi = warp >= 2;
  const int quad = warp & 3;
  const int rl = quad * 32 + lane;  // row within the tile
  const int r = m0 + rl;
  const bool row_ok = epi && r < M;
  float x[32];
  if (epi) {
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    uint32_t v[32], v2[32];
    const uint32_t tacc = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    tmem_ld_32x32(tacc, v);
    tmem_ld_32x32(tacc + BN, v2);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]) + __uint_as_float(v2[j]);
    if (bias) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(bias + n0 + j));
        x[j] += b.x; x[j + 1] += b.y; x[j + 2] += b.z; x[j + 3] += b.w;
      }
    }
    float s = 0.f;
    if (row_ok) {
      const float* rr = res + static_cast<long long>(r) * res_ld + n0;
      float* oo = out_f32 + static_cast<long long>(r) * out_ld + n0;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 t = *reinterpret_cast<const float4*>(rr + j);
        x[j] += t.x; x[j + 1] += t.y; x[j + 2] += t.z; x[j + 3] += t.w;
      }
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(oo + j) = make_float4(x[j], x[j + 1], x[j + 2], x[j + 3]);
#pragma unroll
      for (int j = 0; j < 32; ++j) s += x[j];
    }
    part1[rl] = s;
  }
  cluster_sync_all();
  float mean = 0.f;
  if (epi) {
    float t = 0.f;
#pragma unroll
    for (uint32_t rk = 0; rk < 8; ++rk) t += ld_dsmem_f32(part1 + rl, rk);
    mean = t * (1.f / 256.f);
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float d = x[j] - mean;
      sq = fmaf(d, d, sq);
    }
    part2[rl] = row_ok ? sq : 0.f;
  }
  cluster_sync_all();
  if (epi) {
    float t = 0.f;
#pragma unroll
    for (uint32_t rk = 0; rk < 8; ++rk) t += ld_dsmem_f32(part2 + rl, rk);
    const float rstd = rsqrtf(t * (1.f / 256.f) + eps);
    if (row_ok) {
      uint32_t hw[16], lw[16];
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma + n0 + j));
        const float4 bt = __ldg(reinterpret_cast<const float4*>(beta + n0 + j));
        const float y0 = (x[j] - mean) * rstd * g.x + bt.x, y1 = (x[j + 1] - mean) * rstd * g.y + bt.y;
        const float y2 = (x[j + 2] - mean) * rstd * g.z + bt.z, y3 = (x[j + 3] - mean) * rstd * g.w + bt.w;
        split2_bf16(y0, y1, hw[j >> 1], lw[j >> 1]);
        split2_bf16(y2, y3, hw[(j >> 1) + 1], lw[(j >> 1) + 1]);
      }
      __nv_bfloat16* oh = ln_split + static_cast<long long>(r) * 256 + n0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        *reinterpret_cast<uint4*>(oh + 8 * j) = make_uint4(hw[4 * j], hw[4 * j + 1], hw[4 * j + 2], hw[4 * j + 3]);
        *reinterpret_cast<uint4*>(oh + ln_plane + 8 * j) = make_uint4(lw[4 * j], lw[4 * j + 1], lw[4 * j + 2], lw[4 * j + 3]);
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();  // no CTA of the cluster leaves while its partials may still be read
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<64>(tmem_base);
  }
}


// ------------------------------------------------------------------------------------------------
// LayerNorm fused into the GEMM's A operand (decode path, K = 256 = d_model; one CTA per (n-tile, 128-row tile)).
// Measured at the bench shape (M = 1024) in round 2: 188.2 vs 168.5 ms per step -- the serial per-thread row
// normalisation sits on every GEMM's critical path and costs more than the 5 us LayerNorm launch it removes.  Opt-in.
//   D[M,N] = LN(x)[M,256] . W[N,256]^T  (+ epilogue)      x fp32 (row stride ldx), gamma/beta fp32 [256]
// The 128 epilogue threads first normalise one row each (two sweeps of independent LDG.128s), split the
// result into bf16 hi/lo and store it straight into shared memory in the K-major SWIZZLE_128B
// layout the UMMA descriptor expects (16-byte chunk index XOR (row & 7)) -- exactly what a TMA load of a
// pre-normalised tensor would have produced, minus the separate LayerNorm kernel and its HBM round trip.
// W streams in by TMA (all 4 k-blocks issued up front).  One CTA per n-tile.
// Replaces nn.LayerNorm + nn.Linear pairs of the pre-LN decoder layers (common/common.py:26-41).
// ------------------------------------------------------------------------------------------------
template <int BN>
__global__ void __launch_bounds__(192, 1)
gemm_ln_kernel(const __grid_constant__ CUtensorMap tmB, const float* __restrict__ x, const int ldx,
               const float* __restrict__ gamma, const float* __restrict__ beta, const float eps, const GemmEpi ep,
               const int M, const int N) {
  constexpr int KB = 4;                      // K = 256 -> 4 k-blocks of 64
  constexpr int A_PLANE = 128 * 128;         // 16 KB per k-block per plane
  constexpr int A_BYTES = KB * 2 * A_PLANE;  // 128 KB
  constexpr int B_PLANE = BN * 128;
  constexpr uint32_t TMEM_COLS = BN < 32 ? 32 : BN;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base_u32 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base_u32 - smem_u32(smem_raw));
  uint8_t* smem_b = smem + A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + KB * 2 * B_PLANE);
  uint64_t* aready_bar = full_bar + KB;
  uint64_t* tfull_bar = aready_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);
  uint4* stage_all = reinterpret_cast<uint4*>(smem_b + KB * 2 * B_PLANE + 256);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmB);
    for (int kb = 0; kb < KB; ++kb) mbar_init(&full_bar[kb], 1);
    mbar_init(aready_bar, 128);
    mbar_init(tfull_bar, 1);
    fence_mbar_init();
    for (int kb = 0; kb < KB; ++kb) {
      mbar_expect_tx(&full_bar[kb], 2 * B_PLANE);
      tma_load_3d(&tmB, &full_bar[kb], smem_b + (kb * 2 + 0) * B_PLANE, kb * 64, n0, 0);
      tma_load_3d(&tmB, &full_bar[kb], smem_b + (kb * 2 + 1) * B_PLANE, kb * 64, n0, 1);
    }
  }
  if (warp == 1) tmem_alloc<TMEM_COLS>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(1, 128, BN);
      mbar_wait(aready_bar, 0);
      tc_fence_after();
      for (int kb = 0; kb < KB; ++kb) {
        mbar_wait(&full_bar[kb], 0);
        tc_fence_after();
        const uint32_t a_hi = base_u32 + kb * 2 * A_PLANE;
        const uint32_t b_hi = base_u32 + A_BYTES + kb * 2 * B_PLANE;
        const uint64_t da_hi = make_sw128_kmajor_desc(a_hi), da_lo = make_sw128_kmajor_desc(a_hi + A_PLANE);
        const uint64_t db_hi = make_sw128_kmajor_desc(b_hi), db_lo = make_sw128_kmajor_desc(b_hi + B_PLANE);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t koff = static_cast<uint64_t>(2 * k);
          mma_bf16_ss(tmem_base, da_hi + koff, db_lo + koff, idesc, (kb | k) != 0);
          mma_bf16_ss(tmem_base, da_lo + koff, db_hi + koff, idesc, 1);
          mma_bf16_ss(tmem_base, da_hi + koff, db_hi + koff, idesc, 1);
        }
      }
      tc_commit(tfull_bar);
    }
  } else if (warp >= 2) {
    // ---- LayerNorm prologue: one thread per row (128 epilogue threads <-> 128 rows).  Sweep 1 streams the row
    // (64 independent LDG.128 in flight per thread) for sum / sum of squares; sweep 2 re-reads it (L2 hit), normalises,
    // splits to bf16 hi/lo and stores 16-byte chunks into the SWIZZLE_128B K-major layout. ----
    const int r = static_cast<int>(threadIdx.x) - 64;   // row inside this CTA's 128-row tile (blockIdx.y)
    const int gr = static_cast<int>(blockIdx.y) * 128 + r;
    const float* xr = x + static_cast<long long>(gr < M ? gr : 0) * ldx;
    float s = 0.f, ss = 0.f;
    if (gr < M) {
#pragma unroll 16
      for (int c = 0; c < 64; ++c) {
        const float4 f = *reinterpret_cast<const float4*>(xr + 4 * c);
        s += (f.x + f.y) + (f.z + f.w);
        ss = fmaf(f.x, f.x, ss); ss = fmaf(f.y, f.y, ss); ss = fmaf(f.z, f.z, ss); ss = fmaf(f.w, f.w, ss);
      }
    }
    const float mean = s * (1.f / 256.f);
    const float rstd = rsqrtf(fmaxf(ss * (1.f / 256.f) - mean * mean, 0.f) + eps);
#pragma unroll 4
    for (int ch32 = 0; ch32 < 32; ++ch32) {  // 32 chunks of 8 columns
      uint32_t hw[4] = {0, 0, 0, 0}, lw[4] = {0, 0, 0, 0};
      if (gr < M) {
        const float4 f0 = *reinterpret_cast<const float4*>(xr + 8 * ch32);
        const float4 f1 = *reinterpret_cast<const float4*>(xr + 8 * ch32 + 4);
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + 8 * ch32));
        const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + 8 * ch32 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + 8 * ch32));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + 8 * ch32 + 4));
        const float v[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
        const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bt[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          split2_bf16((v[i] - mean) * rstd * g[i] + bt[i], (v[i + 1] - mean) * rstd * g[i + 1] + bt[i + 1], hw[i >> 1], lw[i >> 1]);
        }
      }
      const int kb = ch32 >> 3, chunk = ch32 & 7;
      uint8_t* dst = smem + kb * 2 * A_PLANE + r * 128 + ((chunk ^ (r & 7)) << 4);
      *reinterpret_cast<uint4*>(dst) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
      *reinterpret_cast<uint4*>(dst + A_PLANE) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
    }
    fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor-core (async proxy) reads
    mbar_arrive(aready_bar);
    // ---- epilogue ----
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    gemm_epilogue_tile<BN>(ep, tmem_base, warp & 3, lane, static_cast<int>(blockIdx.y) * 128, n0, M, N,
                           stage_all + (warp & 3) * 256);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// ------------------------------------------------------------------------------------------------
// host side: tensor-map construction (driver entry point, cached) and launch
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

struct TmapKey {
  const void* ptr;
  uint64_t d0, d1, d2, ld, plane_stride;
  uint32_t b0, b1, esz;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && ld == o.ld && plane_stride == o.plane_stride &&
           b0 == o.b0 && b1 == o.b1 && esz == o.esz;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    auto mix = [&h](uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(k.d0); mix(k.d1); mix(k.d2); mix(k.ld); mix(k.plane_stride); mix(k.b0); mix(k.b1); mix(k.esz);
    return h;
  }
};

// K-major operand [planes, rows, K] (K contiguous; row stride ld elements; plane stride given in
// elements), box = 128 bytes along K x box_rows rows x 1 plane, SWIZZLE_128B, OOB -> zeros.
int make_kmajor_tmap(CUtensorMap* out, const void* ptr, int elem_bytes, uint64_t K, uint64_t rows,
                     uint64_t planes, uint64_t ld, uint64_t plane_stride, uint32_t box_rows) {
  static std::mutex mu;
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  TmapKey key{ptr, K, rows, planes, ld, plane_stride, (uint32_t)(128 / elem_bytes), box_rows,
              (uint32_t)elem_bytes};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return 0; }
  }
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return RALF_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * elem_bytes) & 15) || ((plane_stride * elem_bytes) & 15))
    return RALF_ERR_ALIGN;
  cuuint64_t gdim[3] = {K, rows, planes};
  cuuint64_t gstr[2] = {ld * elem_bytes, (planes > 1 ? plane_stride : ld * rows) * elem_bytes};
  cuuint32_t box[3] = {(cuuint32_t)(128 / elem_bytes), box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = fn(out, dt, 3, const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "ralf_b200: cuTensorMapEncodeTiled failed (%d) K=%llu rows=%llu planes=%llu ld=%llu\n",
            (int)r, (unsigned long long)K, (unsigned long long)rows, (unsigned long long)planes,
            (unsigned long long)ld);
    return RALF_ERR_DRIVER;
  }
  std::lock_guard<std::mutex> g(mu);
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, *out);
  return 0;
}

template <int BN, int NPASS, int FOLD>
static int launch_gemm_v(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpi& ep, int M, int N, int K,
                         cudaStream_t st, const ConvGeom& cg, int splits, long long split_stride) {
  using Cfg = GemmCfg<BN, NPASS>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_kernel<BN, NPASS, FOLD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::smem_bytes(Cfg::MAX_STAGES));
    if (e != cudaSuccess) return set_cuda_error(e);
    // several CTAs per SM only materialise when the SM is configured with the full shared-memory carve-out
    e = cudaFuncSetAttribute(gemm_bf16_kernel<BN, NPASS, FOLD>, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return set_cuda_error(e);
    attr_set = true;
  }
  const long long tiles_m = cg.enabled ? static_cast<long long>((cg.B + cg.NB - 1) / cg.NB) * cg.hblocks : (M + 127) / 128;
  const long long num_tiles = static_cast<long long>((N + BN - 1) / BN) * tiles_m * splits;
  const int sms = num_sms();
  const int nkb = (K + 63) / 64;
  // The smem ring may span tiles (RALF_GEMM_MIN_STAGES > k-blocks: the producer runs ahead into the next tiles).  Measured
  // (profiles/r1_gemm_epilogue_e_notes.md): no gain for the short-K GEMMs -- their epilogue, not the operand fetch, is
  // the limiter -- so the default keeps the small ring.
  static const int min_stages = getenv("RALF_GEMM_MIN_STAGES") ? atoi(getenv("RALF_GEMM_MIN_STAGES")) : 1;
  const bool multi_tile = num_tiles > static_cast<long long>(sms);
  const int nkb_per = (nkb + splits - 1) / splits;
  int stages = nkb_per < Cfg::MAX_STAGES ? nkb_per : Cfg::MAX_STAGES;
  if (multi_tile && stages < min_stages) stages = min_stages < Cfg::MAX_STAGES ? min_stages : Cfg::MAX_STAGES;
  // Short-K GEMMs are bound by their epilogue's memory traffic, not by the MMAs: their small smem ring lets several
  // persistent CTAs share an SM (limited by registers -- see MINB above --, shared memory, 512 TMEM columns and an env
  // override for A/B runs), which multiplies the loads / stores in flight per SM.
  static const int occ_cap = getenv("RALF_GEMM_OCC") ? atoi(getenv("RALF_GEMM_OCC")) : 4;
  static int occ_cache[16] = {0};
  int occ = occ_cache[stages];
  if (occ == 0) {
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gemm_bf16_kernel<BN, NPASS, FOLD>, 320,
                                                                  Cfg::smem_bytes(stages));
    if (e != cudaSuccess) return set_cuda_error(e);
    constexpr int tcols = FOLD ? Cfg::TMEM_COLS_FOLD : Cfg::TMEM_COLS;
    if (occ > static_cast<int>(512 / tcols)) occ = 512 / tcols;
    if (occ < 1) occ = 1;
    occ_cache[stages] = occ;
  }
  if (occ > occ_cap) occ = occ_cap;
  // RALF_GEMM_RESERVE_SMS=n (A/B knob for the two-batches-in-flight pipeline, default 0): multi-wave persistent grids leave
  // n SMs unclaimed, so the small-grid kernels of a concurrent latency-bound stream (the decode loop of the previous
  // batch) find an SM at once instead of queueing behind a whole GEMM.
  static const int reserve_sms = getenv("RALF_GEMM_RESERVE_SMS") ? atoi(getenv("RALF_GEMM_RESERVE_SMS")) : 0;
  const int grid_sms = (multi_tile && reserve_sms > 0 && reserve_sms < sms) ? sms - reserve_sms : sms;
  const long long max_ctas = static_cast<long long>(grid_sms) * occ;
  dim3 grid(static_cast<unsigned>(num_tiles < max_ctas ? num_tiles : max_ctas));
  const cudaError_t le = launch_pdl(gemm_bf16_kernel<BN, NPASS, FOLD>, grid, dim3(320), Cfg::smem_bytes(stages), st, ta, tb,
                                    ep, M, N, K, stages, cg, splits, split_stride);
  return set_cuda_error(le != cudaSuccess ? le : cudaGetLastError());
}

template <int BN, int NPASS>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmEpi& ep, int M, int N, int K,
                       cudaStream_t st, const ConvGeom& cg, int splits = 1, long long split_stride = 0) {
  // RALF_GEMM_FOLD=0 restores three MMAs per k-step (A/B runs)
  static const bool fold = !(getenv("RALF_GEMM_FOLD") && atoi(getenv("RALF_GEMM_FOLD")) == 0);
  if constexpr (NPASS == 3 && BN <= 128) {
    if (fold) return launch_gemm_v<BN, NPASS, 1>(ta, tb, ep, M, N, K, st, cg, splits, split_stride);
  }
  return launch_gemm_v<BN, NPASS, 0>(ta, tb, ep, M, N, K, st, cg, splits, split_stride);
}

}  // namespace ralf

using namespace ralf;

// NHWC split activation [planes][B, H, W, C] -> 5-D tensor map (C, W, H, B, plane); box = 64 channels x Wo x BH x NB.
static int make_conv_tmap(CUtensorMap* out, const void* ptr, uint64_t C_, uint64_t W_, uint64_t H_, uint64_t B_,
                          uint64_t planes, uint64_t plane_stride, uint32_t bw, uint32_t bh, uint32_t nb, uint32_t stride = 1) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return RALF_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((C_ * 2) & 15) || ((plane_stride * 2) & 15)) return RALF_ERR_ALIGN;
  cuuint64_t gdim[5] = {C_, W_, H_, B_, planes};
  cuuint64_t gstr[4] = {C_ * 2, W_ * C_ * 2, H_ * W_ * C_ * 2, (planes > 1 ? plane_stride : B_ * H_ * W_ * C_) * 2};
  // element strides: a box of bw * stride traversed positions loads every stride-th one = bw of them
  cuuint32_t box[5] = {64, bw * stride, bh * stride, nb, 1};
  cuuint32_t estr[5] = {1, stride, stride, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(ptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "ralf_b200: cuTensorMapEncodeTiled (conv) failed (%d)\n", (int)r);
    return RALF_ERR_DRIVER;
  }
  return 0;
}


// Epilogue operand of gemm_bf16_tepi_kernel: split bf16 tensor [2 planes][rows, cols] (row stride ld, plane stride in
// elements) as a 3-D map (cols, rows, plane) with boxes of 32 columns x 32 rows x both planes, SWIZZLE_64B.
static int make_chunk_tmap(CUtensorMap* out, const void* ptr, uint64_t cols, uint64_t rows, uint64_t ld, uint64_t plane_stride) {
  static std::mutex mu;
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  TmapKey key{ptr, cols, rows, 2, ld, plane_stride, 32u, 32u, 0xC2u};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return 0; }
  }
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return RALF_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * 2) & 15) || ((plane_stride * 2) & 15)) return RALF_ERR_ALIGN;
  cuuint64_t gdim[3] = {cols, rows, 2};
  cuuint64_t gstr[2] = {ld * 2, plane_stride * 2};
  cuuint32_t box[3] = {32, 32, 2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "ralf_b200: cuTensorMapEncodeTiled (epilogue chunks) failed (%d)\n", (int)r);
    return RALF_ERR_DRIVER;
  }
  std::lock_guard<std::mutex> g(mu);
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, *out);
  return 0;
}

// fp32 flavour: [rows, cols] fp32 (row stride ld elements) as a 2-D map with 32 x 32 boxes (128-byte rows), SWIZZLE_128B.
static int make_chunk_tmap_f32(CUtensorMap* out, const void* ptr, uint64_t cols, uint64_t rows, uint64_t ld) {
  static std::mutex mu;
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  TmapKey key{ptr, cols, rows, 1, ld, 0, 32u, 32u, 0xC4u};
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return 0; }
  }
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return RALF_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) || ((ld * 4) & 15)) return RALF_ERR_ALIGN;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstr[1] = {ld * 4};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "ralf_b200: cuTensorMapEncodeTiled (fp32 epilogue chunks) failed (%d)\n", (int)r);
    return RALF_ERR_DRIVER;
  }
  std::lock_guard<std::mutex> g(mu);
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, *out);
  return 0;
}

template <int BN, int NBUF, int F32 = 0>
static int launch_tepi(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tr, const CUtensorMap& to,
                       const float* bias, int act, int has_res, int post_relu, int M, int N, int K, int stages, const ConvGeom& cg,
                       cudaStream_t st) {
  static int attr_bytes = 0;
  const int smem = TepiCfg<BN>::smem_bytes(stages, NBUF);
  if (smem > 232448) return RALF_ERR_SHAPE;
  if (smem > attr_bytes) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tepi_kernel<BN, NBUF, F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_cuda_error(e);
    attr_bytes = smem;
  }
  const long long tiles_m = cg.enabled ? static_cast<long long>(cg.B / cg.NB) * cg.hblocks : (M + 127) / 128;
  const long long num_tiles = static_cast<long long>(N / BN) * tiles_m;
  const int sms = num_sms();
  dim3 grid(static_cast<unsigned>(num_tiles < sms ? num_tiles : sms));
  const cudaError_t le = launch_pdl(gemm_bf16_tepi_kernel<BN, NBUF, F32>, grid, dim3(320), smem, st, ta, tb, tr, to, bias, act, has_res,
                                    post_relu, M, N, K, stages, cg);
  return set_cuda_error(le != cudaSuccess ? le : cudaGetLastError());
}

__global__ void splitk_reduce_kernel(const float* __restrict__ ws, int S, long long stride, int M, int N,
                                     float* __restrict__ out, int out_ld) {
  const long long total = static_cast<long long>(M) * N;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < S; ++s) acc += ws[s * stride + i];  // fixed order: deterministic
    out[(i / N) * out_ld + (i % N)] = acc;
  }
}

// Split-K plan for a plain fp32-output GEMM whose M x N tiles cannot fill the GPU (weight gradients): number of k
// slices (1 = no split) such that tiles * slices ~ #SMs and every slice keeps >= 8 k-blocks.
static int splitk_plan(int M, int N, int K, int bn) {
  const long long tiles = static_cast<long long>((M + 127) / 128) * ((N + bn - 1) / bn);
  const int nkb = (K + 63) / 64;
  const int sms = num_sms();
  if (tiles * 2 > sms || nkb < 32) return 1;
  int s = static_cast<int>(sms / tiles);
  if (s > nkb / 8) s = nkb / 8;
  if (s > 64) s = 64;
  if (s < 2) return 1;
  const int per = (nkb + s - 1) / s;
  return (nkb + per - 1) / per;  // no empty slice
}

extern "C" size_t ralf_gemm_splitk_workspace_bytes(int M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  const int bn = N >= 64 ? 64 : 32;
  const int s = splitk_plan(M, N, K, bn);
  return s > 1 ? static_cast<size_t>(s) * M * N * sizeof(float) : 0;
}

static int gemm_dispatch(const RalfGemmArgs* a, const CUtensorMap& ta, int bn, const ConvGeom& cg, void* stream) {
  const int planes = a->npass == 3 ? 2 : 1;
  CUtensorMap tb;
  int rc = make_kmajor_tmap(&tb, a->W, 2, a->K, a->N, planes, a->ldw, a->w_plane, bn);
  if (rc) return rc;
  GemmEpi ep;
  ep.bias = a->bias;
  ep.res = a->res;
  ep.res_split = reinterpret_cast<const __nv_bfloat16*>(a->res_split);
  ep.res_plane = a->res_plane;
  ep.res_ld = a->res_ld;
  ep.res_row_mod = a->res_row_mod;
  ep.out_f32 = a->out_f32;
  ep.out_split = reinterpret_cast<__nv_bfloat16*>(a->out_split);
  ep.out_plane = a->out_plane;
  ep.out_ld = a->out_ld;
  ep.out_col0 = a->out_col0;
  ep.rows_per_group = a->rows_per_group > 0 ? a->rows_per_group : a->M;
  ep.group_stride = a->rows_per_group > 0 ? a->group_stride : 0;
  ep.group_offset = a->rows_per_group > 0 ? a->group_offset : 0;
  ep.act = a->act;
  ep.post_relu = a->post_relu;
  ep.split_lo = a->out_split_lo;
  ep.out_kv24 = reinterpret_cast<uint8_t*>(a->out_kv24);
  ep.kv_fmt = a->out_kv_fmt == 16 ? 16 : 24;
  ep.kv24_ld = ep.kv_fmt == 16 ? 1088 : 1536;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  ep.vec_ok = (a->out_ld % 8 == 0) && (a->out_col0 % 8 == 0) && al16(a->out_f32) && al16(a->out_split) &&
              (a->out_plane % 8 == 0) && al16(a->bias) && al16(a->res) && al16(a->res_split) &&
              (a->res_ld % 8 == 0 || (!a->res && !a->res_split)) && (a->res_plane % 8 == 0);
  if (a->out_kv24 && (a->N != 512 || !ep.vec_ok || !al16(a->out_kv24))) return RALF_ERR_SHAPE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int np = a->npass;
  if (a->splitk_ws && !cg.enabled && a->out_f32 && !a->out_split && !a->bias && !a->act && !a->res && !a->res_split &&
      a->rows_per_group <= 0 && np == 3) {
    const int sbn = a->N >= 64 ? 64 : 32;
    const int S = splitk_plan(a->M, a->N, a->K, sbn);
    if (S > 1 && a->splitk_ws_bytes >= static_cast<size_t>(S) * a->M * a->N * sizeof(float)) {
      CUtensorMap tbs;
      rc = make_kmajor_tmap(&tbs, a->W, 2, a->K, a->N, planes, a->ldw, a->w_plane, sbn);
      if (rc) return rc;
      GemmEpi es = ep;
      es.out_f32 = reinterpret_cast<float*>(a->splitk_ws);
      es.out_ld = a->N;
      es.out_col0 = 0;
      es.vec_ok = (a->N % 8 == 0) && al16(a->splitk_ws);
      const long long stride = static_cast<long long>(a->M) * a->N;
      rc = sbn == 64 ? launch_gemm<64, 3>(ta, tbs, es, a->M, a->N, a->K, st, cg, S, stride)
                     : launch_gemm<32, 3>(ta, tbs, es, a->M, a->N, a->K, st, cg, S, stride);
      if (rc) return rc;
      const long long total = stride;
      const int blocks = static_cast<int>((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
      splitk_reduce_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float*>(a->splitk_ws), S, stride, a->M, a->N,
                                                   a->out_f32 + a->out_col0, a->out_ld);
      return set_cuda_error(cudaGetLastError());
    }
  }
  // Bottleneck-tail shape class -> the TMA-epilogue kernel (RALF_GEMM_TEPI=0 keeps the register-staged epilogue: A/B runs;
  // RALF_TEPI_KMAX = longest K it takes, its operand ring being two stages deep)
  static const bool tepi = !(getenv("RALF_GEMM_TEPI") && atoi(getenv("RALF_GEMM_TEPI")) == 0);
  static const bool fold_on = !(getenv("RALF_GEMM_FOLD") && atoi(getenv("RALF_GEMM_FOLD")) == 0);
  static const int tepi_kmax = getenv("RALF_TEPI_KMAX") ? atoi(getenv("RALF_TEPI_KMAX")) : 256;
  static const int tepi_k64 = getenv("RALF_TEPI_K64") ? atoi(getenv("RALF_TEPI_K64")) : 14;  // stages * 10 + chunk buffers
  // Without a residual one chunk buffer per warp is enough, which leaves room for a THREE-stage operand ring
  // (RALF_TEPI_DEEP=0: keep two stages + three buffers); RALF_TEPI_KMAX_NORES = longest K of that variant.
  static const bool tepi_deep = !(getenv("RALF_TEPI_DEEP") && atoi(getenv("RALF_TEPI_DEEP")) == 0);
  static const int tepi_kmax_nores = getenv("RALF_TEPI_KMAX_NORES") ? atoi(getenv("RALF_TEPI_KMAX_NORES")) : 256;
  const int tepi_klim = (tepi_deep && !a->res && !a->res_split && tepi_kmax_nores > tepi_kmax) ? tepi_kmax_nores : tepi_kmax;
  // implicit convolutions qualify when every M tile is a full, contiguous block of 128 output rows
  const bool cg_full = !cg.enabled || (cg.Wo * cg.BH * cg.NB == 128 && cg.Ho % cg.BH == 0 && cg.B % cg.NB == 0);
  // (A BN = 64 flavour for the 64-channel layers -- stem, layer-1 conv1 / conv2 -- was measured: stem 300.0 vs 300.0 us,
  // conv2 140.3 vs 140.3 us.  Those layers are bound by the tensor core's shared-memory operand reads -- 14 KB per k-step
  // for 128 x 64 outputs --, not by their epilogue; ncu: tensor pipe 32-40 %, DRAM 20-25 %.  Not instantiated.)
  const long long tepi_tiles = (cg.enabled ? static_cast<long long>(cg.B / (cg.NB > 0 ? cg.NB : 1)) * cg.hblocks : (a->M + 127) / 128) *
                               (a->N / 128);
  if (tepi && fold_on && bn == 128 && np == 3 && cg_full && a->N % 128 == 0 && a->K <= tepi_klim &&
      tepi_tiles >= 2 * num_sms() && ep.out_split && ep.split_lo && !ep.out_f32 && !ep.out_kv24 && !ep.res &&
      a->rows_per_group <= 0 && ep.res_row_mod <= 0 && ep.vec_ok && ep.act != 2) {
    CUtensorMap tr, to;
    rc = make_chunk_tmap(&to, ep.out_split + ep.out_col0, a->N, a->M, ep.out_ld, ep.out_plane);
    if (rc) return rc;
    tr = to;
    if (ep.res_split) {
      rc = make_chunk_tmap(&tr, ep.res_split, a->N, a->M, ep.res_ld, ep.res_plane);
      if (rc) return rc;
    }
    const int nkb = (a->K + 63) / 64;
    const int has_res = ep.res_split != nullptr;
    if (nkb == 1 && tepi_k64 == 14)
      return launch_tepi<128, 4>(ta, tb, tr, to, ep.bias, ep.act, has_res, ep.post_relu, a->M, a->N, a->K, 1, cg, st);
    if (!has_res && tepi_deep && nkb >= 3)
      return launch_tepi<128, 1>(ta, tb, tr, to, ep.bias, ep.act, 0, ep.post_relu, a->M, a->N, a->K, 3, cg, st);
    return launch_tepi<128, 3>(ta, tb, tr, to, ep.bias, ep.act, has_res, ep.post_relu, a->M, a->N, a->K, 2, cg, st);
  }
  // fp32 flavour of the same kernel: fp32 output (+ optional fp32 residual), no split output (transformer encoder qkv / out-proj,
  // FIDNetV3 layers); RALF_GEMM_TEPI_F32=0 switches it off
  static const bool tepi_f32 = !(getenv("RALF_GEMM_TEPI_F32") && atoi(getenv("RALF_GEMM_TEPI_F32")) == 0);
  if (tepi && tepi_f32 && fold_on && bn == 128 && np == 3 && !cg.enabled && a->N % 128 == 0 && a->K <= tepi_klim &&
      tepi_tiles >= 2 * num_sms() && ep.out_f32 && !ep.out_split && !ep.out_kv24 && !ep.res_split && a->rows_per_group <= 0 &&
      ep.res_row_mod <= 0 && ep.vec_ok && ep.act != 2 && !a->splitk_ws) {
    CUtensorMap tr, to;
    rc = make_chunk_tmap_f32(&to, ep.out_f32 + ep.out_col0, a->N, a->M, ep.out_ld);
    if (rc) return rc;
    tr = to;
    if (ep.res) {
      rc = make_chunk_tmap_f32(&tr, ep.res, a->N, a->M, ep.res_ld);
      if (rc) return rc;
    }
    const int nkb = (a->K + 63) / 64;
    const int has_res = ep.res != nullptr;
    if (nkb == 1) return launch_tepi<128, 4, 1>(ta, tb, tr, to, ep.bias, ep.act, has_res, ep.post_relu, a->M, a->N, a->K, 1, cg, st);
    if (!has_res && tepi_deep && nkb >= 3)
      return launch_tepi<128, 1, 1>(ta, tb, tr, to, ep.bias, ep.act, 0, ep.post_relu, a->M, a->N, a->K, 3, cg, st);
    return launch_tepi<128, 3, 1>(ta, tb, tr, to, ep.bias, ep.act, has_res, ep.post_relu, a->M, a->N, a->K, 2, cg, st);
  }
#define RALF_GEMM_CASE(BN_, NP_) \
  if (bn == BN_ && np == NP_) return launch_gemm<BN_, NP_>(ta, tb, ep, a->M, a->N, a->K, st, cg);
  RALF_GEMM_CASE(32, 3)
  RALF_GEMM_CASE(64, 3)
  RALF_GEMM_CASE(128, 3)
  RALF_GEMM_CASE(256, 3)
  RALF_GEMM_CASE(32, 1)
  RALF_GEMM_CASE(64, 1)
  RALF_GEMM_CASE(128, 1)
  RALF_GEMM_CASE(256, 1)
#undef RALF_GEMM_CASE
  return RALF_ERR_SHAPE;
}

static int gemm_pick_bn(const RalfGemmArgs* a, long long mt) {
  int bn = a->block_n;
  if (bn == 0) {  // widest tile that still yields ~one tile per SM; skinny (decode, M <= 128) GEMMs get BN = 32
    if (mt * ((a->N + 255) / 256) >= 120 && a->N >= 256 && a->npass == 1) bn = 256;
    else if (mt * ((a->N + 127) / 128) >= 120 && a->N >= 128) bn = 128;
    else if (mt * ((a->N + 63) / 64) >= 60 && a->N >= 64) bn = 64;
    else bn = 32;
  }
  return bn;
}

extern "C" int ralf_gemm(const RalfGemmArgs* a, void* stream) {
  if (!a || a->M <= 0 || a->N <= 0 || a->K <= 0) return RALF_ERR_SHAPE;
  if (a->npass != 1 && a->npass != 3) return RALF_ERR_SHAPE;
  if (!a->A || !a->W) return RALF_ERR_NULL;
  if ((a->lda % 8) || (a->ldw % 8)) return RALF_ERR_ALIGN;
  const int planes = a->npass == 3 ? 2 : 1;
  const int bn = gemm_pick_bn(a, (a->M + 127) / 128);
  if (bn != 32 && bn != 64 && bn != 128 && bn != 256) return RALF_ERR_SHAPE;
  CUtensorMap ta;
  int rc = make_kmajor_tmap(&ta, a->A, 2, a->K, a->M, planes, a->lda, a->a_plane, 128);
  if (rc) return rc;
  ConvGeom cg;
  cg.enabled = 0;
  cg.B = cg.Ho = cg.Wo = cg.BH = cg.NB = cg.hblocks = cg.cblocks = cg.KW = cg.stride = 1;
  cg.pad = 0;
  return gemm_dispatch(a, ta, bn, cg, stream);
}

// Convolution (stride 1 or 2, padding KH / 2) as an implicit GEMM: args->A = NHWC split activation [planes][B*H*W, C]
// (dense rows, lda == C), args->W = [planes][N, KH*KW*C] with k = (kh*KW + kw)*C + c, args->M = B*Ho*Wo output positions
// (Ho = (H + 2*(KH/2) - KH) / stride + 1), args->K = KH*KW*C.  The epilogue fields mean what they mean for ralf_gemm.
// C % 64 == 0, Wo <= 128.  Stride 2 uses the tensor map's element strides: a tap's box skips every other input position.
extern "C" int ralf_conv_gemm_strided(const RalfGemmArgs* a, int B, int H, int W, int C, int KH, int KW, int stride,
                                      void* stream) {
  if (!a || a->M <= 0 || a->N <= 0 || a->K <= 0) return RALF_ERR_SHAPE;
  if (a->npass != 1 && a->npass != 3) return RALF_ERR_SHAPE;
  if (!a->A || !a->W) return RALF_ERR_NULL;
  if (stride != 1 && stride != 2) return RALF_ERR_SHAPE;
  if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || (C % 64) || KH != KW || !(KH & 1) || KH > 7) return RALF_ERR_SHAPE;
  const int pad = KH / 2;
  const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
  if (Wo > 128 || a->M != B * Ho * Wo || a->K != KH * KW * C || a->lda != C) return RALF_ERR_SHAPE;
  if (a->ldw % 8) return RALF_ERR_ALIGN;
  const int planes = a->npass == 3 ? 2 : 1;
  ConvGeom cg;
  cg.enabled = 1;
  cg.B = B; cg.Ho = Ho; cg.Wo = Wo;
  cg.BH = 128 / Wo < Ho ? 128 / Wo : Ho;
  cg.NB = cg.BH == Ho ? 128 / (Wo * Ho) : 1;
  if (cg.NB < 1) cg.NB = 1;
  if (cg.NB > B) cg.NB = B;
  cg.hblocks = (Ho + cg.BH - 1) / cg.BH;
  cg.cblocks = C / 64;
  cg.KW = KW;
  cg.pad = pad;
  cg.stride = stride;
  const long long mt = static_cast<long long>((B + cg.NB - 1) / cg.NB) * cg.hblocks;
  const int bn = gemm_pick_bn(a, mt);
  if (bn != 32 && bn != 64 && bn != 128 && bn != 256) return RALF_ERR_SHAPE;
  CUtensorMap ta;
  int rc = make_conv_tmap(&ta, a->A, C, W, H, B, planes, a->a_plane, Wo, cg.BH, cg.NB, stride);
  if (rc) return rc;
  return gemm_dispatch(a, ta, bn, cg, stream);
}

extern "C" int ralf_conv_gemm(const RalfGemmArgs* a, int B, int H, int W, int C, int KH, int KW, void* stream) {
  return ralf_conv_gemm_strided(a, B, H, W, C, KH, KW, 1, stream);
}

// ResNet stem as an implicit GEMM over the space-to-depth image of ralf_stem_s2d (nn_kernels.cu):
// A = [planes][B, Hp, Wp, 16] (Hp = Ho + 3, Wp = Wo + 3, zero border), W = [planes][N, 256] with
// k = kh'*64 + kw'*16 + (dy*2+dx)*4 + c.  k-block kh' of output pixel (b, ho, wo) is the 128 contiguous bytes starting
// at pixel (b, ho + kh', wo): the tensor map's W dimension steps by ONE pixel (32 bytes) while its innermost extent is
// four pixels (overlapping windows), so a (64 x Wo x BH) box is the whole A tile of a k-block.
extern "C" int ralf_stem_gemm(const RalfGemmArgs* a, int B, int Ho, int Wo, void* stream) {
  if (!a || a->M <= 0 || a->N <= 0) return RALF_ERR_SHAPE;
  if (a->npass != 1 && a->npass != 3) return RALF_ERR_SHAPE;
  if (!a->A || !a->W) return RALF_ERR_NULL;
  if (B <= 0 || Ho <= 0 || Wo <= 0 || Wo > 128 || a->M != B * Ho * Wo || a->K != 256) return RALF_ERR_SHAPE;
  if (a->ldw % 8) return RALF_ERR_ALIGN;
  const int planes = a->npass == 3 ? 2 : 1;
  const int Hp = Ho + 3, Wp = Wo + 3;
  ConvGeom cg;
  cg.enabled = 1;
  cg.B = B; cg.Ho = Ho; cg.Wo = Wo;
  cg.BH = 128 / Wo < Ho ? 128 / Wo : Ho;
  cg.NB = 1;  // a box must not run over the bottom border rows into the next image
  cg.hblocks = (Ho + cg.BH - 1) / cg.BH;
  cg.cblocks = 1;  // k-block = kh' (tap = kb, kw = 0, channel block 0)
  cg.KW = 1;
  cg.pad = 0;
  cg.stride = 1;
  const long long mt = static_cast<long long>(B) * cg.hblocks;
  const int bn = gemm_pick_bn(a, mt);
  if (bn != 32 && bn != 64 && bn != 128 && bn != 256) return RALF_ERR_SHAPE;
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return RALF_ERR_DRIVER;
  if ((reinterpret_cast<uintptr_t>(a->A) & 15) || ((a->a_plane * 2) & 15)) return RALF_ERR_ALIGN;
  CUtensorMap ta;
  cuuint64_t gdim[5] = {64, static_cast<cuuint64_t>(Wo), static_cast<cuuint64_t>(Hp), static_cast<cuuint64_t>(B),
                        static_cast<cuuint64_t>(planes)};
  cuuint64_t gstr[4] = {32, static_cast<cuuint64_t>(Wp) * 32, static_cast<cuuint64_t>(Hp) * Wp * 32,
                        static_cast<cuuint64_t>(planes > 1 ? a->a_plane : static_cast<long long>(B) * Hp * Wp * 16) * 2};
  cuuint32_t box[5] = {64, static_cast<cuuint32_t>(Wo), static_cast<cuuint32_t>(cg.BH), 1, 1};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(a->A), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "ralf_b200: cuTensorMapEncodeTiled (stem) failed (%d)\n", (int)r);
    return RALF_ERR_DRIVER;
  }
  return gemm_dispatch(a, ta, bn, cg, stream);
}


// x_new = A . W^T + bias + res -> out_f32 (may alias res), LayerNorm(x_new; gamma, beta) -> ln_split ([2][M, 256] split rows,
// plane stride ln_plane elements).  N = 256, K a multiple of 64, bf16x3.  `a` supplies A, W, bias, res / res_ld,
// out_f32 / out_ld (its other epilogue fields must be unset).  Cluster kernel: gemm_resln_kernel.
extern "C" int ralf_gemm_res_ln(const RalfGemmArgs* a, const float* gamma, const float* beta, float eps, void* ln_split,
                                long long ln_plane, void* stream) {
  if (!a || !a->A || !a->W || !a->res || !a->out_f32 || !gamma || !beta || !ln_split) return RALF_ERR_NULL;
  if (a->M <= 0 || a->N != 256 || a->K <= 0 || (a->K % 64) || a->npass != 3) return RALF_ERR_SHAPE;
  if (a->res_split || a->out_split || a->out_kv24 || a->act || a->post_relu || a->rows_per_group > 0 || a->res_row_mod > 0 ||
      a->out_col0)
    return RALF_ERR_SHAPE;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if ((a->lda % 8) || (a->ldw % 8) || (a->res_ld % 4) || (a->out_ld % 4) || (ln_plane % 8) || !al16(a->res) || !al16(a->out_f32) ||
      !al16(a->bias) || !al16(gamma) || !al16(beta) || !al16(ln_split))
    return RALF_ERR_ALIGN;
  CUtensorMap ta, tb;
  int rc = make_kmajor_tmap(&ta, a->A, 2, a->K, a->M, 2, a->lda, a->a_plane, 128);
  if (rc) return rc;
  rc = make_kmajor_tmap(&tb, a->W, 2, a->K, a->N, 2, a->ldw, a->w_plane, 32);
  if (rc) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_resln_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RLN_SMEM);
    if (e != cudaSuccess) return set_cuda_error(e);
    attr_set = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(8, (a->M + 127) / 128);
  cfg.blockDim = dim3(192);
  cfg.dynamicSmemBytes = RLN_SMEM;
  cfg.stream = reinterpret_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 8;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  const cudaError_t le = cudaLaunchKernelEx(&cfg, gemm_resln_kernel, ta, tb, a->bias, a->res, a->res_ld, a->out_f32, a->out_ld,
                                            gamma, beta, eps, reinterpret_cast<__nv_bfloat16*>(ln_split), ln_plane, a->M, a->K);
  return set_cuda_error(le != cudaSuccess ? le : cudaGetLastError());
}

template <int BN>
static int launch_gemm_ln(const CUtensorMap& tb, const float* x, int ldx, const float* gamma, const float* beta, float eps,
                          const GemmEpi& ep, int M, int N, cudaStream_t st) {
  constexpr int SMEM = 4 * 2 * 128 * 128 + 4 * 2 * BN * 128 + 1024 + 256 + 16384;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_ln_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e != cudaSuccess) return set_cuda_error(e);
    attr_set = true;
  }
  gemm_ln_kernel<BN><<<dim3((N + BN - 1) / BN, (M + 127) / 128), 192, SMEM, st>>>(tb, x, ldx, gamma, beta, eps, ep, M, N);
  return set_cuda_error(cudaGetLastError());
}

// LayerNorm(x) . W^T with the LayerNorm computed inside the GEMM (M <= 128, K = 256, bf16x3).  `a` supplies W and the
// epilogue fields (its A / lda / a_plane are ignored).
extern "C" int ralf_gemm_ln(const float* x, int ldx, const float* gamma, const float* beta, float eps,
                            const RalfGemmArgs* a, void* stream) {
  if (!x || !gamma || !beta || !a || !a->W) return RALF_ERR_NULL;
  if (a->M <= 0 || a->K != 256 || a->N <= 0 || a->npass != 3) return RALF_ERR_SHAPE;
  if ((ldx % 4) || (a->ldw % 8) || (reinterpret_cast<uintptr_t>(x) & 15)) return RALF_ERR_ALIGN;
  const int bn = a->N >= 1024 ? 64 : 32;
  CUtensorMap tb;
  int rc = make_kmajor_tmap(&tb, a->W, 2, a->K, a->N, 2, a->ldw, a->w_plane, bn);
  if (rc) return rc;
  GemmEpi ep;
  ep.bias = a->bias;
  ep.res = a->res;
  ep.res_split = reinterpret_cast<const __nv_bfloat16*>(a->res_split);
  ep.res_plane = a->res_plane;
  ep.res_ld = a->res_ld;
  ep.res_row_mod = a->res_row_mod;
  ep.out_f32 = a->out_f32;
  ep.out_split = reinterpret_cast<__nv_bfloat16*>(a->out_split);
  ep.out_plane = a->out_plane;
  ep.out_ld = a->out_ld;
  ep.out_col0 = a->out_col0;
  ep.rows_per_group = a->rows_per_group > 0 ? a->rows_per_group : a->M;
  ep.group_stride = a->rows_per_group > 0 ? a->group_stride : 0;
  ep.group_offset = a->rows_per_group > 0 ? a->group_offset : 0;
  ep.act = a->act;
  ep.post_relu = a->post_relu;
  ep.split_lo = a->out_split_lo;
  ep.out_kv24 = nullptr;
  ep.kv24_ld = 0;
  ep.kv_fmt = 24;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  ep.vec_ok = (a->out_ld % 8 == 0) && (a->out_col0 % 8 == 0) && al16(a->out_f32) && al16(a->out_split) &&
              (a->out_plane % 8 == 0) && al16(a->bias) && al16(a->res) && al16(a->res_split) &&
              (a->res_ld % 8 == 0 || (!a->res && !a->res_split)) && (a->res_plane % 8 == 0);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (bn == 64) return launch_gemm_ln<64>(tb, x, ldx, gamma, beta, eps, ep, a->M, a->N, st);
  return launch_gemm_ln<32>(tb, x, ldx, gamma, beta, eps, ep, a->M, a->N, st);
}
