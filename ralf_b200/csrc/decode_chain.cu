// Fused decode-step GEMM chains (greedy loop of BaseDecoder, common/common.py:84-135 with KV caches;
// retrieval_augmented_autoreg.py:271-297): LayerNorm -> Linear -> (+bias, ReLU, +residual) -> LayerNorm -> Linear ...
// for the ONE new token of every canvas, as a single kernel per chain instead of one kernel per op.
//
// Why: at M = B (one row per canvas) every linear layer of the decoder is a skinny GEMM whose time is launch + pipeline
// fill (11-20 us each, 69 launches per token, profiles/r1_launches_l_summary.md).  All ops of a decoder layer except the
// two attentions are ROW-LOCAL, so a CTA that owns a few canvases can run the whole chain without talking to any other
// CTA: no grid barriers, the residual row never leaves shared memory.
//
// Mapping ("swap A/B"): a CTA owns BT = 16 canvases.  The tensor core computes  D^T[feature, canvas] = W[feature, k] .
// X^T[k, canvas]:  A operand = a 128-feature x 64-k tile of the split-bf16 weight matrix streamed by TMA through a
// shared-memory ring (the whole chain's weight schedule is static, so the producer warp runs ahead across stage
// boundaries), B operand = the 16 activation rows (split bf16, K-major, SWIZZLE_128B) written by the compute warps,
// accumulator = 16 TMEM columns per 128-feature tile.  bf16x3: per k-step  w_lo.x_hi + w_hi.x_lo + w_hi.x_hi  in the
// same order as gemm_bf16_kernel issues them.  Epilogue thread = one output feature (TMEM lane) x 16 canvases:
// bias, ReLU, residual add from the shared-memory row buffer, global fp32 store (coalesced along features), or split
// straight into the next GEMM's operand buffer (FFN hidden layer, never leaves the SM).
//
// Bound (measured, profiles/r2_decode_chain.md): MMA issue.  An SS-mode tcgen05.mma at N = 16 costs ~92 cycles (fetch of
// the 128-row weight slice from shared memory), 12 per 32 KB weight tile, and every CTA must push ALL of the chain's
// weights through its own tensor core: ~0.6 us per tile, 66 us for the 96-tile FFN chain -- slower than the per-op GEMMs
// it replaces, which split the weights over n-tiles.  Kept opt-in (RALF_DECODE_CHAIN=1) with its parity test.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "ralf_internal.h"

namespace ralf {

constexpr int CH_BT = 16;            // canvases per CTA = UMMA N
constexpr int CH_D = 256;            // d_model: width of the residual row buffer
constexpr int CH_RING = 4;           // weight ring stages (3 when tiles are issued one at a time)
constexpr int CH_WT = 32768;         // bytes per ring stage: [hi 128 x 128 B][lo 128 x 128 B]
constexpr int CH_ACTB = 2 * 4 * 2048;    // operand buffer for K = 256 : [plane][kb][16 rows x 128 B]
constexpr int CH_ACTF = 2 * 16 * 2048;   // operand buffer for K = 1024
constexpr int CH_XS = CH_BT * CH_D * 4;  // fp32 residual rows
constexpr int CH_MAX_STAGES = 4;
constexpr int CH_SMEM = CH_RING * CH_WT + CH_ACTB + CH_ACTF + CH_XS + 1024 /*align*/ + 256 /*barriers*/;

struct ChainStageDev {
  int n_out;      // output features
  int n_tiles;    // ceil(n_out / 128)
  int nkb;        // K / 64 (4 or 16)
  int in_mode;    // 0: operand written by the previous stage's epilogue, 1: LayerNorm(x rows), 2: split rows from global
  const float* gamma;
  const float* beta;
  float eps;
  const __nv_bfloat16* in_split;  // [2, B, 256] hi plane; lo at + in_plane
  long long in_plane;
  int in_ld;
  const float* bias;
  int act;          // 1: ReLU
  int add_x;        // + residual row
  int to_x;         // result becomes the residual row (n_out == 256)
  int out_operand;  // result (split) is the next stage's operand (its K = n_out)
  float* out_f32;   // global fp32 [B, out_ld] or null
  int out_ld;
};
struct ChainArgsDev {
  ChainStageDev st[CH_MAX_STAGES];
  int n_stages;
  // MMA issue order.  A tile's 12 MMAs (4 k-steps x 3 passes) into ONE accumulator form a dependent chain, and at N = 16
  // an MMA occupies the tensor pipe for ~8 cycles but its accumulator is only ready for the next one after ~120 (measured:
  // 1440 cycles per tile, tensor pipe 7 % busy, profiles/r2_chain_ncu.md).  nacc = 3 gives each bf16x3 pass its own
  // accumulator (summed in the epilogue), pair = 2 interleaves two feature tiles: up to 6 independent chains in flight.
  int nacc;   // 1 (bit-identical to gemm_bf16_kernel's order) or 3
  int pair;   // feature tiles issued together: 1 or 2
  int l2_prefetch;  // CTAs share out an up-front L2 prefetch of the whole weight schedule
  int debug;  // RALF_CHAIN_DEBUG (profiling only, results are garbage): 1 = no MMAs (TMA + barriers only), 2 = no TMA
              // (the producer only arrives), 3 = one MMA per tile instead of 12
  const float* x;  // residual rows [B, ldx] fp32 (read once) or null
  int ldx;
  int B;
};
struct ChainMaps {
  CUtensorMap m[CH_MAX_STAGES];
};

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_l2_3d(const CUtensorMap* m, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(reinterpret_cast<uint64_t>(m)),
               "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void compute_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ float ch_warp_sum(float v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}
// byte offset of element (row r, k) inside an operand buffer [plane][kb][16 rows x 128 B], SWIZZLE_128B (what TMA would write)
__device__ __forceinline__ uint32_t opnd_off(int r, int k) {
  const int kb = k >> 6, kk = k & 63;
  return static_cast<uint32_t>(kb * 2048 + r * 128 + ((((kk >> 3) ^ (r & 7)) << 4) | ((kk & 7) << 1)));
}

// NACC / PAIR: see ChainArgsDev (compile-time so that the MMA issue loop is straight-line code).
template <int NACC, int PAIR>
__global__ void __launch_bounds__(320, 1)
decode_chain_kernel(const __grid_constant__ ChainMaps maps, const __grid_constant__ ChainArgsDev args) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base_u32 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base_u32 - smem_u32(smem_raw));
  uint8_t* ring = smem;
  uint8_t* actB = ring + CH_RING * CH_WT;
  uint8_t* actF = actB + CH_ACTB;
  float* xs = reinterpret_cast<float*>(actF + CH_ACTF);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(xs) + CH_XS);
  uint64_t* empty_bar = full_bar + CH_RING;
  uint64_t* bready_bar = empty_bar + CH_RING;          // [stage] operand written (256 compute threads)
  uint64_t* tfull_bar = bready_bar + CH_MAX_STAGES;    // [stage] accumulators complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + CH_MAX_STAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * CH_BT;
  const int nst = args.n_stages;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < nst; ++s) tma_prefetch_desc(&maps.m[s]);
    for (int s = 0; s < CH_RING; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < CH_MAX_STAGES; ++s) {
      mbar_init(&bready_bar[s], 256);
      mbar_init(&tfull_bar[s], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ---------------- weight producer: the static tile schedule of the whole chain ----------------
    if (lane == 0) {
      if (args.l2_prefetch) {
        // The decoder weights do not survive in L2 between tokens (the K/V streams of a token are ~1 GB), and all CTAs
        // walk the same schedule in lockstep, so without this the whole grid waits on ONE HBM fetch at a time.  Every CTA
        // prefetches its share of the schedule up front: the full 0.5-3 MB is in flight at once, later TMA loads hit L2.
        int i = 0;
        for (int st = 0; st < nst; ++st) {
          const ChainStageDev& S = args.st[st];
          for (int mt = 0; mt < S.n_tiles; ++mt)
            for (int kb = 0; kb < S.nkb; ++kb, ++i)
              if (i % static_cast<int>(gridDim.x) == static_cast<int>(blockIdx.x)) {
                tma_prefetch_l2_3d(&maps.m[st], kb * 64, mt * 128, 0);
                tma_prefetch_l2_3d(&maps.m[st], kb * 64, mt * 128, 1);
              }
        }
      }
      int s = 0;
      uint32_t ph = 0;
      for (int st = 0; st < nst; ++st) {
        const ChainStageDev& S = args.st[st];
        for (int mt0 = 0; mt0 < S.n_tiles; mt0 += PAIR) {
          const int g = min(PAIR, S.n_tiles - mt0);
          for (int kb = 0; kb < S.nkb; ++kb) {
            for (int j = 0; j < g; ++j) {  // the tiles of a group sit in consecutive ring stages
              mbar_wait(&empty_bar[s], ph ^ 1);
              if (args.debug == 2) {
                mbar_arrive(&full_bar[s]);
              } else {
                mbar_expect_tx(&full_bar[s], CH_WT);
                uint8_t* dst = ring + s * CH_WT;
                tma_load_3d(&maps.m[st], &full_bar[s], dst, kb * 64, (mt0 + j) * 128, 0);
                tma_load_3d(&maps.m[st], &full_bar[s], dst + CH_WT / 2, kb * 64, (mt0 + j) * 128, 1);
              }
              if (++s == CH_RING) { s = 0; ph ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(1, 128, CH_BT);
      int s = 0;
      uint32_t ph = 0;
      for (int st = 0; st < nst; ++st) {
        const ChainStageDev& S = args.st[st];
        mbar_wait(&bready_bar[st], 0);
        tc_fence_after();
        const uint32_t opnd = base_u32 + static_cast<uint32_t>((S.nkb == 4 ? actB : actF) - smem);
        const uint32_t plane = static_cast<uint32_t>(S.nkb) * 2048u;
        for (int mt0 = 0; mt0 < S.n_tiles; mt0 += PAIR) {
          const bool two = PAIR == 2 && mt0 + 1 < S.n_tiles;
          // accumulator of (feature tile mt, pass p): TMEM columns (mt * NACC + (NACC == 3 ? p : 0)) * 16 ...
          // Stages reuse the same columns: a stage's MMAs start only after the previous epilogue has drained them.
          const uint32_t t0 = tmem_base + static_cast<uint32_t>(mt0 * NACC * CH_BT);
          const uint32_t t1 = t0 + NACC * CH_BT;
          constexpr uint32_t P1 = NACC == 3 ? CH_BT : 0, P2 = NACC == 3 ? 2 * CH_BT : 0;
          for (int kb = 0; kb < S.nkb; ++kb) {
            const int s0 = s;
            mbar_wait(&full_bar[s0], ph);
            if (++s == CH_RING) { s = 0; ph ^= 1; }
            const int s1 = s;
            if (two) {
              mbar_wait(&full_bar[s1], ph);
              if (++s == CH_RING) { s = 0; ph ^= 1; }
            }
            tc_fence_after();
            const uint32_t w0 = base_u32 + s0 * CH_WT, w1 = base_u32 + s1 * CH_WT;
            const uint64_t d0_hi = make_sw128_kmajor_desc(w0), d0_lo = make_sw128_kmajor_desc(w0 + CH_WT / 2);
            const uint64_t d1_hi = make_sw128_kmajor_desc(w1), d1_lo = make_sw128_kmajor_desc(w1 + CH_WT / 2);
            const uint64_t dx_hi = make_sw128_kmajor_desc(opnd + kb * 2048);
            const uint64_t dx_lo = make_sw128_kmajor_desc(opnd + plane + kb * 2048);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (args.debug == 1 || (args.debug == 3 && k > 0)) break;
              const uint64_t koff = static_cast<uint64_t>(2 * k);
              const uint32_t acc = (kb | k) != 0;
              const uint32_t acc12 = NACC == 3 ? acc : 1u;
              mma_bf16_ss(t0, d0_lo + koff, dx_hi + koff, idesc, acc);             // x_hi . w_lo
              if (two) mma_bf16_ss(t1, d1_lo + koff, dx_hi + koff, idesc, acc);
              if (args.debug == 3) continue;
              mma_bf16_ss(t0 + P1, d0_hi + koff, dx_lo + koff, idesc, acc12);      // x_lo . w_hi
              if (two) mma_bf16_ss(t1 + P1, d1_hi + koff, dx_lo + koff, idesc, acc12);
              mma_bf16_ss(t0 + P2, d0_hi + koff, dx_hi + koff, idesc, acc12);      // x_hi . w_hi
              if (two) mma_bf16_ss(t1 + P2, d1_hi + koff, dx_hi + koff, idesc, acc12);
            }
            tc_commit(&empty_bar[s0]);
            if (two) tc_commit(&empty_bar[s1]);
          }
        }
        tc_commit(&tfull_bar[st]);
      }
    }
  } else {
    // ---------------- compute warps: operand preparation + epilogues ----------------
    const int ct = static_cast<int>(threadIdx.x) - 64;  // 0..255
    const int cw = ct >> 5;                             // 0..7
    const int quad = warp & 3;                          // TMEM lane quadrant this warp may read
    const int group = cw >> 2;                          // warps 2-5 take even feature tiles, 6-9 odd ones
    if (args.x) {
      for (int i = ct; i < CH_BT * CH_D; i += 256) {
        const int c = i >> 8, f = i & 255;
        xs[i] = (row0 + c < args.B) ? args.x[static_cast<long long>(row0 + c) * args.ldx + f] : 0.f;
      }
    }
    compute_bar();
    for (int st = 0; st < nst; ++st) {
      const ChainStageDev& S = args.st[st];
      // ---- operand of this stage ----
      if (S.in_mode == 1) {
        // LayerNorm of the residual rows: warp cw owns rows 2cw, 2cw+1; arithmetic = layernorm_kernel<8> (nn_kernels.cu)
#pragma unroll 1
        for (int rr = 0; rr < 2; ++rr) {
          const int c = cw * 2 + rr;
          float v[8];
          float s = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            v[i] = xs[c * CH_D + lane + 32 * i];
            s += v[i];
          }
          const float mean = ch_warp_sum(s) / 256.f;
          float sq = 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float dlt = v[i] - mean;
            sq += dlt * dlt;
          }
          const float rstd = rsqrtf(ch_warp_sum(sq) / 256.f + S.eps);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int k = lane + 32 * i;
            const float y = (v[i] - mean) * rstd * S.gamma[k] + S.beta[k];
            __nv_bfloat16 h, l;
            split_bf16(y, h, l);
            const uint32_t off = opnd_off(c, k);
            *reinterpret_cast<__nv_bfloat16*>(actB + off) = h;
            *reinterpret_cast<__nv_bfloat16*>(actB + CH_ACTB / 2 + off) = l;
          }
        }
      } else if (S.in_mode == 2) {
        // split rows from global (attention output): 16 rows x 2 planes x 32 chunks of 16 B
        for (int i = ct; i < CH_BT * 2 * 32; i += 256) {
          const int ch = i & 31, p = (i >> 5) & 1, c = i >> 6;
          uint4 val = make_uint4(0u, 0u, 0u, 0u);
          if (row0 + c < args.B)
            val = *reinterpret_cast<const uint4*>(S.in_split + p * S.in_plane + static_cast<long long>(row0 + c) * S.in_ld + ch * 8);
          const int kb = ch >> 3, cc = ch & 7;
          *reinterpret_cast<uint4*>(actB + p * (CH_ACTB / 2) + kb * 2048 + c * 128 + ((cc ^ (c & 7)) << 4)) = val;
        }
      }
      fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's (async proxy) operand reads
      mbar_arrive(&bready_bar[st]);
      // ---- epilogue ----
      mbar_wait(&tfull_bar[st], 0);
      tc_fence_after();
      for (int mt = group; mt < S.n_tiles; mt += 2) {
        uint32_t v[16];
        const uint32_t tq = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
        tmem_ld_32x16(tq + static_cast<uint32_t>(mt * NACC * CH_BT), v);
        if (NACC == 3) {  // (x_hi.w_lo + x_lo.w_hi) + x_hi.w_hi
          uint32_t v1[16], v2[16];
          tmem_ld_32x16(tq + static_cast<uint32_t>((mt * 3 + 1) * CH_BT), v1);
          tmem_ld_32x16(tq + static_cast<uint32_t>((mt * 3 + 2) * CH_BT), v2);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < CH_BT; ++c)
            v[c] = __float_as_uint((__uint_as_float(v[c]) + __uint_as_float(v1[c])) + __uint_as_float(v2[c]));
        } else {
          tmem_ld_wait();
        }
        const int f = mt * 128 + quad * 32 + lane;
        if (f < S.n_out) {
          const float b = S.bias ? S.bias[f] : 0.f;
#pragma unroll
          for (int c = 0; c < CH_BT; ++c) {
            float y = __uint_as_float(v[c]) + b;
            if (S.act == 1) y = fmaxf(y, 0.f);
            if (S.add_x) y += xs[c * CH_D + f];
            if (S.to_x) xs[c * CH_D + f] = y;
            if (S.out_f32 && row0 + c < args.B) S.out_f32[static_cast<long long>(row0 + c) * S.out_ld + f] = y;
            if (S.out_operand) {
              __nv_bfloat16 h, l;
              split_bf16(y, h, l);
              const uint32_t off = opnd_off(c, f);
              *reinterpret_cast<__nv_bfloat16*>(actF + off) = h;
              *reinterpret_cast<__nv_bfloat16*>(actF + CH_ACTF / 2 + off) = l;
            }
          }
        }
      }
      tc_fence_before();
      compute_bar();  // residual rows / next operand complete before the next stage reads them
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace ralf

using namespace ralf;

extern "C" int ralf_decode_chain(const float* x, int ldx, int B, const RalfChainStage* stages, int n_stages,
                                 void* stream) {
  if (!stages) return RALF_ERR_NULL;
  if (B <= 0 || n_stages <= 0 || n_stages > CH_MAX_STAGES) return RALF_ERR_SHAPE;
  ChainMaps maps;
  ChainArgsDev a;
  a.n_stages = n_stages;
  a.x = x;
  a.ldx = ldx;
  a.B = B;
  bool need_x = false;
  for (int s = 0; s < n_stages; ++s) {
    const RalfChainStage& h = stages[s];
    ChainStageDev& d = a.st[s];
    if (!h.W) return RALF_ERR_NULL;
    if (h.n_out <= 0 || h.n_out > 1024 || (h.k_in != 256 && h.k_in != 1024)) return RALF_ERR_SHAPE;
    if (h.ldw % 8) return RALF_ERR_ALIGN;
    if (h.in_mode == 0 && (s == 0 || !stages[s - 1].out_operand || stages[s - 1].n_out != h.k_in || h.k_in != 1024))
      return RALF_ERR_SHAPE;
    if (h.in_mode != 0 && h.k_in != 256) return RALF_ERR_SHAPE;
    if (h.in_mode == 1 && (!h.gamma || !h.beta)) return RALF_ERR_NULL;
    if (h.in_mode == 2 && !h.in_split) return RALF_ERR_NULL;
    if (h.in_mode == 2 && ((h.in_ld % 8) || (h.in_plane % 8) || (reinterpret_cast<uintptr_t>(h.in_split) & 15))) return RALF_ERR_ALIGN;
    if ((h.to_x || h.add_x) && h.n_out != CH_D) return RALF_ERR_SHAPE;
    if (h.out_operand && (h.n_out != 1024 || s + 1 >= n_stages)) return RALF_ERR_SHAPE;
    need_x = need_x || h.in_mode == 1 || h.add_x;
    int rc = make_kmajor_tmap(&maps.m[s], h.W, 2, h.k_in, h.n_out, 2, h.ldw, h.w_plane, 128);
    if (rc) return rc;
    d.n_out = h.n_out;
    d.n_tiles = (h.n_out + 127) / 128;
    d.nkb = h.k_in / 64;
    d.in_mode = h.in_mode;
    d.gamma = h.gamma;
    d.beta = h.beta;
    d.eps = h.eps;
    d.in_split = reinterpret_cast<const __nv_bfloat16*>(h.in_split);
    d.in_plane = h.in_plane;
    d.in_ld = h.in_ld;
    d.bias = h.bias;
    d.act = h.act;
    d.add_x = h.add_x;
    d.to_x = h.to_x;
    d.out_operand = h.out_operand;
    d.out_f32 = h.out_f32;
    d.out_ld = h.out_ld;
  }
  for (int s = n_stages; s < CH_MAX_STAGES; ++s) {
    a.st[s] = a.st[0];
    maps.m[s] = maps.m[0];
  }
  if (need_x && !x) return RALF_ERR_NULL;
  static const int env_nacc = getenv("RALF_CHAIN_ACC") ? atoi(getenv("RALF_CHAIN_ACC")) : 1;
  static const int env_pair = getenv("RALF_CHAIN_PAIR") ? atoi(getenv("RALF_CHAIN_PAIR")) : 1;
  static const int env_pf = getenv("RALF_CHAIN_PREFETCH") ? atoi(getenv("RALF_CHAIN_PREFETCH")) : 0;
  // defaults: the issue order of gemm_bf16_kernel (bit-identical results); the variants measured the same speed
  a.nacc = env_nacc == 3 ? 3 : 1;
  a.pair = env_pair == 2 ? 2 : 1;
  a.l2_prefetch = env_pf != 0;
  static const int env_dbg = getenv("RALF_CHAIN_DEBUG") ? atoi(getenv("RALF_CHAIN_DEBUG")) : 0;
  a.debug = env_dbg;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int grid = (B + CH_BT - 1) / CH_BT;
#define RALF_CHAIN_CASE(NA_, PR_)                                                                                          \
  if (a.nacc == NA_ && a.pair == PR_) {                                                                                    \
    static bool attr_set = false;                                                                                          \
    if (!attr_set) {                                                                                                       \
      cudaError_t e = cudaFuncSetAttribute(decode_chain_kernel<NA_, PR_>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                           CH_SMEM);                                                                       \
      if (e != cudaSuccess) return set_cuda_error(e);                                                                      \
      attr_set = true;                                                                                                     \
    }                                                                                                                      \
    decode_chain_kernel<NA_, PR_><<<grid, 320, CH_SMEM, st>>>(maps, a);                                                    \
    return set_cuda_error(cudaGetLastError());                                                                             \
  }
  RALF_CHAIN_CASE(1, 1)
  RALF_CHAIN_CASE(1, 2)
  RALF_CHAIN_CASE(3, 1)
  RALF_CHAIN_CASE(3, 2)
#undef RALF_CHAIN_CASE
  return RALF_ERR_SHAPE;
}
