// Multi-head self-attention of the image encoder on the 5th-gen tensor cores (tcgen05 + TMEM), fp32-faithful.
//
// Replaces the scalar attention of nn.MultiheadAttention inside nn.TransformerEncoderLayer
// (image2layout/train/models/retrieval_augmented_autoreg.py:116-126; 8 heads x 32, T = h*w image tokens) for the
// shape class that dominates the step: head_dim 32, no mask, Tk <= 256 keys (one key tile), any Tq.
//
// One CTA (128 threads) per (batch, head); it keeps the head's K and V^T in shared memory and walks the queries in
// tiles of 128:
//   S[128, 256]  = Q K^T           tcgen05.mma kind::f16, operands "split" bf16 (hi | lo of every fp32 value, 3 passes:
//                                  hi.lo + lo.hi + hi.hi  ->  ~2^-17 relative, see gemm.cu), fp32 accumulators in TMEM
//   P            = exp(scale (S - rowmax))   thread r owns row r: tcgen05.ld, 2 sweeps, P written BACK TO TMEM as
//                                  split bf16 (tcgen05.st) -- it never touches shared or global memory
//   O[128, 32]   = P V             tcgen05.mma with the A operand read from TMEM (P), B = V^T from shared memory
//   out          = O / rowsum      split bf16 rows for the out-projection GEMM
// TMEM columns: [0,256) S, later O in [0,32);  [256,384) P hi;  [384,512) P lo  (bf16 pairs, 2 keys per column).
// Shared-memory operands are K-major rows of 128 bytes in the SWIZZLE_128B layout; Q and K rows hold
// [hi(32) | lo(32)] so one tile serves both planes (descriptor start offset 0 / 64 bytes).
#include <stdlib.h>

#include "common.cuh"
#include "ralf_internal.h"

namespace ralf {

__device__ __forceinline__ void mma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr int ATC_Q_BYTES = 128 * 128;       // 128 queries x [hi 64 B | lo 64 B]
constexpr int ATC_K_BYTES = 256 * 128;       // 256 keys
constexpr int ATC_VT_PLANE = 4 * 32 * 128;   // 4 key blocks x 32 d-rows x 128 B (64 keys)
constexpr int ATC_SMEM = ATC_Q_BYTES + ATC_K_BYTES + 2 * ATC_VT_PLANE + 1024 + 64;

// fp32 row segment (8 values) -> hi / lo bf16 chunks (16 B each)
__device__ __forceinline__ void split8(const float4 a, const float4 b, uint4& hi, uint4& lo) {
  const float f[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    split2_bf16(f[2 * i], f[2 * i + 1], h[i], l[i]);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__global__ void __launch_bounds__(128, 2)
attention_tc_kernel(const float* __restrict__ q, const int ldq, const float* __restrict__ k,
                    const float* __restrict__ v, const int ldk, const int Tq, const int Tk, const float scale,
                    __nv_bfloat16* __restrict__ out_split, const long long out_plane, float* __restrict__ out_f32,
                    const int ldo) {
  constexpr int DH = 32;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base_u32 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base_u32 - smem_u32(smem_raw));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + ATC_Q_BYTES;
  uint8_t* sVt = sK + ATC_K_BYTES;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sVt + 2 * ATC_VT_PLANE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int h = blockIdx.x, b = blockIdx.y;

  // ---- K rows -> [hi | lo] swizzled rows; V -> V^T key blocks (both planes); rows >= Tk are zero.
  // Every thread first requests ALL its global data (2 rows x 128 B), then converts: one memory round trip. ----
  {
    float4 kr[2][8];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int j = tid + 128 * u;
      const float* src = k + (static_cast<long long>(b) * Tk + j) * ldk + h * DH;
#pragma unroll
      for (int g = 0; g < 8; ++g)
        kr[u][g] = (j < Tk) ? *reinterpret_cast<const float4*>(src + 4 * g) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int j = tid + 128 * u;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 hi, lo;
        split8(kr[u][2 * g], kr[u][2 * g + 1], hi, lo);
        *reinterpret_cast<uint4*>(sK + j * 128 + ((g ^ (j & 7)) << 4)) = hi;
        *reinterpret_cast<uint4*>(sK + j * 128 + (((g + 4) ^ (j & 7)) << 4)) = lo;
      }
    }
  }
  {
    float4 vr[2][8];  // key pair (2 tid, 2 tid + 1)
    const int j = 2 * tid;
    const float* src = v + (static_cast<long long>(b) * Tk + j) * ldk + h * DH;
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int g = 0; g < 8; ++g)
        vr[u][g] = (j + u < Tk) ? *reinterpret_cast<const float4*>(src + static_cast<long long>(u) * ldk + 4 * g)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
    const int kb = j >> 6, jj = j & 63;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const float fa[4] = {vr[0][g].x, vr[0][g].y, vr[0][g].z, vr[0][g].w};
      const float fc[4] = {vr[1][g].x, vr[1][g].y, vr[1][g].z, vr[1][g].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int d = 4 * g + e;
        __nv_bfloat16 h0, l0, h1, l1;
        split_bf16(fa[e], h0, l0);
        split_bf16(fc[e], h1, l1);
        // a warp writes 32 consecutive key pairs of one d-row: consecutive words, conflict free
        const int off = kb * 4096 + d * 128 + (((jj >> 3) ^ (d & 7)) << 4) + (jj & 7) * 2;
        *reinterpret_cast<uint32_t*>(sVt + off) = pack_bf16(h0, h1);
        *reinterpret_cast<uint32_t*>(sVt + ATC_VT_PLANE + off) = pack_bf16(l0, l1);
      }
    }
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  // TMEM is requested only now: a second CTA resident on this SM converts its K / V while the first one computes
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tPh = tmem_base + 256, tPl = tmem_base + 384;
  const uint32_t trow = static_cast<uint32_t>(warp * 32) << 16;  // this warp's TMEM lane quadrant
  uint32_t phase = 0;
  const int nks = (Tk + 15) >> 4;  // P.V k-steps that hold real keys

  for (int q0 = 0; q0 < Tq; q0 += 128) {
    // ---- Q tile -> [hi | lo] rows (thread = row) ----
    {
      const int r = tid;
      float4 qr[8];
      const float* src = q + (static_cast<long long>(b) * Tq + q0 + r) * ldq + h * DH;
#pragma unroll
      for (int g = 0; g < 8; ++g)
        qr[g] = (q0 + r < Tq) ? *reinterpret_cast<const float4*>(src + 4 * g) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 hi, lo;
        split8(qr[2 * g], qr[2 * g + 1], hi, lo);
        *reinterpret_cast<uint4*>(sQ + r * 128 + ((g ^ (r & 7)) << 4)) = hi;
        *reinterpret_cast<uint4*>(sQ + r * 128 + (((g + 4) ^ (r & 7)) << 4)) = lo;
      }
    }
    fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core
    tc_fence_before();         // (also orders the previous tile's TMEM reads before the MMAs below)
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      constexpr uint32_t idesc_s = make_idesc(1, 128, 256);
      const uint64_t dq = make_sw128_kmajor_desc(smem_u32(sQ));
      const uint64_t dk = make_sw128_kmajor_desc(smem_u32(sK));
#pragma unroll
      for (int s = 0; s < 2; ++s) {  // K = 32 = 2 k-steps of 16; +4 = the lo half of the row (64 B)
        mma_bf16_ss(tS, dq + 2 * s, dk + 4 + 2 * s, idesc_s, s != 0);
        mma_bf16_ss(tS, dq + 4 + 2 * s, dk + 2 * s, idesc_s, 1);
        mma_bf16_ss(tS, dq + 2 * s, dk + 2 * s, idesc_s, 1);
      }
      tc_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- softmax: thread = query row ----
    float mx = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < 256; c += 32) {
      if (c >= Tk) break;
      uint32_t sv[32];
      tmem_ld_32x32(tS + trow + c, sv);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, (c + j < Tk) ? __uint_as_float(sv[j]) : -INFINITY);
    }
    float lsum = 0.f;
    const float sl2 = scale * 1.4426950408889634f, mxs = mx * sl2;  // exp(scale (s - mx)) = 2^(s sl2 - mx sl2)
#pragma unroll 1
    for (int c = 0; c < 256; c += 32) {
      uint32_t ph[16], pl[16];
      if (c < Tk) {
        uint32_t sv[32];
        tmem_ld_32x32(tS + trow + c, sv);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float p0 = (c + j < Tk) ? exp2f(fmaf(__uint_as_float(sv[j]), sl2, -mxs)) : 0.f;
          const float p1 = (c + j + 1 < Tk) ? exp2f(fmaf(__uint_as_float(sv[j + 1]), sl2, -mxs)) : 0.f;
          lsum += p0 + p1;
          split2_bf16(p0, p1, ph[j >> 1], pl[j >> 1]);
        }
      } else {
        if (c >= nks * 16) break;  // k-steps past the last key are never issued
#pragma unroll
        for (int j = 0; j < 16; ++j) ph[j] = pl[j] = 0u;
      }
      tmem_st_32x16(tPh + trow + (c >> 1), ph);
      tmem_st_32x16(tPl + trow + (c >> 1), pl);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      constexpr uint32_t idesc_o = make_idesc(1, 128, 32);
      for (int ks = 0; ks < nks; ++ks) {
        const uint32_t vt = smem_u32(sVt) + (ks >> 2) * 4096;
        const uint64_t bh = make_sw128_kmajor_desc(vt) + 2 * (ks & 3);
        const uint64_t bl = make_sw128_kmajor_desc(vt + ATC_VT_PLANE) + 2 * (ks & 3);
        mma_bf16_ts(tS, tPh + ks * 8, bl, idesc_o, ks != 0);  // O overwrites S columns [0, 32): S is consumed
        mma_bf16_ts(tS, tPl + ks * 8, bh, idesc_o, 1);
        mma_bf16_ts(tS, tPh + ks * 8, bh, idesc_o, 1);
      }
      tc_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    {
      uint32_t ov[32];
      tmem_ld_32x32(tS + trow, ov);
      tmem_ld_wait();
      const int r = q0 + tid;
      if (r < Tq) {
        const float inv = 1.f / lsum;
        const long long orow = (static_cast<long long>(b) * Tq + r) * ldo + h * DH;
        if (out_f32) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(out_f32 + orow + j) =
                make_float4(__uint_as_float(ov[j]) * inv, __uint_as_float(ov[j + 1]) * inv,
                            __uint_as_float(ov[j + 2]) * inv, __uint_as_float(ov[j + 3]) * inv);
        }
        if (out_split) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              split2_bf16(__uint_as_float(ov[j + 2 * e]) * inv, __uint_as_float(ov[j + 2 * e + 1]) * inv, hw[e], lw[e]);
            }
            *reinterpret_cast<uint4*>(out_split + orow + j) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            *reinterpret_cast<uint4*>(out_split + out_plane + orow + j) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Variant with TWO threads per query row (256 threads, RALF_ATTN_TC=2; built, not yet measured on hardware).
// The kernel above is bound by its softmax: one warp per scheduler walks 256 scores per row through tcgen05.ld ->
// exp2 -> two bf16 splits -> tcgen05.st, a serial chain the tensor pipe (4 % busy, profiles/r1_attn_tc_ncu.md) waits
// for.  Here warps 0-3 own key columns [0, 128) and warps 4-7 columns [128, 256) of the SAME 128 rows (a warp reaches
// the TMEM lane quadrant warp % 4, so both halves of a row live in one quadrant); row maxima and sums are combined
// through 2 KB of shared memory.  Operand conversion (K: one row per thread; V^T, Q, output: half a row per thread) is
// split the same way.  MMA issue, descriptors, TMEM layout and the order of the floating-point operations inside one
// half are those of the kernel above; only the row sum is now (left half) + (right half).
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}

constexpr int ATC2_SMEM = ATC_SMEM + 2 * 2 * 128 * 4;  // + row max / row sum exchange, [2 halves][128 rows] each

__global__ void __launch_bounds__(256, 2)
attention_tc2_kernel(const float* __restrict__ q, const int ldq, const float* __restrict__ k,
                     const float* __restrict__ v, const int ldk, const int Tq, const int Tk, const float scale,
                     __nv_bfloat16* __restrict__ out_split, const long long out_plane, float* __restrict__ out_f32,
                     const int ldo) {
  constexpr int DH = 32;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base_u32 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base_u32 - smem_u32(smem_raw));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + ATC_Q_BYTES;
  uint8_t* sVt = sK + ATC_K_BYTES;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sVt + 2 * ATC_VT_PLANE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  float* red_max = reinterpret_cast<float*>(sVt + 2 * ATC_VT_PLANE + 64);  // [2][128]
  float* red_sum = red_max + 256;                                          // [2][128]
  const int tid = threadIdx.x, warp = tid >> 5;
  const int half = warp >> 2;      // 0: key columns [0, 128), 1: [128, 256)
  const int row = tid & 127;       // query row inside the tile = TMEM lane
  const int h = blockIdx.x, b = blockIdx.y;

  // ---- K: one row per thread -> [hi | lo] swizzled row; rows >= Tk are zero ----
  {
    float4 kr[8];
    const int j = tid;
    const float* src = k + (static_cast<long long>(b) * Tk + j) * ldk + h * DH;
#pragma unroll
    for (int g = 0; g < 8; ++g)
      kr[g] = (j < Tk) ? *reinterpret_cast<const float4*>(src + 4 * g) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      uint4 hi, lo;
      split8(kr[2 * g], kr[2 * g + 1], hi, lo);
      *reinterpret_cast<uint4*>(sK + j * 128 + ((g ^ (j & 7)) << 4)) = hi;
      *reinterpret_cast<uint4*>(sK + j * 128 + (((g + 4) ^ (j & 7)) << 4)) = lo;
    }
  }
  // ---- V -> V^T key blocks (both planes): key pair (2 row, 2 row + 1), d range [16 half, 16 half + 16) ----
  {
    float4 vr[2][4];
    const int j = 2 * row;
    const float* src = v + (static_cast<long long>(b) * Tk + j) * ldk + h * DH + 16 * half;
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int g = 0; g < 4; ++g)
        vr[u][g] = (j + u < Tk) ? *reinterpret_cast<const float4*>(src + static_cast<long long>(u) * ldk + 4 * g)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
    const int kb = j >> 6, jj = j & 63;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const float fa[4] = {vr[0][g].x, vr[0][g].y, vr[0][g].z, vr[0][g].w};
      const float fc[4] = {vr[1][g].x, vr[1][g].y, vr[1][g].z, vr[1][g].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int d = 16 * half + 4 * g + e;
        __nv_bfloat16 h0, l0, h1, l1;
        split_bf16(fa[e], h0, l0);
        split_bf16(fc[e], h1, l1);
        const int off = kb * 4096 + d * 128 + (((jj >> 3) ^ (d & 7)) << 4) + (jj & 7) * 2;
        *reinterpret_cast<uint32_t*>(sVt + off) = pack_bf16(h0, h1);
        *reinterpret_cast<uint32_t*>(sVt + ATC_VT_PLANE + off) = pack_bf16(l0, l1);
      }
    }
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tPh = tmem_base + 256, tPl = tmem_base + 384;
  const uint32_t trow = static_cast<uint32_t>((warp & 3) * 32) << 16;  // this warp's TMEM lane quadrant
  uint32_t phase = 0;
  const int nks = (Tk + 15) >> 4;  // P.V k-steps that hold real keys
  const int c_lo = 128 * half, c_hi = c_lo + 128;

  for (int q0 = 0; q0 < Tq; q0 += 128) {
    // ---- Q tile -> [hi | lo] rows: thread (row, half) converts elements [16 half, 16 half + 16) ----
    {
      float4 qr[4];
      const float* src = q + (static_cast<long long>(b) * Tq + q0 + row) * ldq + h * DH + 16 * half;
#pragma unroll
      for (int g = 0; g < 4; ++g)
        qr[g] = (q0 + row < Tq) ? *reinterpret_cast<const float4*>(src + 4 * g) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int gg = 0; gg < 2; ++gg) {
        const int g = 2 * half + gg;  // 16-byte chunk of the hi half of the row; its lo twin is chunk g + 4
        uint4 hi, lo;
        split8(qr[2 * gg], qr[2 * gg + 1], hi, lo);
        *reinterpret_cast<uint4*>(sQ + row * 128 + ((g ^ (row & 7)) << 4)) = hi;
        *reinterpret_cast<uint4*>(sQ + row * 128 + (((g + 4) ^ (row & 7)) << 4)) = lo;
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      constexpr uint32_t idesc_s = make_idesc(1, 128, 256);
      const uint64_t dq = make_sw128_kmajor_desc(smem_u32(sQ));
      const uint64_t dk = make_sw128_kmajor_desc(smem_u32(sK));
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        mma_bf16_ss(tS, dq + 2 * s, dk + 4 + 2 * s, idesc_s, s != 0);
        mma_bf16_ss(tS, dq + 4 + 2 * s, dk + 2 * s, idesc_s, 1);
        mma_bf16_ss(tS, dq + 2 * s, dk + 2 * s, idesc_s, 1);
      }
      tc_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- softmax: thread = (query row, key half) ----
    float mx = -INFINITY;
#pragma unroll 1
    for (int c = c_lo; c < c_hi; c += 32) {
      if (c >= Tk) break;
      uint32_t sv[32];
      tmem_ld_32x32(tS + trow + c, sv);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, (c + j < Tk) ? __uint_as_float(sv[j]) : -INFINITY);
    }
    red_max[half * 128 + row] = mx;
    __syncthreads();
    mx = fmaxf(red_max[row], red_max[128 + row]);  // keys [0, 128) always hold a real key (Tk >= 64): finite
    float lsum = 0.f;
    const float sl2 = scale * 1.4426950408889634f, mxs = mx * sl2;
#pragma unroll 1
    for (int c = c_lo; c < c_hi; c += 32) {
      uint32_t ph[16], pl[16];
      if (c < Tk) {
        uint32_t sv[32];
        tmem_ld_32x32(tS + trow + c, sv);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float p0 = (c + j < Tk) ? exp2f(fmaf(__uint_as_float(sv[j]), sl2, -mxs)) : 0.f;
          const float p1 = (c + j + 1 < Tk) ? exp2f(fmaf(__uint_as_float(sv[j + 1]), sl2, -mxs)) : 0.f;
          lsum += p0 + p1;
          split2_bf16(p0, p1, ph[j >> 1], pl[j >> 1]);
        }
      } else {
        if (c >= nks * 16) break;
#pragma unroll
        for (int j = 0; j < 16; ++j) ph[j] = pl[j] = 0u;
      }
      tmem_st_32x16(tPh + trow + (c >> 1), ph);
      tmem_st_32x16(tPl + trow + (c >> 1), pl);
    }
    red_sum[half * 128 + row] = lsum;
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      constexpr uint32_t idesc_o = make_idesc(1, 128, 32);
      for (int ks = 0; ks < nks; ++ks) {
        const uint32_t vt = smem_u32(sVt) + (ks >> 2) * 4096;
        const uint64_t bh = make_sw128_kmajor_desc(vt) + 2 * (ks & 3);
        const uint64_t bl = make_sw128_kmajor_desc(vt + ATC_VT_PLANE) + 2 * (ks & 3);
        mma_bf16_ts(tS, tPh + ks * 8, bl, idesc_o, ks != 0);
        mma_bf16_ts(tS, tPl + ks * 8, bh, idesc_o, 1);
        mma_bf16_ts(tS, tPh + ks * 8, bh, idesc_o, 1);
      }
      tc_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    {
      uint32_t ov[16];  // O columns [16 half, 16 half + 16) of this row
      tmem_ld_32x16(tS + trow + 16 * half, ov);
      tmem_ld_wait();
      const int r = q0 + row;
      if (r < Tq) {
        const float inv = 1.f / (red_sum[row] + red_sum[128 + row]);
        const long long orow = (static_cast<long long>(b) * Tq + r) * ldo + h * DH + 16 * half;
        if (out_f32) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(out_f32 + orow + j) =
                make_float4(__uint_as_float(ov[j]) * inv, __uint_as_float(ov[j + 1]) * inv,
                            __uint_as_float(ov[j + 2]) * inv, __uint_as_float(ov[j + 3]) * inv);
        }
        if (out_split) {
#pragma unroll
          for (int j = 0; j < 16; j += 8) {
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              split2_bf16(__uint_as_float(ov[j + 2 * e]) * inv, __uint_as_float(ov[j + 2 * e + 1]) * inv, hw[e], lw[e]);
            }
            *reinterpret_cast<uint4*>(out_split + orow + j) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            *reinterpret_cast<uint4*>(out_split + out_plane + orow + j) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
          }
        }
      }
    }
    // red_max / red_sum of this tile are read above; the next tile's Q-conversion barrier orders them before reuse
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Variant for 256 < Tk <= 480 keys (default for these shapes since round 2; RALF_ATTN_TC_BIG=0 switches it off): the reference's real canvases
// are 350 x 240 -> 22 x 15 = 330 image tokens, which the kernels above hand to the CUDA-core fallback.
// S[128, Tk] no longer leaves room for separate P planes in the 512 TMEM columns, so P is written IN PLACE: a thread
// that has read its 32 S columns [c, c + 32) of a chunk stores the chunk's P as 16 hi columns [c, c + 16) and 16 lo
// columns [c + 16, c + 32) over them (same lane, so no other thread or MMA is affected), and the P.V MMAs take their
// A operand of k-step ks (keys 16 ks .. 16 ks + 15) from column 32 (ks / 2) + 8 (ks % 2) (+ 16 for lo).  O gets its own
// columns [480, 512).  S is issued as two MMA groups (an MMA is at most 256 wide): keys [0, 256) and [256, 256 + N2).
// ---------------------------------------------------------------------------------------------------------------
constexpr int ATCB_MAX_KEYS = 480;
constexpr int ATCB_K_BYTES = ATCB_MAX_KEYS * 128;
constexpr int ATCB_VT_PLANE = 8 * 32 * 128;  // 8 key blocks of 64 keys
constexpr int ATCB_SMEM = ATC_Q_BYTES + ATCB_K_BYTES + 2 * ATCB_VT_PLANE + 1024 + 64;

__global__ void __launch_bounds__(128, 1)
attention_tc_big_kernel(const float* __restrict__ q, const int ldq, const float* __restrict__ k,
                        const float* __restrict__ v, const int ldk, const int Tq, const int Tk, const float scale,
                        __nv_bfloat16* __restrict__ out_split, const long long out_plane, float* __restrict__ out_f32,
                        const int ldo) {
  constexpr int DH = 32;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base_u32 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base_u32 - smem_u32(smem_raw));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + ATC_Q_BYTES;
  uint8_t* sVt = sK + ATCB_K_BYTES;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sVt + 2 * ATCB_VT_PLANE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int h = blockIdx.x, b = blockIdx.y;
  const int n2 = ((Tk - 256) + 15) & ~15;  // width of the second S group (keys 256 .. 256 + n2), 16 .. 224
  const int nks = (Tk + 15) >> 4;          // P.V k-steps that hold real keys
  const int krows = 256 + n2;              // K rows the S MMAs read

  // ---- K rows -> [hi | lo] swizzled rows, zero beyond Tk ----
#pragma unroll 1
  for (int j = tid; j < krows; j += 128) {
    float4 kr[8];
    const float* src = k + (static_cast<long long>(b) * Tk + j) * ldk + h * DH;
#pragma unroll
    for (int g = 0; g < 8; ++g)
      kr[g] = (j < Tk) ? *reinterpret_cast<const float4*>(src + 4 * g) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      uint4 hi, lo;
      split8(kr[2 * g], kr[2 * g + 1], hi, lo);
      *reinterpret_cast<uint4*>(sK + j * 128 + ((g ^ (j & 7)) << 4)) = hi;
      *reinterpret_cast<uint4*>(sK + j * 128 + (((g + 4) ^ (j & 7)) << 4)) = lo;
    }
  }
  // ---- V -> V^T key blocks (both planes), key pairs; zero beyond Tk up to the last k-step ----
#pragma unroll 1
  for (int j = 2 * tid; j < nks * 16; j += 256) {
    float4 vr[2][8];
    const float* src = v + (static_cast<long long>(b) * Tk + j) * ldk + h * DH;
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int g = 0; g < 8; ++g)
        vr[u][g] = (j + u < Tk) ? *reinterpret_cast<const float4*>(src + static_cast<long long>(u) * ldk + 4 * g)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
    const int kb = j >> 6, jj = j & 63;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const float fa[4] = {vr[0][g].x, vr[0][g].y, vr[0][g].z, vr[0][g].w};
      const float fc[4] = {vr[1][g].x, vr[1][g].y, vr[1][g].z, vr[1][g].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int d = 4 * g + e;
        __nv_bfloat16 h0, l0, h1, l1;
        split_bf16(fa[e], h0, l0);
        split_bf16(fc[e], h1, l1);
        const int off = kb * 4096 + d * 128 + (((jj >> 3) ^ (d & 7)) << 4) + (jj & 7) * 2;
        *reinterpret_cast<uint32_t*>(sVt + off) = pack_bf16(h0, h1);
        *reinterpret_cast<uint32_t*>(sVt + ATCB_VT_PLANE + off) = pack_bf16(l0, l1);
      }
    }
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tO = tmem_base + ATCB_MAX_KEYS;
  const uint32_t trow = static_cast<uint32_t>(warp * 32) << 16;
  uint32_t phase = 0;

  for (int q0 = 0; q0 < Tq; q0 += 128) {
    {
      const int r = tid;
      float4 qr[8];
      const float* src = q + (static_cast<long long>(b) * Tq + q0 + r) * ldq + h * DH;
#pragma unroll
      for (int g = 0; g < 8; ++g)
        qr[g] = (q0 + r < Tq) ? *reinterpret_cast<const float4*>(src + 4 * g) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 hi, lo;
        split8(qr[2 * g], qr[2 * g + 1], hi, lo);
        *reinterpret_cast<uint4*>(sQ + r * 128 + ((g ^ (r & 7)) << 4)) = hi;
        *reinterpret_cast<uint4*>(sQ + r * 128 + (((g + 4) ^ (r & 7)) << 4)) = lo;
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();  // (also orders the previous tile's TMEM reads before the MMAs below)
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint64_t dq = make_sw128_kmajor_desc(smem_u32(sQ));
      for (int grp = 0; grp < 2; ++grp) {
        const uint32_t idesc_s = make_idesc(1, 128, grp == 0 ? 256u : static_cast<uint32_t>(n2));
        const uint64_t dk = make_sw128_kmajor_desc(smem_u32(sK) + grp * 256 * 128);
        const uint32_t td = tS + grp * 256;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          mma_bf16_ss(td, dq + 2 * s, dk + 4 + 2 * s, idesc_s, s != 0);
          mma_bf16_ss(td, dq + 4 + 2 * s, dk + 2 * s, idesc_s, 1);
          mma_bf16_ss(td, dq + 2 * s, dk + 2 * s, idesc_s, 1);
        }
      }
      tc_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- softmax: thread = query row; P overwrites S chunk by chunk ----
    float mx = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < Tk; c += 32) {
      uint32_t sv[32];
      tmem_ld_32x32(tS + trow + c, sv);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) mx = fmaxf(mx, (c + j < Tk) ? __uint_as_float(sv[j]) : -INFINITY);
    }
    float lsum = 0.f;
    const float sl2 = scale * 1.4426950408889634f, mxs = mx * sl2;
#pragma unroll 1
    for (int c = 0; c < nks * 16; c += 32) {
      uint32_t ph[16], pl[16];
      uint32_t sv[32];
      tmem_ld_32x32(tS + trow + c, sv);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float p0 = (c + j < Tk) ? exp2f(fmaf(__uint_as_float(sv[j]), sl2, -mxs)) : 0.f;
        const float p1 = (c + j + 1 < Tk) ? exp2f(fmaf(__uint_as_float(sv[j + 1]), sl2, -mxs)) : 0.f;
        lsum += p0 + p1;
        split2_bf16(p0, p1, ph[j >> 1], pl[j >> 1]);
      }
      tmem_st_32x16(tS + trow + c, ph);       // hi of keys c .. c + 31 (2 keys per column)
      tmem_st_32x16(tS + trow + c + 16, pl);  // lo
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      constexpr uint32_t idesc_o = make_idesc(1, 128, 32);
      for (int ks = 0; ks < nks; ++ks) {
        const uint32_t vt = smem_u32(sVt) + (ks >> 2) * 4096;
        const uint64_t bh = make_sw128_kmajor_desc(vt) + 2 * (ks & 3);
        const uint64_t bl = make_sw128_kmajor_desc(vt + ATCB_VT_PLANE) + 2 * (ks & 3);
        const uint32_t ah = tS + 32 * (ks >> 1) + 8 * (ks & 1), al = ah + 16;
        mma_bf16_ts(tO, ah, bl, idesc_o, ks != 0);
        mma_bf16_ts(tO, al, bh, idesc_o, 1);
        mma_bf16_ts(tO, ah, bh, idesc_o, 1);
      }
      tc_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    {
      uint32_t ov[32];
      tmem_ld_32x32(tO + trow, ov);
      tmem_ld_wait();
      const int r = q0 + tid;
      if (r < Tq) {
        const float inv = 1.f / lsum;
        const long long orow = (static_cast<long long>(b) * Tq + r) * ldo + h * DH;
        if (out_f32) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(out_f32 + orow + j) =
                make_float4(__uint_as_float(ov[j]) * inv, __uint_as_float(ov[j + 1]) * inv,
                            __uint_as_float(ov[j + 2]) * inv, __uint_as_float(ov[j + 3]) * inv);
        }
        if (out_split) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              split2_bf16(__uint_as_float(ov[j + 2 * e]) * inv, __uint_as_float(ov[j + 2 * e + 1]) * inv, hw[e], lw[e]);
            }
            *reinterpret_cast<uint4*>(out_split + orow + j) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            *reinterpret_cast<uint4*>(out_split + out_plane + orow + j) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Variant with HALF the tensor memory (RALF_ATTN_TC=3, default since round 2): attention_tc_kernel allocates all 512
// TMEM columns, so the second CTA resident on an SM blocks in tcgen05.alloc until the first one is done -- one 4-warp CTA
// per SM then serialises operand conversion -> MMA -> softmax -> MMA with the tensor pipe 4 % busy and the issue slots
// 19 % (profiles/r1_attn_tc_ncu.md).  Here a CTA takes 256 columns, so TWO CTAs compute on an SM at the same time and
// fill each other's MMA round trips:
//   * the keys are processed in two halves of 128: S_h = Q K_h^T is a 128 x 128 tile (columns [0, 128));
//   * P_h = exp(scale (S_h - rowmax_h)) is written IN PLACE over S_h (the 32 fp32 scores of a chunk become 16 columns
//     of hi pairs + 16 columns of lo pairs);
//   * each half has its own accumulator O_h (columns 128 + 64 h ..), and V^T is stored [hi rows ; lo rows] per key
//     block so that ONE MMA with N = 64 yields P_hi.V_hi | P_hi.V_lo and a second (N = 32) adds P_lo.V_hi: two MMAs per
//     k-step instead of three (an MMA at M = 128 costs ~92 cycles however narrow N is);
//   * the halves are combined in registers: out = (a_0 O_0 + a_1 O_1) / (a_0 l_0 + a_1 l_1), a_h = exp(scale (m_h - m)).
// Same operand precision (split bf16, fp32 accumulation) as the kernel above; the two-half softmax differs from the
// single-pass one by fp32 rounding only.
// ---------------------------------------------------------------------------------------------------------------
constexpr int ATC3_VT_BLOCK = 2 * 32 * 128;  // one 64-key block: [hi: 32 d-rows x 128 B][lo: 32 d-rows x 128 B]
constexpr int ATC3_SMEM = ATC_Q_BYTES + ATC_K_BYTES + 4 * ATC3_VT_BLOCK + 1024 + 64;

__global__ void __launch_bounds__(128, 2)
attention_tc3_kernel(const float* __restrict__ q, const int ldq, const float* __restrict__ k,
                     const float* __restrict__ v, const int ldk, const int Tq, const int Tk, const float scale,
                     __nv_bfloat16* __restrict__ out_split, const long long out_plane, float* __restrict__ out_f32,
                     const int ldo) {
  constexpr int DH = 32;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base_u32 = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base_u32 - smem_u32(smem_raw));
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + ATC_Q_BYTES;
  uint8_t* sVt = sK + ATC_K_BYTES;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sVt + 4 * ATC3_VT_BLOCK);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5;
  const int h = blockIdx.x, b = blockIdx.y;

  {  // K rows -> [hi | lo] swizzled rows (rows >= Tk are zero)
    float4 kr[2][8];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int j = tid + 128 * u;
      const float* src = k + (static_cast<long long>(b) * Tk + j) * ldk + h * DH;
#pragma unroll
      for (int g = 0; g < 8; ++g)
        kr[u][g] = (j < Tk) ? *reinterpret_cast<const float4*>(src + 4 * g) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int j = tid + 128 * u;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 hi, lo;
        split8(kr[u][2 * g], kr[u][2 * g + 1], hi, lo);
        *reinterpret_cast<uint4*>(sK + j * 128 + ((g ^ (j & 7)) << 4)) = hi;
        *reinterpret_cast<uint4*>(sK + j * 128 + (((g + 4) ^ (j & 7)) << 4)) = lo;
      }
    }
  }
  {  // V -> V^T key blocks, hi rows then lo rows; thread = key pair (2 tid, 2 tid + 1)
    float4 vr[2][8];
    const int j = 2 * tid;
    const float* src = v + (static_cast<long long>(b) * Tk + j) * ldk + h * DH;
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
      for (int g = 0; g < 8; ++g)
        vr[u][g] = (j + u < Tk) ? *reinterpret_cast<const float4*>(src + static_cast<long long>(u) * ldk + 4 * g)
                                : make_float4(0.f, 0.f, 0.f, 0.f);
    const int kb = j >> 6, jj = j & 63;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const float fa[4] = {vr[0][g].x, vr[0][g].y, vr[0][g].z, vr[0][g].w};
      const float fc[4] = {vr[1][g].x, vr[1][g].y, vr[1][g].z, vr[1][g].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int d = 4 * g + e;
        uint32_t hp, lp;
        split2_bf16(fa[e], fc[e], hp, lp);
        const int off = kb * ATC3_VT_BLOCK + d * 128 + (((jj >> 3) ^ (d & 7)) << 4) + (jj & 7) * 2;
        *reinterpret_cast<uint32_t*>(sVt + off) = hp;
        *reinterpret_cast<uint32_t*>(sVt + 4096 + off) = lp;
      }
    }
  }
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base;
  const uint32_t trow = static_cast<uint32_t>(warp * 32) << 16;  // this warp's TMEM lane quadrant
  uint32_t phase = 0;
  const int nhalf = Tk > 128 ? 2 : 1;
  const float sl2 = scale * 1.4426950408889634f;

  for (int q0 = 0; q0 < Tq; q0 += 128) {
    {  // Q tile -> [hi | lo] rows (thread = row)
      const int r = tid;
      float4 qr[8];
      const float* src = q + (static_cast<long long>(b) * Tq + q0 + r) * ldq + h * DH;
#pragma unroll
      for (int g = 0; g < 8; ++g)
        qr[g] = (q0 + r < Tq) ? *reinterpret_cast<const float4*>(src + 4 * g) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        uint4 hi, lo;
        split8(qr[2 * g], qr[2 * g + 1], hi, lo);
        *reinterpret_cast<uint4*>(sQ + r * 128 + ((g ^ (r & 7)) << 4)) = hi;
        *reinterpret_cast<uint4*>(sQ + r * 128 + (((g + 4) ^ (r & 7)) << 4)) = lo;
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    float mh[2] = {-INFINITY, -INFINITY}, lh[2] = {0.f, 0.f};
#pragma unroll
    for (int kh = 0; kh < 2; ++kh) {
      if (kh < nhalf) {
        const int tkh = min(Tk - kh * 128, 128);   // keys of this half
        const int nks = (tkh + 15) >> 4;           // P.V k-steps that hold real keys
        if (tid == 0) {
          tc_fence_after();
          constexpr uint32_t idesc_s = make_idesc(1, 128, 128);
          const uint64_t dq = make_sw128_kmajor_desc(smem_u32(sQ));
          const uint64_t dk = make_sw128_kmajor_desc(smem_u32(sK) + kh * 128 * 128);
#pragma unroll
          for (int s = 0; s < 2; ++s) {  // K = 32 = 2 k-steps of 16; +4 = the lo half of the row (64 B)
            mma_bf16_ss(tS, dq + 2 * s, dk + 4 + 2 * s, idesc_s, s != 0);
            mma_bf16_ss(tS, dq + 4 + 2 * s, dk + 2 * s, idesc_s, 1);
            mma_bf16_ss(tS, dq + 2 * s, dk + 2 * s, idesc_s, 1);
          }
          tc_commit(bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after();
        float mx = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < 128; c += 32) {
          if (c >= tkh) break;
          uint32_t sv[32];
          tmem_ld_32x32(tS + trow + c, sv);
          tmem_ld_wait();
          if (c + 32 <= tkh) {  // warp-uniform: only the last chunk of a ragged key count needs the per-key mask
#pragma unroll
            for (int j = 0; j < 32; ++j) mx = fmaxf(mx, __uint_as_float(sv[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) mx = fmaxf(mx, (c + j < tkh) ? __uint_as_float(sv[j]) : -INFINITY);
          }
        }
        float lsum = 0.f;
        const float mxs = mx * sl2;
#pragma unroll 1
        for (int c = 0; c < 128; c += 32) {
          uint32_t ph[16], pl[16];
          if (c < tkh) {
            uint32_t sv[32];
            tmem_ld_32x32(tS + trow + c, sv);
            tmem_ld_wait();
            if (c + 32 <= tkh) {
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                const float p0 = ex2_ftz(fmaf(__uint_as_float(sv[j]), sl2, -mxs));
                const float p1 = ex2_ftz(fmaf(__uint_as_float(sv[j + 1]), sl2, -mxs));
                lsum += p0 + p1;
                split2_bf16(p0, p1, ph[j >> 1], pl[j >> 1]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                const float p0 = (c + j < tkh) ? ex2_ftz(fmaf(__uint_as_float(sv[j]), sl2, -mxs)) : 0.f;
                const float p1 = (c + j + 1 < tkh) ? ex2_ftz(fmaf(__uint_as_float(sv[j + 1]), sl2, -mxs)) : 0.f;
                lsum += p0 + p1;
                split2_bf16(p0, p1, ph[j >> 1], pl[j >> 1]);
              }
            }
          } else {
            if (c >= nks * 16) break;  // k-steps past the last key are never issued
#pragma unroll
            for (int j = 0; j < 16; ++j) ph[j] = pl[j] = 0u;
          }
          tmem_st_32x16(tS + trow + c, ph);       // in place: hi pairs over the chunk's first 16 columns,
          tmem_st_32x16(tS + trow + c + 16, pl);  // lo pairs over the other 16
        }
        mh[kh] = mx;
        lh[kh] = lsum;
        tmem_st_wait();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
          tc_fence_after();
          constexpr uint32_t idesc_64 = make_idesc(1, 128, 64);
          constexpr uint32_t idesc_32 = make_idesc(1, 128, 32);
          const uint32_t tO = tmem_base + 128 + 64 * kh;
          for (int ks = 0; ks < nks; ++ks) {
            const uint32_t vt = smem_u32(sVt) + (kh * 2 + (ks >> 2)) * ATC3_VT_BLOCK;
            const uint64_t bd = make_sw128_kmajor_desc(vt) + 2 * (ks & 3);
            const uint32_t a_hi = tS + (ks >> 1) * 32 + (ks & 1) * 8;
            mma_bf16_ts(tO, a_hi, bd, idesc_64, ks != 0);      // [P_hi.V_hi | P_hi.V_lo]
            mma_bf16_ts(tO + 32, a_hi + 16, bd, idesc_32, 1);  // + P_lo.V_hi
          }
          tc_commit(bar);
        }
        mbar_wait(bar, phase);  // P is consumed: the next half's S may overwrite it
        phase ^= 1;
        tc_fence_after();
      }
    }
    {
      const float m = fmaxf(mh[0], mh[1]);
      const float a0 = ex2_ftz((mh[0] - m) * sl2);
      const float a1 = nhalf > 1 ? ex2_ftz((mh[1] - m) * sl2) : 0.f;
      const float inv = 1.f / (a0 * lh[0] + a1 * lh[1]);
      float o[32];
      {
        uint32_t u0[32], u1[32];
        tmem_ld_32x32(tmem_base + 128 + trow, u0);
        tmem_ld_32x32(tmem_base + 160 + trow, u1);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] = a0 * (__uint_as_float(u0[j]) + __uint_as_float(u1[j]));
      }
      if (nhalf > 1) {
        uint32_t u0[32], u1[32];
        tmem_ld_32x32(tmem_base + 192 + trow, u0);
        tmem_ld_32x32(tmem_base + 224 + trow, u1);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) o[j] = fmaf(a1, __uint_as_float(u0[j]) + __uint_as_float(u1[j]), o[j]);
      }
      const int r = q0 + tid;
      if (r < Tq) {
        const long long orow = (static_cast<long long>(b) * Tq + r) * ldo + h * DH;
        if (out_f32) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(out_f32 + orow + j) = make_float4(o[j] * inv, o[j + 1] * inv, o[j + 2] * inv, o[j + 3] * inv);
        }
        if (out_split) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint32_t hw[4], lw[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              split2_bf16(o[j + 2 * e] * inv, o[j + 2 * e + 1] * inv, hw[e], lw[e]);
            }
            *reinterpret_cast<uint4*>(out_split + orow + j) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
            *reinterpret_cast<uint4*>(out_split + out_plane + orow + j) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

// Returns 1 when the tensor-core kernel took the call, 0 when the shape is outside its class, < 0 on error.
int attention_tc_try(const float* q, int ldq, const float* k, const float* v, int ldk, const unsigned char* mask, int B,
                     int H, int Tq, int Tk, int head_dim, int causal, float scale, void* out_split,
                     long long out_plane, float* out_f32, int ldo, cudaStream_t st) {
  // 0 off, 1 = round-1 kernel (512 TMEM columns), 2 = 8-warp kernel, 3 = half-TMEM kernel, two computing CTAs per SM (default)
  static const int variant = getenv("RALF_ATTN_TC") ? atoi(getenv("RALF_ATTN_TC")) : 3;
  const bool enabled = variant != 0;
  // 256 < Tk <= 480 (the reference's real 350 x 240 canvases: 330 image tokens) on the tensor cores: default since round 2
  // (verified against fp64 on hardware; encode of 128 real-size canvases 15.4 vs 18.0 ms); RALF_ATTN_TC_BIG=0 = CUDA cores
  static const bool big = !(getenv("RALF_ATTN_TC_BIG") && atoi(getenv("RALF_ATTN_TC_BIG")) == 0);
  const bool use_big = big && Tk > 256 && Tk <= ATCB_MAX_KEYS;
  if (!enabled || mask || causal || head_dim != 32 || (Tk > 256 && !use_big) || Tk < 64 || Tq < 64) return 0;
  if ((ldq & 3) || (ldk & 3) || (ldo & 7) || (reinterpret_cast<uintptr_t>(out_split) & 15) ||
      (reinterpret_cast<uintptr_t>(out_f32) & 15) || (out_plane & 7))
    return 0;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC_SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                               cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return set_cuda_error(e);
    attr_set = true;
  }
  if (use_big) {
    static bool attrb_set = false;
    if (!attrb_set) {
      cudaError_t e = cudaFuncSetAttribute(attention_tc_big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATCB_SMEM);
      if (e != cudaSuccess) return set_cuda_error(e);
      attrb_set = true;
    }
    attention_tc_big_kernel<<<dim3(H, B), 128, ATCB_SMEM, st>>>(q, ldq, k, v, ldk, Tq, Tk, scale,
                                                                reinterpret_cast<__nv_bfloat16*>(out_split), out_plane,
                                                                out_f32, ldo);
    const int rcb = set_cuda_error(cudaGetLastError());
    return rcb ? rcb : 1;
  }
  if (variant == 3) {
    static bool attr3_set = false;
    if (!attr3_set) {
      cudaError_t e = cudaFuncSetAttribute(attention_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC3_SMEM);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(attention_tc3_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared);
      if (e != cudaSuccess) return set_cuda_error(e);
      attr3_set = true;
    }
    attention_tc3_kernel<<<dim3(H, B), 128, ATC3_SMEM, st>>>(q, ldq, k, v, ldk, Tq, Tk, scale,
                                                             reinterpret_cast<__nv_bfloat16*>(out_split), out_plane,
                                                             out_f32, ldo);
    const int rc3 = set_cuda_error(cudaGetLastError());
    return rc3 ? rc3 : 1;
  }
  if (variant == 2) {
    static bool attr2_set = false;
    if (!attr2_set) {
      cudaError_t e = cudaFuncSetAttribute(attention_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATC2_SMEM);
      if (e == cudaSuccess)
        e = cudaFuncSetAttribute(attention_tc2_kernel, cudaFuncAttributePreferredSharedMemoryCarveout,
                                 cudaSharedmemCarveoutMaxShared);
      if (e != cudaSuccess) return set_cuda_error(e);
      attr2_set = true;
    }
    attention_tc2_kernel<<<dim3(H, B), 256, ATC2_SMEM, st>>>(q, ldq, k, v, ldk, Tq, Tk, scale,
                                                             reinterpret_cast<__nv_bfloat16*>(out_split), out_plane,
                                                             out_f32, ldo);
    const int rc2 = set_cuda_error(cudaGetLastError());
    return rc2 ? rc2 : 1;
  }
  attention_tc_kernel<<<dim3(H, B), 128, ATC_SMEM, st>>>(q, ldq, k, v, ldk, Tq, Tk, scale,
                                                          reinterpret_cast<__nv_bfloat16*>(out_split), out_plane,
                                                          out_f32, ldo);
  const int rc = set_cuda_error(cudaGetLastError());
  return rc ? rc : 1;
}

}  // namespace ralf
