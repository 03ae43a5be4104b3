// Non-GEMM kernels of the RALF forward / generate path: LayerNorm, multi-head attention (full and
// KV-cached single-query), im2col for the ResNet convolutions, max-pool, FPN merge, embeddings,
// row glue, masked arg-max.  All fp32 arithmetic on CUDA cores (these ops are HBM/latency bound;
// the dense contractions around them run on tcgen05 in gemm.cu).  Activations that feed a GEMM are
// emitted in the split-bf16 format (hi plane, lo plane) by the producing kernel.
#include <float.h>
#include <math.h>

#include <stdlib.h>

#include "common.cuh"
#include "ralf_internal.h"

namespace ralf {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, off));
  return v;
}
__device__ __forceinline__ void store_split(__nv_bfloat16* hi_plane, long long plane, long long off, float x) {
  __nv_bfloat16 h, l;
  split_bf16(x, h, l);
  hi_plane[off] = h;
  hi_plane[plane + off] = l;
}
__device__ __forceinline__ float load_split(const __nv_bfloat16* hi_plane, long long plane, long long off) {
  return __bfloat162float(hi_plane[off]) + __bfloat162float(hi_plane[plane + off]);
}

// ------------------------------------------------------------------------------------------------
// LayerNorm over the last dim (nn.LayerNorm, eps 1e-5; common/attention.py:20, nn.Transformer*Layer
// norms).  One warp per row; D <= 1024, D % 32 == 0.
// ------------------------------------------------------------------------------------------------
// PER = D / 32 as a template parameter keeps the row in registers (a run-time trip count indexes v[] dynamically and
// pushes it to local memory).
template <int PER>
__global__ void layernorm_kernel(const float* __restrict__ x, long long in_ld, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, float eps, int M, int D,
                                 float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_split,
                                 long long out_plane) {
  pdl_trigger();
  pdl_wait();  // every thread, before the early return: keeps the completion order of a PDL chain
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const float* xr = x + static_cast<long long>(row) * in_ld;
  float v[PER];
  constexpr int per = PER;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < per; ++i) {
    v[i] = xr[lane + 32 * i];
    s += v[i];
  }
  const float mean = warp_sum(s) / static_cast<float>(D);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < per; ++i) {
    const float dlt = v[i] - mean;
    sq += dlt * dlt;
  }
  const float rstd = rsqrtf(warp_sum(sq) / static_cast<float>(D) + eps);
#pragma unroll
  for (int i = 0; i < per; ++i) {
    const int c = lane + 32 * i;
    const float y = (v[i] - mean) * rstd * gamma[c] + beta[c];
    const long long off = static_cast<long long>(row) * D + c;
    if (out_f32) out_f32[off] = y;
    if (out_split) store_split(out_split, out_plane, off, y);
  }
}

// ------------------------------------------------------------------------------------------------
// Multi-head attention, fp32, online softmax.  One thread per query, K/V tiles of 64 keys staged in
// shared memory (broadcast reads).  Covers nn.MultiheadAttention inside nn.TransformerEncoderLayer /
// DecoderLayer (common/common.py:26-35, retrieval_augmented_autoreg.py:116-126, fid/model.py:26-33)
// and the fusion Attention (common/attention.py:49-71).
//   q row (b, t): q + (b*Tq + t)*ldq + h*DH ; k/v row (b, j): k + (b*Tk + j)*ldk + h*DH
//   mask: uint8 [B, Tk], nonzero = key is padding (-inf);  causal: key j > query t masked.
// ------------------------------------------------------------------------------------------------
template <int DH>
__global__ void __launch_bounds__(128)
attention_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k, const float* __restrict__ v,
                 int ldk, const unsigned char* __restrict__ mask, int Tq, int Tk, int causal, float scale,
                 __nv_bfloat16* __restrict__ out_split, long long out_plane, float* __restrict__ out_f32, int ldo,
                 DropArgs da, float* __restrict__ lse_out) {
  constexpr int KT = 64;
  __shared__ __align__(16) float ks[KT][DH];
  __shared__ __align__(16) float vs[KT][DH];
  __shared__ unsigned char ms[KT];
  const int b = blockIdx.z, h = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = t < Tq;
  float qr[DH], o[DH];
  if (active) {
    const float* qp = q + (static_cast<long long>(b) * Tq + t) * ldq + h * DH;
#pragma unroll
    for (int i = 0; i < DH; i += 4) {
      const float4 f = *reinterpret_cast<const float4*>(qp + i);
      qr[i] = f.x * scale; qr[i + 1] = f.y * scale; qr[i + 2] = f.z * scale; qr[i + 3] = f.w * scale;
    }
  }
#pragma unroll
  for (int i = 0; i < DH; ++i) o[i] = 0.f;
  float mrun = -INFINITY, lrun = 0.f;
  // attention-probability dropout (nn.MultiheadAttention(dropout=p) in training): O = (P o M / keep) V, P normalised
  // over ALL keys; mask element index = ((b*H + h)*Tq + t)*Tk + j
  const unsigned long long dstream = da.thresh24 ? drop_stream(*da.seed, da.site) : 0ull;
  const unsigned long long drow = ((static_cast<unsigned long long>(b) * gridDim.y + h) * Tq + t) * Tk;
  const int kmax = causal ? min(Tk, (blockIdx.x + 1) * static_cast<int>(blockDim.x)) : Tk;
  for (int j0 = 0; j0 < kmax; j0 += KT) {
    __syncthreads();
    const int nk = min(KT, Tk - j0);
    for (int i = threadIdx.x; i < KT * (DH / 4); i += blockDim.x) {
      const int r = i / (DH / 4), c = (i % (DH / 4)) * 4;
      float4 fk = make_float4(0.f, 0.f, 0.f, 0.f), fv = fk;
      if (r < nk) {
        const long long off = (static_cast<long long>(b) * Tk + j0 + r) * ldk + h * DH + c;
        fk = *reinterpret_cast<const float4*>(k + off);
        fv = *reinterpret_cast<const float4*>(v + off);
      }
      *reinterpret_cast<float4*>(&ks[r][c]) = fk;
      *reinterpret_cast<float4*>(&vs[r][c]) = fv;
    }
    for (int i = threadIdx.x; i < KT; i += blockDim.x)
      ms[i] = (i < nk) ? (mask ? mask[static_cast<long long>(b) * Tk + j0 + i] : 0) : 1;
    __syncthreads();
    if (!active) continue;
    for (int c0 = 0; c0 < nk; c0 += 8) {
      float s[8];
      float cmax = -INFINITY;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const int j = c0 + jj;
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < DH; i += 4) {
          const float4 f = *reinterpret_cast<const float4*>(&ks[j & (KT - 1)][i]);
          acc = fmaf(qr[i], f.x, acc); acc = fmaf(qr[i + 1], f.y, acc);
          acc = fmaf(qr[i + 2], f.z, acc); acc = fmaf(qr[i + 3], f.w, acc);
        }
        const bool dead = (j >= nk) || ms[j & (KT - 1)] || (causal && (j0 + j) > t);
        s[jj] = dead ? -INFINITY : acc;
        cmax = fmaxf(cmax, s[jj]);
      }
      if (cmax == -INFINITY) continue;
      const float mnew = fmaxf(mrun, cmax);
      const float corr = __expf(mrun - mnew);  // mrun = -inf -> 0
      lrun *= corr;
#pragma unroll
      for (int i = 0; i < DH; ++i) o[i] *= corr;
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        float p = __expf(s[jj] - mnew);  // -inf -> 0
        lrun += p;
        if (da.thresh24) p = drop_keep(dstream, drow + j0 + c0 + jj, da.thresh24) ? p * da.inv_keep : 0.f;
        const int j = (c0 + jj) & (KT - 1);
#pragma unroll
        for (int i = 0; i < DH; i += 4) {
          const float4 f = *reinterpret_cast<const float4*>(&vs[j][i]);
          o[i] = fmaf(p, f.x, o[i]); o[i + 1] = fmaf(p, f.y, o[i + 1]);
          o[i + 2] = fmaf(p, f.z, o[i + 2]); o[i + 3] = fmaf(p, f.w, o[i + 3]);
        }
      }
      mrun = mnew;
    }
  }
  if (!active) return;
  // training: the log-sum-exp of the row, so that the backward (attn_bwd_dq_kernel) need not sweep the keys twice
  if (lse_out) lse_out[(static_cast<long long>(b) * gridDim.y + h) * Tq + t] = mrun + __logf(lrun);
  const float inv = 1.f / lrun;
  const long long orow = (static_cast<long long>(b) * Tq + t) * ldo + h * DH;
#pragma unroll
  for (int i = 0; i < DH; ++i) {
    const float y = o[i] * inv;
    if (out_f32) out_f32[orow + i] = y;
    if (out_split) store_split(out_split, out_plane, orow + i, y);
  }
}

// ------------------------------------------------------------------------------------------------
// Attention over a handful of keys (the fusion Attention of RALF: 256-330 image tokens attend to the 16 retrieved
// layouts, 8 heads x 64; common/attention.py:49-71).  The generic kernel above is bound by its scattered 2-byte
// output stores here (HBM floor 20 us, measured 198 us per 128 canvases); this one keeps a 32-query x H-head tile per
// CTA (warp = head, lane = query), reads K/V through L1 broadcast loads, and stages the output tile in shared memory
// so that every global store is a full 128/256-byte row segment.
// Also the FIDNetV3 layout encoder (fid/model.py:26-33: B*16 sequences of <= 11 tokens, 4 heads x 64, key padding).
//   q row (b, t): q + (b*Tq + t)*ldq + h*64 ; k/v row (b, j): base + (b*Tk + j)*ldk + h*64 ; Tk <= 16, not causal.
// ------------------------------------------------------------------------------------------------
constexpr int kFewKeysMax = 16;
__global__ void __launch_bounds__(256)
attention_fewkeys_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k, const float* __restrict__ v,
                         int ldk, const unsigned char* __restrict__ mask, int Tq, int Tk, float scale,
                         __nv_bfloat16* __restrict__ out_split, long long out_plane, float* __restrict__ out_f32, int ldo) {
  constexpr int DH = 64, PITCH = DH + 4;  // 68 words: float4 accesses of 8 consecutive rows cover all 32 banks
  extern __shared__ __align__(16) float stage[];  // [H][32][PITCH]
  const int b = blockIdx.y, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t0 = blockIdx.x * 32, t = t0 + lane;
  const bool active = t < Tq;
  float qr[DH], o[DH], s[kFewKeysMax];
  if (active) {
    const float* qp = q + (static_cast<long long>(b) * Tq + t) * ldq + h * DH;
#pragma unroll
    for (int i = 0; i < DH; i += 4) {
      const float4 f = *reinterpret_cast<const float4*>(qp + i);
      qr[i] = f.x * scale; qr[i + 1] = f.y * scale; qr[i + 2] = f.z * scale; qr[i + 3] = f.w * scale;
    }
  } else {
#pragma unroll
    for (int i = 0; i < DH; ++i) qr[i] = 0.f;
  }
  const float* kb = k + static_cast<long long>(b) * Tk * ldk + h * DH;
  const float* vb = v + static_cast<long long>(b) * Tk * ldk + h * DH;
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < kFewKeysMax; ++j) {
    float acc = 0.f;
    if (j < Tk) {
      const float* kp = kb + static_cast<long long>(j) * ldk;  // same address in every lane: one broadcast transaction
#pragma unroll
      for (int i = 0; i < DH; i += 4) {
        const float4 f = __ldg(reinterpret_cast<const float4*>(kp + i));
        acc = fmaf(qr[i], f.x, acc); acc = fmaf(qr[i + 1], f.y, acc);
        acc = fmaf(qr[i + 2], f.z, acc); acc = fmaf(qr[i + 3], f.w, acc);
      }
      if (mask && mask[static_cast<long long>(b) * Tk + j]) acc = -INFINITY;  // key is padding
      mx = fmaxf(mx, acc);
    }
    s[j] = acc;
  }
  float l = 0.f;
#pragma unroll
  for (int i = 0; i < DH; ++i) o[i] = 0.f;
#pragma unroll
  for (int j = 0; j < kFewKeysMax; ++j) {
    if (j < Tk) {
      const float p = __expf(s[j] - mx);
      l += p;
      const float* vp = vb + static_cast<long long>(j) * ldk;
#pragma unroll
      for (int i = 0; i < DH; i += 4) {
        const float4 f = __ldg(reinterpret_cast<const float4*>(vp + i));
        o[i] = fmaf(p, f.x, o[i]); o[i + 1] = fmaf(p, f.y, o[i + 1]);
        o[i + 2] = fmaf(p, f.z, o[i + 2]); o[i + 3] = fmaf(p, f.w, o[i + 3]);
      }
    }
  }
  const float inv = 1.f / l;
  float* mine = stage + (h * 32 + lane) * PITCH;
#pragma unroll
  for (int i = 0; i < DH; i += 4)
    *reinterpret_cast<float4*>(mine + i) = make_float4(o[i] * inv, o[i + 1] * inv, o[i + 2] * inv, o[i + 3] * inv);
  __syncwarp();
  // warp h stores its 32 x 64 tile: half a warp per row -> 256 B (fp32) / 128 B (each bf16 plane) contiguous segments
  const int c4 = (lane & 15) * 4;
  for (int r = lane >> 4; r < 32; r += 2) {
    if (t0 + r >= Tq) break;
    const float4 f = *reinterpret_cast<const float4*>(stage + (h * 32 + r) * PITCH + c4);
    const long long off = (static_cast<long long>(b) * Tq + t0 + r) * ldo + h * DH + c4;
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + off) = f;
    if (out_split) {
      __align__(8) __nv_bfloat16 hi[4];
      __align__(8) __nv_bfloat16 lo[4];
      split_bf16(f.x, hi[0], lo[0]); split_bf16(f.y, hi[1], lo[1]);
      split_bf16(f.z, hi[2], lo[2]); split_bf16(f.w, hi[3], lo[3]);
      *reinterpret_cast<uint2*>(out_split + off) = *reinterpret_cast<const uint2*>(hi);
      *reinterpret_cast<uint2*>(out_split + out_plane + off) = *reinterpret_cast<const uint2*>(lo);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Few-keys attention, second version: K/V of a CTA's groups staged ONCE in shared memory.
// The kernel above reads every K/V element with a dependent, warp-broadcast __ldg per float4 at ~170 registers per
// thread (one CTA per SM): 157 us per FIDNetV3 layer and 184 us for the fusion Attention at 128 canvases against HBM
// floors of 11 / 16 us (profiles/r2_launches_f_chain_on_summary.md) -- latency bound.  Here a CTA owns G "groups" (a group
// = one K/V set: a layout sequence of FIDNetV3, or a canvas for the fusion Attention) x a tile of QT queries per group:
//   1. K and V rows of the G groups -> shared memory [g][h][j][64 + 4] with coalesced float4 loads (all 256 threads);
//   2. thread = (g, h, query): the same arithmetic, in the same order, as attention_fewkeys_kernel (bit-identical
//      results), operands from shared memory (lanes of a (g, h) share the address: broadcast);
//   3. the output tile is staged in shared memory and stored as whole rows.
// QP = lanes reserved per (g, h) (16 for the short sequences, 32 for image tokens); NT = query tiles per CTA (the fusion
// Attention reuses a canvas's K/V for 4 tiles of 32 queries).
// ------------------------------------------------------------------------------------------------
template <int H, int QP, int NTH>
__global__ void __launch_bounds__(NTH, 256 / NTH)
attention_kvsmem_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k, const float* __restrict__ v,
                        int ldk, const unsigned char* __restrict__ mask, int B, int Tq, int Tk, float scale,
                        __nv_bfloat16* __restrict__ out_split, long long out_plane, float* __restrict__ out_f32, int ldo,
                        int NT) {
  constexpr int DH = 64, PITCH = DH + 4;
  constexpr int G = NTH / (H * QP);        // groups per CTA
  constexpr int OP = H * DH + 4;           // staged output row pitch
  extern __shared__ __align__(16) float sm[];
  float* ks = sm;                                     // [G][H][Tk][PITCH]
  float* vs = ks + G * H * kFewKeysMax * PITCH;
  float* stage = vs + G * H * kFewKeysMax * PITCH;    // [G][QP][OP]
  unsigned char* ms = reinterpret_cast<unsigned char*>(stage + G * QP * OP);  // [G][16]
  const int tid = threadIdx.x;
  const int g0 = blockIdx.y * G;
  // ---- 1. stage K / V (and the key-padding mask) of the CTA's groups
  constexpr int row_f4 = H * DH / 4;  // float4 per K (or V) row
  {
    // thread = (row within a pass, float4 of the row); up to 8 rows of K and V are REQUESTED before the first is stored (the
    // first version issued one dependent load pair per iteration: 13 global round trips per CTA, which was its run time)
    constexpr int RPP = NTH / row_f4;  // rows per pass
    constexpr int UNR = 8;
    const int c4 = tid % row_f4, r0 = tid / row_f4;
    const int h = (c4 * 4) / DH, d = (c4 * 4) % DH;
    const int R = G * Tk;
#pragma unroll 1
    for (int rb = 0; rb < R; rb += RPP * UNR) {
      float4 fk[UNR], fv[UNR];
      int so[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int r = rb + u * RPP + r0;
        int g = 0;  // r / Tk without a division: G <= 4
#pragma unroll
        for (int gg = 1; gg < G; ++gg) g += (r >= gg * Tk) ? 1 : 0;
        const int j = r - g * Tk;
        so[u] = -1;
        if (r < R && g0 + g < B) {
          const long long off = (static_cast<long long>(g0 + g) * Tk + j) * ldk + c4 * 4;
          fk[u] = *reinterpret_cast<const float4*>(k + off);
          fv[u] = *reinterpret_cast<const float4*>(v + off);
          so[u] = ((g * H + h) * kFewKeysMax + j) * PITCH + d;
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        if (so[u] >= 0) {
          *reinterpret_cast<float4*>(ks + so[u]) = fk[u];
          *reinterpret_cast<float4*>(vs + so[u]) = fv[u];
        }
      }
    }
  }
  if (tid < G * kFewKeysMax) {
    const int g = tid / kFewKeysMax, j = tid % kFewKeysMax;
    ms[tid] = (mask && g0 + g < B && j < Tk) ? mask[static_cast<long long>(g0 + g) * Tk + j] : 0;
  }
  __syncthreads();
  const int tl = tid % QP, gh = tid / QP;
  const int h = gh % H, g = gh / H;
  const bool group_ok = g0 + g < B;
  for (int nt = 0; nt < NT; ++nt) {
    const int t0 = (blockIdx.x * NT + nt) * QP;
    if (t0 >= Tq) break;  // CTA-uniform
    const int t = t0 + tl;
    const bool active = group_ok && t < Tq;
    float qr[DH], o[DH], sc[kFewKeysMax];
    if (active) {
      const float* qp = q + (static_cast<long long>(g0 + g) * Tq + t) * ldq + h * DH;
#pragma unroll
      for (int i = 0; i < DH; i += 4) {
        const float4 f = *reinterpret_cast<const float4*>(qp + i);
        qr[i] = f.x * scale; qr[i + 1] = f.y * scale; qr[i + 2] = f.z * scale; qr[i + 3] = f.w * scale;
      }
    } else {
#pragma unroll
      for (int i = 0; i < DH; ++i) qr[i] = 0.f;
    }
    const float* kb = ks + (g * H + h) * kFewKeysMax * PITCH;
    const float* vb = vs + (g * H + h) * kFewKeysMax * PITCH;
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < kFewKeysMax; ++j) {
      float acc = 0.f;
      if (j < Tk) {
#pragma unroll
        for (int i = 0; i < DH; i += 4) {
          const float4 f = *reinterpret_cast<const float4*>(kb + j * PITCH + i);
          acc = fmaf(qr[i], f.x, acc); acc = fmaf(qr[i + 1], f.y, acc);
          acc = fmaf(qr[i + 2], f.z, acc); acc = fmaf(qr[i + 3], f.w, acc);
        }
        if (ms[g * kFewKeysMax + j]) acc = -INFINITY;
        mx = fmaxf(mx, acc);
      }
      sc[j] = acc;
    }
    float l = 0.f;
#pragma unroll
    for (int i = 0; i < DH; ++i) o[i] = 0.f;
#pragma unroll
    for (int j = 0; j < kFewKeysMax; ++j) {
      if (j < Tk) {
        const float p = __expf(sc[j] - mx);
        l += p;
#pragma unroll
        for (int i = 0; i < DH; i += 4) {
          const float4 f = *reinterpret_cast<const float4*>(vb + j * PITCH + i);
          o[i] = fmaf(p, f.x, o[i]); o[i + 1] = fmaf(p, f.y, o[i + 1]);
          o[i + 2] = fmaf(p, f.z, o[i + 2]); o[i + 3] = fmaf(p, f.w, o[i + 3]);
        }
      }
    }
    const float inv = 1.f / l;
    float* mine = stage + (g * QP + tl) * OP + h * DH;
#pragma unroll
    for (int i = 0; i < DH; i += 4)
      *reinterpret_cast<float4*>(mine + i) = make_float4(o[i] * inv, o[i + 1] * inv, o[i + 2] * inv, o[i + 3] * inv);
    __syncthreads();
    // ---- 3. whole output rows: row (g, tl) = H*64 contiguous floats
    for (int i = tid; i < G * QP * row_f4; i += NTH) {
      const int c4 = i % row_f4, r = i / row_f4;
      const int rl = r % QP, rg = r / QP;
      if (g0 + rg >= B || t0 + rl >= Tq) continue;
      const float4 f = *reinterpret_cast<const float4*>(stage + (rg * QP + rl) * OP + c4 * 4);
      const long long off = (static_cast<long long>(g0 + rg) * Tq + t0 + rl) * ldo + c4 * 4;
      if (out_f32) *reinterpret_cast<float4*>(out_f32 + off) = f;
      if (out_split) {
        __align__(8) __nv_bfloat16 hi[4];
        __align__(8) __nv_bfloat16 lo[4];
        split_bf16(f.x, hi[0], lo[0]); split_bf16(f.y, hi[1], lo[1]);
        split_bf16(f.z, hi[2], lo[2]); split_bf16(f.w, hi[3], lo[3]);
        *reinterpret_cast<uint2*>(out_split + off) = *reinterpret_cast<const uint2*>(hi);
        *reinterpret_cast<uint2*>(out_split + out_plane + off) = *reinterpret_cast<const uint2*>(lo);
      }
    }
    __syncthreads();  // the stage is rewritten by the next query tile
  }
}
template <int H, int QP, int NTH>
static int launch_kvsmem_n(const float* q, int ldq, const float* k, const float* v, int ldk, const unsigned char* mask, int B,
                           int Tq, int Tk, float scale, void* out_split, long long out_plane, float* out_f32, int ldo,
                           cudaStream_t st) {
  constexpr int G = NTH / (H * QP);
  constexpr int SMEM = (2 * G * H * kFewKeysMax * 68 + G * QP * (H * 64 + 4)) * 4 + G * kFewKeysMax;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attention_kvsmem_kernel<H, QP, NTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attention_kvsmem_kernel<H, QP, NTH>, cudaFuncAttributePreferredSharedMemoryCarveout,
                               cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return set_cuda_error(e);
    attr_set = true;
  }
  const int tiles = (Tq + QP - 1) / QP;
  const int NT = tiles >= 4 ? 4 : tiles;  // query tiles per CTA (K/V staged once for all of them)
  dim3 grid((tiles + NT - 1) / NT, (B + G - 1) / G);
  attention_kvsmem_kernel<H, QP, NTH><<<grid, NTH, SMEM, st>>>(q, ldq, k, v, ldk, mask, B, Tq, Tk, scale,
                                                              reinterpret_cast<__nv_bfloat16*>(out_split), out_plane, out_f32,
                                                              ldo, NT);
  return set_cuda_error(cudaGetLastError());
}
// RALF_KVSMEM_THREADS: 128 (default where two groups fit a CTA: FIDNetV3's 4 heads x 16 query lanes -- 70 KB of shared memory,
// three CTAs per SM overlap each other's load -> compute -> store phases) or 256 (round-2 first version, one CTA per SM).
// Same arithmetic per thread either way.
template <int H, int QP>
static int launch_kvsmem(const float* q, int ldq, const float* k, const float* v, int ldk, const unsigned char* mask, int B,
                         int Tq, int Tk, float scale, void* out_split, long long out_plane, float* out_f32, int ldo,
                         cudaStream_t st) {
  static const int nth = getenv("RALF_KVSMEM_THREADS") ? atoi(getenv("RALF_KVSMEM_THREADS")) : 128;
  if constexpr (H * QP <= 128) {
    if (nth == 128)
      return launch_kvsmem_n<H, QP, 128>(q, ldq, k, v, ldk, mask, B, Tq, Tk, scale, out_split, out_plane, out_f32, ldo, st);
  }
  return launch_kvsmem_n<H, QP, 256>(q, ldq, k, v, ldk, mask, B, Tq, Tk, scale, out_split, out_plane, out_f32, ldo, st);
}

// ------------------------------------------------------------------------------------------------
// Single-query attention against a K/V cache (greedy decode).  One warp per (batch, head).
//   q: [B, ldq] fp32 (+ h*DH);  K/V row j of batch b: base + (b*kv_bstride + j)*ldk + h*DH
//   mask: uint8 [B, mask_ld] or null; Tk keys.  Output: split bf16 [2, B, H*DH].
// Replaces the full-prefix recompute of retrieval_augmented_autoreg.py:271-297.
// ------------------------------------------------------------------------------------------------
template <int DH>
__global__ void __launch_bounds__(128)
attention_decode_kernel(const float* __restrict__ q, int ldq, const float* k /* may alias kc_w */,
                        const float* v /* may alias vc_w */, long long kv_bstride, int ldk,
                        const unsigned char* __restrict__ mask, int mask_ld, int Tk, int B, int H, float scale,
                        __nv_bfloat16* __restrict__ out_split, long long out_plane, int ldo,
                        const float* __restrict__ knew, const float* __restrict__ vnew, int ldnew,
                        float* kc_w, float* vc_w) {
  // One CTA (4 warps) per (batch, head); each warp owns a contiguous quarter of the keys (flash-decoding split),
  // partial (max, sum, o) are combined through shared memory.  Lane mapping inside a warp: CPL = DH/4 lanes cover
  // one key row (16 B each), KPI = 32/CPL keys per iteration, 8 iterations in flight -> every warp-wide LDG.128
  // reads KPI full contiguous rows and ~4 KB per warp are outstanding (the cross-attention K/V stream is HBM bound).
  constexpr int CPL = DH / 4;
  constexpr int KPI = 32 / CPL;
  constexpr int UN = 8;
  extern __shared__ float probs[];  // [Tk_pad]
  __shared__ float red_m[4], red_l[4];
  __shared__ float red_o[4][DH];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bh = blockIdx.x;
  const int b = bh / H, h = bh - b * H;
  const int kk = lane / CPL, c = lane % CPL;
  if (knew) {  // append this step's K/V row (position Tk-1) to the cache before attending over it
    if (threadIdx.x < DH) {
      const long long dst = (static_cast<long long>(b) * kv_bstride + (Tk - 1)) * ldk + h * DH + threadIdx.x;
      kc_w[dst] = knew[static_cast<long long>(b) * ldnew + h * DH + threadIdx.x];
      vc_w[dst] = vnew[static_cast<long long>(b) * ldnew + h * DH + threadIdx.x];
    }
    __syncthreads();
  }
  const int seg = (((Tk + 3) / 4) + KPI - 1) / KPI * KPI;  // keys per warp, multiple of KPI
  const int j_lo = wid * seg, j_hi = min(Tk, j_lo + seg);
  float4 q4 = *reinterpret_cast<const float4*>(q + static_cast<long long>(b) * ldq + h * DH + 4 * c);
  q4.x *= scale; q4.y *= scale; q4.z *= scale; q4.w *= scale;
  const float* kb = k + static_cast<long long>(b) * kv_bstride * ldk + h * DH + 4 * c;
  const float* vb = v + static_cast<long long>(b) * kv_bstride * ldk + h * DH + 4 * c;
  float lmax = -INFINITY;
  for (int j0 = j_lo; j0 < j_hi; j0 += KPI * UN) {
    float part[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int j = j0 + u * KPI + kk;
      float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j < j_hi) f = *reinterpret_cast<const float4*>(kb + static_cast<long long>(j) * ldk);
      part[u] = q4.x * f.x + q4.y * f.y + q4.z * f.z + q4.w * f.w;
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
#pragma unroll
      for (int off = CPL >> 1; off >= 1; off >>= 1) part[u] += __shfl_xor_sync(0xffffffffu, part[u], off);
      const int j = j0 + u * KPI + kk;
      if (j < j_hi) {
        const float sc = (mask && mask[static_cast<long long>(b) * mask_ld + j]) ? -INFINITY : part[u];
        if (c == 0) probs[j] = sc;
        lmax = fmaxf(lmax, sc);
      }
    }
  }
  lmax = warp_max(lmax);
  if (lane == 0) red_m[wid] = lmax;
  __syncthreads();
  const float gmax = fmaxf(fmaxf(red_m[0], red_m[1]), fmaxf(red_m[2], red_m[3]));
  // o[4c..4c+3] over this lane's keys, exponentials computed once per key (lanes of a key share it)
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
  float lsum = 0.f;
  for (int j0 = j_lo; j0 < j_hi; j0 += KPI * UN) {
    float4 f[UN];
    float pj[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int j = j0 + u * KPI + kk;
      f[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      pj[u] = 0.f;
      if (j < j_hi) {
        f[u] = *reinterpret_cast<const float4*>(vb + static_cast<long long>(j) * ldk);
        pj[u] = __expf(probs[j] - gmax);
      }
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      o.x = fmaf(pj[u], f[u].x, o.x); o.y = fmaf(pj[u], f[u].y, o.y);
      o.z = fmaf(pj[u], f[u].z, o.z); o.w = fmaf(pj[u], f[u].w, o.w);
      if (c == 0) lsum += pj[u];
    }
  }
#pragma unroll
  for (int off = CPL; off < 32; off <<= 1) {
    o.x += __shfl_xor_sync(0xffffffffu, o.x, off); o.y += __shfl_xor_sync(0xffffffffu, o.y, off);
    o.z += __shfl_xor_sync(0xffffffffu, o.z, off); o.w += __shfl_xor_sync(0xffffffffu, o.w, off);
  }
  lsum = warp_sum(lsum);
  if (kk == 0) {
    red_o[wid][4 * c + 0] = o.x; red_o[wid][4 * c + 1] = o.y;
    red_o[wid][4 * c + 2] = o.z; red_o[wid][4 * c + 3] = o.w;
  }
  if (lane == 0) red_l[wid] = lsum;
  __syncthreads();
  if (threadIdx.x < DH) {
    const float inv = 1.f / (red_l[0] + red_l[1] + red_l[2] + red_l[3]);
    const float y = (red_o[0][threadIdx.x] + red_o[1][threadIdx.x] + red_o[2][threadIdx.x] + red_o[3][threadIdx.x]) * inv;
    store_split(out_split, out_plane, static_cast<long long>(b) * ldo + h * DH + threadIdx.x, y);
  }
}

// ------------------------------------------------------------------------------------------------
// Cross-attention decode step over the memory K/V cache, single pass ("flash decoding" per head): one CTA per
// canvas, one warp per head.  A warp walks its head's keys 32 at a time: the K AND V rows of the batch are requested
// together (16 x LDG.128 per lane = 8 KB per warp in flight), scores, online-softmax rescale and the P.V update follow
// from registers -- half the dependent HBM round trips of the two-pass kernel above, and the eight warps of a CTA
// consume the canvas's K/V rows (layer-major cache: 2 KB contiguous per memory token) as one sequential stream.
// No key-padding mask (BaseDecoder passes none for the memory, common/common.py:123-131).
// ------------------------------------------------------------------------------------------------
template <int DH>
__global__ void __launch_bounds__(256)
attention_decode_stream_kernel(const float* __restrict__ q, int ldq, const float* k, const float* v,
                               long long kv_bstride, int ldk, int Tk, int H, float scale,
                               __nv_bfloat16* __restrict__ out_split, long long out_plane, int ldo,
                               const unsigned char* __restrict__ mask, int mask_ld, const float* __restrict__ knew,
                               const float* __restrict__ vnew, int ldnew, float* kc_w, float* vc_w) {
  constexpr int CPL = DH / 4;    // lanes per key row (16 B each)
  constexpr int KPI = 32 / CPL;  // keys per warp-wide load
  constexpr int UN = 32 / KPI;   // loads per batch of 32 keys
  pdl_trigger();
  pdl_wait();
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x;
  if (h >= H) return;
  const int kk = lane / CPL, c = lane % CPL;
  float4 q4 = *reinterpret_cast<const float4*>(q + static_cast<long long>(b) * ldq + h * DH + 4 * c);
  q4.x *= scale; q4.y *= scale; q4.z *= scale; q4.w *= scale;
  const float* kb = k + static_cast<long long>(b) * kv_bstride * ldk + h * DH + 4 * c;
  const float* vb = v + static_cast<long long>(b) * kv_bstride * ldk + h * DH + 4 * c;
  if (knew) {  // self-attention step: append this step's K/V row (position Tk-1) of this head, then attend over it
    if (kk == 0) {
      const long long dst = (static_cast<long long>(b) * kv_bstride + (Tk - 1)) * ldk + h * DH + 4 * c;
      const long long src = static_cast<long long>(b) * ldnew + h * DH + 4 * c;
      *reinterpret_cast<float4*>(kc_w + dst) = *reinterpret_cast<const float4*>(knew + src);
      *reinterpret_cast<float4*>(vc_w + dst) = *reinterpret_cast<const float4*>(vnew + src);
    }
    __syncwarp();
  }
  const unsigned char* mrow = mask ? mask + static_cast<long long>(b) * mask_ld : nullptr;
  float m = -INFINITY, l = 0.f;
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
  for (int j0 = 0; j0 < Tk; j0 += 32) {
    float4 kf[UN], vf[UN];
    bool live[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int j = j0 + u * KPI + kk;
      kf[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      vf[u] = kf[u];
      live[u] = j < Tk;
      if (j < Tk) {
        kf[u] = __ldcs(reinterpret_cast<const float4*>(kb + static_cast<long long>(j) * ldk));
        vf[u] = __ldcs(reinterpret_cast<const float4*>(vb + static_cast<long long>(j) * ldk));
        if (mrow && mrow[j]) live[u] = false;  // padded key
      }
    }
    float s[UN];
    float bm = -INFINITY;
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      float d = q4.x * kf[u].x + q4.y * kf[u].y + q4.z * kf[u].z + q4.w * kf[u].w;
#pragma unroll
      for (int off = CPL >> 1; off >= 1; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
      s[u] = live[u] ? d : -INFINITY;
      bm = fmaxf(bm, s[u]);
    }
#pragma unroll
    for (int off = CPL; off < 32; off <<= 1) bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, off));
    if (bm == -INFINITY) continue;     // a batch of masked keys only (warp-uniform)
    const float m_new = fmaxf(m, bm);
    const float corr = __expf(m - m_new);
    l *= corr;
    o.x *= corr; o.y *= corr; o.z *= corr; o.w *= corr;
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const float p = __expf(s[u] - m_new);  // -inf -> 0
      l += p;
      o.x = fmaf(p, vf[u].x, o.x); o.y = fmaf(p, vf[u].y, o.y);
      o.z = fmaf(p, vf[u].z, o.z); o.w = fmaf(p, vf[u].w, o.w);
    }
    m = m_new;
  }
#pragma unroll
  for (int off = CPL; off < 32; off <<= 1) {
    o.x += __shfl_xor_sync(0xffffffffu, o.x, off); o.y += __shfl_xor_sync(0xffffffffu, o.y, off);
    o.z += __shfl_xor_sync(0xffffffffu, o.z, off); o.w += __shfl_xor_sync(0xffffffffu, o.w, off);
    l += __shfl_xor_sync(0xffffffffu, l, off);
  }
  if (kk == 0) {
    const float inv = 1.f / l;
    const long long off0 = static_cast<long long>(b) * ldo + h * DH + 4 * c;
    store_split(out_split, out_plane, off0 + 0, o.x * inv);
    store_split(out_split, out_plane, off0 + 1, o.y * inv);
    store_split(out_split, out_plane, off0 + 2, o.z * inv);
    store_split(out_split, out_plane, off0 + 3, o.w * inv);
  }
}

// The same single-pass cross-attention over the 24-bit K/V cache (ralf_gemm out_kv24): a cache row is 1536 bytes
// [K hi 256 x u16 | V hi 256 x u16 | K lo 256 x u8 | V lo 256 x u8]; a value is (hi << 16 | lo << 8) as fp32 bits.
// 25 % fewer bytes on the stream that bounds the decode loop.  8 lanes cover one key of one head (4 channels each:
// 8 B of hi + 4 B of lo per operand), 4 keys per warp-wide load, 32 keys per batch.
__device__ __forceinline__ float4 kv24_unpack(const uint2 hi, const uint32_t lo) {
  float4 f;
  f.x = __uint_as_float((hi.x << 16) | ((lo & 0xffu) << 8));
  f.y = __uint_as_float((hi.x & 0xffff0000u) | (lo & 0xff00u));
  f.z = __uint_as_float((hi.y << 16) | ((lo >> 8) & 0xff00u));
  f.w = __uint_as_float((hi.y & 0xffff0000u) | ((lo >> 16) & 0xff00u));
  return f;
}

__global__ void __launch_bounds__(256)
attention_decode_kv24_kernel(const float* __restrict__ q, int ldq, const uint8_t* __restrict__ kv, long long kv_bstride,
                             int Tk, int H, float scale, __nv_bfloat16* __restrict__ out_split, long long out_plane,
                             int ldo) {
  constexpr int DH = 32, CPL = 8, KPI = 4, UN = 8, ROW = 1536;
  pdl_trigger();
  pdl_wait();
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x;
  if (h >= H) return;
  const int kk = lane / CPL, c = lane % CPL;
  float4 q4 = *reinterpret_cast<const float4*>(q + static_cast<long long>(b) * ldq + h * DH + 4 * c);
  q4.x *= scale; q4.y *= scale; q4.z *= scale; q4.w *= scale;
  const uint8_t* base = kv + static_cast<long long>(b) * kv_bstride * ROW;
  const int o_khi = h * 64 + c * 8, o_vhi = 512 + h * 64 + c * 8, o_klo = 1024 + h * 32 + c * 4,
            o_vlo = 1280 + h * 32 + c * 4;
  float m = -INFINITY, l = 0.f;
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
  for (int j0 = 0; j0 < Tk; j0 += 32) {
    uint2 kh[UN], vh[UN];
    uint32_t kl[UN], vl[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int j = j0 + u * KPI + kk;
      kh[u] = vh[u] = make_uint2(0u, 0u);
      kl[u] = vl[u] = 0u;
      if (j < Tk) {
        const uint8_t* r = base + static_cast<long long>(j) * ROW;
        kh[u] = __ldcs(reinterpret_cast<const uint2*>(r + o_khi));
        vh[u] = __ldcs(reinterpret_cast<const uint2*>(r + o_vhi));
        kl[u] = __ldcs(reinterpret_cast<const uint32_t*>(r + o_klo));
        vl[u] = __ldcs(reinterpret_cast<const uint32_t*>(r + o_vlo));
      }
    }
    float s[UN];
    float bm = -INFINITY;
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const float4 kf = kv24_unpack(kh[u], kl[u]);
      float d = q4.x * kf.x + q4.y * kf.y + q4.z * kf.z + q4.w * kf.w;
#pragma unroll
      for (int off = CPL >> 1; off >= 1; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
      s[u] = (j0 + u * KPI + kk < Tk) ? d : -INFINITY;
      bm = fmaxf(bm, s[u]);
    }
#pragma unroll
    for (int off = CPL; off < 32; off <<= 1) bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, off));
    const float m_new = fmaxf(m, bm);
    const float corr = __expf(m - m_new);
    l *= corr;
    o.x *= corr; o.y *= corr; o.z *= corr; o.w *= corr;
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const float p = __expf(s[u] - m_new);
      const float4 vf = kv24_unpack(vh[u], vl[u]);
      l += p;
      o.x = fmaf(p, vf.x, o.x); o.y = fmaf(p, vf.y, o.y);
      o.z = fmaf(p, vf.z, o.z); o.w = fmaf(p, vf.w, o.w);
    }
    m = m_new;
  }
#pragma unroll
  for (int off = CPL; off < 32; off <<= 1) {
    o.x += __shfl_xor_sync(0xffffffffu, o.x, off); o.y += __shfl_xor_sync(0xffffffffu, o.y, off);
    o.z += __shfl_xor_sync(0xffffffffu, o.z, off); o.w += __shfl_xor_sync(0xffffffffu, o.w, off);
    l += __shfl_xor_sync(0xffffffffu, l, off);
  }
  if (kk == 0) {
    const float inv = 1.f / l;
    const long long off0 = static_cast<long long>(b) * ldo + h * DH + 4 * c;
    store_split(out_split, out_plane, off0 + 0, o.x * inv);
    store_split(out_split, out_plane, off0 + 1, o.y * inv);
    store_split(out_split, out_plane, off0 + 2, o.z * inv);
    store_split(out_split, out_plane, off0 + 3, o.w * inv);
  }
}

// The same single-pass cross-attention over the 16-bit per-head-scaled K/V cache (ralf_gemm out_kv24, out_kv_fmt = 16):
// a cache row is 1088 bytes [K 256 x u16 | V 256 x u16 | 8 x (K scale, V scale) f32 pairs], a value is
// (u16 - 32768) * scale[head] (the two scales of a head sit together: one 8-byte load per key and lane).  29 % fewer bytes than the 24-bit rows on the stream that bounds the decode loop.
// u16 -> float without a conversion instruction: 0x4B000000 | u16 is the float 2^23 + u16, minus (2^23 + 32768).
__device__ __forceinline__ float4 kv16_unpack(const uint2 w) {
  constexpr float BIAS = 8388608.f + 32768.f;
  float4 f;
  f.x = __uint_as_float(0x4B000000u | (w.x & 0xffffu)) - BIAS;
  f.y = __uint_as_float(__byte_perm(w.x, 0x4B000000u, 0x7632)) - BIAS;
  f.z = __uint_as_float(0x4B000000u | (w.y & 0xffffu)) - BIAS;
  f.w = __uint_as_float(__byte_perm(w.y, 0x4B000000u, 0x7632)) - BIAS;
  return f;
}

__global__ void __launch_bounds__(256)
attention_decode_kv16_kernel(const float* __restrict__ q, int ldq, const uint8_t* __restrict__ kv, long long kv_bstride,
                             int Tk, int H, float scale, __nv_bfloat16* __restrict__ out_split, long long out_plane,
                             int ldo) {
  constexpr int DH = 32, CPL = 8, KPI = 4, UN = 8, ROW = 1088;
  pdl_trigger();
  pdl_wait();
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x;
  if (h >= H) return;
  const int kk = lane / CPL, c = lane % CPL;
  float4 q4 = *reinterpret_cast<const float4*>(q + static_cast<long long>(b) * ldq + h * DH + 4 * c);
  q4.x *= scale; q4.y *= scale; q4.z *= scale; q4.w *= scale;
  const uint8_t* base = kv + static_cast<long long>(b) * kv_bstride * ROW;
  const int o_k = h * 64 + c * 8, o_v = 512 + h * 64 + c * 8, o_s = 1024 + h * 8;
  float m = -INFINITY, l = 0.f;
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
  for (int j0 = 0; j0 < Tk; j0 += 32) {
    uint2 kw[UN], vw[UN];
    float ksc[UN], vsc[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int j = j0 + u * KPI + kk;
      kw[u] = vw[u] = make_uint2(0x80008000u, 0x80008000u);  // offset-binary zero
      ksc[u] = vsc[u] = 0.f;
      if (j < Tk) {
        const uint8_t* r = base + static_cast<long long>(j) * ROW;
        kw[u] = __ldcs(reinterpret_cast<const uint2*>(r + o_k));
        vw[u] = __ldcs(reinterpret_cast<const uint2*>(r + o_v));
        const float2 sc2 = __ldcs(reinterpret_cast<const float2*>(r + o_s));
        ksc[u] = sc2.x;
        vsc[u] = sc2.y;
      }
    }
    float s[UN];
    float bm = -INFINITY;
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const float4 kf = kv16_unpack(kw[u]);
      float d = q4.x * kf.x + q4.y * kf.y + q4.z * kf.z + q4.w * kf.w;
#pragma unroll
      for (int off = CPL >> 1; off >= 1; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
      s[u] = (j0 + u * KPI + kk < Tk) ? d * ksc[u] : -INFINITY;
      bm = fmaxf(bm, s[u]);
    }
#pragma unroll
    for (int off = CPL; off < 32; off <<= 1) bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, off));
    const float m_new = fmaxf(m, bm);
    const float corr = __expf(m - m_new);
    l *= corr;
    o.x *= corr; o.y *= corr; o.z *= corr; o.w *= corr;
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const float p = __expf(s[u] - m_new);
      const float4 vf = kv16_unpack(vw[u]);
      const float pv = p * vsc[u];
      l += p;
      o.x = fmaf(pv, vf.x, o.x); o.y = fmaf(pv, vf.y, o.y);
      o.z = fmaf(pv, vf.z, o.z); o.w = fmaf(pv, vf.w, o.w);
    }
    m = m_new;
  }
#pragma unroll
  for (int off = CPL; off < 32; off <<= 1) {
    o.x += __shfl_xor_sync(0xffffffffu, o.x, off); o.y += __shfl_xor_sync(0xffffffffu, o.y, off);
    o.z += __shfl_xor_sync(0xffffffffu, o.z, off); o.w += __shfl_xor_sync(0xffffffffu, o.w, off);
    l += __shfl_xor_sync(0xffffffffu, l, off);
  }
  if (kk == 0) {
    const float inv = 1.f / l;
    const long long off0 = static_cast<long long>(b) * ldo + h * DH + 4 * c;
    store_split(out_split, out_plane, off0 + 0, o.x * inv);
    store_split(out_split, out_plane, off0 + 1, o.y * inv);
    store_split(out_split, out_plane, off0 + 2, o.z * inv);
    store_split(out_split, out_plane, off0 + 3, o.w * inv);
  }
}

// ------------------------------------------------------------------------------------------------
// ResNet stem im2col: image fp32 NCHW [B, 4, H, W] -> split rows [B*Ho*Wo, KP] with k = (kh*7+kw)*4+c
// for the 7x7 / stride 2 / pad 3 convolution (common/image.py:69-77), zero padded to KP columns.
// ------------------------------------------------------------------------------------------------
__global__ void stem_im2col_kernel(const float* __restrict__ img, int B, int H, int W, int Ho, int Wo, int KP,
                                   __nv_bfloat16* __restrict__ out, long long plane) {
  const long long total = static_cast<long long>(B) * Ho * Wo * 50;  // 49 taps + 1 pad group of 4
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int tap = static_cast<int>(i % 50);
    const long long pix = i / 50;
    const int ox = static_cast<int>(pix % Wo);
    const int oy = static_cast<int>((pix / Wo) % Ho);
    const int b = static_cast<int>(pix / (static_cast<long long>(Wo) * Ho));
    float vals[4] = {0.f, 0.f, 0.f, 0.f};
    if (tap < 49) {
      const int kh = tap / 7, kw = tap % 7;
      const int iy = oy * 2 - 3 + kh, ix = ox * 2 - 3 + kw;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
#pragma unroll
        for (int c = 0; c < 4; ++c) vals[c] = img[((static_cast<long long>(b) * 4 + c) * H + iy) * W + ix];
      }
    }
    const long long off = pix * KP + tap * 4;
    if (tap * 4 + 4 <= KP) {
      __nv_bfloat16 h[4], l[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) split_bf16(vals[c], h[c], l[c]);
      *reinterpret_cast<uint2*>(out + off) = make_uint2(pack_bf16(h[0], h[1]), pack_bf16(h[2], h[3]));
      *reinterpret_cast<uint2*>(out + plane + off) = make_uint2(pack_bf16(l[0], l[1]), pack_bf16(l[2], l[3]));
    }
  }
}

// ResNet stem, space-to-depth form: the 7x7 / stride 2 / pad 3 convolution on 4 channels equals a 4x4 / stride 1
// convolution on the 2x2-blocked image with 16 channels (s2d channel = (dy*2+dx)*4 + c; taps (kh', dy) = ((kh+1)/2,
// (kh+1)%2), the 8th tap row/column carries zero weights).  In NHWC the four w-taps x 16 channels of one kh' row are 128
// CONTIGUOUS bytes -- exactly one tensor-core k-block -- so the stem runs as an implicit GEMM (ralf_stem_gemm) over this
// zero-bordered buffer: out [2][B, H/2 + 3, W/2 + 3, 16] split bf16, 2 border pixels left / top, 1 right / bottom.
__global__ void stem_s2d_kernel(const float* __restrict__ img, int B, int H, int W, int Hp, int Wp,
                                __nv_bfloat16* __restrict__ out, long long plane) {
  const long long total = static_cast<long long>(B) * Hp * Wp;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int px = static_cast<int>(i % Wp);
    const int py = static_cast<int>((i / Wp) % Hp);
    const int b = static_cast<int>(i / (static_cast<long long>(Wp) * Hp));
    const int x = px - 2, y = py - 2;
    float v[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) v[e] = 0.f;
    if (x >= 0 && 2 * x < W && y >= 0 && 2 * y < H) {
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
          const float2 f = *reinterpret_cast<const float2*>(
              img + ((static_cast<long long>(b) * 4 + c) * H + 2 * y + dy) * W + 2 * x);
          v[(dy * 2 + 0) * 4 + c] = f.x;
          v[(dy * 2 + 1) * 4 + c] = f.y;
        }
    }
    uint32_t hw[8], lw[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      __nv_bfloat16 h0, l0, h1, l1;
      split_bf16(v[2 * e], h0, l0);
      split_bf16(v[2 * e + 1], h1, l1);
      hw[e] = pack_bf16(h0, h1);
      lw[e] = pack_bf16(l0, l1);
    }
    uint4* oh = reinterpret_cast<uint4*>(out + i * 16);
    uint4* ol = reinterpret_cast<uint4*>(out + plane + i * 16);
    oh[0] = make_uint4(hw[0], hw[1], hw[2], hw[3]);
    oh[1] = make_uint4(hw[4], hw[5], hw[6], hw[7]);
    ol[0] = make_uint4(lw[0], lw[1], lw[2], lw[3]);
    ol[1] = make_uint4(lw[4], lw[5], lw[6], lw[7]);
  }
}

// Generic NHWC im2col on split activations: in [2][B,H,W,C] -> out [2][B*Ho*Wo, KH*KW*C]
// (k = (kh*KW + kw)*C + c), 8 channels (16 bytes) per thread.  Used for the 3x3 convolutions and
// the strided 1x1 downsample convolutions of ResNet50 / the FPN 3x3.
__global__ void im2col_kernel(const __nv_bfloat16* __restrict__ in, long long in_plane, int B, int H, int W, int C,
                              int KH, int KW, int stride, int pad, int Ho, int Wo,
                              __nv_bfloat16* __restrict__ out, long long out_plane) {
  const int c8 = C >> 3;
  const long long total = static_cast<long long>(B) * Ho * Wo * KH * KW * c8;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cc = static_cast<int>(i % c8);
    long long r = i / c8;
    const int kw = static_cast<int>(r % KW); r /= KW;
    const int kh = static_cast<int>(r % KH); r /= KH;
    const long long pix = r;
    const int ox = static_cast<int>(pix % Wo);
    const int oy = static_cast<int>((pix / Wo) % Ho);
    const int b = static_cast<int>(pix / (static_cast<long long>(Wo) * Ho));
    const int iy = oy * stride - pad + kh, ix = ox * stride - pad + kw;
    uint4 h = make_uint4(0, 0, 0, 0), l = h;
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
      const long long src = ((static_cast<long long>(b) * H + iy) * W + ix) * C + cc * 8;
      h = *reinterpret_cast<const uint4*>(in + src);
      l = *reinterpret_cast<const uint4*>(in + in_plane + src);
    }
    const long long dst = pix * (static_cast<long long>(KH) * KW * C) + (static_cast<long long>(kh) * KW + kw) * C + cc * 8;
    *reinterpret_cast<uint4*>(out + dst) = h;
    *reinterpret_cast<uint4*>(out + out_plane + dst) = l;
  }
}

// 3x3 / stride 2 / pad 1 max-pool on split NHWC (ResNet stem pool).  2 channels per thread.
__global__ void maxpool_kernel(const __nv_bfloat16* __restrict__ in, long long in_plane, int B, int H, int W, int C,
                               int Ho, int Wo, __nv_bfloat16* __restrict__ out, long long out_plane) {
  // 8 channels (16 bytes per plane) per thread
  const int c8 = C >> 3;
  const long long total = static_cast<long long>(B) * Ho * Wo * c8;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cc = static_cast<int>(i % c8);
    const long long pix = i / c8;
    const int ox = static_cast<int>(pix % Wo);
    const int oy = static_cast<int>((pix / Wo) % Ho);
    const int b = static_cast<int>(pix / (static_cast<long long>(Wo) * Ho));
    float m[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) m[e] = -INFINITY;
#pragma unroll
    for (int kh = 0; kh < 3; ++kh) {
      const int iy = oy * 2 - 1 + kh;
      if (iy < 0 || iy >= H) continue;
#pragma unroll
      for (int kw = 0; kw < 3; ++kw) {
        const int ix = ox * 2 - 1 + kw;
        if (ix < 0 || ix >= W) continue;
        const long long src = ((static_cast<long long>(b) * H + iy) * W + ix) * C + cc * 8;
        const uint4 h4 = *reinterpret_cast<const uint4*>(in + src);
        const uint4 l4 = *reinterpret_cast<const uint4*>(in + in_plane + src);
        const uint32_t hw[4] = {h4.x, h4.y, h4.z, h4.w}, lw[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          m[2 * e] = fmaxf(m[2 * e], __uint_as_float(hw[e] << 16) + __uint_as_float(lw[e] << 16));
          m[2 * e + 1] = fmaxf(m[2 * e + 1], __uint_as_float(hw[e] & 0xffff0000u) + __uint_as_float(lw[e] & 0xffff0000u));
        }
      }
    }
    uint32_t oh[4], ol[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      __nv_bfloat16 h0, l0, h1, l1;
      split_bf16(m[2 * e], h0, l0);
      split_bf16(m[2 * e + 1], h1, l1);
      oh[e] = pack_bf16(h0, h1);
      ol[e] = pack_bf16(l0, l1);
    }
    const long long dst = pix * C + cc * 8;
    *reinterpret_cast<uint4*>(out + dst) = make_uint4(oh[0], oh[1], oh[2], oh[3]);
    *reinterpret_cast<uint4*>(out + out_plane + dst) = make_uint4(ol[0], ol[1], ol[2], ol[3]);
  }
}

// FPN merge (common/image.py:103-111): up = nearest(c5 -> h4 x w4); fused[:, 0:C] = up;
// sum = up + c4.  c5 [B,h5,w5,C] fp32, c4 [B,h4,w4,C] fp32; outputs split, fused row stride ldf.
__global__ void fpn_merge_kernel(const float* __restrict__ c5, const float* __restrict__ c4, int B, int h5, int w5,
                                 int h4, int w4, int C, __nv_bfloat16* __restrict__ fused, long long fused_plane,
                                 int ldf, __nv_bfloat16* __restrict__ sum, long long sum_plane) {
  const long long total = static_cast<long long>(B) * h4 * w4 * C;
  const float sy = static_cast<float>(h5) / static_cast<float>(h4);
  const float sx = static_cast<float>(w5) / static_cast<float>(w4);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const long long pix = i / C;
    const int x = static_cast<int>(pix % w4);
    const int y = static_cast<int>((pix / w4) % h4);
    const int b = static_cast<int>(pix / (static_cast<long long>(w4) * h4));
    // F.interpolate(mode="nearest"): src = min(floor(dst * in/out), in - 1)
    const int yy = min(static_cast<int>(floorf(static_cast<float>(y) * sy)), h5 - 1);
    const int xx = min(static_cast<int>(floorf(static_cast<float>(x) * sx)), w5 - 1);
    const float up = c5[((static_cast<long long>(b) * h5 + yy) * w5 + xx) * C + c];
    store_split(fused, fused_plane, pix * ldf + c, up);
    store_split(sum, sum_plane, pix * C + c, up + c4[pix * C + c]);
  }
}

// Row glue: out[map(r), :] = in[r*in_ld + :] * scale + add + table[(r % tab_mod), :]
// (positional encodings, concatenations, CLS-token fill).  in may be null (zeros).
__global__ void rows_affine_kernel(const float* __restrict__ in, long long in_ld, int M, int D, float scale,
                                   float add, const float* __restrict__ table, int tab_mod, int rpg, int gs, int go,
                                   float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_split,
                                   long long out_plane, int out_ld) {
  const long long total = static_cast<long long>(M) * D;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % D);
    const int r = static_cast<int>(i / D);
    float y = (in ? in[static_cast<long long>(r) * in_ld + c] : 0.f) * scale + add;
    if (table) y += table[static_cast<long long>(r % tab_mod) * D + c];
    const long long orow = static_cast<long long>(r / rpg) * gs + go + r % rpg;
    if (out_f32) out_f32[orow * out_ld + c] = y;
    if (out_split) store_split(out_split, out_plane, orow * out_ld + c, y);
  }
}

// Token embedding (BaseDecoder / UserConstraintTransformerEncoder front end, common/common.py:99-100,
// 243-245): out[r, :] = emb[tok[r], :] * scale + pe[(pos0 + r % S), :]
__global__ void embed_kernel(const long long* __restrict__ tok, long long tok_ld, int tok_col, int Bn, int S,
                             const float* __restrict__ emb, int D, float scale, const float* __restrict__ pe,
                             int pos0, float* __restrict__ out) {
  const long long total = static_cast<long long>(Bn) * S * D;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % D);
    const long long r = i / D;
    const int s = static_cast<int>(r % S);
    const int b = static_cast<int>(r / S);
    const long long t = tok[static_cast<long long>(b) * tok_ld + tok_col + s];
    out[i] = emb[t * D + c] * scale + pe[static_cast<long long>(pos0 + s) * D + c];
  }
}

// FIDNetV3 input (fid/model.py:95-101): row = cat[fc_bbox(cx,cy,w,h), emb_label[label]] -> split [rows, 2*D]
__global__ void fid_embed_kernel(const float* __restrict__ cx, const float* __restrict__ cy, const float* __restrict__ w,
                                 const float* __restrict__ h, const long long* __restrict__ label, int rows, int D,
                                 const float* __restrict__ fc_w /*[D,4]*/, const float* __restrict__ fc_b,
                                 const float* __restrict__ emb /*[L,D]*/, __nv_bfloat16* __restrict__ out,
                                 long long plane) {
  const long long total = static_cast<long long>(rows) * 2 * D;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % (2 * D));
    const long long r = i / (2 * D);
    float y;
    if (c < D) {
      // nn.Linear(4, D) on BBOX_KEYS order (center_x, center_y, width, height), fid/data.py
      float acc = fc_b[c];
      acc = fmaf(cx[r], fc_w[c * 4 + 0], acc);
      acc = fmaf(cy[r], fc_w[c * 4 + 1], acc);
      acc = fmaf(w[r], fc_w[c * 4 + 2], acc);
      acc = fmaf(h[r], fc_w[c * 4 + 3], acc);
      y = acc;
    } else {
      y = emb[label[r] * D + (c - D)];
    }
    store_split(out, plane, i, y);
  }
}


// Exemplar fetch (helpers/retrieval_dataset_wrapper.py:89-148 as an index gather): packed layout table
// [N, 6, E] fp32 (label, mask, center_x, center_y, width, height) -> out [rows, 6, E] for idx [rows]
// (negative / missing index -> empty layout).
__global__ void gather_layouts_kernel(const float* __restrict__ table, const long long* __restrict__ idx, int rows,
                                      int row_elems, long long n_table, long long index_base, float* __restrict__ out) {
  const long long total = static_cast<long long>(rows) * row_elems;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % row_elems);
    const long long r = i / row_elems;
    const long long src = idx[r] - index_base;
    out[i] = (src >= 0 && src < n_table) ? table[src * row_elems + c] : 0.f;
  }
}

// FIDNetV3 input from packed layouts [rows_seq, 6, E]: for element (n, e) emit
// cat[fc_bbox(cx,cy,w,h), emb_label[label]] (split, [rows_seq*E, 2D]) and the key-padding mask row
// [rows_seq, E+1] (column 0 = CLS, never padded; fid/model.py:34-50,88-103).
__global__ void fid_embed_packed_kernel(const float* __restrict__ packed, int nseq, int E, int D,
                                        const float* __restrict__ fc_w, const float* __restrict__ fc_b,
                                        const float* __restrict__ emb, int num_labels, __nv_bfloat16* __restrict__ out,
                                        long long plane, unsigned char* __restrict__ pad) {
  const long long total = static_cast<long long>(nseq) * E * 2 * D;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % (2 * D));
    const long long r = i / (2 * D);  // n * E + e
    const int e = static_cast<int>(r % E);
    const long long n = r / E;
    const float* row = packed + n * 6 * E;
    float y;
    if (c < D) {
      float acc = fc_b[c];
      acc = fmaf(row[2 * E + e], fc_w[c * 4 + 0], acc);
      acc = fmaf(row[3 * E + e], fc_w[c * 4 + 1], acc);
      acc = fmaf(row[4 * E + e], fc_w[c * 4 + 2], acc);
      acc = fmaf(row[5 * E + e], fc_w[c * 4 + 3], acc);
      y = acc;
    } else {
      int lab = static_cast<int>(row[e]);
      lab = lab < 0 ? 0 : (lab >= num_labels ? num_labels - 1 : lab);
      y = emb[static_cast<long long>(lab) * D + (c - D)];
    }
    store_split(out, plane, i, y);
    if (c == 0) {
      pad[n * (E + 1) + 1 + e] = (row[E + e] != 0.f) ? 0 : 1;
      if (e == 0) pad[n * (E + 1)] = 0;
    }
  }
}

// Greedy step tail (retrieval_augmented_autoreg.py:281-297, helpers/sampling.py:24-25):
// logits[b, allowed == 0] = -inf; tok = argmax (first max wins, like torch.argmax); seq[b, pos] = tok;
// pad_mask[b, pos] = (tok == pad); optionally x_next[b, :] = emb[tok]*scale + pe[pos].
__global__ void argmax_next_kernel(const float* __restrict__ logits, int ldl, int V,
                                   const unsigned char* __restrict__ allowed, long long* __restrict__ seq, int seq_ld,
                                   int pos, unsigned char* __restrict__ pad_mask, int mask_ld, long long pad_id,
                                   const float* __restrict__ emb, int D, float scale, const float* __restrict__ pe,
                                   float* __restrict__ x_next) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __shared__ float bv[8];
  __shared__ int bi[8];
  __shared__ int win;
  float best = -INFINITY;
  int arg = 0x7fffffff;
  for (int c = threadIdx.x; c < V; c += blockDim.x) {
    const float x = allowed[c] ? logits[static_cast<long long>(b) * ldl + c] : -INFINITY;
    if (x > best || (x == best && c < arg)) { best = x; arg = c; }
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, off);
    const int oi = __shfl_xor_sync(0xffffffffu, arg, off);
    if (ov > best || (ov == best && oi < arg)) { best = ov; arg = oi; }
  }
  if (lane == 0) { bv[wid] = best; bi[wid] = arg; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float v = bv[0];
    int a = bi[0];
    for (int w = 1; w < static_cast<int>(blockDim.x >> 5); ++w)
      if (bv[w] > v || (bv[w] == v && bi[w] < a)) { v = bv[w]; a = bi[w]; }
    win = a;
    seq[static_cast<long long>(b) * seq_ld + pos] = a;
    if (pad_mask) pad_mask[static_cast<long long>(b) * mask_ld + pos] = (a == pad_id) ? 1 : 0;
  }
  __syncthreads();
  if (x_next) {
    const int t = win;
    for (int c = threadIdx.x; c < D; c += blockDim.x)
      x_next[static_cast<long long>(b) * D + c] = emb[static_cast<long long>(t) * D + c] * scale + pe[static_cast<long long>(pos) * D + c];
  }
}

// Stochastic / constrained step tail.  One CTA per canvas.
//   forced[b, step] >= 0  -> that token (DECODE_SPACE_RESTRICTION for c / cwh / refinement and the teacher-forced prefix of
//                            `partial`, precomputed on the host as one int32 table: decoding_space_restriction.py:5-106)
//   mode 0                -> greedy (first maximum of the masked logits, helpers/sampling.py:24-25)
//   mode 1 random, 2 top_k, 3 top_p, 4 gumbel (helpers/sampling.py:28-58): x = logits / temperature (+ gumbel noise),
//     sorted descending by (x, then lower index); top_k keeps every x >= the k-th largest; top_p drops every rank r > 0
//     whose inclusive cumulative softmax mass exceeds top_p; the token is the first rank whose cumulative mass among the
//     kept ranks exceeds uniform[b] * (kept mass)  (an inverse-CDF draw: same distribution as torch.multinomial).
// Then, like argmax_next_kernel: seq[b, pos] = tok; pad_mask; x_next = emb[tok] * scale + pe[pos].
constexpr int kSampleSlots = 1024;
__global__ void __launch_bounds__(256) sample_next_kernel(
    const float* __restrict__ logits, int ldl, int V, const unsigned char* __restrict__ allowed,
    const int* __restrict__ forced, int forced_ld, int step, int mode, float temperature, int top_k, float top_p,
    const float* __restrict__ uniform, const float* __restrict__ noise, int noise_ld, long long* __restrict__ seq,
    int seq_ld, int pos, unsigned char* __restrict__ pad_mask, int mask_ld, long long pad_id,
    const float* __restrict__ emb, int D, float scale, const float* __restrict__ pe, float* __restrict__ x_next) {
  __shared__ float val[kSampleSlots];
  __shared__ int idx[kSampleSlots];
  __shared__ int win;
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x, tid = threadIdx.x;
  const int f = forced ? forced[static_cast<long long>(b) * forced_ld + step] : -1;
  if (f >= 0) {
    if (tid == 0) win = f;
  } else {
    for (int c = tid; c < kSampleSlots; c += blockDim.x) {
      float x = -INFINITY;
      if (c < V && allowed[c]) {
        x = logits[static_cast<long long>(b) * ldl + c];
        if (mode != 0) {
          x = x / temperature;
          if (mode == 4) {
            const float u = noise[static_cast<long long>(b) * noise_ld + c];
            x += -logf(-logf(u + 1e-30f) + 1e-30f);
          }
        }
      }
      val[c] = x;
      idx[c] = c;
    }
    __syncthreads();
    if (mode == 0) {  // greedy: block arg-max, first maximum wins
      for (int s = kSampleSlots / 2; s >= 1; s >>= 1) {
        for (int i = tid; i < s; i += blockDim.x) {
          const float a = val[i], o = val[i + s];
          const int ia = idx[i], io = idx[i + s];
          if (o > a || (o == a && io < ia)) { val[i] = o; idx[i] = io; }
        }
        __syncthreads();
      }
      if (tid == 0) win = idx[0];
    } else {
      // bitonic sort, descending by (value, -index)
      for (int k = 2; k <= kSampleSlots; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
          for (int i = tid; i < kSampleSlots; i += blockDim.x) {
            const int l = i ^ j;
            if (l > i) {
              const float a = val[i], o = val[l];
              const int ia = idx[i], io = idx[l];
              const bool i_first = (a > o) || (a == o && ia < io);  // i already ranks before l in descending order
              const bool desc = ((i & k) == 0);
              if (desc ? !i_first : i_first) { val[i] = o; val[l] = a; idx[i] = io; idx[l] = ia; }
            }
          }
          __syncthreads();
        }
      }
      const float m = val[0];
      __shared__ int nvalid_s;
      if (tid == 0) nvalid_s = 0;
      __syncthreads();
      int cnt = 0;
      for (int i = tid; i < kSampleSlots; i += blockDim.x) cnt += (val[i] > -INFINITY) ? 1 : 0;
      if (cnt) atomicAdd(&nvalid_s, cnt);
      __syncthreads();
      const int nvalid = nvalid_s;
      float kth = -INFINITY;
      if (mode == 2 && top_k >= 1) kth = val[min(top_k, kSampleSlots) - 1];
      __syncthreads();
      for (int i = tid; i < kSampleSlots; i += blockDim.x) {  // unnormalised probabilities, in sorted order
        const float x = val[i];
        val[i] = (x > -INFINITY) ? expf(x - m) : 0.f;
        if (mode == 2 && x < kth) val[i] = 0.f;
      }
      __syncthreads();
      if (tid == 0) {
        int nkeep = nvalid;
        if (nvalid == 0) {
          win = idx[0];
        } else {
          if (mode == 3) {
            float total = 0.f;
            for (int r = 0; r < nvalid; ++r) total += val[r];
            float cum = 0.f;
            nkeep = 1;
            for (int r = 0; r < nvalid; ++r) {
              cum += val[r] / total;
              if (r > 0 && cum > top_p) break;
              nkeep = r + 1;
            }
          }
          float kept = 0.f;
          for (int r = 0; r < nkeep; ++r) kept += val[r];
          const float target = uniform[b] * kept;
          float acc = 0.f;
          int pick = 0;
          for (int r = 0; r < nkeep; ++r) {
            if (val[r] <= 0.f) break;
            pick = r;
            acc += val[r];
            if (acc > target) break;
          }
          win = idx[pick];
        }
      }
    }
  }
  __syncthreads();
  const int t = win;
  if (tid == 0) {
    seq[static_cast<long long>(b) * seq_ld + pos] = t;
    if (pad_mask) pad_mask[static_cast<long long>(b) * mask_ld + pos] = (t == pad_id) ? 1 : 0;
  }
  if (x_next)
    for (int c = tid; c < D; c += blockDim.x)
      x_next[static_cast<long long>(b) * D + c] = emb[static_cast<long long>(t) * D + c] * scale + pe[static_cast<long long>(pos) * D + c];
}

// Scatter this step's K and V (columns [D, 3D) of the fused QKV projection) into the self-attention cache.
__global__ void kv_append_kernel(const float* __restrict__ qkv, int B, int D, float* __restrict__ kcache,
                                 float* __restrict__ vcache, int S, int pos) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * D) return;
  const int b = i / D, c = i - b * D;
  kcache[(static_cast<long long>(b) * S + pos) * D + c] = qkv[static_cast<long long>(b) * 3 * D + D + c];
  vcache[(static_cast<long long>(b) * S + pos) * D + c] = qkv[static_cast<long long>(b) * 3 * D + 2 * D + c];
}


// Label-smoothed cross entropy (nn.CrossEntropyLoss(label_smoothing=eps, ignore_index=pad), mean over
// non-ignored targets; retrieval_augmented_autoreg.py:140-142,213-214).  One warp per row:
//   row_loss = (1-eps) * (lse - x[t]) + eps * (lse - mean_c x[c]);  rows with t == ignore contribute 0.
__global__ void ce_rows_kernel(const float* __restrict__ logits, int ldl, const long long* __restrict__ tgt, int M, int V,
                               float eps, long long ignore, float* __restrict__ row_loss, float* __restrict__ row_valid) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const float* x = logits + static_cast<long long>(row) * ldl;
  float mx = -INFINITY;
  for (int c = lane; c < V; c += 32) mx = fmaxf(mx, x[c]);
  mx = warp_max(mx);
  float se = 0.f, sx = 0.f;
  for (int c = lane; c < V; c += 32) {
    se += expf(x[c] - mx);
    sx += x[c];
  }
  se = warp_sum(se);
  sx = warp_sum(sx);
  if (lane == 0) {
    const long long t = tgt[row];
    const float lse = mx + logf(se);
    const bool ok = (t != ignore);
    row_loss[row] = ok ? ((1.f - eps) * (lse - x[t]) + eps * (lse - sx / static_cast<float>(V))) : 0.f;
    row_valid[row] = ok ? 1.f : 0.f;
  }
}
// Deterministic final reduction: one block, fixed order.
__global__ void ce_reduce_kernel(const float* __restrict__ row_loss, const float* __restrict__ row_valid, int M,
                                 float* __restrict__ out) {
  __shared__ float sl[32], sv[32];
  float a = 0.f, b = 0.f;
  for (int i = threadIdx.x; i < M; i += blockDim.x) { a += row_loss[i]; b += row_valid[i]; }
  a = warp_sum(a);
  b = warp_sum(b);
  if ((threadIdx.x & 31) == 0) { sl[threadIdx.x >> 5] = a; sv[threadIdx.x >> 5] = b; }
  __syncthreads();
  if (threadIdx.x < 32) {
    a = threadIdx.x < (blockDim.x >> 5) ? sl[threadIdx.x] : 0.f;
    b = threadIdx.x < (blockDim.x >> 5) ? sv[threadIdx.x] : 0.f;
    a = warp_sum(a);
    b = warp_sum(b);
    if (threadIdx.x == 0) out[0] = a / b;
  }
}

static inline int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace ralf

using namespace ralf;
#define ST(s) reinterpret_cast<cudaStream_t>(s)
#define BF(p) reinterpret_cast<__nv_bfloat16*>(p)
#define CBF(p) reinterpret_cast<const __nv_bfloat16*>(p)

extern "C" int ralf_layernorm(const float* x, long long in_ld, const float* gamma, const float* beta, float eps,
                              int M, int D, float* out_f32, void* out_split, long long out_plane, void* stream) {
  if (!x || !gamma || !beta) return RALF_ERR_NULL;
  if (M <= 0 || D <= 0 || D > 1024 || (D & 31)) return RALF_ERR_SHAPE;
#define RALF_LN_CASE(P)                                                                                          \
  if (D == 32 * P) {                                                                                            \
    const cudaError_t e = launch_pdl(layernorm_kernel<P>, dim3((M + 7) / 8), dim3(256), 0, ST(stream), x, in_ld, \
                                     gamma, beta, eps, M, D, out_f32, BF(out_split), out_plane);                \
    return set_cuda_error(e != cudaSuccess ? e : cudaGetLastError());                                           \
  }
  RALF_LN_CASE(8)   // d_model = 256: every LayerNorm of the RALF path
  RALF_LN_CASE(1) RALF_LN_CASE(2) RALF_LN_CASE(4) RALF_LN_CASE(16) RALF_LN_CASE(32)
#undef RALF_LN_CASE
  return RALF_ERR_SHAPE;  // D / 32 must be a power of two <= 32
}

static int attention_impl(const float* q, int ldq, const float* k, const float* v, int ldk,
                          const unsigned char* key_padding_mask, int B, int H, int Tq, int Tk, int head_dim, int causal,
                          float scale, void* out_split, long long out_plane, float* out_f32, int ldo, DropArgs da,
                          void* stream, float* lse_out = nullptr) {
  if (!q || !k || !v) return RALF_ERR_NULL;
  if (B <= 0 || H <= 0 || Tq <= 0 || Tk <= 0 || (head_dim != 32 && head_dim != 64)) return RALF_ERR_SHAPE;
  if ((ldq & 3) || (ldk & 3)) return RALF_ERR_ALIGN;
  if (!da.thresh24) {
    // image-encoder shape class (8 x 32 heads, no mask, <= 256 keys): tcgen05 kernel (attention_tc.cu)
    const int tc = attention_tc_try(q, ldq, k, v, ldk, key_padding_mask, B, H, Tq, Tk, head_dim, causal, scale, out_split,
                                    out_plane, out_f32, ldo, ST(stream));
    if (tc != 0) return tc < 0 ? tc : 0;
  }
  // few keys, wide heads (fusion Attention over the 16 retrieved layouts): tile kernel with coalesced stores
  // RALF_ATTN_FEWKEYS: 2 (default) = K/V staged in shared memory (attention_kvsmem_kernel), 1 = the round-1 tile kernel
  // (broadcast __ldg), 0 = generic kernel.  1 and 2 give bit-identical results.
  static const int fewkeys_mode = getenv("RALF_ATTN_FEWKEYS") ? atoi(getenv("RALF_ATTN_FEWKEYS")) : 2;
  const bool fewkeys_on = fewkeys_mode != 0;
  if (fewkeys_mode == 2 && !da.thresh24 && head_dim == 64 && Tk <= kFewKeysMax && (H == 8 || H == 4) && !causal &&
      (ldo & 3) == 0 && (out_plane & 3) == 0 && (reinterpret_cast<uintptr_t>(k) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(v) & 15) == 0 && (reinterpret_cast<uintptr_t>(q) & 15) == 0) {
    if (H == 8) {
      if (Tq > 16)
        return launch_kvsmem<8, 32>(q, ldq, k, v, ldk, key_padding_mask, B, Tq, Tk, scale, out_split, out_plane, out_f32, ldo,
                                    ST(stream));
      return launch_kvsmem<8, 16>(q, ldq, k, v, ldk, key_padding_mask, B, Tq, Tk, scale, out_split, out_plane, out_f32, ldo,
                                  ST(stream));
    }
    if (Tq > 16)
      return launch_kvsmem<4, 32>(q, ldq, k, v, ldk, key_padding_mask, B, Tq, Tk, scale, out_split, out_plane, out_f32, ldo,
                                  ST(stream));
    return launch_kvsmem<4, 16>(q, ldq, k, v, ldk, key_padding_mask, B, Tq, Tk, scale, out_split, out_plane, out_f32, ldo,
                                ST(stream));
  }
  if (fewkeys_on && !da.thresh24 && head_dim == 64 && Tk <= kFewKeysMax && H <= 8 && !causal &&
      (ldo & 3) == 0 && (out_plane & 3) == 0) {
    static bool attr_set = false;
    const int smem = H * 32 * (64 + 4) * static_cast<int>(sizeof(float));
    if (!attr_set) {
      cudaError_t e = cudaFuncSetAttribute(attention_fewkeys_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 32 * 68 * 4);
      if (e != cudaSuccess) return set_cuda_error(e);
      attr_set = true;
    }
    dim3 g2((Tq + 31) / 32, B);
    attention_fewkeys_kernel<<<g2, 32 * H, smem, ST(stream)>>>(q, ldq, k, v, ldk, key_padding_mask, Tq, Tk, scale,
                                                              BF(out_split), out_plane, out_f32, ldo);
    return set_cuda_error(cudaGetLastError());
  }
  const int threads = Tq >= 128 ? 128 : ((Tq + 31) / 32) * 32;
  dim3 grid((Tq + threads - 1) / threads, H, B);
  if (head_dim == 32)
    attention_kernel<32><<<grid, threads, 0, ST(stream)>>>(q, ldq, k, v, ldk, key_padding_mask, Tq, Tk, causal, scale,
                                                          BF(out_split), out_plane, out_f32, ldo, da, lse_out);
  else
    attention_kernel<64><<<grid, threads, 0, ST(stream)>>>(q, ldq, k, v, ldk, key_padding_mask, Tq, Tk, causal, scale,
                                                          BF(out_split), out_plane, out_f32, ldo, da, lse_out);
  return set_cuda_error(cudaGetLastError());
}

extern "C" int ralf_attention(const float* q, int ldq, const float* k, const float* v, int ldk,
                              const unsigned char* key_padding_mask, int B, int H, int Tq, int Tk, int head_dim,
                              int causal, float scale, void* out_split, long long out_plane, float* out_f32, int ldo,
                              void* stream) {
  return attention_impl(q, ldq, k, v, ldk, key_padding_mask, B, H, Tq, Tk, head_dim, causal, scale, out_split, out_plane,
                        out_f32, ldo, make_drop_args(nullptr, 0, 0.f), stream);
}

extern "C" int ralf_attention_dropout(const float* q, int ldq, const float* k, const float* v, int ldk,
                                      const unsigned char* key_padding_mask, int B, int H, int Tq, int Tk, int head_dim,
                                      int causal, float scale, void* out_split, long long out_plane, float* out_f32,
                                      int ldo, const unsigned long long* seed, unsigned int site, float p, float* lse_out,
                                      void* stream) {
  if (!seed) return RALF_ERR_NULL;
  if (!(p >= 0.f && p < 1.f)) return RALF_ERR_SHAPE;
  if (lse_out && !(p > 0.f)) return RALF_ERR_SHAPE;  // p = 0 may take a kernel that does not produce the log-sum-exp
  return attention_impl(q, ldq, k, v, ldk, key_padding_mask, B, H, Tq, Tk, head_dim, causal, scale, out_split, out_plane,
                        out_f32, ldo, make_drop_args(seed, site, p), stream, lse_out);
}

static int attention_decode_impl(const float* q, int ldq, const float* k, const float* v, long long kv_bstride, int ldk,
                                 const unsigned char* key_padding_mask, int mask_ld, int Tk, int B, int H, int head_dim,
                                 float scale, void* out_split, long long out_plane, int ldo, const float* knew,
                                 const float* vnew, int ldnew, float* kc_w, float* vc_w, void* stream) {
  if (!q || !k || !v || !out_split) return RALF_ERR_NULL;
  if (B <= 0 || H <= 0 || Tk <= 0 || Tk > 2048 || (head_dim != 32 && head_dim != 64)) return RALF_ERR_SHAPE;
  if ((ldq & 3) || (ldk & 3)) return RALF_ERR_ALIGN;
  // Single-pass streaming kernel (CTA per canvas, warp per head): the memory cross-attention (no mask, long Tk) and the
  // decode self-attention over the short token cache (append fused, key-padding mask).  RALF_DECODE_2PASS=1 keeps the
  // two-pass kernel for the masked / appended case (A/B runs).
  static const bool two_pass = getenv("RALF_DECODE_2PASS") && atoi(getenv("RALF_DECODE_2PASS")) != 0;
  const bool plain = !key_padding_mask && !knew;
  if (H <= 8 && (plain ? Tk >= 64 : !two_pass)) {
    cudaError_t e;
    if (head_dim == 32)
      e = launch_pdl(attention_decode_stream_kernel<32>, dim3(B), dim3(32 * H), 0, ST(stream), q, ldq, k, v, kv_bstride,
                     ldk, Tk, H, scale, BF(out_split), out_plane, ldo, key_padding_mask, mask_ld, knew, vnew, ldnew, kc_w,
                     vc_w);
    else
      e = launch_pdl(attention_decode_stream_kernel<64>, dim3(B), dim3(32 * H), 0, ST(stream), q, ldq, k, v, kv_bstride,
                     ldk, Tk, H, scale, BF(out_split), out_plane, ldo, key_padding_mask, mask_ld, knew, vnew, ldnew, kc_w,
                     vc_w);
    return set_cuda_error(e != cudaSuccess ? e : cudaGetLastError());
  }
  const size_t smem = static_cast<size_t>((Tk + 31) & ~31) * sizeof(float);
  const int grid = B * H;
  const int warps = 4;
  if (head_dim == 32)
    attention_decode_kernel<32><<<grid, warps * 32, smem, ST(stream)>>>(q, ldq, k, v, kv_bstride, ldk, key_padding_mask,
                                                                       mask_ld, Tk, B, H, scale, BF(out_split),
                                                                       out_plane, ldo, knew, vnew, ldnew, kc_w, vc_w);
  else
    attention_decode_kernel<64><<<grid, warps * 32, smem, ST(stream)>>>(q, ldq, k, v, kv_bstride, ldk, key_padding_mask,
                                                                       mask_ld, Tk, B, H, scale, BF(out_split),
                                                                       out_plane, ldo, knew, vnew, ldnew, kc_w, vc_w);
  return set_cuda_error(cudaGetLastError());
}

extern "C" int ralf_attention_decode(const float* q, int ldq, const float* k, const float* v, long long kv_bstride,
                                     int ldk, const unsigned char* key_padding_mask, int mask_ld, int Tk, int B, int H,
                                     int head_dim, float scale, void* out_split, long long out_plane, int ldo,
                                     void* stream) {
  return attention_decode_impl(q, ldq, k, v, kv_bstride, ldk, key_padding_mask, mask_ld, Tk, B, H, head_dim, scale,
                               out_split, out_plane, ldo, nullptr, nullptr, 0, nullptr, nullptr, stream);
}

extern "C" int ralf_attention_decode_kv24(const float* q, int ldq, const void* kv24, long long kv_bstride, int Tk, int B,
                                          int H, float scale, void* out_split, long long out_plane, int ldo,
                                          void* stream) {
  if (!q || !kv24 || !out_split) return RALF_ERR_NULL;
  if (B <= 0 || H != 8 || Tk <= 0) return RALF_ERR_SHAPE;  // row format: 8 heads x 32 (d_model 256)
  if ((ldq & 3) || (reinterpret_cast<uintptr_t>(kv24) & 15)) return RALF_ERR_ALIGN;
  const cudaError_t e = launch_pdl(attention_decode_kv24_kernel, dim3(B), dim3(32 * H), 0, ST(stream), q, ldq,
                                   reinterpret_cast<const uint8_t*>(kv24), kv_bstride, Tk, H, scale, BF(out_split),
                                   out_plane, ldo);
  return set_cuda_error(e != cudaSuccess ? e : cudaGetLastError());
}

// Four lanes per key (8 channels = one 16-byte load each for K and for V): per cache value a third of the shuffles and
// half of the load / scale instructions of the eight-lane kernel above, 8 keys per warp-wide load, 64 keys per batch.
__device__ __forceinline__ void kv16_unpack8(const uint4 w, float (&f)[8]) {
  constexpr float BIAS = 8388608.f + 32768.f;
  const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    f[2 * e] = __uint_as_float(0x4B000000u | (ww[e] & 0xffffu)) - BIAS;
    f[2 * e + 1] = __uint_as_float(__byte_perm(ww[e], 0x4B000000u, 0x7632)) - BIAS;
  }
}
__global__ void __launch_bounds__(256)
attention_decode_kv16x4_kernel(const float* __restrict__ q, int ldq, const uint8_t* __restrict__ kv, long long kv_bstride,
                               int Tk, int H, float scale, __nv_bfloat16* __restrict__ out_split, long long out_plane,
                               int ldo) {
  constexpr int DH = 32, CPL = 4, KPI = 8, UN = 8, ROW = 1088;
  pdl_trigger();
  pdl_wait();
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x;
  if (h >= H) return;
  const int kk = lane / CPL, c = lane % CPL;
  float qr[8];
  {
    const float* qp = q + static_cast<long long>(b) * ldq + h * DH + 8 * c;
    const float4 a = *reinterpret_cast<const float4*>(qp), bq = *reinterpret_cast<const float4*>(qp + 4);
    qr[0] = a.x * scale; qr[1] = a.y * scale; qr[2] = a.z * scale; qr[3] = a.w * scale;
    qr[4] = bq.x * scale; qr[5] = bq.y * scale; qr[6] = bq.z * scale; qr[7] = bq.w * scale;
  }
  const uint8_t* base = kv + static_cast<long long>(b) * kv_bstride * ROW;
  const int o_k = h * 64 + c * 16, o_v = 512 + h * 64 + c * 16, o_s = 1024 + h * 8;
  float m = -INFINITY, l = 0.f;
  float o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = 0.f;
#pragma unroll 1
  for (int j0 = 0; j0 < Tk; j0 += KPI * UN) {
    uint4 kw[UN], vw[UN];
    float ksc[UN], vsc[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int j = j0 + u * KPI + kk;
      kw[u] = vw[u] = make_uint4(0x80008000u, 0x80008000u, 0x80008000u, 0x80008000u);  // offset-binary zero
      ksc[u] = vsc[u] = 0.f;
      if (j < Tk) {
        const uint8_t* r = base + static_cast<long long>(j) * ROW;
        kw[u] = __ldcs(reinterpret_cast<const uint4*>(r + o_k));
        vw[u] = __ldcs(reinterpret_cast<const uint4*>(r + o_v));
        const float2 sc2 = __ldcs(reinterpret_cast<const float2*>(r + o_s));
        ksc[u] = sc2.x;
        vsc[u] = sc2.y;
      }
    }
    float s[UN];
    float bm = -INFINITY;
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      float kf[8];
      kv16_unpack8(kw[u], kf);
      float d = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) d = fmaf(qr[i], kf[i], d);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      s[u] = (j0 + u * KPI + kk < Tk) ? d * ksc[u] : -INFINITY;
      bm = fmaxf(bm, s[u]);
    }
#pragma unroll
    for (int off = CPL; off < 32; off <<= 1) bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, off));
    const float m_new = fmaxf(m, bm);
    const float corr = __expf(m - m_new);
    l *= corr;
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] *= corr;
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const float p = __expf(s[u] - m_new);
      float vf[8];
      kv16_unpack8(vw[u], vf);
      const float pv = p * vsc[u];
      l += p;
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = fmaf(pv, vf[i], o[i]);
    }
    m = m_new;
  }
#pragma unroll
  for (int off = CPL; off < 32; off <<= 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] += __shfl_xor_sync(0xffffffffu, o[i], off);
    l += __shfl_xor_sync(0xffffffffu, l, off);
  }
  if (kk == 0) {
    const float inv = 1.f / l;
    const long long off0 = static_cast<long long>(b) * ldo + h * DH + 8 * c;
#pragma unroll
    for (int i = 0; i < 8; ++i) store_split(out_split, out_plane, off0 + i, o[i] * inv);
  }
}

extern "C" int ralf_attention_decode_kv16(const float* q, int ldq, const void* kv16, long long kv_bstride, int Tk, int B,
                                          int H, float scale, void* out_split, long long out_plane, int ldo,
                                          void* stream) {
  if (!q || !kv16 || !out_split) return RALF_ERR_NULL;
  if (B <= 0 || H != 8 || Tk <= 0) return RALF_ERR_SHAPE;  // row format: 8 heads x 32 (d_model 256)
  if ((ldq & 3) || (reinterpret_cast<uintptr_t>(kv16) & 15)) return RALF_ERR_ALIGN;
  // (A lane-per-key variant -- whole 64-byte head rows per lane, no shuffles, private online softmax -- was measured in
  // round 2: 0.155 ms against 0.105 ms for this one; 32 different rows per load instruction cost more L1 wavefronts than
  // the shuffles it saved.  Removed; profiles/r2_kv16_ncu.md.)
  // RALF_KV16_LANES: 4 (default) = four lanes per key, 16-byte loads; 8 = the first version (eight lanes, 8-byte loads)
  static const int lanes = getenv("RALF_KV16_LANES") ? atoi(getenv("RALF_KV16_LANES")) : 4;
  const cudaError_t e =
      lanes == 8 ? launch_pdl(attention_decode_kv16_kernel, dim3(B), dim3(32 * H), 0, ST(stream), q, ldq,
                              reinterpret_cast<const uint8_t*>(kv16), kv_bstride, Tk, H, scale, BF(out_split), out_plane, ldo)
                 : launch_pdl(attention_decode_kv16x4_kernel, dim3(B), dim3(32 * H), 0, ST(stream), q, ldq,
                              reinterpret_cast<const uint8_t*>(kv16), kv_bstride, Tk, H, scale, BF(out_split), out_plane, ldo);
  return set_cuda_error(e != cudaSuccess ? e : cudaGetLastError());
}

extern "C" int ralf_attention_decode_append(const float* qkv, int ldqkv, float* kcache, float* vcache, int S, int pos,
                                            const unsigned char* key_padding_mask, int mask_ld, int B, int H,
                                            int head_dim, float scale, void* out_split, long long out_plane, int ldo,
                                            void* stream) {
  if (!qkv || !kcache || !vcache) return RALF_ERR_NULL;
  if (pos < 0 || pos >= S) return RALF_ERR_SHAPE;
  const int Dm = H * head_dim;
  return attention_decode_impl(qkv, ldqkv, kcache, vcache, S, Dm, key_padding_mask, mask_ld, pos + 1, B, H, head_dim,
                               scale, out_split, out_plane, ldo, qkv + Dm, qkv + 2 * Dm, ldqkv, kcache, vcache, stream);
}

extern "C" int ralf_stem_im2col(const float* img, int B, int H, int W, int KP, void* out, long long out_plane,
                                void* stream) {
  if (!img || !out) return RALF_ERR_NULL;
  if (B <= 0 || H <= 0 || W <= 0 || KP < 196 || KP > 200 || (KP & 7)) return RALF_ERR_SHAPE;
  const int Ho = (H + 6 - 7) / 2 + 1, Wo = (W + 6 - 7) / 2 + 1;
  const long long total = static_cast<long long>(B) * Ho * Wo * 50;
  stem_im2col_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(img, B, H, W, Ho, Wo, KP, BF(out), out_plane);
  return set_cuda_error(cudaGetLastError());
}

extern "C" int ralf_stem_s2d(const float* img, int B, int H, int W, void* out, long long out_plane, void* stream) {
  if (!img || !out) return RALF_ERR_NULL;
  if (B <= 0 || H <= 0 || W <= 0 || (H & 1) || (W & 1)) return RALF_ERR_SHAPE;
  const int Hp = H / 2 + 3, Wp = W / 2 + 3;
  const long long total = static_cast<long long>(B) * Hp * Wp;
  stem_s2d_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(img, B, H, W, Hp, Wp, BF(out), out_plane);
  return set_cuda_error(cudaGetLastError());
}

extern "C" int ralf_im2col(const void* in, long long in_plane, int B, int H, int W, int C, int KH, int KW, int stride,
                           int pad, void* out, long long out_plane, void* stream) {
  if (!in || !out) return RALF_ERR_NULL;
  if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 7) || KH <= 0 || KW <= 0 || stride <= 0) return RALF_ERR_SHAPE;
  const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
  const long long total = static_cast<long long>(B) * Ho * Wo * KH * KW * (C >> 3);
  im2col_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(CBF(in), in_plane, B, H, W, C, KH, KW, stride, pad, Ho,
                                                             Wo, BF(out), out_plane);
  return set_cuda_error(cudaGetLastError());
}

extern "C" int ralf_maxpool3x3s2(const void* in, long long in_plane, int B, int H, int W, int C, void* out,
                                 long long out_plane, void* stream) {
  if (!in || !out) return RALF_ERR_NULL;
  if (B <= 0 || H <= 0 || W <= 0 || C <= 0 || (C & 7)) return RALF_ERR_SHAPE;
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long long total = static_cast<long long>(B) * Ho * Wo * (C >> 3);
  maxpool_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(CBF(in), in_plane, B, H, W, C, Ho, Wo, BF(out),
                                                              out_plane);
  return set_cuda_error(cudaGetLastError());
}

extern "C" int ralf_fpn_merge(const float* c5, const float* c4, int B, int h5, int w5, int h4, int w4, int C,
                              void* fused, long long fused_plane, int ldf, void* sum, long long sum_plane,
                              void* stream) {
  if (!c5 || !c4 || !fused || !sum) return RALF_ERR_NULL;
  if (B <= 0 || h5 <= 0 || w5 <= 0 || h4 <= 0 || w4 <= 0 || C <= 0) return RALF_ERR_SHAPE;
  const long long total = static_cast<long long>(B) * h4 * w4 * C;
  fpn_merge_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(c5, c4, B, h5, w5, h4, w4, C, BF(fused), fused_plane,
                                                                ldf, BF(sum), sum_plane);
  return set_cuda_error(cudaGetLastError());
}

extern "C" int ralf_rows_affine(const float* in, long long in_ld, int M, int D, float scale, float add,
                                const float* table, int tab_mod, int rows_per_group, int group_stride,
                                int group_offset, float* out_f32, void* out_split, long long out_plane, int out_ld,
                                void* stream) {
  if (M <= 0 || D <= 0) return RALF_ERR_SHAPE;
  if (!out_f32 && !out_split) return RALF_ERR_NULL;
  const int rpg = rows_per_group > 0 ? rows_per_group : M;
  const int gs = rows_per_group > 0 ? group_stride : 0;
  const int go = rows_per_group > 0 ? group_offset : 0;
  const long long total = static_cast<long long>(M) * D;
  rows_affine_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(in, in_ld, M, D, scale, add, table,
                                                                  tab_mod > 0 ? tab_mod : 1, rpg, gs, go, out_f32,
                                                                  BF(out_split), out_plane, out_ld);
  return set_cuda_error(cudaGetLastError());
}

extern "C" int ralf_embed(const long long* tok, long long tok_ld, int tok_col, int B, int S, const float* emb, int D,
                          float scale, const float* pe, int pos0, float* out, void* stream) {
  if (!tok || !emb || !pe || !out) return RALF_ERR_NULL;
  if (B <= 0 || S <= 0 || D <= 0) return RALF_ERR_SHAPE;
  const long long total = static_cast<long long>(B) * S * D;
  embed_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(tok, tok_ld, tok_col, B, S, emb, D, scale, pe, pos0, out);
  return set_cuda_error(cudaGetLastError());
}

extern "C" int ralf_fid_embed(const float* cx, const float* cy, const float* w, const float* h, const long long* label,
                              int rows, int D, const float* fc_w, const float* fc_b, const float* emb, void* out,
                              long long out_plane, void* stream) {
  if (!cx || !cy || !w || !h || !label || !fc_w || !fc_b || !emb || !out) return RALF_ERR_NULL;
  if (rows <= 0 || D <= 0) return RALF_ERR_SHAPE;
  const long long total = static_cast<long long>(rows) * 2 * D;
  fid_embed_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(cx, cy, w, h, label, rows, D, fc_w, fc_b, emb, BF(out),
                                                                out_plane);
  return set_cuda_error(cudaGetLastError());
}

extern "C" int ralf_argmax_next(const float* logits, int ldl, int B, int V, const unsigned char* allowed,
                                long long* seq, int seq_ld, int pos, unsigned char* pad_mask, int mask_ld,
                                long long pad_id, const float* emb, int D, float scale, const float* pe, float* x_next,
                                void* stream) {
  if (!logits || !allowed || !seq) return RALF_ERR_NULL;
  if (B <= 0 || V <= 0) return RALF_ERR_SHAPE;
  const cudaError_t e = launch_pdl(argmax_next_kernel, dim3(B), dim3(128), 0, ST(stream), logits, ldl, V, allowed, seq, seq_ld,
                                   pos, pad_mask, mask_ld, pad_id, emb, D, scale, pe, x_next);
  return set_cuda_error(e != cudaSuccess ? e : cudaGetLastError());
}

extern "C" int ralf_sample_next(const float* logits, int ldl, int B, int V, const unsigned char* allowed,
                                const int* forced, int forced_ld, int step, int mode, float temperature, int top_k,
                                float top_p, const float* uniform, const float* noise, int noise_ld, long long* seq,
                                int seq_ld, int pos, unsigned char* pad_mask, int mask_ld, long long pad_id,
                                const float* emb, int D, float scale, const float* pe, float* x_next, void* stream) {
  if (!logits || !allowed || !seq) return RALF_ERR_NULL;
  if (B <= 0 || V <= 0 || V > kSampleSlots) return RALF_ERR_SHAPE;
  if (mode < 0 || mode > 4) return RALF_ERR_SHAPE;
  if (mode != 0 && (!uniform || !(temperature > 0.f))) return RALF_ERR_NULL;
  if (mode == 4 && !noise) return RALF_ERR_NULL;
  if (mode == 2 && top_k < 1) return RALF_ERR_SHAPE;
  if (mode == 3 && !(top_p > 0.f && top_p <= 1.f)) return RALF_ERR_SHAPE;
  if (x_next && (!emb || !pe)) return RALF_ERR_NULL;
  const cudaError_t e = launch_pdl(sample_next_kernel, dim3(B), dim3(256), 0, ST(stream), logits, ldl, V, allowed, forced,
                                   forced_ld, step, mode, temperature, top_k, top_p, uniform, noise, noise_ld, seq, seq_ld,
                                   pos, pad_mask, mask_ld, pad_id, emb, D, scale, pe, x_next);
  return set_cuda_error(e != cudaSuccess ? e : cudaGetLastError());
}

extern "C" int ralf_kv_append(const float* qkv, int B, int D, float* kcache, float* vcache, int S, int pos,
                              void* stream) {
  if (!qkv || !kcache || !vcache) return RALF_ERR_NULL;
  if (B <= 0 || D <= 0 || pos < 0 || pos >= S) return RALF_ERR_SHAPE;
  kv_append_kernel<<<(B * D + 255) / 256, 256, 0, ST(stream)>>>(qkv, B, D, kcache, vcache, S, pos);
  return set_cuda_error(cudaGetLastError());
}

extern "C" int ralf_ce_label_smooth(const float* logits, int ldl, const long long* targets, int M, int V, float eps,
                                    long long ignore_index, float* workspace /* 2*M floats */, float* out_loss,
                                    void* stream) {
  if (!logits || !targets || !workspace || !out_loss) return RALF_ERR_NULL;
  if (M <= 0 || V <= 0) return RALF_ERR_SHAPE;
  ce_rows_kernel<<<(M + 7) / 8, 256, 0, ST(stream)>>>(logits, ldl, targets, M, V, eps, ignore_index, workspace,
                                                     workspace + M);
  ce_reduce_kernel<<<1, 256, 0, ST(stream)>>>(workspace, workspace + M, M, out_loss);
  return set_cuda_error(cudaGetLastError());
}

extern "C" int ralf_gather_layouts(const float* table, const long long* idx, int rows, int row_elems,
                                   long long n_table, long long index_base, float* out, void* stream) {
  if (!table || !idx || !out) return RALF_ERR_NULL;
  if (rows <= 0 || row_elems <= 0) return RALF_ERR_SHAPE;
  const long long total = static_cast<long long>(rows) * row_elems;
  gather_layouts_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(table, idx, rows, row_elems, n_table, index_base,
                                                                     out);
  return set_cuda_error(cudaGetLastError());
}

extern "C" int ralf_fid_embed_packed(const float* packed, int nseq, int E, int D, const float* fc_w, const float* fc_b,
                                     const float* emb, int num_labels, void* out, long long out_plane,
                                     unsigned char* pad_mask, void* stream) {
  if (!packed || !fc_w || !fc_b || !emb || !out || !pad_mask) return RALF_ERR_NULL;
  if (nseq <= 0 || E <= 0 || D <= 0 || num_labels <= 0) return RALF_ERR_SHAPE;
  const long long total = static_cast<long long>(nseq) * E * 2 * D;
  fid_embed_packed_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>(packed, nseq, E, D, fc_w, fc_b, emb, num_labels,
                                                                       BF(out), out_plane, pad_mask);
  return set_cuda_error(cudaGetLastError());
}
