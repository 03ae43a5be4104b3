// Training-side kernels: operand transposes for the weight-gradient GEMMs,
// bias-gradient reductions, LayerNorm / attention / cross-entropy / BatchNorm / pooling backward, dropout,
// embedding scatter, global gradient norm and the fused clip + AdamW update.  The dense contractions of the
// backward pass (dX = dY.W, dW = dY^T.X) reuse the tcgen05 GEMM in gemm.cu with K-major operands produced here.
// fp32 arithmetic; GEMM operands are emitted as split bf16 (hi, lo planes).
#include <algorithm>
#include <float.h>
#include <math.h>

#include "common.cuh"
#include "ralf_internal.h"

namespace ralf {

__device__ __forceinline__ float t_warp_sum(float v) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}
__device__ __forceinline__ void t_store_split(__nv_bfloat16* hi_plane, long long plane, long long off, float x) {
  __nv_bfloat16 h, l;
  split_bf16(x, h, l);
  hi_plane[off] = h;
  hi_plane[plane + off] = l;
}
__device__ __forceinline__ float t_load_split(const __nv_bfloat16* hi_plane, long long plane, long long off) {
  return __bfloat162float(hi_plane[off]) + __bfloat162float(hi_plane[plane + off]);
}

// ------------------------------------------------------------------------------------------------
// Transpose with optional format change: in [R, C] (fp32, row stride ld_in) or split -> out split [C, R]
// (row stride ld_out >= R, padded columns zero-filled by the caller's allocation).  32x32 smem tiles.
// ------------------------------------------------------------------------------------------------
__global__ void transpose_to_split_kernel(const float* __restrict__ in_f32, const __nv_bfloat16* __restrict__ in_split,
                                          long long in_plane, long long ld_in, int R, int C,
                                          __nv_bfloat16* __restrict__ out, long long out_plane, long long ld_out) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    float v = 0.f;
    if (r < R && c < C) {
      const long long off = static_cast<long long>(r) * ld_in + c;
      v = in_f32 ? in_f32[off] : t_load_split(in_split, in_plane, off);
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < C && r < R) t_store_split(out, out_plane, static_cast<long long>(c) * ld_out + r, tile[threadIdx.x][i]);
  }
}

// Multi-tensor weight refresh (after every optimiser step): for each task, fp32 W [R, C] (row stride ld_in) ->
// split W [2, R, w_ld] AND split W^T [2, C, wt_ld] in one pass over 32x32 tiles.  One launch for all ~140 GEMM weights of
// the model instead of two launches per weight (274 of the 2330 kernels of a training step).
struct RefreshTask {
  const float* src;
  long long ld_in;
  int R, C;
  __nv_bfloat16* w;
  long long w_plane, w_ld;
  __nv_bfloat16* wt;
  long long wt_plane, wt_ld;
  int tile0;    // first global tile index of this task
  int tiles_x;  // ceil(C / 32)
};
__global__ void refresh_operands_kernel(const RefreshTask* __restrict__ tasks, int ntasks) {
  __shared__ float tile[32][33];
  const int g = blockIdx.x;
  int lo = 0, hi = ntasks - 1;
  while (lo < hi) {  // last task with tile0 <= g
    const int mid = (lo + hi + 1) >> 1;
    if (tasks[mid].tile0 <= g) lo = mid; else hi = mid - 1;
  }
  const RefreshTask t = tasks[lo];
  const int local = g - t.tile0;
  const int c0 = (local % t.tiles_x) * 32, r0 = (local / t.tiles_x) * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    float v = 0.f;
    if (r < t.R && c < t.C) {
      v = t.src[static_cast<long long>(r) * t.ld_in + c];
      t_store_split(t.w, t.w_plane, static_cast<long long>(r) * t.w_ld + c, v);
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < t.C && r < t.R) t_store_split(t.wt, t.wt_plane, static_cast<long long>(c) * t.wt_ld + r, tile[threadIdx.x][i]);
  }
}

// One tensor, both operands: fp32 dY [R, C] -> split dY [2, R, w_ld] (dgrad A operand) and split dY^T [2, C, wt_ld]
// (wgrad A operand) in one pass -- the backward of every linear layer needs both.
__global__ void split_and_transpose_kernel(const float* __restrict__ src, long long ld_in, int R, int C,
                                           __nv_bfloat16* __restrict__ w, long long w_plane, long long w_ld,
                                           __nv_bfloat16* __restrict__ wt, long long wt_plane, long long wt_ld) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    float v = 0.f;
    if (r < R && c < C) {
      v = src[static_cast<long long>(r) * ld_in + c];
      t_store_split(w, w_plane, static_cast<long long>(r) * w_ld + c, v);
    }
    tile[i][threadIdx.x] = v;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (c < C && r < R) t_store_split(wt, wt_plane, static_cast<long long>(c) * wt_ld + r, tile[threadIdx.x][i]);
  }
}

// fp32 [M, C] -> split [M, C] (same layout): GEMM A operand from an fp32 gradient.
__global__ void to_split_kernel(const float* __restrict__ in, long long total, __nv_bfloat16* __restrict__ out,
                                long long plane) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    t_store_split(out, plane, i, in[i]);
}

// Column sums of an fp32 [M, C] matrix (bias gradients): out[c] (+)= sum_r in[r, c].  Deterministic:
// grid.x column tiles of 32, one block walks all rows with 32x8 threads, smem reduce.
__global__ void colsum_kernel(const float* __restrict__ in, long long ld, int M, int C, float* __restrict__ out,
                              int accumulate) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (c < C)
    for (int r = threadIdx.y; r < M; r += 8) acc += in[static_cast<long long>(r) * ld + c];
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
    out[c] = accumulate ? out[c] + s : s;
  }
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward.  y = (x - mean) * rstd * g + b.  One warp per row:
//   dx = rstd * (dy*g - mean(dy*g) - xhat * mean(dy*g*xhat))        (+ add_to, the residual-stream gradient)
// dgamma / dbeta partials are accumulated per block into part[blockIdx][2][D]; ln_bwd_reduce sums them.
// ------------------------------------------------------------------------------------------------
__global__ void ln_bwd_kernel(const float* __restrict__ x, long long x_ld, const float* __restrict__ dy,
                              const float* __restrict__ gamma, float eps, int M, int D, const float* __restrict__ add_to,
                              float* __restrict__ dx, float* __restrict__ part) {
  extern __shared__ float sh[];  // [warps][2][D] per-warp partials (deterministic: no atomics, fixed reduction order)
  const int warps = blockDim.x >> 5, wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = D >> 5;
  float pg[32], pb[32];
  for (int i = 0; i < per; ++i) { pg[i] = 0.f; pb[i] = 0.f; }
  for (int row = blockIdx.x * warps + wid; row < M; row += gridDim.x * warps) {
    const float* xr = x + static_cast<long long>(row) * x_ld;
    const float* dyr = dy + static_cast<long long>(row) * D;
    float xv[32], gv[32];
    float s = 0.f;
    for (int i = 0; i < per; ++i) { xv[i] = xr[lane + 32 * i]; s += xv[i]; }
    const float mean = t_warp_sum(s) / D;
    float sq = 0.f;
    for (int i = 0; i < per; ++i) { const float d = xv[i] - mean; sq += d * d; }
    const float rstd = rsqrtf(t_warp_sum(sq) / D + eps);
    float s1 = 0.f, s2 = 0.f;
    for (int i = 0; i < per; ++i) {
      const int c = lane + 32 * i;
      const float xh = (xv[i] - mean) * rstd;
      const float d = dyr[c];
      pg[i] += d * xh;
      pb[i] += d;
      gv[i] = d * gamma[c];
      xv[i] = xh;
      s1 += gv[i];
      s2 += gv[i] * xh;
    }
    s1 = t_warp_sum(s1) / D;
    s2 = t_warp_sum(s2) / D;
    for (int i = 0; i < per; ++i) {
      const int c = lane + 32 * i;
      float v = rstd * (gv[i] - s1 - xv[i] * s2);
      const long long off = static_cast<long long>(row) * D + c;
      if (add_to) v += add_to[off];
      dx[off] = v;
    }
  }
  for (int i = 0; i < per; ++i) {
    sh[(wid * 2 + 0) * D + lane + 32 * i] = pg[i];
    sh[(wid * 2 + 1) * D + lane + 32 * i] = pb[i];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < warps; ++w) s += sh[w * 2 * D + i];
    part[static_cast<long long>(blockIdx.x) * 2 * D + i] = s;
  }
}
__global__ void ln_bwd_reduce_kernel(const float* __restrict__ part, int nblocks, int D, float* __restrict__ dgamma,
                                     float* __restrict__ dbeta) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * D) return;
  float s = 0.f;
  for (int b = 0; b < nblocks; ++b) s += part[static_cast<long long>(b) * 2 * D + i];
  if (i < D) dgamma[i] = s; else dbeta[i - D] = s;
}


// ------------------------------------------------------------------------------------------------
// Attention backward (fp32), same addressing conventions as attention_kernel in nn_kernels.cu.
//   P_ij = exp(q_i.k_j*scale - lse_i) (0 where masked), delta_i = sum_d dO_i[d]*O_i[d]
//   dV_j = sum_i P_ij dO_i ; dS_ij = P_ij (dO_i.v_j - delta_i) ; dQ_i = scale sum_j dS_ij k_j ; dK_j = scale sum_i dS_ij q_i
// Kernel 1: thread per query -> dQ, delta.   Kernel 2: thread per key -> dK, dV (two passes to bound registers).
// ------------------------------------------------------------------------------------------------
template <int DH>
__global__ void __launch_bounds__(128)
attn_bwd_dq_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k, const float* __restrict__ v, int ldk,
                   const unsigned char* __restrict__ mask, int Tq, int Tk, int causal, float scale,
                   const __nv_bfloat16* __restrict__ o_split, long long o_plane, const float* __restrict__ dO, int ldo,
                   float* __restrict__ lse, float* __restrict__ delta, float* __restrict__ dq, int lddq, DropArgs da,
                   int lse_given) {
  constexpr int KT = 64;
  __shared__ __align__(16) float ks[KT][DH];
  __shared__ __align__(16) float vs[KT][DH];
  __shared__ unsigned char ms[KT];
  const int b = blockIdx.z, h = blockIdx.y, H = gridDim.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = t < Tq;
  float qr[DH], dor[DH], acc[DH];
  float my_lse = 0.f, dl = 0.f;
  // with attention-probability dropout (mask M' = keep / (1-p)): O = (P o M') V, so dP = M' o (dO V^T) and
  // delta = sum_j P_ij dP_ij = dO_i . O_i still holds; dS = P o (dP - delta)
  const unsigned long long dstream = da.thresh24 ? drop_stream(*da.seed, da.site) : 0ull;
  const unsigned long long drow = ((static_cast<unsigned long long>(b) * H + h) * Tq + t) * Tk;
  if (active) {
    const long long qrow = static_cast<long long>(b) * Tq + t;
    const float* qp = q + qrow * ldq + h * DH;
    const float* dp = dO + qrow * ldo + h * DH;
#pragma unroll
    for (int i = 0; i < DH; ++i) {
      qr[i] = qp[i] * scale;
      dor[i] = dp[i];
      dl = fmaf(dor[i], t_load_split(o_split, o_plane, qrow * ldo + h * DH + i), dl);
      acc[i] = 0.f;
    }
    delta[(static_cast<long long>(b) * H + h) * Tq + t] = dl;
  }
  const int kmax = causal ? min(Tk, (blockIdx.x + 1) * static_cast<int>(blockDim.x)) : Tk;
  // pass 0: log-sum-exp of this query's scores -- skipped when the forward kernel saved it (lse_given: training with
  // dropout always runs attention_kernel, which writes it; halves this kernel's sweep over the keys)
  if (lse_given) {
    if (active) my_lse = lse[(static_cast<long long>(b) * H + h) * Tq + t];
  } else {
    float mrun = -INFINITY, lrun = 0.f;
    for (int j0 = 0; j0 < kmax; j0 += KT) {
      __syncthreads();
      const int nk = min(KT, Tk - j0);
      for (int i = threadIdx.x; i < KT * (DH / 4); i += blockDim.x) {
        const int r = i / (DH / 4), c = (i % (DH / 4)) * 4;
        float4 fk = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nk) fk = *reinterpret_cast<const float4*>(k + (static_cast<long long>(b) * Tk + j0 + r) * ldk + h * DH + c);
        *reinterpret_cast<float4*>(&ks[r][c]) = fk;
      }
      for (int i = threadIdx.x; i < KT; i += blockDim.x)
        ms[i] = (i < nk) ? (mask ? mask[static_cast<long long>(b) * Tk + j0 + i] : 0) : 1;
      __syncthreads();
      if (!active) continue;
      for (int j = 0; j < nk; ++j) {
        if (ms[j] || (causal && (j0 + j) > t)) continue;
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < DH; ++i) s = fmaf(qr[i], ks[j][i], s);
        const float mnew = fmaxf(mrun, s);
        lrun = lrun * __expf(mrun - mnew) + __expf(s - mnew);
        mrun = mnew;
      }
    }
    my_lse = mrun + __logf(lrun);
    if (active) lse[(static_cast<long long>(b) * H + h) * Tq + t] = my_lse;
  }
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
    for (int j = 0; j < nk; ++j) {
      if (ms[j] || (causal && (j0 + j) > t)) continue;
      float s = 0.f, dpv = 0.f;
#pragma unroll
      for (int i = 0; i < DH; ++i) {
        s = fmaf(qr[i], ks[j][i], s);
        dpv = fmaf(dor[i], vs[j][i], dpv);
      }
      if (da.thresh24) dpv = drop_keep(dstream, drow + j0 + j, da.thresh24) ? dpv * da.inv_keep : 0.f;
      const float ds = __expf(s - my_lse) * (dpv - dl) * scale;
#pragma unroll
      for (int i = 0; i < DH; ++i) acc[i] = fmaf(ds, ks[j][i], acc[i]);
    }
  }
  if (!active) return;
  float* out = dq + (static_cast<long long>(b) * Tq + t) * lddq + h * DH;
#pragma unroll
  for (int i = 0; i < DH; ++i) out[i] = acc[i];
}

template <int DH>
__global__ void __launch_bounds__(128)
attn_bwd_dkv_kernel(const float* __restrict__ q, int ldq, const float* __restrict__ k, const float* __restrict__ v, int ldk,
                    const unsigned char* __restrict__ mask, int Tq, int Tk, int causal, float scale,
                    const float* __restrict__ dO, int ldo, const float* __restrict__ lse, const float* __restrict__ delta,
                    float* __restrict__ dk, float* __restrict__ dv, int lddk, DropArgs da) {
  constexpr int QT = 64;
  __shared__ __align__(16) float qs[QT][DH];
  __shared__ __align__(16) float ds_[QT][DH];  // dO tile
  __shared__ float ls[QT], dls[QT];
  const int b = blockIdx.z, h = blockIdx.y, H = gridDim.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = j < Tk;
  const bool dead = active && mask && mask[static_cast<long long>(b) * Tk + j];
  const unsigned long long dstream = da.thresh24 ? drop_stream(*da.seed, da.site) : 0ull;
  const unsigned long long dhead = (static_cast<unsigned long long>(b) * H + h) * Tq;
  // one sweep over the queries: dV and dK accumulate together, K and V rows of this thread's key live in registers
  float kr[DH], vr[DH], accv[DH], acck[DH];
  if (active) {
    const float* kp = k + (static_cast<long long>(b) * Tk + j) * ldk + h * DH;
    const float* vp = v + (static_cast<long long>(b) * Tk + j) * ldk + h * DH;
#pragma unroll
    for (int i = 0; i < DH; i += 4) {
      const float4 fk = *reinterpret_cast<const float4*>(kp + i);
      const float4 fv = *reinterpret_cast<const float4*>(vp + i);
      kr[i] = fk.x * scale; kr[i + 1] = fk.y * scale; kr[i + 2] = fk.z * scale; kr[i + 3] = fk.w * scale;
      vr[i] = fv.x; vr[i + 1] = fv.y; vr[i + 2] = fv.z; vr[i + 3] = fv.w;
    }
  }
#pragma unroll
  for (int i = 0; i < DH; ++i) accv[i] = acck[i] = 0.f;
  const int i_begin = causal ? (blockIdx.x * static_cast<int>(blockDim.x)) / QT * QT : 0;  // queries t >= j only
  for (int i0 = i_begin; i0 < Tq; i0 += QT) {
    __syncthreads();
    const int nq = min(QT, Tq - i0);
    for (int e = threadIdx.x; e < QT * (DH / 4); e += blockDim.x) {
      const int r = e / (DH / 4), c = (e % (DH / 4)) * 4;
      float4 fq = make_float4(0.f, 0.f, 0.f, 0.f), fd = fq;
      if (r < nq) {
        const long long row = static_cast<long long>(b) * Tq + i0 + r;
        fq = *reinterpret_cast<const float4*>(q + row * ldq + h * DH + c);
        fd = *reinterpret_cast<const float4*>(dO + row * ldo + h * DH + c);
      }
      *reinterpret_cast<float4*>(&qs[r][c]) = fq;
      *reinterpret_cast<float4*>(&ds_[r][c]) = fd;
    }
    for (int e = threadIdx.x; e < QT; e += blockDim.x) {
      const long long idx = (static_cast<long long>(b) * H + h) * Tq + i0 + e;
      ls[e] = (e < nq) ? lse[idx] : 0.f;
      dls[e] = (e < nq) ? delta[idx] : 0.f;
    }
    __syncthreads();
    if (!active || dead) continue;
    for (int r = 0; r < nq; ++r) {
      if (causal && j > (i0 + r)) continue;
      float s = 0.f, dpv = 0.f;
#pragma unroll
      for (int i = 0; i < DH; i += 4) {
        const float4 fq = *reinterpret_cast<const float4*>(&qs[r][i]);
        const float4 fd = *reinterpret_cast<const float4*>(&ds_[r][i]);
        s = fmaf(fq.x, kr[i], s); s = fmaf(fq.y, kr[i + 1], s); s = fmaf(fq.z, kr[i + 2], s); s = fmaf(fq.w, kr[i + 3], s);
        dpv = fmaf(fd.x, vr[i], dpv); dpv = fmaf(fd.y, vr[i + 1], dpv);
        dpv = fmaf(fd.z, vr[i + 2], dpv); dpv = fmaf(fd.w, vr[i + 3], dpv);
      }
      const float p = __expf(s - ls[r]);
      float pm = p;  // P o M' (what multiplied V in the forward)
      if (da.thresh24) {
        const float mk = drop_keep(dstream, (dhead + i0 + r) * Tk + j, da.thresh24) ? da.inv_keep : 0.f;
        pm = p * mk;
        dpv *= mk;
      }
      const float dsv = p * (dpv - dls[r]) * scale;
#pragma unroll
      for (int i = 0; i < DH; i += 4) {
        const float4 fq = *reinterpret_cast<const float4*>(&qs[r][i]);
        const float4 fd = *reinterpret_cast<const float4*>(&ds_[r][i]);
        accv[i] = fmaf(pm, fd.x, accv[i]); accv[i + 1] = fmaf(pm, fd.y, accv[i + 1]);
        accv[i + 2] = fmaf(pm, fd.z, accv[i + 2]); accv[i + 3] = fmaf(pm, fd.w, accv[i + 3]);
        acck[i] = fmaf(dsv, fq.x, acck[i]); acck[i + 1] = fmaf(dsv, fq.y, acck[i + 1]);
        acck[i + 2] = fmaf(dsv, fq.z, acck[i + 2]); acck[i + 3] = fmaf(dsv, fq.w, acck[i + 3]);
      }
    }
  }
  if (active) {
    float* ov = dv + (static_cast<long long>(b) * Tk + j) * lddk + h * DH;
    float* ok = dk + (static_cast<long long>(b) * Tk + j) * lddk + h * DH;
#pragma unroll
    for (int i = 0; i < DH; i += 4) {
      *reinterpret_cast<float4*>(ov + i) = make_float4(accv[i], accv[i + 1], accv[i + 2], accv[i + 3]);
      *reinterpret_cast<float4*>(ok + i) = make_float4(acck[i], acck[i + 1], acck[i + 2], acck[i + 3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Cross-entropy (label smoothing) backward: dlogits = (softmax - ((1-eps)*onehot + eps/V)) / n_valid,
// zero for ignored rows.  One warp per row; n_valid read from the forward's workspace sum.
// ------------------------------------------------------------------------------------------------
__global__ void ce_bwd_kernel(const float* __restrict__ logits, int ldl, const long long* __restrict__ tgt, int M, int V,
                              float eps, long long ignore, const float* __restrict__ row_valid, float scale,
                              float* __restrict__ dlogits, int ldd) {
  __shared__ float nvalid_s;
  if (threadIdx.x == 0) {
    float n = 0.f;
    for (int i = 0; i < M; ++i) n += row_valid[i];
    nvalid_s = n;
  }
  __syncthreads();
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const float* x = logits + static_cast<long long>(row) * ldl;
  float* d = dlogits + static_cast<long long>(row) * ldd;
  const long long t = tgt[row];
  if (t == ignore) {
    for (int c = lane; c < V; c += 32) d[c] = 0.f;
    return;
  }
  float mx = -INFINITY;
  for (int c = lane; c < V; c += 32) mx = fmaxf(mx, x[c]);
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  float se = 0.f;
  for (int c = lane; c < V; c += 32) se += expf(x[c] - mx);
  se = t_warp_sum(se);
  const float inv = scale / nvalid_s;
  for (int c = lane; c < V; c += 32) {
    const float p = expf(x[c] - mx) / se;
    d[c] = (p - ((c == t ? (1.f - eps) : 0.f) + eps / V)) * inv;
  }
}

// ------------------------------------------------------------------------------------------------
// Global gradient norm + fused clip + AdamW (train.py:450-454; torch.optim.AdamW semantics).
// ------------------------------------------------------------------------------------------------
__global__ void sqnorm_partial_kernel(const float* __restrict__ g, long long n, float* __restrict__ part) {
  __shared__ float red[32];
  float acc = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    acc = fmaf(g[i], g[i], acc);
  acc = t_warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = t_warp_sum(v);
    if (threadIdx.x == 0) part[blockIdx.x] = v;
  }
}
__global__ void sqnorm_final_kernel(const float* __restrict__ part, int n, float* __restrict__ out_norm) {
  __shared__ float red[32];
  float acc = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += part[i];
  acc = t_warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
    v = t_warp_sum(v);
    if (threadIdx.x == 0) out_norm[0] = sqrtf(v);
  }
}
// p, g, m, v: contiguous range of one (lr, weight_decay) group.  clip = min(1, max_norm / (norm + 1e-6))
// (torch.nn.utils.clip_grad_norm_); decoupled weight decay then Adam with bias correction.
__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, long long n, const float* __restrict__ norm, float max_norm,
                             float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2,
                             const float* __restrict__ dyn) {
  if (dyn) {  // per-step scalars from device memory (CUDA-graph replay): lr scale, bias corrections
    lr *= dyn[0];
    bc1 = dyn[1];
    bc2 = dyn[2];
  }
  const float nrm = norm ? norm[0] : 0.f;
  const float clip = (norm && max_norm > 0.f) ? fminf(1.f, max_norm / (nrm + 1e-6f)) : 1.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float gi = g[i] * clip;
    float pi = p[i];
    pi -= lr * wd * pi;
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
    pi -= (lr / bc1) * (mi / denom);
    p[i] = pi;
  }
}


// dy *= (y > 0) in place (ReLU backward; y = saved post-activation, split hi plane sign is enough)
__global__ void relu_bwd_kernel(float* __restrict__ dy, const __nv_bfloat16* __restrict__ y_hi, long long total) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    if (!(__bfloat162float(y_hi[i]) > 0.f)) dy[i] = 0.f;
}
// GELU (erf) forward on fp32 pre-activations -> split; backward dz = dy * gelu'(z) in place.
__global__ void gelu_fwd_kernel(const float* __restrict__ z, long long total, __nv_bfloat16* __restrict__ out,
                                long long plane) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float x = z[i];
    t_store_split(out, plane, i, 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)));
  }
}
__global__ void gelu_bwd_kernel(float* __restrict__ dy, const float* __restrict__ z, long long total) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float x = z[i];
    const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
    dy[i] *= cdf + x * pdf;
  }
}
// a += alpha * b
__global__ void axpy_kernel(float* __restrict__ a, const float* __restrict__ b, float alpha, long long total) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    a[i] = fmaf(alpha, b[i], a[i]);
}
// Backward of rows_affine / concatenation: dst[r, :] (+)= scale * src[map(r), :],  map(r) = (r/rpg)*gs + go + r%rpg
__global__ void rows_gather_kernel(const float* __restrict__ src, long long src_ld, int M, int D, float scale, int rpg,
                                   int gs, int go, float* __restrict__ dst, int accumulate) {
  const long long total = static_cast<long long>(M) * D;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % D);
    const int r = static_cast<int>(i / D);
    const long long srow = static_cast<long long>(r / rpg) * gs + go + r % rpg;
    const float v = scale * src[srow * src_ld + c];
    dst[i] = accumulate ? dst[i] + v : v;
  }
}
// Embedding backward: demb[tok[r], :] += scale * dy[r, :].  Deterministic gather form (no atomics): one block per vocabulary
// row t, thread per column; the block walks the B*S token ids (broadcast loads) and accumulates the matching dy rows in order.
__global__ void embed_bwd_kernel(const long long* __restrict__ tok, long long tok_ld, int tok_col, int Bn, int S,
                                 const float* __restrict__ dy, int D, float scale, float* __restrict__ demb) {
  const long long t = blockIdx.x;
  const int rows = Bn * S;
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float acc = 0.f;
    bool any = false;
    for (int r = 0; r < rows; ++r) {
      if (tok[(r / S) * tok_ld + tok_col + (r % S)] == t) {
        acc += dy[static_cast<long long>(r) * D + c];
        any = true;
      }
    }
    if (any) demb[t * D + c] += scale * acc;
  }
}

static inline int t_grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace ralf

using namespace ralf;
#define ST(s) reinterpret_cast<cudaStream_t>(s)
#define BF(p) reinterpret_cast<__nv_bfloat16*>(p)
#define CBF(p) reinterpret_cast<const __nv_bfloat16*>(p)

extern "C" int ralf_transpose_to_split(const float* in_f32, const void* in_split, long long in_plane, long long ld_in,
                                       int R, int C, void* out, long long out_plane, long long ld_out, void* stream) {
  if ((!in_f32 && !in_split) || !out) return RALF_ERR_NULL;
  if (R <= 0 || C <= 0 || ld_out < R) return RALF_ERR_SHAPE;
  dim3 grid((C + 31) / 32, (R + 31) / 32), block(32, 8);
  transpose_to_split_kernel<<<grid, block, 0, ST(stream)>>>(in_f32, CBF(in_split), in_plane, ld_in, R, C, BF(out),
                                                           out_plane, ld_out);
  return set_cuda_error(cudaGetLastError());
}

extern "C" int ralf_split_and_transpose(const float* in, long long ld_in, int R, int C, void* out, long long out_plane,
                                        long long out_ld, void* outT, long long outT_plane, long long outT_ld, void* stream) {
  if (!in || !out || !outT) return RALF_ERR_NULL;
  if (R <= 0 || C <= 0 || out_ld < C || outT_ld < R) return RALF_ERR_SHAPE;
  dim3 grid((C + 31) / 32, (R + 31) / 32), block(32, 8);
  split_and_transpose_kernel<<<grid, block, 0, ST(stream)>>>(in, ld_in, R, C, BF(out), out_plane, out_ld, BF(outT), outT_plane,
                                                            outT_ld);
  return set_cuda_error(cudaGetLastError());
}

static_assert(sizeof(RefreshTask) == sizeof(RalfRefreshTask), "RalfRefreshTask layout");
extern "C" int ralf_refresh_operands(const RalfRefreshTask* tasks_dev, int ntasks, int total_tiles, void* stream) {
  if (!tasks_dev) return RALF_ERR_NULL;
  if (ntasks <= 0 || total_tiles <= 0) return RALF_ERR_SHAPE;
  refresh_operands_kernel<<<total_tiles, dim3(32, 8), 0, ST(stream)>>>(reinterpret_cast<const RefreshTask*>(tasks_dev), ntasks);
  return set_cuda_error(cudaGetLastError());
}

extern "C" int ralf_to_split(const float* in, long long total, void* out, long long out_plane, void* stream) {
  if (!in || !out) return RALF_ERR_NULL;
  if (total <= 0) return RALF_ERR_SHAPE;
  to_split_kernel<<<t_grid_for(total, 256), 256, 0, ST(stream)>>>(in, total, BF(out), out_plane);
  return set_cuda_error(cudaGetLastError());
}

// Tall matrices (M = batch x tokens or batch x pixels): the single-pass kernel above has only C/32 CTAs (57 us per call,
// 11 % of the round-2 training step, profiles/r2_train_slice_summary.md).  Two deterministic passes instead: slabs of
// 512 rows reduced by (C/32 x slabs) CTAs into the workspace, then summed in slab order.
__global__ void colsum_slab_kernel(const float* __restrict__ in, long long ld, int M, int C, int rows_per_slab,
                                   float* __restrict__ part) {
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int r_lo = blockIdx.y * rows_per_slab, r_hi = min(M, r_lo + rows_per_slab);
  float acc = 0.f;
  if (c < C)
    for (int r = r_lo + threadIdx.y; r < r_hi; r += 8) acc += in[static_cast<long long>(r) * ld + c];
  red[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += red[i][threadIdx.x];
    part[static_cast<long long>(blockIdx.y) * C + c] = s;
  }
}
__global__ void colsum_final_kernel(const float* __restrict__ part, int nslabs, int C, float* __restrict__ out, int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.f;
  for (int b = 0; b < nslabs; ++b) s += part[static_cast<long long>(b) * C + c];
  out[c] = accumulate ? out[c] + s : s;
}
constexpr int COLSUM_SLAB = 512;

extern "C" size_t ralf_colsum_workspace_bytes(int M, int C) {
  if (M <= 2 * COLSUM_SLAB || C <= 0) return 0;
  return static_cast<size_t>((M + COLSUM_SLAB - 1) / COLSUM_SLAB) * C * sizeof(float);
}
extern "C" int ralf_colsum(const float* in, long long ld, int M, int C, float* out, int accumulate, float* workspace,
                           void* stream) {
  if (!in || !out) return RALF_ERR_NULL;
  if (M <= 0 || C <= 0) return RALF_ERR_SHAPE;
  if (workspace && M > 2 * COLSUM_SLAB) {
    const int nslabs = (M + COLSUM_SLAB - 1) / COLSUM_SLAB;
    colsum_slab_kernel<<<dim3((C + 31) / 32, nslabs), dim3(32, 8), 0, ST(stream)>>>(in, ld, M, C, COLSUM_SLAB, workspace);
    colsum_final_kernel<<<(C + 127) / 128, 128, 0, ST(stream)>>>(workspace, nslabs, C, out, accumulate);
  } else {
    colsum_kernel<<<(C + 31) / 32, dim3(32, 8), 0, ST(stream)>>>(in, ld, M, C, out, accumulate);
  }
  return set_cuda_error(cudaGetLastError());
}

extern "C" int ralf_layernorm_bwd(const float* x, long long x_ld, const float* dy, const float* gamma, float eps, int M,
                                  int D, const float* add_to, float* dx, float* dgamma, float* dbeta,
                                  float* workspace /* 2*D*nblocks floats, nblocks = min(M/8+1, 4*SMs) */,
                                  void* stream) {
  if (!x || !dy || !gamma || !dx || !dgamma || !dbeta || !workspace) return RALF_ERR_NULL;
  if (M <= 0 || D <= 0 || D > 1024 || (D & 31)) return RALF_ERR_SHAPE;
  int nblocks = (M + 7) / 8;
  const int cap = num_sms() * 4;
  if (nblocks > cap) nblocks = cap;
  ln_bwd_kernel<<<nblocks, 256, 8 * 2 * D * sizeof(float), ST(stream)>>>(x, x_ld, dy, gamma, eps, M, D, add_to, dx, workspace);
  ln_bwd_reduce_kernel<<<(2 * D + 255) / 256, 256, 0, ST(stream)>>>(workspace, nblocks, D, dgamma, dbeta);
  return set_cuda_error(cudaGetLastError());
}

extern "C" int ralf_ce_label_smooth_bwd(const float* logits, int ldl, const long long* targets, int M, int V, float eps,
                                        long long ignore_index, const float* fwd_workspace, float scale, float* dlogits,
                                        int ldd, void* stream) {
  if (!logits || !targets || !fwd_workspace || !dlogits) return RALF_ERR_NULL;
  if (M <= 0 || V <= 0) return RALF_ERR_SHAPE;
  ce_bwd_kernel<<<(M + 7) / 8, 256, 0, ST(stream)>>>(logits, ldl, targets, M, V, eps, ignore_index, fwd_workspace + M,
                                                    scale, dlogits, ldd);
  return set_cuda_error(cudaGetLastError());
}

extern "C" int ralf_grad_norm(const float* grads, long long n, float* workspace /* 1024 floats */, float* out_norm,
                              void* stream) {
  if (!grads || !workspace || !out_norm) return RALF_ERR_NULL;
  if (n <= 0) return RALF_ERR_SHAPE;
  int blocks = t_grid_for(n, 256);
  if (blocks > 1024) blocks = 1024;
  sqnorm_partial_kernel<<<blocks, 256, 0, ST(stream)>>>(grads, n, workspace);
  sqnorm_final_kernel<<<1, 256, 0, ST(stream)>>>(workspace, blocks, out_norm);
  return set_cuda_error(cudaGetLastError());
}

extern "C" int ralf_adamw_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                               const float* grad_norm, float max_norm, float lr, float beta1, float beta2, float eps,
                               float weight_decay, int step, void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq) return RALF_ERR_NULL;
  if (n <= 0 || step <= 0) return RALF_ERR_SHAPE;
  const float bc1 = 1.f - powf(beta1, static_cast<float>(step));
  const float bc2 = 1.f - powf(beta2, static_cast<float>(step));
  adamw_kernel<<<t_grid_for(n, 256), 256, 0, ST(stream)>>>(params, grads, exp_avg, exp_avg_sq, n, grad_norm, max_norm, lr,
                                                         beta1, beta2, eps, weight_decay, bc1, bc2, nullptr);
  return set_cuda_error(cudaGetLastError());
}

// Same step with the per-step scalars read from device memory, so a captured CUDA graph of the training step can be
// replayed: dyn = {lr scale (scheduler), 1 - beta1^t, 1 - beta2^t}, written by the host before each replay.
extern "C" int ralf_adamw_step_dyn(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                                   const float* grad_norm, float max_norm, float lr, float beta1, float beta2, float eps,
                                   float weight_decay, const float* dyn, void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || !dyn) return RALF_ERR_NULL;
  if (n <= 0) return RALF_ERR_SHAPE;
  adamw_kernel<<<t_grid_for(n, 256), 256, 0, ST(stream)>>>(params, grads, exp_avg, exp_avg_sq, n, grad_norm, max_norm, lr,
                                                         beta1, beta2, eps, weight_decay, 1.f, 1.f, dyn);
  return set_cuda_error(cudaGetLastError());
}

static int attention_bwd_impl(const float* q, int ldq, const float* k, const float* v, int ldk,
                                  const unsigned char* key_padding_mask, int B, int H, int Tq, int Tk, int head_dim,
                                  int causal, float scale, const void* o_split, long long o_plane, const float* dO,
                                  int ldo, float* lse_ws, float* delta_ws, float* dq, int lddq, float* dk, float* dv,
                                  int lddk, DropArgs da, void* stream, int lse_given = 0) {
  float* lse = lse_ws;
  if (!q || !k || !v || !o_split || !dO || !lse || !delta_ws || !dq || !dk || !dv) return RALF_ERR_NULL;
  if (B <= 0 || H <= 0 || Tq <= 0 || Tk <= 0 || (head_dim != 32 && head_dim != 64)) return RALF_ERR_SHAPE;
  if ((ldq & 3) || (ldk & 3) || (ldo & 3) || (lddk & 3) || (lddq & 3)) return RALF_ERR_ALIGN;
  const int tq = Tq >= 128 ? 128 : ((Tq + 31) / 32) * 32;
  const int tk = Tk >= 128 ? 128 : ((Tk + 31) / 32) * 32;
  dim3 g1((Tq + tq - 1) / tq, H, B), g2((Tk + tk - 1) / tk, H, B);
  if (head_dim == 32) {
    attn_bwd_dq_kernel<32><<<g1, tq, 0, ST(stream)>>>(q, ldq, k, v, ldk, key_padding_mask, Tq, Tk, causal, scale,
                                                     CBF(o_split), o_plane, dO, ldo, lse, delta_ws, dq, lddq, da, lse_given);
    attn_bwd_dkv_kernel<32><<<g2, tk, 0, ST(stream)>>>(q, ldq, k, v, ldk, key_padding_mask, Tq, Tk, causal, scale, dO, ldo,
                                                      lse, delta_ws, dk, dv, lddk, da);
  } else {
    attn_bwd_dq_kernel<64><<<g1, tq, 0, ST(stream)>>>(q, ldq, k, v, ldk, key_padding_mask, Tq, Tk, causal, scale,
                                                     CBF(o_split), o_plane, dO, ldo, lse, delta_ws, dq, lddq, da, lse_given);
    attn_bwd_dkv_kernel<64><<<g2, tk, 0, ST(stream)>>>(q, ldq, k, v, ldk, key_padding_mask, Tq, Tk, causal, scale, dO, ldo,
                                                      lse, delta_ws, dk, dv, lddk, da);
  }
  return set_cuda_error(cudaGetLastError());
}

extern "C" int ralf_attention_bwd(const float* q, int ldq, const float* k, const float* v, int ldk,
                                  const unsigned char* key_padding_mask, int B, int H, int Tq, int Tk, int head_dim,
                                  int causal, float scale, const void* o_split, long long o_plane, const float* dO,
                                  int ldo, float* lse_ws, float* delta_ws, float* dq, int lddq, float* dk, float* dv,
                                  int lddk, void* stream) {
  return attention_bwd_impl(q, ldq, k, v, ldk, key_padding_mask, B, H, Tq, Tk, head_dim, causal, scale, o_split, o_plane, dO,
                            ldo, lse_ws, delta_ws, dq, lddq, dk, dv, lddk, make_drop_args(nullptr, 0, 0.f), stream);
}

extern "C" int ralf_attention_bwd_dropout(const float* q, int ldq, const float* k, const float* v, int ldk,
                                          const unsigned char* key_padding_mask, int B, int H, int Tq, int Tk,
                                          int head_dim, int causal, float scale, const void* o_split, long long o_plane,
                                          const float* dO, int ldo, float* lse_ws, float* delta_ws, float* dq, int lddq,
                                          float* dk, float* dv, int lddk, const unsigned long long* seed,
                                          unsigned int site, float p, int lse_given, void* stream) {
  if (!seed) return RALF_ERR_NULL;
  if (!(p >= 0.f && p < 1.f)) return RALF_ERR_SHAPE;
  return attention_bwd_impl(q, ldq, k, v, ldk, key_padding_mask, B, H, Tq, Tk, head_dim, causal, scale, o_split, o_plane, dO,
                            ldo, lse_ws, delta_ws, dq, lddq, dk, dv, lddk, make_drop_args(seed, site, p), stream, lse_given);
}

// ------------------------------------------------------------------------------------------------
// Elementwise dropout (nn.Dropout(p) in training; same op for the backward applied to the gradient):
//   y[i] = (res ? res[i] : 0) + (keep(i) ? x[i] / (1-p) : 0),  i = linear element index of the [M, C] tensor.
// x is fp32 or split bf16 (hi + lo); y is written as fp32 and / or split.  In-place use (out == in) is fine.
// ------------------------------------------------------------------------------------------------
__global__ void dropout_kernel(const float* __restrict__ in_f32, const __nv_bfloat16* __restrict__ in_split,
                               long long in_plane, const float* __restrict__ res, long long total, DropArgs da,
                               float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_split, long long out_plane) {
  const unsigned long long dstream = da.thresh24 ? drop_stream(*da.seed, da.site) : 0ull;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float x = in_f32 ? in_f32[i] : (__bfloat162float(in_split[i]) + __bfloat162float(in_split[in_plane + i]));
    if (da.thresh24) x = drop_keep(dstream, static_cast<unsigned long long>(i), da.thresh24) ? x * da.inv_keep : 0.f;
    if (res) x += res[i];
    if (out_f32) out_f32[i] = x;
    if (out_split) t_store_split(out_split, out_plane, i, x);
  }
}

__global__ void dropout_mask_kernel(DropArgs da, long long total, unsigned char* __restrict__ out) {
  const unsigned long long dstream = drop_stream(*da.seed, da.site);
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    out[i] = drop_keep(dstream, static_cast<unsigned long long>(i), da.thresh24) ? 1 : 0;
}

extern "C" int ralf_dropout(const float* in_f32, const void* in_split, long long in_plane, const float* res,
                            long long total, const unsigned long long* seed, unsigned int site, float p,
                            float* out_f32, void* out_split, long long out_plane, void* stream) {
  if ((!in_f32 && !in_split) || (!out_f32 && !out_split) || !seed) return RALF_ERR_NULL;
  if (total <= 0 || !(p >= 0.f && p < 1.f)) return RALF_ERR_SHAPE;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148LL * 16));
  dropout_kernel<<<blocks, 256, 0, ST(stream)>>>(in_f32, CBF(in_split), in_plane, res, total, make_drop_args(seed, site, p),
                                                out_f32, BF(out_split), out_plane);
  return set_cuda_error(cudaGetLastError());
}

/* Test / debugging aid: the keep mask (1 = kept) the kernels derive for elements [0, total) of a site. */
extern "C" int ralf_dropout_mask(const unsigned long long* seed, unsigned int site, float p, long long total,
                                 unsigned char* out, void* stream) {
  if (!seed || !out) return RALF_ERR_NULL;
  if (total <= 0 || !(p > 0.f && p < 1.f)) return RALF_ERR_SHAPE;
  const int blocks = static_cast<int>(std::min<long long>((total + 255) / 256, 148LL * 16));
  dropout_mask_kernel<<<blocks, 256, 0, ST(stream)>>>(make_drop_args(seed, site, p), total, out);
  return set_cuda_error(cudaGetLastError());
}

extern "C" int ralf_relu_bwd(float* dy, const void* y_split_hi, long long total, void* stream) {
  if (!dy || !y_split_hi) return RALF_ERR_NULL;
  if (total <= 0) return RALF_ERR_SHAPE;
  relu_bwd_kernel<<<t_grid_for(total, 256), 256, 0, ST(stream)>>>(dy, CBF(y_split_hi), total);
  return set_cuda_error(cudaGetLastError());
}
extern "C" int ralf_gelu_fwd(const float* z, long long total, void* out_split, long long out_plane, void* stream) {
  if (!z || !out_split) return RALF_ERR_NULL;
  if (total <= 0) return RALF_ERR_SHAPE;
  gelu_fwd_kernel<<<t_grid_for(total, 256), 256, 0, ST(stream)>>>(z, total, BF(out_split), out_plane);
  return set_cuda_error(cudaGetLastError());
}
extern "C" int ralf_gelu_bwd(float* dy, const float* z, long long total, void* stream) {
  if (!dy || !z) return RALF_ERR_NULL;
  if (total <= 0) return RALF_ERR_SHAPE;
  gelu_bwd_kernel<<<t_grid_for(total, 256), 256, 0, ST(stream)>>>(dy, z, total);
  return set_cuda_error(cudaGetLastError());
}
extern "C" int ralf_axpy(float* a, const float* b, float alpha, long long total, void* stream) {
  if (!a || !b) return RALF_ERR_NULL;
  if (total <= 0) return RALF_ERR_SHAPE;
  axpy_kernel<<<t_grid_for(total, 256), 256, 0, ST(stream)>>>(a, b, alpha, total);
  return set_cuda_error(cudaGetLastError());
}
extern "C" int ralf_rows_gather(const float* src, long long src_ld, int M, int D, float scale, int rows_per_group,
                                int group_stride, int group_offset, float* dst, int accumulate, void* stream) {
  if (!src || !dst) return RALF_ERR_NULL;
  if (M <= 0 || D <= 0) return RALF_ERR_SHAPE;
  const int rpg = rows_per_group > 0 ? rows_per_group : M;
  rows_gather_kernel<<<t_grid_for(static_cast<long long>(M) * D, 256), 256, 0, ST(stream)>>>(
      src, src_ld, M, D, scale, rpg, rows_per_group > 0 ? group_stride : 0, rows_per_group > 0 ? group_offset : 0, dst,
      accumulate);
  return set_cuda_error(cudaGetLastError());
}
extern "C" int ralf_embed_bwd(const long long* tok, long long tok_ld, int tok_col, int B, int S, const float* dy, int D,
                              float scale, float* demb, int V, void* stream) {
  if (!tok || !dy || !demb) return RALF_ERR_NULL;
  if (B <= 0 || S <= 0 || D <= 0 || V <= 0) return RALF_ERR_SHAPE;
  embed_bwd_kernel<<<V, D < 256 ? D : 256, 0, ST(stream)>>>(tok, tok_ld, tok_col, B, S, dy, D, scale, demb);
  return set_cuda_error(cudaGetLastError());
}
