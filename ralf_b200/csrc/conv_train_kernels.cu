// Training kernels for the ResNet50-FPN trunk: BatchNorm in training mode (batch
// statistics, running-stat update), its backward, col2im for the 3x3 / strided convolutions, max-pool and
// nearest-upsample backward, and the weight-layout permutes between nn.Conv2d's [Cout, Cin, KH, KW] and the
// GEMM's [Cout, (kh, kw, cin)].  Convolutions themselves are im2col + the tcgen05 GEMM (forward),
// dW = dZ^T . Xcol and dXcol = dZ . W (backward) -- see ralf_b200/train_conv.py.
#include <float.h>
#include <math.h>

#include <stdlib.h>

#include "common.cuh"
#include "ralf_internal.h"

namespace ralf {

__device__ __forceinline__ void c_store_split(__nv_bfloat16* hi_plane, long long plane, long long off, float x) {
  __nv_bfloat16 h, l;
  split_bf16(x, h, l);
  hi_plane[off] = h;
  hi_plane[plane + off] = l;
}
__device__ __forceinline__ float c_load_split(const __nv_bfloat16* hi_plane, long long plane, long long off) {
  return __bfloat162float(hi_plane[off]) + __bfloat162float(hi_plane[plane + off]);
}

// ------------------------------------------------------------------------------------------------
// Column statistics of fp32 [M, C] matrices: part[blk][which][c] partial sums over a row slab.
//   mode 0: sum(a), sum(a*a)                       (BatchNorm forward statistics)
//   mode 1: sum(a), sum(a * (z - mean) * rstd)     (BatchNorm backward: a = dY, needs z / mean / rstd)
// grid = (ceil(C/32), nslabs), block = (32, 8).  A second kernel reduces the slabs in fixed order.
// ------------------------------------------------------------------------------------------------
__global__ void colstats_partial_kernel(const float* __restrict__ a, const float* __restrict__ z,
                                        const float* __restrict__ mean, const float* __restrict__ rstd, int mode, int M,
                                        int C, int rows_per_slab, float* __restrict__ part) {
  __shared__ float r0[8][33], r1[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int row_lo = blockIdx.y * rows_per_slab;
  const int row_hi = min(M, row_lo + rows_per_slab);
  float s0 = 0.f, s1 = 0.f;
  if (c < C) {
    const float mu = mode ? mean[c] : 0.f, rs = mode ? rstd[c] : 0.f;
    for (int r = row_lo + threadIdx.y; r < row_hi; r += 8) {
      const float v = a[static_cast<long long>(r) * C + c];
      s0 += v;
      s1 += mode ? v * (z[static_cast<long long>(r) * C + c] - mu) * rs : v * v;
    }
  }
  r0[threadIdx.y][threadIdx.x] = s0;
  r1[threadIdx.y][threadIdx.x] = s1;
  __syncthreads();
  if (threadIdx.y == 0 && c < C) {
    float t0 = 0.f, t1 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { t0 += r0[i][threadIdx.x]; t1 += r1[i][threadIdx.x]; }
    part[(static_cast<long long>(blockIdx.y) * 2 + 0) * C + c] = t0;
    part[(static_cast<long long>(blockIdx.y) * 2 + 1) * C + c] = t1;
  }
}
// mode 0: mean, rstd (biased variance) + running-stat update (momentum; unbiased variance like torch).
// mode 1: writes the two sums (dbeta, dgamma).
__global__ void colstats_final_kernel(const float* __restrict__ part, int nslabs, int C, int M, int mode, float eps,
                                      float momentum, float* __restrict__ out0, float* __restrict__ out1,
                                      float* __restrict__ running_mean, float* __restrict__ running_var) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s0 = 0.0, s1 = 0.0;
  for (int b = 0; b < nslabs; ++b) {
    s0 += part[(static_cast<long long>(b) * 2 + 0) * C + c];
    s1 += part[(static_cast<long long>(b) * 2 + 1) * C + c];
  }
  if (mode == 0) {
    const double mu = s0 / M;
    double var = s1 / M - mu * mu;
    if (var < 0.0) var = 0.0;
    out0[c] = static_cast<float>(mu);
    out1[c] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    if (running_mean) {
      running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * static_cast<float>(mu);
      const double unb = M > 1 ? var * M / (M - 1) : var;
      running_var[c] = (1.f - momentum) * running_var[c] + momentum * static_cast<float>(unb);
    }
  } else {
    out0[c] = static_cast<float>(s0);
    out1[c] = static_cast<float>(s1);
  }
}

// BatchNorm(train) apply: y = [relu]( gamma*(z-mean)*rstd + beta (+ residual) ) -> split (and/or fp32)
__global__ void bn_apply_kernel(const float* __restrict__ z, const float* __restrict__ mean, const float* __restrict__ rstd,
                                const float* __restrict__ gamma, const float* __restrict__ beta,
                                const __nv_bfloat16* __restrict__ res, long long res_plane, int relu, long long total, int C,
                                __nv_bfloat16* __restrict__ out_split, long long out_plane, float* __restrict__ out_f32) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    float y = gamma[c] * (z[i] - mean[c]) * rstd[c] + beta[c];
    if (res) y += c_load_split(res, res_plane, i);
    if (relu) y = fmaxf(y, 0.f);
    if (out_split) c_store_split(out_split, out_plane, i, y);
    if (out_f32) out_f32[i] = y;
  }
}
// dY masked by the ReLU of the saved output (split hi-plane sign), in place; also used for the residual branch.
__global__ void relu_mask_kernel(float* __restrict__ dy, const __nv_bfloat16* __restrict__ y_hi, long long total) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    if (!(__bfloat162float(y_hi[i]) > 0.f)) dy[i] = 0.f;
}
// BatchNorm backward apply: dz = gamma*rstd * (dy - sum_dy/M - xhat * sum_dy_xhat/M)
__global__ void bn_bwd_apply_kernel(const float* __restrict__ dy, const float* __restrict__ z,
                                    const float* __restrict__ mean, const float* __restrict__ rstd,
                                    const float* __restrict__ gamma, const float* __restrict__ sum_dy,
                                    const float* __restrict__ sum_dy_xhat, long long total, int C, int M,
                                    float* __restrict__ dz) {
  const float invM = 1.f / static_cast<float>(M);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const float xh = (z[i] - mean[c]) * rstd[c];
    dz[i] = gamma[c] * rstd[c] * (dy[i] - sum_dy[c] * invM - xh * sum_dy_xhat[c] * invM);
  }
}

// col2im (gather form): dx[b,iy,ix,c] (+)= sum over taps of dcol[(b,oy,ox), (kh*KW+kw)*C + c]
__global__ void col2im_kernel(const float* __restrict__ dcol, int B, int H, int W, int C, int KH, int KW, int stride,
                              int pad, int Ho, int Wo, float* __restrict__ dx, int accumulate) {
  const long long total = static_cast<long long>(B) * H * W * C;
  const long long Kd = static_cast<long long>(KH) * KW * C;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    long long r = i / C;
    const int ix = static_cast<int>(r % W); r /= W;
    const int iy = static_cast<int>(r % H);
    const int b = static_cast<int>(r / H);
    float acc = 0.f;
    for (int kh = 0; kh < KH; ++kh) {
      const int ty = iy + pad - kh;
      if (ty < 0 || ty % stride) continue;
      const int oy = ty / stride;
      if (oy >= Ho) continue;
      for (int kw = 0; kw < KW; ++kw) {
        const int tx = ix + pad - kw;
        if (tx < 0 || tx % stride) continue;
        const int ox = tx / stride;
        if (ox >= Wo) continue;
        acc += dcol[((static_cast<long long>(b) * Ho + oy) * Wo + ox) * Kd + (static_cast<long long>(kh) * KW + kw) * C + c];
      }
    }
    dx[i] = accumulate ? dx[i] + acc : acc;
  }
}

// float4 form (C % 4 == 0, 16-byte aligned rows): four channels per thread -- a quarter of the index arithmetic and
// 16-byte loads (the scalar kernel ran at a third of the HBM roofline, profiles/r2_train.md).
__global__ void col2im_vec4_kernel(const float4* __restrict__ dcol, int B, int H, int W, int C4, int KH, int KW, int stride,
                                   int pad, int Ho, int Wo, float4* __restrict__ dx, int accumulate) {
  const long long total = static_cast<long long>(B) * H * W * C4;
  const long long Kd = static_cast<long long>(KH) * KW * C4;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C4);
    long long r = i / C4;
    const int ix = static_cast<int>(r % W); r /= W;
    const int iy = static_cast<int>(r % H);
    const int b = static_cast<int>(r / H);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int kh = 0; kh < KH; ++kh) {
      const int ty = iy + pad - kh;
      if (ty < 0 || ty % stride) continue;
      const int oy = ty / stride;
      if (oy >= Ho) continue;
      for (int kw = 0; kw < KW; ++kw) {
        const int tx = ix + pad - kw;
        if (tx < 0 || tx % stride) continue;
        const int ox = tx / stride;
        if (ox >= Wo) continue;
        const float4 v = dcol[((static_cast<long long>(b) * Ho + oy) * Wo + ox) * Kd + (static_cast<long long>(kh) * KW + kw) * C4 + c];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;  // same tap order as the scalar kernel: identical sums
      }
    }
    if (accumulate) {
      const float4 o = dx[i];
      acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
    }
    dx[i] = acc;
  }
}

// max-pool 3x3/s2/p1 backward, deterministic (no atomics): (1) per output element the tap (kh*3+kw) of its FIRST maximum --
// torch's rule -- as a byte; (2) gather: every input pixel sums dy over the <= 4 windows whose recorded tap points at it, in a
// fixed order.  dx is written completely (no zero-initialisation needed).
__global__ void maxpool_bwd_arg_kernel(const __nv_bfloat16* __restrict__ x, long long x_plane, int B, int H, int W, int C,
                                       int Ho, int Wo, uint8_t* __restrict__ tap) {
  const long long total = static_cast<long long>(B) * Ho * Wo * C;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const long long pix = i / C;
    const int ox = static_cast<int>(pix % Wo);
    const int oy = static_cast<int>((pix / Wo) % Ho);
    const int b = static_cast<int>(pix / (static_cast<long long>(Wo) * Ho));
    float best = -INFINITY;
    int arg = 255;
    for (int kh = 0; kh < 3; ++kh) {
      const int iy = oy * 2 - 1 + kh;
      if (iy < 0 || iy >= H) continue;
      for (int kw = 0; kw < 3; ++kw) {
        const int ix = ox * 2 - 1 + kw;
        if (ix < 0 || ix >= W) continue;
        const long long src = ((static_cast<long long>(b) * H + iy) * W + ix) * C + c;
        const float v = c_load_split(x, x_plane, src);
        if (v > best) { best = v; arg = kh * 3 + kw; }
      }
    }
    tap[i] = static_cast<uint8_t>(arg);
  }
}
__global__ void maxpool_bwd_gather_kernel(const uint8_t* __restrict__ tap, const float* __restrict__ dy, int B, int H, int W,
                                          int C, int Ho, int Wo, float* __restrict__ dx) {
  const long long total = static_cast<long long>(B) * H * W * C;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const long long pix = i / C;
    const int ix = static_cast<int>(pix % W);
    const int iy = static_cast<int>((pix / W) % H);
    const int b = static_cast<int>(pix / (static_cast<long long>(W) * H));
    float acc = 0.f;
    for (int kh = 0; kh < 3; ++kh) {  // iy = 2*oy - 1 + kh
      const int ty = iy + 1 - kh;
      if (ty < 0 || (ty & 1)) continue;
      const int oy = ty >> 1;
      if (oy >= Ho) continue;
      for (int kw = 0; kw < 3; ++kw) {
        const int tx = ix + 1 - kw;
        if (tx < 0 || (tx & 1)) continue;
        const int ox = tx >> 1;
        if (ox >= Wo) continue;
        const long long o = ((static_cast<long long>(b) * Ho + oy) * Wo + ox) * C + c;
        if (tap[o] == kh * 3 + kw) acc += dy[o];
      }
    }
    dx[i] = acc;
  }
}

// nearest-upsample backward: d_small[b,y5,x5,c] = sum of d_big over the children of (y5,x5)
__global__ void upsample_bwd_kernel(const float* __restrict__ dbig, long long ld_big, int B, int h5, int w5, int h4, int w4,
                                    int C, float* __restrict__ dsmall) {
  const long long total = static_cast<long long>(B) * h5 * w5 * C;
  const float sy = static_cast<float>(h5) / static_cast<float>(h4);
  const float sx = static_cast<float>(w5) / static_cast<float>(w4);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    long long r = i / C;
    const int x5 = static_cast<int>(r % w5); r /= w5;
    const int y5 = static_cast<int>(r % h5);
    const int b = static_cast<int>(r / h5);
    float acc = 0.f;
    for (int y = 0; y < h4; ++y) {
      if (min(static_cast<int>(floorf(static_cast<float>(y) * sy)), h5 - 1) != y5) continue;
      for (int x = 0; x < w4; ++x) {
        if (min(static_cast<int>(floorf(static_cast<float>(x) * sx)), w5 - 1) != x5) continue;
        acc += dbig[((static_cast<long long>(b) * h4 + y) * w4 + x) * ld_big + c];
      }
    }
    dsmall[i] = acc;
  }
}

// nn.Conv2d weight [N, C, T] (T = KH*KW) -> GEMM operands W [N, T*C] split and (optionally) its transpose
// [T*C, Np] split;  and the reverse for gradients: dWg fp32 [N, T*C] -> dW [N, C, T].
__global__ void conv_weight_to_gemm_kernel(const float* __restrict__ w, int N, int C, int T, int Kp,
                                           __nv_bfloat16* __restrict__ out, long long out_plane,
                                           __nv_bfloat16* __restrict__ outT, long long outT_plane, int Np) {
  const long long total = static_cast<long long>(N) * T * C;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % C);
    const int t = static_cast<int>((i / C) % T);
    const int n = static_cast<int>(i / (static_cast<long long>(C) * T));
    const float v = w[(static_cast<long long>(n) * C + c) * T + t];
    const long long k = static_cast<long long>(t) * C + c;
    c_store_split(out, out_plane, static_cast<long long>(n) * Kp + k, v);
    if (outT) c_store_split(outT, outT_plane, k * Np + n, v);
  }
}
__global__ void conv_grad_from_gemm_kernel(const float* __restrict__ dwg, int N, int C, int T, int ldg,
                                           float* __restrict__ dw) {
  const long long total = static_cast<long long>(N) * C * T;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int t = static_cast<int>(i % T);
    const int c = static_cast<int>((i / T) % C);
    const int n = static_cast<int>(i / (static_cast<long long>(C) * T));
    dw[i] = dwg[static_cast<long long>(n) * ldg + static_cast<long long>(t) * C + c];
  }
}

static inline int c_grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  return static_cast<int>(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace ralf

using namespace ralf;
#define ST(s) reinterpret_cast<cudaStream_t>(s)
#define BF(p) reinterpret_cast<__nv_bfloat16*>(p)
#define CBF(p) reinterpret_cast<const __nv_bfloat16*>(p)

// mode 0: BatchNorm statistics of z -> mean, rstd (+ running stats).  mode 1: (sum a, sum a*xhat) for the backward.
// workspace: 2 * C * ceil(M / 256) floats.
extern "C" int ralf_bn_colstats(const float* a, const float* z, const float* mean, const float* rstd, int mode, int M, int C,
                                float eps, float momentum, float* out0, float* out1, float* running_mean,
                                float* running_var, float* workspace, void* stream) {
  if (!a || !out0 || !out1 || !workspace) return RALF_ERR_NULL;
  if (mode && (!z || !mean || !rstd)) return RALF_ERR_NULL;
  if (M <= 0 || C <= 0) return RALF_ERR_SHAPE;
  // 512-row slabs: 4x the CTAs of round 1's 2048-row slabs (the kernel was latency bound: 55 us per call).  The slab size only
  // changes the fp32 summation order; RALF_BN_SLAB (>= 256, the workspace bound) exists because the gradient-parity test on a
  // 2-sample BatchNorm batch is chaotic in that order (median per-tensor max-rel error vs fp64: 1.43e-4 / 1.69e-4 / 2.33e-4 at
  // 512 / 2048 / 256 rows -- ReLU kinks flip on 1e-7 perturbations of the statistics, tests/test_train_gpu.py).
  static const int rows_per_slab = getenv("RALF_BN_SLAB") ? atoi(getenv("RALF_BN_SLAB")) : 512;
  const int nslabs = (M + rows_per_slab - 1) / rows_per_slab;
  colstats_partial_kernel<<<dim3((C + 31) / 32, nslabs), dim3(32, 8), 0, ST(stream)>>>(a, z, mean, rstd, mode, M, C,
                                                                                      rows_per_slab, workspace);
  colstats_final_kernel<<<(C + 127) / 128, 128, 0, ST(stream)>>>(workspace, nslabs, C, M, mode, eps, momentum, out0, out1,
                                                                running_mean, running_var);
  return set_cuda_error(cudaGetLastError());
}

extern "C" int ralf_bn_apply(const float* z, const float* mean, const float* rstd, const float* gamma, const float* beta,
                             const void* res_split, long long res_plane, int relu, int M, int C, void* out_split,
                             long long out_plane, float* out_f32, void* stream) {
  if (!z || !mean || !rstd || !gamma || !beta) return RALF_ERR_NULL;
  if (M <= 0 || C <= 0) return RALF_ERR_SHAPE;
  const long long total = static_cast<long long>(M) * C;
  bn_apply_kernel<<<c_grid_for(total, 256), 256, 0, ST(stream)>>>(z, mean, rstd, gamma, beta, CBF(res_split), res_plane, relu,
                                                                 total, C, BF(out_split), out_plane, out_f32);
  return set_cuda_error(cudaGetLastError());
}

extern "C" int ralf_bn_bwd_apply(const float* dy, const float* z, const float* mean, const float* rstd, const float* gamma,
                                 const float* sum_dy, const float* sum_dy_xhat, int M, int C, float* dz, void* stream) {
  if (!dy || !z || !mean || !rstd || !gamma || !sum_dy || !sum_dy_xhat || !dz) return RALF_ERR_NULL;
  if (M <= 0 || C <= 0) return RALF_ERR_SHAPE;
  const long long total = static_cast<long long>(M) * C;
  bn_bwd_apply_kernel<<<c_grid_for(total, 256), 256, 0, ST(stream)>>>(dy, z, mean, rstd, gamma, sum_dy, sum_dy_xhat, total, C,
                                                                     M, dz);
  return set_cuda_error(cudaGetLastError());
}

extern "C" int ralf_col2im(const float* dcol, int B, int H, int W, int C, int KH, int KW, int stride, int pad, float* dx,
                           int accumulate, void* stream) {
  if (!dcol || !dx) return RALF_ERR_NULL;
  if (B <= 0 || H <= 0 || W <= 0 || C <= 0) return RALF_ERR_SHAPE;
  const int Ho = (H + 2 * pad - KH) / stride + 1, Wo = (W + 2 * pad - KW) / stride + 1;
  const long long total = static_cast<long long>(B) * H * W * C;
  if (C % 4 == 0 && !(reinterpret_cast<uintptr_t>(dcol) & 15) && !(reinterpret_cast<uintptr_t>(dx) & 15)) {
    col2im_vec4_kernel<<<c_grid_for(total / 4, 256), 256, 0, ST(stream)>>>(reinterpret_cast<const float4*>(dcol), B, H, W, C / 4,
                                                                          KH, KW, stride, pad, Ho, Wo,
                                                                          reinterpret_cast<float4*>(dx), accumulate);
    return set_cuda_error(cudaGetLastError());
  }
  col2im_kernel<<<c_grid_for(total, 256), 256, 0, ST(stream)>>>(dcol, B, H, W, C, KH, KW, stride, pad, Ho, Wo, dx, accumulate);
  return set_cuda_error(cudaGetLastError());
}

extern "C" int ralf_maxpool3x3s2_bwd(const void* x_split, long long x_plane, const float* dy, int B, int H, int W, int C,
                                     float* dx, void* workspace /* B*Ho*Wo*C bytes */, void* stream) {
  if (!x_split || !dy || !dx || !workspace) return RALF_ERR_NULL;
  if (B <= 0 || H <= 0 || W <= 0 || C <= 0) return RALF_ERR_SHAPE;
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  const long long total = static_cast<long long>(B) * Ho * Wo * C;
  uint8_t* tap = reinterpret_cast<uint8_t*>(workspace);
  maxpool_bwd_arg_kernel<<<c_grid_for(total, 256), 256, 0, ST(stream)>>>(CBF(x_split), x_plane, B, H, W, C, Ho, Wo, tap);
  maxpool_bwd_gather_kernel<<<c_grid_for(static_cast<long long>(B) * H * W * C, 256), 256, 0, ST(stream)>>>(tap, dy, B, H, W, C,
                                                                                                           Ho, Wo, dx);
  return set_cuda_error(cudaGetLastError());
}

extern "C" int ralf_upsample_nearest_bwd(const float* dbig, long long ld_big, int B, int h5, int w5, int h4, int w4, int C,
                                         float* dsmall, void* stream) {
  if (!dbig || !dsmall) return RALF_ERR_NULL;
  if (B <= 0 || h5 <= 0 || w5 <= 0 || h4 <= 0 || w4 <= 0 || C <= 0) return RALF_ERR_SHAPE;
  const long long total = static_cast<long long>(B) * h5 * w5 * C;
  upsample_bwd_kernel<<<c_grid_for(total, 256), 256, 0, ST(stream)>>>(dbig, ld_big, B, h5, w5, h4, w4, C, dsmall);
  return set_cuda_error(cudaGetLastError());
}

extern "C" int ralf_conv_weight_to_gemm(const float* w, int N, int C, int T, int Kp, void* out, long long out_plane,
                                        void* outT, long long outT_plane, int Np, void* stream) {
  if (!w || !out) return RALF_ERR_NULL;
  if (N <= 0 || C <= 0 || T <= 0 || Kp < C * T) return RALF_ERR_SHAPE;
  const long long total = static_cast<long long>(N) * T * C;
  conv_weight_to_gemm_kernel<<<c_grid_for(total, 256), 256, 0, ST(stream)>>>(w, N, C, T, Kp, BF(out), out_plane, BF(outT),
                                                                            outT_plane, Np);
  return set_cuda_error(cudaGetLastError());
}

extern "C" int ralf_conv_grad_from_gemm(const float* dwg, int N, int C, int T, int ldg, float* dw, void* stream) {
  if (!dwg || !dw) return RALF_ERR_NULL;
  if (N <= 0 || C <= 0 || T <= 0) return RALF_ERR_SHAPE;
  const long long total = static_cast<long long>(N) * C * T;
  conv_grad_from_gemm_kernel<<<c_grid_for(total, 256), 256, 0, ST(stream)>>>(dwg, N, C, T, ldg, dw);
  return set_cuda_error(cudaGetLastError());
}
