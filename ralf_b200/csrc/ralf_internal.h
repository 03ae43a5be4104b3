// Internal helpers shared by the .cu files (error plumbing, tensor-map builder).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "../../include/ralf_b200.h"

namespace ralf {
// Records the CUDA error string for ralf_last_cuda_error(); returns 0 or RALF_ERR_CUDA.
int set_cuda_error(cudaError_t e);
// K-major operand [planes, rows, K] -> 3-D tensor map with a (128 bytes x box_rows x 1) box,
// SWIZZLE_128B, zero fill out of bounds.  Cached by (ptr, shape).
int make_kmajor_tmap(CUtensorMap* out, const void* ptr, int elem_bytes, uint64_t K, uint64_t rows,
                     uint64_t planes, uint64_t ld, uint64_t plane_stride, uint32_t box_rows);
int num_sms();
// Programmatic dependent launch mode (RALF_PDL): 0 off, 1 every launch, 2 only single-wave grids (<= #SMs CTAs: the
// latency-bound kernels of the decode loop); see runtime.cu.
int pdl_mode();

// kernel<<<grid, block, smem, st>>>(args...) with the programmatic-stream-serialization attribute (see common.cuh).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  const int mode = pdl_mode();
  cfg.numAttrs = (mode == 1 || (mode == 2 && static_cast<long long>(grid.x) * grid.y * grid.z <= num_sms())) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// attention_tc.cu: tcgen05 attention for (head_dim 32, no mask, 64 <= Tk <= 256); 1 = launched, 0 = not applicable.
int attention_tc_try(const float* q, int ldq, const float* k, const float* v, int ldk, const unsigned char* mask, int B,
                     int H, int Tq, int Tk, int head_dim, int causal, float scale, void* out_split,
                     long long out_plane, float* out_f32, int ldo, cudaStream_t st);
}  // namespace ralf
