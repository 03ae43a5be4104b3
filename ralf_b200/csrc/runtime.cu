// Error plumbing and device queries shared by all entry points.
#include <stdlib.h>
#include <string.h>

#include "ralf_internal.h"

namespace ralf {
static thread_local char g_last_error[256] = "";

int set_cuda_error(cudaError_t e) {
  if (e == cudaSuccess) return RALF_OK;
  strncpy(g_last_error, cudaGetErrorString(e), sizeof(g_last_error) - 1);
  g_last_error[sizeof(g_last_error) - 1] = 0;
  return RALF_ERR_CUDA;
}

int num_sms() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

int pdl_mode() {
  // Opt-in (RALF_PDL=1).  Measured on B200 (capture N, 1024-canvas step under CUDA graphs): 201.2 ms with programmatic
  // edges vs 186.3 ms without -- dependents that become resident early hold shared memory / TMEM next to the still
  // running producer and cost more than the ~2 us of prologue they hide; the graph-replayed training step also drifted
  // 1.5e-4 from the eager one.  Off by default; the kernels keep their (then no-op) griddepcontrol instructions.
  // RALF_PDL=2 (round 2): only for grids of at most one CTA per SM -- the decode loop's kernels, whose dependents find
  // idle SMs to become resident on.
  static const int mode = getenv("RALF_PDL") ? atoi(getenv("RALF_PDL")) : 0;
  return mode;
}
}  // namespace ralf

extern "C" const char* ralf_last_cuda_error(void) { return ralf::g_last_error; }

extern "C" int ralf_check_device(int dev) {
  int major = 0;
  cudaError_t e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return ralf::set_cuda_error(e);
  return major == 10 ? RALF_OK : RALF_ERR_ARCH;
}

extern "C" int ralf_version(void) { return 100; }
