"""CPU: the dropout mask definition.  The kernels' mask functions are __host__ __device__ (csrc/common.cuh), so the
very same source is compiled for the host here and compared with the numpy restatement in oracle/dropout_oracle.py;
tests/test_dropout_gpu.py then compares the device masks with that restatement."""
import os
import shutil
import subprocess

import numpy as np
import pytest

from tests import helpers

SRC = r'''
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
int main(int argc, char** argv) {
  unsigned long long seed = strtoull(argv[1], nullptr, 10);
  unsigned int site = static_cast<unsigned int>(atoi(argv[2]));
  float p = static_cast<float>(atof(argv[3]));
  long long n = atoll(argv[4]);
  ralf::DropArgs a = ralf::make_drop_args(&seed, site, p);
  unsigned long long st = ralf::drop_stream(seed, site);
  for (long long i = 0; i < n; ++i) putchar(ralf::drop_keep(st, static_cast<unsigned long long>(i), a.thresh24) ? '1' : '0');
  return 0;
}
'''


def test_host_compiled_mask_equals_numpy_restatement(tmp_path):
    from oracle import dropout_oracle as D

    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not (os.path.exists(nvcc) or shutil.which("nvcc")):
        pytest.skip("nvcc not available")
    src = tmp_path / "host_mask.cu"
    src.write_text(SRC)
    exe = tmp_path / "host_mask"
    subprocess.run([nvcc, "-std=c++17", "-O2", "-I", os.path.join(helpers.ROOT, "ralf_b200", "csrc"), "-gencode",
                    "arch=compute_100a,code=sm_100a", "-o", str(exe), str(src)], check=True, capture_output=True)
    n = 4096
    for seed, site, p in [(1234567, 3, 0.1), (2 ** 64 - 1, 1, 0.5), (0, 0, 0.1), (-5, 77, 0.25), (9 * 10 ** 18, 200, 0.9)]:
        out = subprocess.run([str(exe), str(seed & (2 ** 64 - 1)), str(site), repr(p), str(n)], check=True,
                             capture_output=True, text=True).stdout
        host = np.frombuffer(out.encode(), dtype=np.uint8) - ord("0")
        assert host.shape == (n,)
        np.testing.assert_array_equal(host, D.keep_mask(seed, site, p, n))


def test_restated_mask_statistics():
    from oracle import dropout_oracle as D

    n = 1 << 18
    for p in (0.1, 0.5):
        m = D.keep_mask(42, 5, p, n)
        assert abs(m.mean() - (1 - p)) < 4 * np.sqrt(p * (1 - p) / n)
    a, b = D.keep_mask(42, 5, 0.1, n), D.keep_mask(42, 6, 0.1, n)
    assert abs((a == b).mean() - (0.9 ** 2 + 0.1 ** 2)) < 5e-3
