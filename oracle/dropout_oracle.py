"""TEST INFRASTRUCTURE ONLY: numpy restatement of the counter-based dropout masks of the training kernels
(ralf_b200/csrc/common.cuh: drop_stream / drop_keep / make_drop_args).  Only tests/ may import this module.

The reference draws its masks from torch's Philox stream (nn.Dropout / nn.MultiheadAttention(dropout=0.1),
image2layout/train/models/retrieval_augmented_autoreg.py:105,116-126), which cannot be reproduced outside torch, so the
product defines its own generator; this file pins that definition independently of the CUDA build:

    stream = seed XOR (0xD1B54A32D192ED03 * (site + 1))                                   (mod 2^64)
    z      = SplitMix64 finaliser of  stream + 0x9E3779B97F4A7C15 * (idx + 1)             (mod 2^64)
    keep   = (z >> 40) >= floor(p * 2^24)          # top 24 bits against the threshold
"""
from __future__ import annotations

import numpy as np

_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def keep_mask(seed: int, site: int, p: float, n: int) -> np.ndarray:
    """uint8 [n]: 1 where element idx in [0, n) of dropout site `site` is kept in the step whose seed is `seed`."""
    with np.errstate(over="ignore"):
        seed = np.uint64(seed & 0xFFFFFFFFFFFFFFFF)  # the kernels read the int64 tensor as unsigned
        stream = seed ^ (np.uint64(0xD1B54A32D192ED03) * np.uint64(site + 1))
        idx = np.arange(n, dtype=np.uint64)
        z = stream + np.uint64(0x9E3779B97F4A7C15) * (idx + np.uint64(1))
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
        thresh = np.uint64(int(np.float32(p) * np.float32(16777216.0)))
        return ((z >> np.uint64(40)) >= thresh).astype(np.uint8)
