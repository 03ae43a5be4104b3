"""Deterministic synthetic weights and inputs shared by the golden-fixture generator, the oracle
and the GPU tests.  TEST INFRASTRUCTURE ONLY (no weights or datasets can be downloaded here).

Weights are drawn per state-dict key from a generator seeded by (seed, crc32(key)), with scales chosen
so activations stay O(1) through ResNet50 + 16 transformer layers.  The key -> shape schema of the
reference classes is committed under tests/golden/schema_*.json (dumped by make_golden.py).
"""
from __future__ import annotations

import math
import zlib

import torch


def sine_pe_1d(max_len: int, d_model: int) -> torch.Tensor:
    """PositionalEncoding1d buffer (common/positional_encoding.py:71-81), shape [1, max_len, d]."""
    position = torch.arange(max_len).unsqueeze(1)
    div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
    pe = torch.zeros(1, max_len, d_model)
    pe[0, :, 0::2] = torch.sin(position * div_term)
    pe[0, :, 1::2] = torch.cos(position * div_term)
    return pe


def _gen(seed: int, key: str) -> torch.Generator:
    return torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 63))


def synth_state_dict(schema: dict, seed: int = 0) -> dict:
    """schema: {key: {"shape": [...], "dtype": "float32"|"int64"|"bool"}} -> state dict (CPU)."""
    sd = {}
    for key, spec in schema.items():
        shape, dtype = tuple(spec["shape"]), spec["dtype"]
        g = _gen(seed, key)
        leaf = key.split(".")[-1]
        if key.endswith(".pe"):
            t = sine_pe_1d(shape[1], shape[2])
        elif dtype == "bool":
            t = torch.zeros(shape, dtype=torch.bool)
        elif dtype == "int64":
            t = torch.ones(shape, dtype=torch.int64) if key == "flag_user_const" else torch.zeros(shape, dtype=torch.int64)
        elif leaf == "running_var":
            t = torch.rand(shape, generator=g) + 0.5
        elif leaf == "running_mean":
            t = torch.randn(shape, generator=g) * 0.1
        elif key == "task_emb.weight":
            t = torch.randn(shape, generator=g) * 0.02
        elif leaf == "token":
            t = torch.randn(shape, generator=g)
        elif len(shape) == 1 and leaf == "weight":  # BatchNorm / LayerNorm gains
            t = torch.rand(shape, generator=g) + 0.5
            if ".bn3." in key:  # keep the residual branches modest so 16 blocks do not blow up
                t = t * 0.25
        elif len(shape) == 1:  # biases
            t = torch.randn(shape, generator=g) * 0.05
        elif "emb" in key:
            t = torch.randn(shape, generator=g) * 0.3
        else:  # conv / linear weights: fan-in scaling
            fan_in = 1
            for s in shape[1:]:
                fan_in *= s
            t = torch.randn(shape, generator=g) * (1.0 / math.sqrt(fan_in))
            if len(shape) == 4:
                t = t * math.sqrt(2.0)
        sd[key] = t
    return sd


def synth_batch(B: int, H: int, W: int, max_elem: int, top_k: int, num_labels: int, seed: int = 0) -> dict:
    """A collated batch with the schema of data.py:42-117 (image, saliency, layout fields, retrieved{...}).
    Elements are drawn like SURVEY.md 8d: n ~ U{1..E}, label uniform, geometry U[0,1), valid-first padding."""
    g = torch.Generator().manual_seed(seed * 7919 + 17)

    def layouts(lead):
        n = torch.randint(1, max_elem + 1, lead, generator=g)
        mask = torch.arange(max_elem).expand(*lead, max_elem) < n[..., None]
        label = torch.randint(0, num_labels, (*lead, max_elem), generator=g)
        out = {"label": torch.where(mask, label, torch.zeros_like(label)), "mask": mask}
        for key in ["center_x", "center_y", "width", "height"]:
            v = torch.rand((*lead, max_elem), generator=g)
            out[key] = torch.where(mask, v, torch.zeros_like(v))
        return out

    batch = layouts((B,))
    batch["image"] = torch.rand((B, 3, H, W), generator=g)
    batch["saliency"] = torch.rand((B, 1, H, W), generator=g)
    batch["id"] = [str(i) for i in range(B)]
    batch["retrieved"] = layouts((B, top_k))
    # the reference asserts retrieved["image"].size(2) == 4 but never reads it (use_reference_image=False)
    batch["retrieved"]["image"] = torch.zeros((B, top_k, 4, 1, 1))
    batch["retrieved"]["saliency"] = torch.zeros((B, top_k, 1, 1, 1))
    return batch
