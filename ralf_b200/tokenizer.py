"""Host-side layout <-> token-sequence conversion (linear bucketizer).

Mirror of the reference's ``LayoutSequenceTokenizer`` (image2layout/train/helpers/layout_tokenizer.py:
293-446) and ``_LinearBucketizer`` (helpers/bucketizer.py:39-81) for the configuration the RALF /
Autoreg experiments use (``train/config/tokenizer.py``): 128 linear bins per geometry variable, no
shared location vocabulary, special tokens pad/bos/eos, element order
(label, width, height, center_x, center_y).  Token ids must be bit-identical to the reference's.
"""
from __future__ import annotations

from typing import Sequence

import torch
from torch import Tensor

GEO_KEYS = ["center_x", "center_y", "width", "height"]  # global_variables.py:1


class LinearBucketizer:
    """helpers/bucketizer.py:39-81: clamp to [0,1], torch.bucketize on right edges i/n; centres (i+.5)/n."""

    def __init__(self, n_boundaries: int = 128) -> None:
        arr = torch.arange(n_boundaries + 1) / n_boundaries
        self.boundaries = arr[1:]
        self.centers = (arr[:-1] + arr[1:]) / 2.0

    def encode(self, data: Tensor) -> Tensor:
        return torch.bucketize(torch.clamp(data, min=0.0, max=1.0), self.boundaries.to(data.device))

    def decode(self, index: Tensor) -> Tensor:
        index = torch.clamp(index, min=0, max=len(self.centers) - 1)
        return self.centers.to(index.device)[index]


class LayoutSequenceTokenizer:
    def __init__(
        self,
        label_names: Sequence[str],
        max_seq_length: int,
        num_bin: int = 128,
        var_order: Sequence[str] = ("label", "width", "height", "center_x", "center_y"),
        special_tokens: Sequence[str] = ("pad", "bos", "eos"),
    ) -> None:
        self.label_names = list(label_names)
        self.max_seq_length = max_seq_length
        self.num_bin = num_bin
        self.var_order = list(var_order)
        self.special_tokens = list(special_tokens)
        self.bucketizers = {k: LinearBucketizer(num_bin) for k in GEO_KEYS}
        self._sp = {t: self.special_tokens.index(t) + self.N_label + self.N_bbox for t in self.special_tokens}

    # --- vocabulary layout (layout_tokenizer.py:246-279) ---
    @property
    def N_label(self) -> int:
        return len(self.label_names)

    @property
    def N_bbox_per_var(self) -> int:
        return self.num_bin

    @property
    def N_bbox(self) -> int:
        return self.num_bin * 4

    @property
    def N_sp_token(self) -> int:
        return len(self.special_tokens)

    @property
    def N_total(self) -> int:
        return self.N_label + self.N_bbox + self.N_sp_token

    @property
    def N_var_per_element(self) -> int:
        return len(self.var_order)

    @property
    def max_token_length(self) -> int:
        return self.max_seq_length * self.N_var_per_element

    def name_to_id(self, name: str) -> int:
        return self._sp[name]

    # --- encode (layout_tokenizer.py:302-360) ---
    def encode(self, inputs: dict) -> dict:
        label = inputs["label"].clone()
        mask = inputs["mask"].clone()
        data = {"label": label}
        for i, key in enumerate(GEO_KEYS):
            data[key] = self.bucketizers[key].encode(inputs[key]) + self.N_label + i * self.N_bbox_per_var
        pad_id = self.name_to_id("pad")
        for key in ["label"] + GEO_KEYS:
            data[key][~mask] = pad_id
        B, S = label.shape
        C = self.N_var_per_element
        seq_len = mask.int().sum(dim=1, keepdim=True)
        seq = torch.stack([data[k] for k in self.var_order], dim=-1).reshape(B, S * C)
        m = mask[:, :, None].expand(B, S, C).reshape(B, S * C).clone()
        indices = torch.arange(0, S * C, device=seq.device)[None]
        eos_mask = seq_len * C == indices
        seq[eos_mask] = self.name_to_id("eos")
        m[eos_mask] = True
        bos = torch.full((B, 1), self.name_to_id("bos"), dtype=seq.dtype, device=seq.device)
        seq = torch.cat([bos, seq], dim=-1)
        m = torch.cat([torch.ones((B, 1), dtype=torch.bool, device=seq.device), m], dim=-1)
        return {"seq": seq, "mask": m}

    # --- decode (layout_tokenizer.py:362-402) ---
    def decode(self, seq: Tensor) -> dict:
        B = seq.size(0)
        s = seq.clone().reshape(B, -1, self.N_var_per_element)
        out = {}
        for i, key in enumerate(self.var_order):
            out[key] = s[..., i].clone()
            if key in GEO_KEYS:
                out[key] = out[key] - self.N_label - GEO_KEYS.index(key) * self.N_bbox_per_var
        invalid = torch.cumsum(out["label"] == self.name_to_id("eos"), dim=1) > 0
        label_valid = (0 <= out["label"]) & (out["label"] < self.N_label)
        geo_valid = torch.ones_like(label_valid)
        for key in GEO_KEYS:
            geo_valid &= (0 <= out[key]) & (out[key] < self.N_bbox)
        invalid = invalid | ~(label_valid & geo_valid)
        for key in GEO_KEYS:
            out[key][invalid] = 0
            out[key] = self.bucketizers[key].decode(out[key])
        for key in self.var_order:
            out[key][invalid] = 0
        out["mask"] = ~invalid
        return out

    # --- per-position vocabulary mask (layout_tokenizer.py:404-446) ---
    @property
    def token_mask(self) -> Tensor:
        last = torch.tensor([t not in ("bos", "mask") for t in self.special_tokens])
        rows = []
        for key in self.var_order:
            if key == "label":
                rows.append(torch.cat([torch.ones(self.N_label, dtype=torch.bool),
                                       torch.zeros(self.N_bbox, dtype=torch.bool), last]))
            else:
                g = torch.zeros(self.N_bbox, dtype=torch.bool)
                i = GEO_KEYS.index(key)
                g[i * self.N_bbox_per_var:(i + 1) * self.N_bbox_per_var] = True
                rows.append(torch.cat([torch.zeros(self.N_label, dtype=torch.bool), g, last]))
        return torch.stack(rows, dim=0).repeat(self.max_seq_length, 1)
