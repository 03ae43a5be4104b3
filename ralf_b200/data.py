"""Input pipeline around the hot path.

f1 -- ``RetrievalDatasetWrapper.__getitem__`` + the "retrieved" branch of ``collate_fn``
(image2layout/train/helpers/retrieval_dataset_wrapper.py:89-148, data.py:42-117) cost the reference 16 random dataset
reads and 32 image decodes per sample to build ``retrieved{...}``, of which the model reads only the layouts
(``use_reference_image=False``, retrieval_augmented_autoreg.py:74,542).  Here the gallery layouts live in ONE packed
table in HBM (:class:`LayoutTable`, [N, 6, E] fp32) and a batch's exemplars are one index gather
(``ralf_gather_layouts``); the dict schema the model classes consume is kept.

f2 -- the reference's retrieval wire formats: the ``.pt`` cache tables ``dict[data_id -> list[db_index]]`` written by
``Retriever.preprocess_retrieval_cache`` (models/retrieval/retriever.py:134-229) and read by ``load_cache_table``
(retrieval_dataset_wrapper.py:17-33), the ``{id: idx}`` pairing table (:167-188) and the YAML export under
``data_splits/retrieval/<dataset>/<split>.yaml`` (``'<query id>': ['<db id>', ... x16]``).  Tables written here load in
the reference and vice versa.
"""
from __future__ import annotations

import os
from typing import Any, Iterable, Optional, Sequence

import torch

from . import ops
from .retrieval import LAYOUT_KEYS, GpuRetriever

GEO = ["center_x", "center_y", "width", "height"]


# --------------------------------------------------------------------------------------------------------------------
# f2: wire formats
# --------------------------------------------------------------------------------------------------------------------
def cache_table_path(dataset_name: str, split: str, retrieval_backbone: str, top_k: int = 32, root: str = "cache") -> str:
    """File name of retriever.py:149 / retrieval_dataset_wrapper.py:59."""
    return os.path.join(root, f"{dataset_name}_{split}_{retrieval_backbone}_wo_head_table_between_dataset_indexes_top_k{top_k}.pt")


def paired_table_path(dataset_name: str, retrieval_backbone: str, root: str = "cache") -> str:
    """retriever.py:163-165: {data_id -> index into the retrieval database}."""
    return os.path.join(root, f"{dataset_name}_{retrieval_backbone}_cache_table_paired.pt")


def save_cache_table(table: dict, path: str) -> None:
    """``torch.save`` of a plain ``dict[data_id, list[int]]`` (the reference pickles a defaultdict; both load alike)."""
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save({k: [int(i) for i in v] for k, v in table.items()}, path)


def load_cache_table(path: str, top_k: int) -> dict:
    """retrieval_dataset_wrapper.py:17-33: load and keep the first ``top_k`` indexes of every entry."""
    if not os.path.exists(path):
        raise ValueError(f"Cache not found in {path}")
    table = torch.load(path, weights_only=False)
    return {k: list(v[:top_k]) for k, v in table.items()}


def export_retrieval_yaml(table: dict, db_ids: Sequence, path: str, top_k: int = 16) -> None:
    """data_splits/retrieval/**.yaml: query id -> the DATA IDS (strings) of its ``top_k`` exemplars."""
    import yaml

    out = {str(k): [str(db_ids[i]) for i in v[:top_k]] for k, v in table.items()}
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "w") as f:
        yaml.safe_dump(out, f, default_flow_style=False)


def load_retrieval_yaml(path: str, id_to_index: Optional[dict] = None) -> dict:
    """Inverse of :func:`export_retrieval_yaml`; with ``id_to_index`` (the pairing table) the data ids are mapped back to
    database indexes, i.e. to the ``.pt`` table format."""
    import yaml

    with open(path) as f:
        raw = yaml.safe_load(f)
    if id_to_index is None:
        return {str(k): [str(x) for x in v] for k, v in raw.items()}
    norm = {str(k): v for k, v in id_to_index.items()}
    return {str(k): [norm[str(x)] for x in v] for k, v in raw.items()}


def build_cache_table(retriever: GpuRetriever, queries: torch.Tensor, query_ids: Sequence, split: str, top_k: int = 32,
                      batch: int = 1024) -> dict:
    """``preprocess_retrieval_cache`` on the GPU: search top_k + 1, drop the query itself on the train split
    (retriever.py:193-213; other splits keep all top_k + 1 like the reference and are cut by ``load_cache_table``)."""
    table: dict = {}
    for s in range(0, len(query_ids), batch):
        part = retriever.build_table(queries[s:s + batch], list(query_ids[s:s + batch]), None, top_k, drop_self=(split == "train"))
        table.update(part)
    return table


# --------------------------------------------------------------------------------------------------------------------
# f1: per-example element ordering (data.transforms of experiment/ralf.yaml:8 = [image, sort_label, sort_lexicographic])
# --------------------------------------------------------------------------------------------------------------------
def _reorder(example: dict, order: Sequence[int]) -> dict:
    """Apply ``order`` to every list field of the example (helpers/hfds_instance_wise_transforms.py:25-37; the keys
    ``transforms``, ``retrieved`` and ``id`` are left alone)."""
    for key, value in example.items():
        if key not in ("transforms", "retrieved", "id") and isinstance(value, list):
            example[key] = [value[i] for i in order]
    return example


def sort_label_transform(example: dict) -> dict:
    """Stable sort of the elements by label id (:58-64)."""
    labels = example["label"]
    return _reorder(example, sorted(range(len(labels)), key=labels.__getitem__)) if labels else example


def lexicographic_order(example: dict) -> list:
    """Element indexes sorted by (top, left) of the boxes (:67-78)."""
    as_float = lambda x: x.tolist() if torch.is_tensor(x) else x  # noqa: E731
    left = [as_float(cx - w / 2.0) for cx, w in zip(example["center_x"], example["width"])]
    top = [as_float(cy - h / 2.0) for cy, h in zip(example["center_y"], example["height"])]
    return sorted(range(len(top)), key=lambda i: (top[i], left[i]))


def sort_lexicographic_transform(example: dict) -> dict:
    return _reorder(example, lexicographic_order(example)) if len(example["center_x"]) else example


def shuffle_transform(example: dict) -> dict:
    """Random element order from python's ``random`` like the reference (:47-55)."""
    import random

    n = len(example["label"])
    return _reorder(example, random.sample(list(range(n)), n)) if n else example


INSTANCE_TRANSFORMS = {"sort_label": sort_label_transform, "sort_lexicographic": sort_lexicographic_transform,
                       "shuffle": shuffle_transform}


def apply_transforms(example: dict, names: Sequence[str] = ("sort_label", "sort_lexicographic")) -> dict:
    """``data.transforms`` minus ``image`` (PIL -> tensor conversion happens where the images are decoded)."""
    for n in names:
        if n != "image":
            example = INSTANCE_TRANSFORMS[n](example)
    return example


# --------------------------------------------------------------------------------------------------------------------
# f1: GPU-resident exemplar layouts
# --------------------------------------------------------------------------------------------------------------------
class LayoutTable:
    """Packed gallery layouts [N, 6, E] fp32 (label, mask, center_x, center_y, width, height; valid-first padding with
    zeros, ``pad()`` of helpers/util.py:52-70) resident on the device."""

    def __init__(self, packed: torch.Tensor, ids: Optional[Sequence] = None) -> None:
        assert packed.dim() == 3 and packed.shape[1] == len(LAYOUT_KEYS)
        self.packed = packed.contiguous()
        self.ids = list(ids) if ids is not None else None

    @classmethod
    def from_rows(cls, rows: Iterable[dict], max_seq_length: int, device=None) -> "LayoutTable":
        """rows: database examples with list fields ``label, center_x, center_y, width, height`` (+ ``id``), i.e. what
        ``db_dataset[i]`` yields in the reference; longer layouts are cut at ``max_seq_length``."""
        rows = list(rows)
        E = max_seq_length
        packed = torch.zeros((len(rows), len(LAYOUT_KEYS), E), dtype=torch.float32)
        ids = []
        for n, r in enumerate(rows):
            m = min(len(r["label"]), E)
            packed[n, 0, :m] = torch.as_tensor(r["label"][:m], dtype=torch.float32)
            packed[n, 1, :m] = 1.0
            for c, key in enumerate(GEO, start=2):
                packed[n, c, :m] = torch.as_tensor(r[key][:m], dtype=torch.float32)
            ids.append(r.get("id", n))
        return cls(packed.to(device) if device is not None else packed, ids)

    @classmethod
    def from_tensors(cls, layouts: dict, device=None) -> "LayoutTable":
        packed = torch.stack([layouts[k].to(torch.float32) for k in LAYOUT_KEYS], dim=1)
        return cls(packed.to(device) if device is not None else packed)

    def __len__(self) -> int:
        return self.packed.shape[0]

    def gather(self, idx: torch.Tensor) -> dict:
        """idx int64 [B, K] (device) -> retrieved{label int64, mask bool, geometry fp32: [B, K, E]; packed; index}."""
        packed = ops.gather_layouts(self.packed, idx.to(self.packed.device).contiguous())
        out = {"packed": packed, "index": idx}
        for c, key in enumerate(LAYOUT_KEYS):
            v = packed[:, :, c]
            out[key] = v.to(torch.int64) if key == "label" else (v != 0 if key == "mask" else v)
        B, K = idx.shape
        # the model classes assert retrieved["image"].size(2) == 4 and never read it (use_reference_image=False)
        out["image"] = torch.zeros((B, K, 4, 1, 1), device=packed.device)
        out["saliency"] = torch.zeros((B, K, 1, 1, 1), device=packed.device)
        return out


class RetrievalCollator:
    """Batch-level replacement of RetrievalDatasetWrapper + collate_fn: pads the main samples on the host (data.py:42-117)
    and attaches ``retrieved`` from the cache table (or from an in-line search when query embeddings are given: row f2's
    "online retrieval") by one device gather."""

    def __init__(self, layouts: LayoutTable, max_seq_length: int, top_k: int = 16, table_idx: Optional[dict] = None,
                 retriever: Optional[GpuRetriever] = None, int_ids: bool = False, transforms: Sequence[str] = (),
                 random_retrieval: bool = False, num_query_rows: Optional[int] = None) -> None:
        """``random_retrieval`` (generator config key; helpers/random_retrieval_dataset_wrapper.py:70-72): exemplars are
        ``torch.randint(0, len(query split), [top_k])`` per sample, drawn on the host generator in batch order;
        ``num_query_rows`` = len of the split being iterated (defaults to the table size)."""
        self.layouts, self.E, self.top_k = layouts, max_seq_length, top_k
        self.random_retrieval, self.num_query_rows = random_retrieval, num_query_rows
        self.table_idx, self.retriever, self.int_ids = table_idx, retriever, int_ids
        self.transforms = tuple(transforms)  # e.g. ("sort_label", "sort_lexicographic") when the dataset is raw

    def indices(self, ids: Sequence) -> torch.Tensor:
        if self.random_retrieval:
            high = self.num_query_rows if self.num_query_rows is not None else len(self.layouts)
            return torch.stack([torch.randint(low=0, high=high, size=[self.top_k]) for _ in ids])
        rows = []
        for i in ids:
            key = int(i) if self.int_ids else i  # "pku" ids are ints in the tables (retrieval_dataset_wrapper.py:103-104)
            row = self.table_idx[key][:self.top_k]
            assert len(row) == self.top_k, f"{len(row)=} != {self.top_k=}"
            rows.append(row)
        return torch.tensor(rows, dtype=torch.int64)

    def collate_main(self, examples: Sequence[dict]) -> dict:
        """Padding / masks of the query samples: list fields padded with 0 to ``max_seq_length``, ``mask`` = valid-first;
        empty layouts get the reference's dummy element (global_variables.DUMMY_LAYOUT)."""
        B, E = len(examples), self.E
        if self.transforms:
            examples = [apply_transforms(dict(ex), self.transforms) for ex in examples]
        out: dict[str, Any] = {"label": torch.zeros((B, E), dtype=torch.int64), "mask": torch.zeros((B, E), dtype=torch.bool)}
        for key in GEO:
            out[key] = torch.zeros((B, E), dtype=torch.float32)
        for b, ex in enumerate(examples):
            n = len(ex["label"])
            if n == 0:
                ex = {**ex, "label": [0], "center_x": [0.5], "center_y": [0.5], "width": [0.05], "height": [0.05]}
                n = 1
            n = min(n, E)
            out["label"][b, :n] = torch.as_tensor(ex["label"][:n], dtype=torch.int64)
            out["mask"][b, :n] = True
            for key in GEO:
                out[key][b, :n] = torch.as_tensor(ex[key][:n], dtype=torch.float32)
        out["id"] = [ex["id"] for ex in examples]
        for key in ("image", "saliency"):
            if key in examples[0] and torch.is_tensor(examples[0][key]):
                out[key] = torch.stack([ex[key] for ex in examples])
        return out

    def __call__(self, examples: Sequence[dict], query_embeddings: Optional[torch.Tensor] = None) -> dict:
        batch = self.collate_main(examples)
        if query_embeddings is not None:
            assert self.retriever is not None, "online retrieval needs a GpuRetriever"
            idx, _ = self.retriever.search(query_embeddings, self.top_k)
        else:
            idx = self.indices(batch["id"])
        batch["retrieved"] = self.layouts.gather(idx)
        return batch
