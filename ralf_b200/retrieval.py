"""GPU retrieval: exact top-k inner-product search over an embedding gallery + exemplar-layout fetch.

Host-side mirror of the reference's retrieval objects for the hot path:

* ``Retriever`` (image2layout/train/models/retrieval/retriever.py:24-229) builds a FAISS
  ``IndexFlat(d, METRIC_INNER_PRODUCT)`` over the training-split embeddings and, per query, searches
  ``top_k + 1`` neighbours, drops the query itself on the train split and stores ``dict[id -> list[idx]]``
  (``preprocess_retrieval_cache`` :134-229).  :class:`GpuRetriever.search` / :meth:`build_table` do the same
  with the sm_100a kernel (csrc/knn.cu), batched over queries and optionally sharded over GPUs.
* ``RetrievalDatasetWrapper.__getitem__`` + ``collate_fn`` (helpers/retrieval_dataset_wrapper.py:89-148,
  data.py:78-88) turn the k indices into ``retrieved{label, mask, center_x, center_y, width, height: [B,K,E]}``.
  :meth:`GpuRetriever.fetch` is an index gather from a GPU-resident packed layout table (exemplar images
  are never read when ``use_reference_image=False``, retrieval_augmented_autoreg.py:74,542).

Multi-GPU: the gallery and layout table are row-sharded; every rank searches its shard for ALL
queries, the per-rank [Q,k] (score, global index) lists are all-gathered (NCCL) and merged with the same
(score desc, index asc) rule -- the result is bit-identical to the unsharded search.
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from . import ops

LAYOUT_KEYS = ["label", "mask", "center_x", "center_y", "width", "height"]


class GpuRetriever:
    def __init__(self, embeddings: torch.Tensor, layouts: Optional[dict] = None, *, device=None, rank: int = 0,
                 world_size: int = 1, index_base: Optional[int] = None, process_group=None) -> None:
        """embeddings: this rank's shard [n_local, d] fp32 (the whole gallery when world_size == 1);
        layouts: dict of [n_total, E] tensors (replicated: it is tiny next to the embeddings)."""
        self.dev = torch.device(device) if device is not None else embeddings.device
        self.emb = embeddings.to(self.dev, torch.float32).contiguous()
        self.rank, self.world = rank, world_size
        self.pg = process_group
        self.index_base = index_base if index_base is not None else 0
        # FAISS does not normalise; the TF32 certificate only needs an upper bound on the row norms
        self.max_norm = float(self.emb.norm(dim=1).max().item()) * 1.0001 if self.emb.numel() else 0.0
        self.table = None  # packed fp32 [n_total, 6, E]: label, mask, center_x, center_y, width, height
        if layouts is not None:
            self.table = torch.stack([layouts[k].to(self.dev, torch.float32) for k in LAYOUT_KEYS], dim=1).contiguous()
        self._ws: Optional[torch.Tensor] = None
        # Opt-in (unmeasured): passes of 128 queries on `knn_ways` parallel streams.  A pass is pre-pass -> threshold ->
        # scan -> re-rank; the scan runs at the HBM roofline, the three small kernels around it are latency bound
        # (51 of the 360 us of a pass, profiles/r1_launches_l_summary.md) and can hide under another pass's scan.
        self.knn_ways = max(1, int(os.environ.get("RALF_KNN_WAYS", "1")))
        self._way_streams: list = []

    # -- search ---------------------------------------------------------------------------------
    def search_local(self, queries: torch.Tensor, k: int):
        """Top-k of this rank's shard for every query: (idx int64 [Q,k] global ids, score fp32 [Q,k]).
        Exact unconditionally: queries the TF32 bound cannot certify are re-run through the exact scan on the device
        (``ops.knn_topk(fixup=True)``: no host sync, so it is part of the captured search graph); afterwards
        ``last_certified`` is 1 (bound) or 2 (exact scan) for every query."""
        q = queries.to(self.dev, torch.float32).contiguous()
        if self.knn_ways > 1 and q.shape[0] > 128:
            return self._search_local_ways(q, k)
        idx, score, cert = ops.knn_topk(self.emb, q, k, index_base=self.index_base, gallery_max_norm=self.max_norm,
                                        workspace=self._ws)
        self.last_certified = cert
        return idx, score

    def _search_local_ways(self, q: torch.Tensor, k: int):
        """The same passes as ralf_knn_topk runs internally (128 queries each, independent of one another), issued
        round-robin on parallel streams with a workspace each; fork / join around them, so inside a graph capture they
        become parallel branches.  Results are identical to the single-stream call."""
        if len(self._way_streams) < self.knn_ways:
            self._way_streams = [torch.cuda.Stream(device=self.dev) for _ in range(self.knn_ways)]
        cur = torch.cuda.current_stream()
        parts = []
        for s in self._way_streams:
            s.wait_stream(cur)
        for n, q0 in enumerate(range(0, q.shape[0], 128)):
            with torch.cuda.stream(self._way_streams[n % self.knn_ways]):
                parts.append(ops.knn_topk(self.emb, q[q0:q0 + 128], k, index_base=self.index_base,
                                          gallery_max_norm=self.max_norm))
        for s in self._way_streams:
            cur.wait_stream(s)
        idx, score, cert = (torch.cat([p[i] for p in parts]) for i in range(3))
        self.last_certified = cert
        return idx, score

    def search(self, queries: torch.Tensor, k: int, *, certify: bool = True):
        """Global top-k (exact: see ``search_local``; ``certify`` is kept for callers of the round-1 signature and no
        longer selects anything -- the device-side fix-up always runs)."""
        idx, score = self.search_local(queries, k)
        if self.world > 1:
            idx, score = exchange_and_merge(idx, score, self.world, self.pg, ops.knn_merge)
        return idx, score

    # -- exemplar fetch -------------------------------------------------------------------------
    def fetch(self, idx: torch.Tensor) -> dict:
        """indices [B, K] -> retrieved{"packed": [B, K, 6, E], label/mask/...: [B, K, E] views} gathered from the
        resident layout table by ralf_gather_layouts (one kernel, no host round trip)."""
        assert self.table is not None, "no layout table attached"
        packed = ops.gather_layouts(self.table, idx)
        out = {"packed": packed}
        for i, k in enumerate(LAYOUT_KEYS):
            out[k] = packed[:, :, i]
        return out

    # -- reference cache-table format -----------------------------------------------------------
    def build_table(self, queries: torch.Tensor, query_ids: list, db_ids: list, k: int, drop_self: bool) -> dict:
        """``dict[data_id -> list[db_index]]`` like preprocess_retrieval_cache (retriever.py:193-221):
        search k+1, and on the train split drop the first hit (the query itself); the other splits keep all k+1 hits
        like the reference (``load_cache_table`` cuts to top_k when reading)."""
        idx, _ = self.search(queries, k + 1)
        idx = idx.cpu().tolist()
        table = {}
        for qid, row in zip(query_ids, idx):
            table[qid] = row[1:] if drop_self else row
        return table


def exchange_and_merge(idx: torch.Tensor, score: torch.Tensor, world: int, pg, merge_fn):
    """The one exchange step of the sharded search: all-gather every rank's [Q,k]
    (score, global index) lists and merge them with ``merge_fn(all_score [W,Q,k], all_idx [W,Q,k])``.
    Backend-agnostic (NCCL on GPUs; gloo in the CPU tests)."""
    import torch.distributed as dist

    all_s = torch.empty((world, *score.shape), dtype=score.dtype, device=score.device)
    all_i = torch.empty((world, *idx.shape), dtype=idx.dtype, device=idx.device)
    if score.is_cuda:
        dist.all_gather_into_tensor(all_s, score.contiguous(), group=pg)
        dist.all_gather_into_tensor(all_i, idx.contiguous(), group=pg)
    else:  # gloo has no all_gather_into_tensor for every dtype: use the list form
        ls = [torch.empty_like(score) for _ in range(world)]
        li = [torch.empty_like(idx) for _ in range(world)]
        dist.all_gather(ls, score.contiguous(), group=pg)
        dist.all_gather(li, idx.contiguous(), group=pg)
        all_s, all_i = torch.stack(ls), torch.stack(li)
    return merge_fn(all_s, all_i)


def shard_bounds(n: int, world: int, rank: int) -> tuple[int, int]:
    return (n * rank) // world, (n * (rank + 1)) // world


class FlatIPIndex:
    """``faiss.IndexFlat(d, faiss.METRIC_INNER_PRODUCT)`` as far as the reference uses it, on the B200.

    The reference never talks to FAISS directly: ``Retriever.__init__`` hands HF ``datasets`` the vectors
    (``add_faiss_index_from_external_arrays(vectors, index_name=..., metric_type=faiss.METRIC_INNER_PRODUCT)``,
    retrieval/retriever.py:79-84) and later calls ``get_nearest_examples(index_name, query, k)`` (:112-114, :200-202);
    ``datasets.search.FaissIndex`` only needs ``index.add(vecs)`` and ``index.search(queries, k) -> (scores, indices)``
    on numpy arrays, and accepts any such object through its ``custom_index=`` keyword.  So the drop-in at the reference's
    own boundary is ONE keyword at retriever.py:79:  ``custom_index=ralf_b200.retrieval.FlatIPIndex(vectors.shape[1])``.

    ``add`` copies host rows into a device-resident gallery (fp32, unnormalised, row order = ids, like IndexFlat);
    ``search`` runs ``ralf_knn_topk`` (uncertified queries re-run through the exact kernel) and returns host arrays:
    scores fp32 [q, k] descending, labels int64 [q, k], ties towards the lower id, ``-1`` / ``-inf`` past ``ntotal``."""

    metric_type = 0  # faiss.METRIC_INNER_PRODUCT
    is_trained = True
    verbose = False

    def __init__(self, d: int, device=None) -> None:
        if d <= 0 or d % 4:
            raise ValueError(f"d = {d}: the search kernel needs a positive multiple of 4 (16-byte rows)")
        self.d = int(d)
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else "cuda"
        self.dev = torch.device(device)
        self._chunks: list = []
        self._retr: Optional[GpuRetriever] = None
        self.ntotal = 0

    def _rows(self, x, what: str) -> torch.Tensor:
        t = torch.as_tensor(x)
        if t.dim() != 2 or t.shape[1] != self.d:
            raise ValueError(f"{what}: expected [n, {self.d}], got {tuple(t.shape)}")
        return t.to(torch.float32)

    def add(self, x) -> None:
        rows = self._rows(x, "add")
        if rows.shape[0] == 0:
            return
        self._chunks.append(rows.to(self.dev, non_blocking=True))
        self.ntotal += rows.shape[0]
        self._retr = None

    def train(self, x) -> None:  # flat indexes have nothing to train
        pass

    def reset(self) -> None:
        self._chunks, self._retr, self.ntotal = [], None, 0

    def _gallery(self) -> GpuRetriever:
        if self._retr is None:
            emb = self._chunks[0] if len(self._chunks) == 1 else torch.cat(self._chunks, dim=0)
            self._chunks = [emb]
            self._retr = GpuRetriever(emb, device=self.dev)
        return self._retr

    def reconstruct(self, i: int):
        return self._gallery().emb[int(i)].cpu().numpy()

    def search(self, x, k: int, **kwargs):
        import numpy as np

        q = self._rows(x, "search")
        k = int(k)
        if k <= 0 or k > 48:
            raise ValueError(f"k = {k}: the fused top-k filter keeps at most 48 results per query (the reference asks for <= 33)")
        scores = np.full((q.shape[0], k), -np.inf, dtype=np.float32)
        labels = np.full((q.shape[0], k), -1, dtype=np.int64)
        if self.ntotal == 0 or q.shape[0] == 0:
            return scores, labels
        kk = min(k, self.ntotal)
        idx, score = self._gallery().search(q.to(self.dev), kk, certify=True)
        scores[:, :kk], labels[:, :kk] = score.cpu().numpy(), idx.cpu().numpy()
        return scores, labels


def coarse_saliency(saliency: torch.Tensor, size: tuple = (16, 16)) -> torch.Tensor:
    """Batched ``coarse_saliency`` (models/retrieval/image.py:35-44): the retrieval feature of
    ``retrieval_backbone="saliency"`` -- a nearest-neighbour ``size`` thumbnail of the saliency map, clamped to [0, 1] and
    mapped to [-1, 1].  saliency [B, 1, H, W] (or [B, H, W]) on any device -> fp32 [B, size[0] * size[1]], ready to be the
    gallery / query rows of the search (d = 256).  Pure index gather (source pixel = floor(dst * in / out), the rule of
    ``F.interpolate(mode="nearest")``), so it runs where the canvases already are."""
    if saliency.dim() == 4:
        assert saliency.size(1) == 1, f"{saliency.shape=}"
        saliency = saliency[:, 0]
    B, H, W = saliency.shape
    rows = torch.div(torch.arange(size[0], device=saliency.device) * H, size[0], rounding_mode="floor")
    cols = torch.div(torch.arange(size[1], device=saliency.device) * W, size[1], rounding_mode="floor")
    thumb = saliency[:, rows][:, :, cols].to(torch.float32)
    return (2.0 * thumb.clamp(0.0, 1.0) - 1.0).reshape(B, -1)


class Retriever:
    """``image2layout.train.models.retrieval.retriever.Retriever`` (retriever.py:24-229): the reference's non-learnable
    generator ("copy the layout of the nearest database canvas") and the builder of its retrieval cache tables -- the class
    SURVEY.md 8 row a1 names.  Same constructor / ``sample`` / ``preprocess_retrieval_cache`` surface, with the FAISS index
    replaced by a device-resident gallery + ``ralf_knn_topk`` and the per-sample python loops by batched calls.

    Features: ``retrieval_backbone="saliency"`` (and ``"random"``, which the reference maps onto it) computes the weight-free
    ``coarse_saliency`` features here.  The deep backbones (dreamsim / clip / vgg) are third-party pretrained embedders that
    are out of scope: pass their outputs as ``embeddings`` [N, d] (database order) and an ``embed_fn`` that
    maps a batch of images [B, 3, H, W] to [B, d]."""

    output_keys = ["label", "mask", "center_x", "center_y", "width", "height"]

    def __init__(self, features, db_dataset, max_seq_length: int, top_k: int = 1, dataset_name: str = "pku",
                 retrieval_backbone: str = "saliency", saliency_k=None, embeddings=None, embed_fn=None, device=None,
                 **kwargs) -> None:
        from .data import LayoutTable

        self.features, self.db_dataset, self.max_seq_length = features, db_dataset, max_seq_length
        self.top_k, self.dataset_name, self.retrieval_backbone = top_k, dataset_name, retrieval_backbone
        self.index_name = "search_feat"
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else "cuda"
        self.device = torch.device(device)
        self.embed_fn = embed_fn
        backbone = "saliency" if retrieval_backbone == "random" else retrieval_backbone
        if "merge" in backbone or "concat" in backbone:
            raise NotImplementedError(f"{retrieval_backbone=}: merged retrieval tables are built offline in the reference")
        rows = [db_dataset[i] for i in range(len(db_dataset))]
        if embeddings is not None:
            vectors = torch.as_tensor(embeddings, dtype=torch.float32)
        elif backbone == "saliency":
            vectors = torch.cat([coarse_saliency(torch.as_tensor(r["saliency"])[None]) for r in rows])
        else:
            raise NotImplementedError(f"{retrieval_backbone=}: pass embeddings= and embed_fn= (pretrained embedders are out of scope)")
        assert vectors.shape[0] == len(rows), f"{vectors.shape[0]} feature rows for {len(rows)} database rows"
        self.layouts = LayoutTable.from_rows(rows, max_seq_length, device=self.device)
        self.retr = GpuRetriever(vectors.to(self.device), device=self.device)
        self.db_ids = [r.get("id", i) for i, r in enumerate(rows)]
        self.table_paired_id_idx = {self._id(i): n for n, i in enumerate(self.db_ids)}

    def _id(self, data_id):
        return int(data_id) if "pku" in self.dataset_name else data_id  # retriever.py:176-178

    def eval(self):
        return self

    def to(self, device):
        return self

    def parameters(self):
        return iter(())

    def get_query(self, image=None, saliency=None) -> torch.Tensor:
        """FeatureExtracterBackbone.get_query (retrieval/image.py:132-139), batched: [B, d] on the device."""
        if self.retrieval_backbone in ("saliency", "random"):
            return coarse_saliency(saliency.to(self.device))
        assert self.embed_fn is not None, "deep retrieval backbones need embed_fn="
        return torch.as_tensor(self.embed_fn(image), dtype=torch.float32).to(self.device)

    @torch.no_grad()
    def sample(self, cond, batch_size=1, sampling_cfg=None, **kwargs):
        """retriever.py:91-132: the layout of the nearest database canvas for every canvas of ``cond``; ``random`` queries
        with the saliency map of a random database row instead (one ``np.random.randint`` per canvas, as there)."""
        import numpy as np

        image = cond.image
        B = image.size(0)
        if self.retrieval_backbone == "random":
            picks = [int(np.random.randint(0, len(self.db_dataset))) for _ in range(B)]
            sal = torch.stack([torch.as_tensor(self.db_dataset[i]["saliency"]) for i in picks])
            query = self.get_query(saliency=sal)
        else:
            query = self.get_query(image=image[:, :-1], saliency=image[:, -1:])
        idx, _ = self.retr.search(query, 1, certify=True)
        got = self.layouts.gather(idx)
        out = {k: got[k][:, 0].cpu() for k in self.output_keys}
        return out, {"total": 1, "viorated": 0}

    def preprocess_retrieval_cache(self, split: str, dataset, top_k: int, run_on_local: bool = True,
                                   save_scores: bool = False, root: str = "cache", queries=None, batch: int = 1024):
        """retriever.py:134-229: ``dict[data_id -> list[database index]]`` of the ``top_k`` (+1, minus the query itself on the
        train split) nearest database rows of every row of ``dataset``, saved under the reference's file names.
        ``queries`` [len(dataset), d] skips the feature extraction (deep backbones)."""
        from .data import cache_table_path, paired_table_path, save_cache_table

        os.makedirs(root, exist_ok=True)
        torch.save(dict(self.table_paired_id_idx), paired_table_path(self.dataset_name, self.retrieval_backbone, root))
        table, scores = {}, {}
        for s in range(0, len(dataset), batch):
            rows = [dataset[i] for i in range(s, min(len(dataset), s + batch))]
            if queries is not None:
                q = torch.as_tensor(queries[s:s + len(rows)], dtype=torch.float32).to(self.device)
            elif self.retrieval_backbone in ("saliency", "random"):
                q = self.get_query(saliency=torch.stack([torch.as_tensor(r["saliency"]) for r in rows]))
            else:
                q = self.get_query(image=torch.stack([torch.as_tensor(r["image"]) for r in rows]))
            idx, score = self.retr.search(q, top_k + 1, certify=True)
            idx, score = idx.cpu().tolist(), score.cpu().tolist()
            for r, i_row, s_row in zip(rows, idx, score):
                if split == "train":  # the first hit is the query itself
                    i_row, s_row = i_row[1:], s_row[1:]
                table[self._id(r["id"])] = i_row
                scores[self._id(r["id"])] = s_row
        path = cache_table_path(self.dataset_name, split, self.retrieval_backbone, top_k, root)
        save_cache_table(table, path)
        if save_scores:
            torch.save(scores, path.replace("indexes", "scores"))
        return table
