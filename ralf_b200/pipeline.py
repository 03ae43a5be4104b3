"""Public end-to-end API: query embeddings + canvases -> retrieved exemplars -> generated layout tokens.

``LayoutPipeline`` chains :class:`ralf_b200.retrieval.GpuRetriever` (top-k search + exemplar fetch) and a
:mod:`ralf_b200.generator` model (encode + KV-cached greedy decode) for a fixed batch shape and captures the
whole chain -- several thousand small kernel launches -- in CUDA graphs, so a step costs one (two when the
gallery is sharded) graph launch instead of thousands of driver calls.  This is the call `bench.py` times for
both `value` (device-resident inputs) and `e2e` (pinned host inputs, H2D and D2H inside the region).

It is what `inference.py:387-443` does per batch in the reference (table lookup -> `model.sample`), with the
lookup replaced by the live k-NN search the north star asks for.
"""
from __future__ import annotations

import contextlib
import os
from typing import Optional

import torch

from .generator import ConditionalInputs
from .retrieval import GpuRetriever

_NVTX = os.environ.get("RALF_NVTX", "0") != "0"


@contextlib.contextmanager
def _stage(name: str):
    """NVTX range around a pipeline stage (RALF_NVTX=1; shows up as search / fetch / encode[i] / decode in a timeline of
    the eager path or of the capture -- host-side markers, nothing is added to the graphs)."""
    if not _NVTX:
        yield
        return
    torch.cuda.nvtx.range_push("ralf." + name)
    try:
        yield
    finally:
        torch.cuda.nvtx.range_pop()


class LayoutPipeline:
    def __init__(self, model, retriever: GpuRetriever, batch: int, height: int, width: int, *, top_k: int = 16,
                 emb_dim: int = 512, use_graph: bool = True, micro_batch: int = 128, decode_ways: int = 1) -> None:
        """``micro_batch``: canvases per encode pass.  Retrieval and the decode loop run over the whole batch; the
        ResNet/encoder activations (im2col buffers, ~60 MB per canvas) only ever exist for one micro-batch, whose
        memory K/V rows land in the batch-wide cross-attention cache.
        ``decode_ways`` (opt-in, unmeasured): the decode loop is a chain of ~70 small dependent kernels per token whose
        grids cover a fraction of the SMs; canvases are independent, so the batch can be cut into ``decode_ways`` groups
        whose chains run on parallel streams (parallel branches of the captured graph) and fill each other's idle SMs.
        Results per canvas do not depend on the grouping."""
        self.model, self.retr = model, retriever
        self.decode_ways = max(1, min(int(decode_ways), batch))
        self._dec_streams = [torch.cuda.Stream(device=model.device) for _ in range(self.decode_ways)] \
            if self.decode_ways > 1 else []
        self.B, self.H, self.W, self.k = batch, height, width, top_k
        self.mb = min(micro_batch, batch)
        self.dev = model.device
        self.world, self.rank = retriever.world, retriever.rank
        self.eng = model.engine()
        tok = model.tokenizer
        self.S = tok.max_token_length
        self.ids = model.special_token_ids
        self.token_mask = tok.token_mask.to(self.dev).to(torch.uint8).contiguous()
        # static inputs / outputs
        self.img = torch.zeros(batch, 4, height, width, device=self.dev)
        self.qry = torch.zeros(batch, emb_dim, device=self.dev)
        self.q_all = torch.zeros(batch * self.world, emb_dim, device=self.dev) if self.world > 1 else self.qry
        const = model.preprocessor(ConditionalInputs(image=self.img))
        self.const_seq = const["seq"].to(self.dev).contiguous()
        self.const_pad = const["pad_mask"].to(self.dev).to(torch.uint8).contiguous()
        self.seq_out = torch.zeros(batch, self.S, dtype=torch.int64, device=self.dev)
        self.idx_out = torch.zeros(batch, top_k, dtype=torch.int64, device=self.dev)
        self.g_search: Optional[torch.cuda.CUDAGraph] = None
        self.g_fetch: Optional[torch.cuda.CUDAGraph] = None
        self.g_enc: list = []
        self.g_dec: Optional[torch.cuda.CUDAGraph] = None
        self.kv, self.Mlen = None, 0
        self.copy_stream = torch.cuda.Stream(device=self.dev)
        self.use_graph = use_graph
        self.kernels_per_step = 0
        self._local = None  # (idx, score) of the local search when sharded
        self._gathered = None
        if use_graph:
            self._capture()

    # ---- stages --------------------------------------------------------------------------------
    def _stage_search(self):
        with _stage("search"):
            idx, score = self.retr.search_local(self.q_all, self.k)
        return idx, score

    def _stage_fetch(self, idx: torch.Tensor):
        """idx: global top-k of this rank's canvases [B, k] -> packed exemplar layouts [B, k, 6, E]."""
        with _stage("fetch"):
            self.idx_out.copy_(idx)
            return self.retr.fetch(idx)["packed"]

    def _stage_encode(self, packed: torch.Tensor, b0: int) -> None:
        """One encoder micro-batch (canvases b0 .. b0+mb): memory -> rows of the batch-wide cross-attention K/V cache."""
        b1 = min(self.B, b0 + self.mb)
        with _stage(f"encode[{b0 // self.mb}]"):
            mem, mem_s = self.eng.encode(self.img[b0:b1], packed[b0:b1], self.const_seq[b0:b1], self.const_pad[b0:b1])
            self.Mlen = mem.shape[1]
            if self.kv is None:
                from .engine import KV24

                self.kv = self.eng.alloc_cross_kv(self.B * self.Mlen, kv24=KV24 and self.eng.npass == 3)
            self.eng.cross_kv(mem_s, out=self.kv, row0=b0 * self.Mlen)

    def _stage_decode(self) -> None:
        with _stage("decode"):
            self._decode()

    def _decode(self) -> None:
        if self.decode_ways == 1:
            seq = self.eng.generate(None, self.B, self.Mlen, self.token_mask, self.ids["bos"], self.ids["pad"], self.S,
                                    kv=self.kv)
            self.seq_out.copy_(seq)
            return
        cur = torch.cuda.current_stream()
        per = (self.B + self.decode_ways - 1) // self.decode_ways
        forked = []
        for w, side in enumerate(self._dec_streams):
            b0, b1 = w * per, min(self.B, (w + 1) * per)
            if b0 >= b1:
                break
            side.wait_stream(cur)  # fork (inside a capture this makes `side` a branch of the same graph)
            forked.append(side)
            with torch.cuda.stream(side):
                kv = [k[b0 * self.Mlen:b1 * self.Mlen] for k in self.kv]
                seq = self.eng.generate(None, b1 - b0, self.Mlen, self.token_mask, self.ids["bos"], self.ids["pad"],
                                        self.S, kv=kv)
                self.seq_out[b0:b1].copy_(seq)
        for side in forked:  # join -- only the branches that exist (a capture must not wait on a stream outside it)
            cur.wait_stream(side)

    def _stage_main(self, idx: torch.Tensor):
        packed = self._stage_fetch(idx)
        for b0 in range(0, self.B, self.mb):
            self._stage_encode(packed, b0)
        self._stage_decode()

    def _merge(self, idx, score):
        from . import ops
        from .retrieval import exchange_and_merge

        gi, _ = exchange_and_merge(idx, score, self.world, self.retr.pg, ops.knn_merge)
        return gi[self.rank * self.B:(self.rank + 1) * self.B]

    def _eager(self):
        idx, score = self._stage_search()
        if self.world > 1:
            idx = self._merge(idx, score)
        self._stage_main(idx)

    def _capture(self):
        # warm-up outside capture: sets kernel attributes, fills the tensor-map cache, sizes the allocator
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                self._eager()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        # two graphs: the k-NN phase (scan + re-rank) and everything after it.  Splitting them costs one extra
        # graph launch and lets bench.py time the k-NN kernel live with CUDA events between the two replays;
        # with a sharded gallery the NCCL all-gather + merge sits between them anyway.
        from . import ops

        n0 = ops.launch_count()
        self.g_search = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_search):
            self._local = self._stage_search()
        self._my_idx = self._local[0] if self.world == 1 else torch.zeros(self.B, self.k, dtype=torch.int64,
                                                                         device=self.dev)
        # everything after the search: fetch, one graph per encoder micro-batch (so the host->device copy of micro-batch
        # i+1 can run under the encode of micro-batch i, see __call__), the decode loop.  The graphs share one memory
        # pool and are always replayed in capture order.
        pool = self.g_search.pool()
        self.g_fetch = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_fetch, pool=pool):
            self._packed = self._stage_fetch(self._my_idx)
        self.g_enc = []
        for b0 in range(0, self.B, self.mb):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=pool):
                self._stage_encode(self._packed, b0)
            self.g_enc.append(g)
        self.g_dec = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_dec, pool=pool):
            self._stage_decode()
        # kernels of OURS recorded into the two graphs = kernels launched per replayed step
        self.kernels_per_step = ops.launch_count() - n0 + (2 if self.world > 1 else 0)
        torch.cuda.synchronize()

    # ---- run -----------------------------------------------------------------------------------
    def step(self, events: Optional[list] = None, copy_events: Optional[list] = None,
             phase_events: Optional[list] = None) -> torch.Tensor:
        """One pass over the static inputs (self.img, self.qry already filled).  Returns token ids [B, S] (device).
        ``events``: optional list that receives a (start, end) CUDA-event pair around the k-NN phase.
        ``phase_events`` (graph path): receives (name, event) marks -- start, search, fetch, encode, decode -- so a caller
        can read the phase split of a step with CUDA events."""

        def mark(name):
            if phase_events is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                phase_events.append((name, ev))

        mark("start")
        if self.world > 1:
            import torch.distributed as dist

            dist.all_gather_into_tensor(self.q_all, self.qry, group=self.retr.pg)
        e0 = e1 = None
        if events is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        if not self.use_graph:
            idx, score = self._stage_search()
            if events is not None:
                e1.record()
                events.append((e0, e1))
            if self.world > 1:
                idx = self._merge(idx, score)
            self._stage_main(idx)
            return self.seq_out
        self.g_search.replay()
        if events is not None:
            e1.record()
            events.append((e0, e1))
        if self.world > 1:
            self._my_idx.copy_(self._merge(*self._local))
        mark("search")
        self.g_fetch.replay()
        mark("fetch")
        for i, g in enumerate(self.g_enc):
            if copy_events is not None:  # micro-batch i's canvases must have landed
                torch.cuda.current_stream().wait_event(copy_events[i])
            g.replay()
        mark("encode")
        self.g_dec.replay()
        mark("decode")
        return self.seq_out

    def __call__(self, image: torch.Tensor, query: torch.Tensor) -> torch.Tensor:
        """image [B,4,H,W] and query [B,d] on host (pinned) or device -> token ids [B, S] on the device.
        The canvases are copied micro-batch by micro-batch on a side stream; the search and the encoder graphs of the
        earlier micro-batches run underneath the later copies."""
        main = torch.cuda.current_stream()
        self.qry.copy_(query, non_blocking=True)
        if not self.use_graph:
            self.img.copy_(image, non_blocking=True)
            return self.step()
        self.copy_stream.wait_stream(main)  # the previous step has consumed self.img
        evs = []
        with torch.cuda.stream(self.copy_stream):
            for b0 in range(0, self.B, self.mb):
                self.img[b0:b0 + self.mb].copy_(image[b0:b0 + self.mb], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
                evs.append(ev)
        return self.step(copy_events=evs)

    def generate_layouts(self, image: torch.Tensor, query: torch.Tensor) -> dict:
        """Host-facing call: returns the decoded layout dict on the CPU like ``model.sample`` does
        (label, mask, center_x, center_y, width, height) plus the token ids and retrieved indices."""
        seq = self(image, query).cpu()
        out = self.model.tokenizer.decode(seq)
        out["seq"] = seq
        out["retrieved_idx"] = self.idx_out.cpu()
        return out


class _Slot:
    """One of the two batches in flight of :class:`OverlappedPipeline`."""

    def __init__(self, kv: list, seq: torch.Tensor, idx: torch.Tensor, enc: list, dec: torch.cuda.CUDAGraph) -> None:
        self.kv, self.seq, self.idx, self.enc, self.dec = kv, seq, idx, enc, dec
        self.enc_done, self.dec_done, self.out_ready = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
        self.host_seq = torch.zeros(seq.shape, dtype=seq.dtype).pin_memory()
        self.host_idx = torch.zeros(idx.shape, dtype=idx.dtype).pin_memory()
        self.busy = False  # submitted and not yet collected (host buffers in use)


class OverlappedPipeline(LayoutPipeline):
    """Two batches in flight: the decode loop of batch i runs on its own stream while search, fetch and the encoder
    micro-batches of batch i+1 run on the caller's stream.

    Why: the decode loop is latency / HBM bound (4 k launches with grids of a few dozen CTAs, plus the K/V stream), the
    encoder is tensor bound; run back to back each leaves most of the other's resource idle ("what comes
    next" item 2).  Nothing changes per canvas -- same graphs, same kernels, same results as :class:`LayoutPipeline`
    (tests/test_pipeline_gpu.py) -- only the order in which the device sees them.

    What that needs: two K/V caches + output buffers (slots, alternating by step), encoder graphs and a decode graph per
    slot (graphs bake addresses in), and a SEPARATE graph memory pool for the decode graphs: graphs that share a pool
    reuse each other's scratch memory, which is only safe while they never run concurrently.
    Ordering: encode(i+1) -> slot s waits for decode(i-1), the last reader of that slot's K/V cache; decode(i) waits
    for encode(i).  ``submit`` enqueues one batch and returns its slot; ``collect(slot)`` hands back its results;
    ``drain`` joins the decode stream into the caller's stream.  ``step`` / ``__call__`` / ``generate_layouts`` keep the
    parent's blocking semantics (submit + wait), so the class is a drop-in for it."""

    def __init__(self, *args, **kwargs) -> None:
        kwargs["use_graph"] = False  # the parent must not capture its single-slot graphs
        super().__init__(*args, **kwargs)
        self.use_graph = True
        self.dec_stream = torch.cuda.Stream(device=self.dev)
        self.slots: list = []
        self._n = 0
        self._capture_overlapped()

    def _capture_overlapped(self) -> None:
        from . import ops

        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                self._eager()  # warm-up; allocates slot 0's K/V cache and fixes self.Mlen
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        n0 = ops.launch_count()
        self.g_search = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_search):
            self._local = self._stage_search()
        self._my_idx = self._local[0] if self.world == 1 else torch.zeros(self.B, self.k, dtype=torch.int64,
                                                                         device=self.dev)
        pool = self.g_search.pool()
        self.g_fetch = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_fetch, pool=pool):
            self._packed = self._stage_fetch(self._my_idx)
        n_front = ops.launch_count() - n0
        dec_pool = torch.cuda.graph_pool_handle()
        kv0 = self.kv
        kv1 = self.eng.alloc_cross_kv(self.B * self.Mlen, kv24=kv0[0].dtype == torch.uint8)
        n_slot = 0
        for kv, seq in ((kv0, self.seq_out), (kv1, torch.zeros_like(self.seq_out))):
            self.kv, self.seq_out = kv, seq  # the stage functions read these; capture bakes the addresses in
            n1 = ops.launch_count()
            enc = []
            for b0 in range(0, self.B, self.mb):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool):
                    self._stage_encode(self._packed, b0)
                enc.append(g)
            dec = torch.cuda.CUDAGraph()
            with torch.cuda.graph(dec, pool=dec_pool):
                self._stage_decode()
            n_slot = ops.launch_count() - n1
            self.slots.append(_Slot(kv, seq, torch.zeros_like(self.idx_out), enc, dec))
        self.kv, self.seq_out = self.slots[0].kv, self.slots[0].seq
        self.kernels_per_step = n_front + n_slot + (2 if self.world > 1 else 0)
        torch.cuda.synchronize()

    # ---- asynchronous interface ------------------------------------------------------------------
    def submit(self, events: Optional[list] = None, copy_events: Optional[list] = None) -> int:
        """Enqueue one batch over the static inputs (self.img / self.qry); returns the slot that will hold its results."""
        main = torch.cuda.current_stream()
        slot_id = self._n % 2
        slot = self.slots[slot_id]
        assert not slot.busy, "collect() the batch submitted two steps ago before submitting another one"
        if self.world > 1:
            import torch.distributed as dist

            dist.all_gather_into_tensor(self.q_all, self.qry, group=self.retr.pg)
        if events is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        self.g_search.replay()
        if events is not None:
            e1.record()
            events.append((e0, e1))
        if self.world > 1:
            self._my_idx.copy_(self._merge(*self._local))
        self.g_fetch.replay()
        main.wait_event(slot.dec_done)  # the decode of two steps ago was the last reader of this slot's K/V cache
        slot.idx.copy_(self.idx_out)
        for i, g in enumerate(slot.enc):
            if copy_events is not None:
                main.wait_event(copy_events[i])
            g.replay()
        slot.enc_done.record(main)
        self.dec_stream.wait_event(slot.enc_done)
        with torch.cuda.stream(self.dec_stream):
            slot.dec.replay()
            slot.dec_done.record(self.dec_stream)
        self._n += 1
        return slot_id

    def submit_host(self, image: torch.Tensor, query: torch.Tensor) -> int:
        """``submit`` with host (pinned) inputs: canvases are copied per micro-batch on the copy stream (parent's
        ``__call__``), so the copy of batch i+1 also runs under the decode of batch i."""
        main = torch.cuda.current_stream()
        self.qry.copy_(query, non_blocking=True)
        self.copy_stream.wait_stream(main)  # the encoder graphs of the previous batch have consumed self.img
        evs = []
        with torch.cuda.stream(self.copy_stream):
            for b0 in range(0, self.B, self.mb):
                self.img[b0:b0 + self.mb].copy_(image[b0:b0 + self.mb], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self.copy_stream)
                evs.append(ev)
        slot_id = self.submit(copy_events=evs)
        slot = self.slots[slot_id]
        with torch.cuda.stream(self.dec_stream):  # device -> host of the results right behind the decode
            slot.host_seq.copy_(slot.seq, non_blocking=True)
            slot.host_idx.copy_(slot.idx, non_blocking=True)
            slot.out_ready.record(self.dec_stream)
        slot.busy = True
        return slot_id

    def collect(self, slot_id: int) -> dict:
        """Results of a batch submitted with ``submit_host``: decoded layout dict on the CPU (like ``generate_layouts``)."""
        slot = self.slots[slot_id]
        assert slot.busy, "nothing submitted into this slot"
        slot.out_ready.synchronize()
        seq = slot.host_seq.clone()
        out = self.model.tokenizer.decode(seq)
        out["seq"] = seq
        out["retrieved_idx"] = slot.host_idx.clone()
        slot.busy = False
        return out

    def drain(self) -> None:
        """Make the caller's stream wait for every decode in flight."""
        main = torch.cuda.current_stream()
        for slot in self.slots:
            main.wait_event(slot.dec_done)

    # ---- blocking interface of the parent ---------------------------------------------------------
    def step(self, events: Optional[list] = None, copy_events: Optional[list] = None) -> torch.Tensor:
        slot = self.slots[self.submit(events=events, copy_events=copy_events)]
        torch.cuda.current_stream().wait_event(slot.dec_done)
        self.idx_out.copy_(slot.idx)
        return slot.seq
