"""Batched tensor API over the C ABI: the throughput path beside the drop-in retrievers.

``NodeWithScore`` lists cannot be produced at 100k queries/s, so the
retriever classes in :mod:`easyrag_b200.retrievers` are thin per-query views over this module.
Everything here takes and returns torch CUDA tensors used purely as device buffers; the
arithmetic is in easyrag_b200/csrc/*.cu.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from .index import Bm25Index, DenseIndex


def _i32(t: Optional[torch.Tensor], device) -> Optional[torch.Tensor]:
    if t is None:
        return None
    return t.to(device=device, dtype=torch.int32).contiguous()


class Workspace:
    """Grow-only device scratch buffer (launch functions never allocate)."""

    def __init__(self, device):
        self.device = device
        self.buf = torch.empty(0, dtype=torch.uint8, device=device)

    def get(self, nbytes: int) -> torch.Tensor:
        if self.buf.numel() < nbytes:
            self.buf = torch.empty(int(nbytes * 1.25) + 256, dtype=torch.uint8, device=self.device)
        return self.buf


@dataclass
class TopK:
    scores: torch.Tensor     # [Q, k] float32 (dense, bm25s) or float64 (Okapi, fusion)
    ids: torch.Tensor        # [Q, k] int32, -1 padded
    counts: torch.Tensor     # [Q] int32


def dense_topk(index: DenseIndex, queries: torch.Tensor, k: int, q_group: Optional[torch.Tensor] = None,
               id_base: Optional[int] = None, ws: Optional[Workspace] = None, stream=None,
               out: Optional[TopK] = None) -> TopK:
    """QdrantRetriever._aretrieve's search for a batch (retrievers.py:37-52): cosine top-k, ids descending on ties."""
    L = _lib.lib()
    dev = index.device
    q = queries
    if q.dtype != torch.bfloat16 or q.device != dev or not q.is_contiguous():
        q = queries.to(device=dev, dtype=torch.bfloat16).contiguous()
    nq, dim = q.shape
    if dim != index.dim:
        raise ValueError(f"query dim {dim} != corpus dim {index.dim}")
    qg = _i32(q_group, dev)
    if qg is not None and index.doc_group is None:
        raise ValueError("q_group given but the index has no doc_group")
    if out is None:
        out = TopK(torch.empty(nq, k, dtype=torch.float32, device=dev), torch.empty(nq, k, dtype=torch.int32, device=dev),
                   torch.empty(nq, dtype=torch.int32, device=dev))
    need = L.ezr_dense_topk_workspace(index.n_rows, dim, nq, k)
    ws = ws or Workspace(dev)
    buf = ws.get(need)
    base = index.row_lo if id_base is None else id_base
    with torch.cuda.device(dev):
        _lib.check(L.ezr_dense_topk(_lib.ptr(index.vectors), index.n_rows, dim, index.vectors.stride(0), _lib.ptr(q), nq,
                                    q.stride(0), k, _lib.ptr(index.doc_group if qg is not None else None), _lib.ptr(qg),
                                    base, _lib.ptr(out.scores), _lib.ptr(out.ids), _lib.ptr(out.counts), _lib.ptr(buf),
                                    buf.numel(), _lib.stream_ptr(stream)), "ezr_dense_topk")
    return out


def bm25_topk(index: Bm25Index, q_ptr: torch.Tensor, q_terms: torch.Tensor, k: int,
              q_group: Optional[torch.Tensor] = None, id_base: Optional[int] = None,
              ws: Optional[Workspace] = None, stream=None, out: Optional[TopK] = None) -> TopK:
    """BM25Retriever.get_scores + filter for a batch (retrievers.py:128-151,191-210)."""
    L = _lib.lib()
    dev = index.device
    qp, qt = _i32(q_ptr, dev), _i32(q_terms, dev)
    nq = qp.numel() - 1
    qg = _i32(q_group, dev)
    if qg is not None and index.doc_group is None:
        raise ValueError("q_group given but the index has no doc_group")
    if out is None:
        out = TopK(torch.empty(nq, k, dtype=index.score_dtype, device=dev),
                   torch.empty(nq, k, dtype=torch.int32, device=dev), torch.empty(nq, dtype=torch.int32, device=dev))
    need = L.ezr_bm25_topk_workspace(index.struct, nq, k)
    ws = ws or Workspace(dev)
    buf = ws.get(need)
    base = index.doc_lo if id_base is None else id_base
    with torch.cuda.device(dev):
        _lib.check(L.ezr_bm25_topk(index.struct, _lib.ptr(qp), _lib.ptr(qt), nq, k, _lib.ptr(qg), base,
                                   _lib.ptr(out.scores), _lib.ptr(out.ids), _lib.ptr(out.counts), _lib.ptr(buf),
                                   buf.numel(), _lib.stream_ptr(stream)), "ezr_bm25_topk")
    return out


def bm25_scores(index: Bm25Index, q_ptr: torch.Tensor, q_terms: torch.Tensor, stream=None) -> torch.Tensor:
    """BM25Retriever.get_scores (retrievers.py:128-151): [Q, n_docs] score rows."""
    L = _lib.lib()
    dev = index.device
    qp, qt = _i32(q_ptr, dev), _i32(q_terms, dev)
    nq = qp.numel() - 1
    out = torch.empty(nq, index.n_docs, dtype=index.score_dtype, device=dev)
    with torch.cuda.device(dev):
        _lib.check(L.ezr_bm25_scores(index.struct, _lib.ptr(qp), _lib.ptr(qt), nq, _lib.ptr(out),
                                     _lib.stream_ptr(stream)), "ezr_bm25_scores")
    return out


def select_rows(scores: torch.Tensor, k: int, positive_only: bool = False, doc_group: Optional[torch.Tensor] = None,
                q_group: Optional[torch.Tensor] = None, id_base: int = 0, ws: Optional[Workspace] = None,
                stream=None) -> TopK:
    L = _lib.lib()
    dev = scores.device
    assert scores.dim() == 2 and scores.stride(1) == 1
    st = _lib.F64 if scores.dtype == torch.float64 else _lib.F32
    if scores.dtype not in (torch.float64, torch.float32):
        raise TypeError("scores must be float32 or float64")
    nq, n = scores.shape
    out = TopK(torch.empty(nq, k, dtype=scores.dtype, device=dev), torch.empty(nq, k, dtype=torch.int32, device=dev),
               torch.empty(nq, dtype=torch.int32, device=dev))
    need = L.ezr_select_rows_workspace(nq, n, k, st)
    ws = ws or Workspace(dev)
    buf = ws.get(need)
    dg, qg = _i32(doc_group, dev), _i32(q_group, dev)
    with torch.cuda.device(dev):
        _lib.check(L.ezr_select_rows(_lib.ptr(scores), st, nq, n, scores.stride(0), k, int(positive_only), _lib.ptr(dg),
                                     _lib.ptr(qg), id_base, _lib.ptr(out.scores), _lib.ptr(out.ids), _lib.ptr(out.counts),
                                     _lib.ptr(buf), buf.numel(), _lib.stream_ptr(stream)), "ezr_select_rows")
    return out


def merge_topk_parts(cand_scores: torch.Tensor, cand_ids: torch.Tensor, n_parts: int, part_stride_bytes: int, k: int,
                     out: Optional[TopK] = None, stream=None) -> TopK:
    """Merge ``n_parts`` per-shard top-k lists per query, read in place from an all-gathered byte record.

    ``cand_scores`` / ``cand_ids``: [Q, k_in] views of part 0 inside the gathered buffer; part p of the same
    arrays lies ``p * part_stride_bytes`` further (easyrag_b200/dist.py)."""
    L = _lib.lib()
    dev = cand_scores.device
    assert cand_scores.shape == cand_ids.shape and cand_scores.stride(1) == 1 and cand_ids.stride(1) == 1
    assert cand_ids.dtype == torch.int32 and cand_scores.stride(0) == cand_ids.stride(0)
    st = _lib.F64 if cand_scores.dtype == torch.float64 else _lib.F32
    nq, c = cand_scores.shape
    if out is None:
        out = TopK(torch.empty(nq, k, dtype=cand_scores.dtype, device=dev),
                   torch.empty(nq, k, dtype=torch.int32, device=dev), torch.empty(nq, dtype=torch.int32, device=dev))
    with torch.cuda.device(dev):
        _lib.check(L.ezr_merge_topk_parts(_lib.ptr(cand_scores), _lib.ptr(cand_ids), st, nq, c, cand_scores.stride(0),
                                          n_parts, part_stride_bytes, k, _lib.ptr(out.scores), _lib.ptr(out.ids),
                                          _lib.ptr(out.counts), _lib.stream_ptr(stream)), "ezr_merge_topk_parts")
    return out


def merge_topk(cand_scores: torch.Tensor, cand_ids: torch.Tensor, k: int, stream=None) -> TopK:
    """Merge candidate lists [Q, C] (id < 0 = empty) into the canonical top-k."""
    L = _lib.lib()
    dev = cand_scores.device
    assert cand_scores.shape == cand_ids.shape and cand_scores.is_contiguous() and cand_ids.is_contiguous()
    st = _lib.F64 if cand_scores.dtype == torch.float64 else _lib.F32
    nq, c = cand_scores.shape
    cand_ids = cand_ids.to(torch.int32)
    out = TopK(torch.empty(nq, k, dtype=cand_scores.dtype, device=dev), torch.empty(nq, k, dtype=torch.int32, device=dev),
               torch.empty(nq, dtype=torch.int32, device=dev))
    with torch.cuda.device(dev):
        _lib.check(L.ezr_merge_topk(_lib.ptr(cand_scores), _lib.ptr(cand_ids), st, nq, c, c, k,
                                    _lib.ptr(out.scores), _lib.ptr(out.ids), _lib.ptr(out.counts), None, 0,
                                    _lib.stream_ptr(stream)), "ezr_merge_topk")
    return out


def rrf_fuse(ids_a: torch.Tensor, cnt_a: torch.Tensor, ids_b: torch.Tensor, cnt_b: torch.Tensor, k_out: int,
             K: int = 60, canon: Optional[torch.Tensor] = None, stream=None, out: Optional[TopK] = None) -> TopK:
    """HybridRetriever.reciprocal_rank_fusion (retrievers.py:256-274) for a batch; list a = sparse, b = dense."""
    L = _lib.lib()
    dev = ids_a.device
    assert ids_a.shape == ids_b.shape and ids_a.dtype == torch.int32 and ids_b.dtype == torch.int32
    nq, stride = ids_a.shape
    if out is None:
        out = TopK(torch.empty(nq, k_out, dtype=torch.float64, device=dev),
                   torch.empty(nq, k_out, dtype=torch.int32, device=dev), torch.empty(nq, dtype=torch.int32, device=dev))
    with torch.cuda.device(dev):
        _lib.check(L.ezr_rrf_fuse(_lib.ptr(ids_a.contiguous()), _lib.ptr(cnt_a), _lib.ptr(ids_b.contiguous()),
                                  _lib.ptr(cnt_b), nq, stride, _lib.ptr(canon), 0, K, k_out, _lib.ptr(out.ids),
                                  _lib.ptr(out.scores), _lib.ptr(out.counts), _lib.stream_ptr(stream)), "ezr_rrf_fuse")
    return out


def fusion_simple(ids_a: torch.Tensor, sc_a: torch.Tensor, cnt_a: torch.Tensor, ids_b: torch.Tensor, sc_b: torch.Tensor,
                  cnt_b: torch.Tensor, k_out: int, canon: Optional[torch.Tensor] = None, stream=None) -> TopK:
    """HybridRetriever.fusion (retrievers.py:239-253) for a batch."""
    L = _lib.lib()
    dev = ids_a.device
    nq, stride = ids_a.shape
    sa = sc_a.to(torch.float64).contiguous()
    sb = sc_b.to(torch.float64).contiguous()
    out = TopK(torch.empty(nq, k_out, dtype=torch.float64, device=dev),
               torch.empty(nq, k_out, dtype=torch.int32, device=dev), torch.empty(nq, dtype=torch.int32, device=dev))
    with torch.cuda.device(dev):
        _lib.check(L.ezr_fusion_simple(_lib.ptr(ids_a.contiguous()), _lib.ptr(sa), _lib.ptr(cnt_a),
                                       _lib.ptr(ids_b.contiguous()), _lib.ptr(sb), _lib.ptr(cnt_b), nq, stride,
                                       _lib.ptr(canon), 0, k_out, _lib.ptr(out.ids), _lib.ptr(out.scores),
                                       _lib.ptr(out.counts), _lib.stream_ptr(stream)), "ezr_fusion_simple")
    return out


def fuse_lists(ids: Sequence[torch.Tensor], counts: Sequence[torch.Tensor], k_out: int, rrf: bool = True, K: int = 60,
               scores: Optional[Sequence[torch.Tensor]] = None, canon: Optional[torch.Tensor] = None, stream=None) -> TopK:
    """``reciprocal_rank_fusion`` / ``fusion`` over ANY number of rank lists (retrievers.py:239-274 loop over a list
    of lists).  ``ids[l]`` int32 [Q, width], ``counts[l]`` int32 [Q], ``scores[l]`` float64 [Q, width] (fusion only);
    all lists share ``width``."""
    import ctypes as C
    L = _lib.lib()
    n = len(ids)
    dev = ids[0].device
    nq, width = ids[0].shape
    ids = [t.contiguous() for t in ids]
    assert all(t.shape == (nq, width) and t.dtype == torch.int32 for t in ids) and len(counts) == n
    sc = None
    if not rrf:
        sc = [t.to(torch.float64).contiguous() for t in scores]
        assert all(t.shape == (nq, width) for t in sc)
    out = TopK(torch.empty(nq, k_out, dtype=torch.float64, device=dev),
               torch.empty(nq, k_out, dtype=torch.int32, device=dev), torch.empty(nq, dtype=torch.int32, device=dev))
    arr = lambda ts: (C.c_void_p * n)(*[t.data_ptr() for t in ts])
    with torch.cuda.device(dev):
        _lib.check(L.ezr_fuse_lists(int(rrf), n, arr(ids), arr(sc) if sc is not None else None, arr(counts), nq, width,
                                    _lib.ptr(canon), 0, K, k_out, _lib.ptr(out.ids), _lib.ptr(out.scores),
                                    _lib.ptr(out.counts), _lib.stream_ptr(stream)), "ezr_fuse_lists")
    return out


class CoarseRanker:
    """dense + BM25 + RRF for query batches on one GPU (``HybridRetriever._aretrieve``, retrievers.py:276-291).

    The two routes are independent until the fusion.  ``overlap=True`` issues them on two CUDA streams and
    lets the RRF kernel join them; the default issues them back to back on the caller's stream (the dense
    kernel occupies all shared memory of every SM, so the routes cannot co-reside anyway and per-kernel
    timing stays clean).  All buffers are preallocated per (batch size, k) and reused.
    """

    def __init__(self, dense: DenseIndex, sparse: Bm25Index, canon: Optional[torch.Tensor] = None,
                 overlap: bool = False, depth: int = 2, serial_routes: bool = False):
        assert dense.device == sparse.device
        self.dense, self.sparse = dense, sparse
        self.device = dense.device
        self.canon = None if canon is None else canon.to(device=self.device, dtype=torch.int32).contiguous()
        self.overlap = overlap
        # submit() only: both routes on ONE side stream (BM25, then dense with its full shared-memory ring) instead of
        # side by side; the join still runs on its own stream under the next batch's routes
        self.serial_routes = bool(serial_routes)
        self.depth = max(int(depth), 1)
        self.s_dense = torch.cuda.Stream(device=self.device) if overlap else None
        self.s_sparse = torch.cuda.Stream(device=self.device) if overlap else None
        # the join of a submitted batch (collective, merges, RRF): short kernels, scheduled ahead of the next
        # batch's route CTAs whenever an SM has room
        self.s_tail = torch.cuda.Stream(device=self.device, priority=-1) if overlap else None
        self.ws_dense = Workspace(self.device)
        self.ws_sparse = Workspace(self.device)
        self._bufs = {}
        self._slots = {}
        self._n_submit = 0

    def _buffers(self, nq, kd, ks, ko):
        key = (nq, kd, ks, ko)
        if key not in self._bufs:
            dev = self.device
            mk = lambda dt, *shape: torch.empty(*shape, dtype=dt, device=dev)
            self._bufs[key] = (
                TopK(mk(torch.float32, nq, kd), mk(torch.int32, nq, kd), mk(torch.int32, nq)),
                TopK(mk(self.sparse.score_dtype, nq, ks), mk(torch.int32, nq, ks), mk(torch.int32, nq)),
                TopK(mk(torch.float64, nq, ko), mk(torch.int32, nq, ko), mk(torch.int32, nq)),
            )
        return self._bufs[key]

    def routes(self, queries: torch.Tensor, q_ptr: torch.Tensor, q_terms: torch.Tensor, k: int, k_out: int,
               q_group: Optional[torch.Tensor] = None, d_out: Optional[TopK] = None,
               s_out: Optional[TopK] = None) -> Tuple[TopK, TopK, TopK]:
        """Both routes over this ranker's (shard of the) corpus -> (dense, sparse, fused-output buffer).

        ``d_out`` / ``s_out``: caller-owned result buffers (the sharded ranker passes views of its exchange
        record, so the kernels write straight into the message)."""
        nq = queries.shape[0]
        d_own, s_own, f_out = self._buffers(nq, k, k, k_out)
        d_out = d_own if d_out is None else d_out
        s_out = s_own if s_out is None else s_out
        cur = torch.cuda.current_stream(self.device)
        if self.overlap:
            self.s_dense.wait_stream(cur)
            self.s_sparse.wait_stream(cur)
            # dense first: its persistent CTAs (one per SM) must be resident before the BM25 grid starts filling
            # whatever shared memory / registers / issue slots they leave free
            with torch.cuda.stream(self.s_dense):
                dense_topk(self.dense, queries, k, q_group=q_group, ws=self.ws_dense, stream=self.s_dense, out=d_out)
            with torch.cuda.stream(self.s_sparse):
                bm25_topk(self.sparse, q_ptr, q_terms, k, q_group=q_group, ws=self.ws_sparse, stream=self.s_sparse,
                          out=s_out)
            cur.wait_stream(self.s_sparse)
            cur.wait_stream(self.s_dense)
        else:
            bm25_topk(self.sparse, q_ptr, q_terms, k, q_group=q_group, ws=self.ws_sparse, stream=cur, out=s_out)
            dense_topk(self.dense, queries, k, q_group=q_group, ws=self.ws_dense, stream=cur, out=d_out)
        return d_out, s_out, f_out

    def hybrid(self, queries: torch.Tensor, q_ptr: torch.Tensor, q_terms: torch.Tensor, k_dense: int = 10,
               k_sparse: int = 10, k_out: int = 10, K: int = 60, q_group: Optional[torch.Tensor] = None
               ) -> Tuple[TopK, TopK, TopK]:
        """Returns (fused, sparse, dense).  Inputs must already be on the device."""
        if k_dense != k_sparse:
            raise ValueError("the fused path keeps both routes at the same k (pad the shorter list upstream)")
        d_out, s_out, f_out = self.routes(queries, q_ptr, q_terms, k_dense, k_out, q_group=q_group)
        rrf_fuse(s_out.ids, s_out.counts, d_out.ids, d_out.counts, k_out, K=K, canon=self.canon, out=f_out)
        return f_out, s_out, d_out

    # ---- batch pipelining: submit() returns before the batch is joined to the caller's stream -------------------
    def _slot(self, key, make):
        """Result buffers + events of the next in-flight batch (``depth`` of them per key, used round robin)."""
        if key not in self._slots:
            self._slots[key] = [dict(make(), ev_in=torch.cuda.Event(), ev_d=torch.cuda.Event(),
                                     ev_s=torch.cuda.Event(), done=torch.cuda.Event(), free=torch.cuda.Event())
                                for _ in range(self.depth)]
        slot = self._slots[key][self._n_submit % self.depth]
        self._n_submit += 1
        return slot

    def launch_routes(self, slot, queries, q_ptr, q_terms, k, q_group, d_out: TopK, s_out: TopK) -> None:
        """Both routes on their streams, ordered after (a) the caller's stream at this point (the inputs), (b) the
        join that last read this slot's buffers and (c) the consumer's release of the slot.  Nothing is joined back
        to the caller's stream: ``slot['ev_d']`` / ``slot['ev_s']`` mark the two routes' results."""
        if not self.overlap:
            raise RuntimeError("submit() needs CoarseRanker(overlap=True): the routes and the join run on own streams")
        cur = torch.cuda.current_stream(self.device)
        slot["ev_in"].record(cur)
        s_sp = self.s_dense if self.serial_routes else self.s_sparse
        for st in ((self.s_dense,) if self.serial_routes else (self.s_dense, self.s_sparse)):
            st.wait_event(slot["ev_in"])
            st.wait_event(slot["done"])          # never-recorded events do not block
            st.wait_event(slot["free"])
        if self.serial_routes:
            with torch.cuda.stream(s_sp):
                bm25_topk(self.sparse, q_ptr, q_terms, k, q_group=q_group, ws=self.ws_sparse, stream=s_sp, out=s_out)
                slot["ev_s"].record(s_sp)
        with torch.cuda.stream(self.s_dense):
            dense_topk(self.dense, queries, k, q_group=q_group, ws=self.ws_dense, stream=self.s_dense, out=d_out)
            slot["ev_d"].record(self.s_dense)
        if not self.serial_routes:
            with torch.cuda.stream(s_sp):
                bm25_topk(self.sparse, q_ptr, q_terms, k, q_group=q_group, ws=self.ws_sparse, stream=s_sp, out=s_out)
                slot["ev_s"].record(s_sp)

    def submit(self, queries: torch.Tensor, q_ptr: torch.Tensor, q_terms: torch.Tensor, k: int = 10, k_out: int = 10,
               K: int = 60, q_group: Optional[torch.Tensor] = None) -> "Ticket":
        """:meth:`hybrid` for a stream of independent batches: the same kernels, but the batch is NOT joined to the
        caller's stream, so the routes of the next submitted batch start while this batch's RRF (and, sharded, its
        all-gather and merges) are still running.  Up to ``depth`` batches are in flight; a slot's buffers are
        rewritten ``depth`` submits later, after its join has finished and -- if the consumer reads them on another
        stream -- after :meth:`Ticket.release`.  Read the results after :meth:`Ticket.wait` or :meth:`join`."""
        nq = queries.shape[0]

        def make():
            mk = lambda dt, *shape: torch.empty(*shape, dtype=dt, device=self.device)
            return dict(d=TopK(mk(torch.float32, nq, k), mk(torch.int32, nq, k), mk(torch.int32, nq)),
                        s=TopK(mk(self.sparse.score_dtype, nq, k), mk(torch.int32, nq, k), mk(torch.int32, nq)),
                        f=TopK(mk(torch.float64, nq, k_out), mk(torch.int32, nq, k_out), mk(torch.int32, nq)))
        slot = self._slot((nq, k, k_out), make)
        self.launch_routes(slot, queries, q_ptr, q_terms, k, q_group, slot["d"], slot["s"])
        with torch.cuda.stream(self.s_tail):
            self.s_tail.wait_event(slot["ev_d"])
            self.s_tail.wait_event(slot["ev_s"])
            rrf_fuse(slot["s"].ids, slot["s"].counts, slot["d"].ids, slot["d"].counts, k_out, K=K, canon=self.canon,
                     out=slot["f"], stream=self.s_tail)
            slot["done"].record(self.s_tail)
        return Ticket(slot["f"], slot["s"], slot["d"], slot)

    def join(self) -> None:
        """The caller's stream waits for every submitted batch."""
        cur = torch.cuda.current_stream(self.device)
        for st in (self.s_tail, self.s_dense, self.s_sparse):
            if st is not None:
                cur.wait_stream(st)


class Ticket:
    """A submitted batch: result buffers (valid once ``done`` has fired) and the slot they live in."""
    __slots__ = ("fused", "sparse", "dense", "_slot")

    def __init__(self, fused: TopK, sparse: TopK, dense: TopK, slot: dict):
        self.fused, self.sparse, self.dense, self._slot = fused, sparse, dense, slot

    @property
    def done(self) -> torch.cuda.Event:
        return self._slot["done"]

    def wait(self, stream: Optional[torch.cuda.Stream] = None) -> None:
        (stream or torch.cuda.current_stream(self.fused.ids.device)).wait_event(self._slot["done"])

    def release(self, stream: Optional[torch.cuda.Stream] = None) -> None:
        """Call on the stream that read the results, after the reads: the slot may be rewritten once they finish."""
        self._slot["free"].record(stream or torch.cuda.current_stream(self.fused.ids.device))


class HostPipeline:
    """Host-buffer front end of the batched path: pinned host inputs in, pinned host results out, every call.

    ``step`` enqueues, without blocking the host: H2D of the query vectors / term pointers / term ids on a copy
    stream, both routes + (all-gather, merge) + RRF on the caller's stream, D2H of the fused ids and float64 scores
    on a second copy stream.  Inputs are double buffered on the device, so the copies of step i+1 run under the
    kernels of step i; a result buffer is only reused once its D2H has completed.  With a ranker built with
    ``overlap=True`` the steps are *submitted* (``CoarseRanker.submit``): the join and the D2H of step i run under the
    route kernels of step i+1.  ``ranker`` is a
    :class:`CoarseRanker` or an :class:`easyrag_b200.dist.ShardedCoarseRanker`.
    """

    def __init__(self, ranker, n_queries: int, dim: int, max_terms: int, k: int = 10, k_out: int = 10, depth: int = 2,
                 pipelined: Optional[bool] = None):
        base = getattr(ranker, "ranker", ranker)
        self.ranker, self.k, self.k_out = ranker, k, k_out
        # pipelined (default whenever the ranker runs its routes on own streams): steps are submitted, not joined
        self.pipelined = bool(base.overlap) if pipelined is None else bool(pipelined)
        self.device = dev = base.device
        self.sharded = base is not ranker
        self.s_in = torch.cuda.Stream(device=dev)
        self.s_out = torch.cuda.Stream(device=dev)
        self.slots = []
        for _ in range(depth):
            self.slots.append(dict(
                qvec=torch.empty(n_queries, dim, dtype=torch.bfloat16, device=dev),
                ptr=torch.empty(n_queries + 1, dtype=torch.int32, device=dev),
                terms=torch.empty(max(max_terms, 1), dtype=torch.int32, device=dev),
                ev_in=torch.cuda.Event(), ev_free=torch.cuda.Event()))
        self.ev_done = torch.cuda.Event()
        self.ev_out = torch.cuda.Event()
        self.n = 0

    def step(self, h_qvec: torch.Tensor, h_ptr: torch.Tensor, h_terms: torch.Tensor, h_ids_out: torch.Tensor,
             h_scores_out: torch.Tensor, q_group: Optional[torch.Tensor] = None) -> None:
        """One batch.  ``h_*`` are pinned host tensors: bf16 [Q, D], int32 [Q+1], int32 [T] in; int32 [Q, k_out] and
        float64 [Q, k_out] out (valid after :meth:`drain` or a synchronize)."""
        slot = self.slots[self.n % len(self.slots)]
        self.n += 1
        nq, nt = h_qvec.shape[0], h_terms.numel()
        if nq != slot["qvec"].shape[0] or nt > slot["terms"].numel():
            raise ValueError("HostPipeline is sized at construction: same batch size, at most max_terms term ids")
        cur = torch.cuda.current_stream(self.device)
        with torch.cuda.stream(self.s_in):
            self.s_in.wait_event(slot["ev_free"])            # the kernels that last read this slot have finished
            slot["qvec"].copy_(h_qvec, non_blocking=True)
            slot["ptr"].copy_(h_ptr, non_blocking=True)
            slot["terms"][:nt].copy_(h_terms, non_blocking=True)
            slot["ev_in"].record(self.s_in)
        cur.wait_event(slot["ev_in"])
        if self.pipelined:
            # the batch is not joined to the caller's stream: the routes of the next step start while this step's
            # join (all-gather, merges, RRF) and its D2H are still running
            if self.sharded:
                t = self.ranker.submit(slot["qvec"], slot["ptr"], slot["terms"], k=self.k, k_out=self.k_out,
                                       q_group=q_group)
            else:
                t = self.ranker.submit(slot["qvec"], slot["ptr"], slot["terms"], self.k, self.k_out, q_group=q_group)
            slot["ev_free"] = t.done                         # both routes have read the inputs once the join has run
            with torch.cuda.stream(self.s_out):
                t.wait(self.s_out)
                h_ids_out.copy_(t.fused.ids, non_blocking=True)
                h_scores_out.copy_(t.fused.scores, non_blocking=True)
                t.release(self.s_out)                        # the result slot may be rewritten after these copies
                self.ev_out.record(self.s_out)
            return
        cur.wait_event(self.ev_out)                          # the previous results have left the fused buffer
        if self.sharded:
            fused = self.ranker.hybrid(slot["qvec"], slot["ptr"], slot["terms"], k=self.k, k_out=self.k_out,
                                       q_group=q_group)[0]
        else:
            fused = self.ranker.hybrid(slot["qvec"], slot["ptr"], slot["terms"], self.k, self.k, self.k_out,
                                       q_group=q_group)[0]
        slot["ev_free"].record(cur)
        self.ev_done.record(cur)
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(self.ev_done)
            h_ids_out.copy_(fused.ids, non_blocking=True)
            h_scores_out.copy_(fused.scores, non_blocking=True)
            self.ev_out.record(self.s_out)

    def drain(self) -> None:
        """Make the caller's stream wait for every copy enqueued so far (then a stream / event sync covers them)."""
        torch.cuda.current_stream(self.device).wait_event(self.ev_out)


def dual_sparse_fusion(chunk_index: Bm25Index, path_index: Bm25Index, q_ptr: torch.Tensor, q_terms: torch.Tensor,
                       path_q_ptr: torch.Tensor, path_q_terms: torch.Tensor, k_chunk: int, k_path: int, k_out: int,
                       canon: Optional[torch.Tensor] = None, q_group: Optional[torch.Tensor] = None,
                       ws: Optional[Workspace] = None) -> TopK:
    """The reference's maintained coarse ranker as one batched op (pipeline.py:357-365): chunk-text BM25
    (k = f_topk_2) and knowledge-path BM25 (k = f_topk_3) over the same nodes, merged by ``HybridRetriever.fusion``
    (text-dedup, stable sort by raw score).  The two indexes have their own vocabularies, hence two term lists."""
    dev = chunk_index.device
    a = bm25_topk(chunk_index, q_ptr, q_terms, k_chunk, q_group=q_group, ws=ws)
    b = bm25_topk(path_index, path_q_ptr, path_q_terms, k_path, q_group=q_group, ws=ws)
    width = max(k_chunk, k_path)

    def pad(t: TopK, k: int):
        if k == width:
            return t.ids, t.scores.to(torch.float64)
        ids = torch.full((t.ids.shape[0], width), -1, dtype=torch.int32, device=dev)
        sc = torch.zeros(t.ids.shape[0], width, dtype=torch.float64, device=dev)
        ids[:, :k] = t.ids
        sc[:, :k] = t.scores
        return ids, sc
    ia, sa = pad(a, k_chunk)
    ib, sb = pad(b, k_path)
    return fusion_simple(ia, sa, a.counts, ib, sb, b.counts, k_out, canon=canon)
