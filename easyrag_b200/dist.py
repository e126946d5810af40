"""Row-sharded coarse ranking across the GPUs of one box.

The reference is single-process; this is new.  Corpus rows (and the BM25 document axis) are cut
into contiguous shards, one per rank; queries are replicated.  Corpus-global BM25 statistics
(idf, avgdl) are shared, so every shard scores exactly as the unsharded index would.  Each rank
computes its local dense and BM25 top-k with *global* ids straight into one byte record and a
SINGLE all-gather (NCCL over NVLink on GPUs, gloo in the CPU tests) exchanges the records;
every rank then merges G*k candidates per route under the canonical order -- the same order
the 1-GPU path uses, hence identical rank lists -- reading the gathered buffer in place, and
runs RRF.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_rows: int, world: int, rank: int, align: int = 1) -> Tuple[int, int]:
    """Contiguous, ordered, exhaustive, BALANCED split of ``n_rows``.

    The rows are cut into ``ceil(n_rows / align)`` units of ``align`` rows (the last may be short); every rank gets
    ``units // world`` of them and the first ``units % world`` ranks one more, so shard sizes differ by at most
    ``align`` rows.  Shard-local indexes use local ids (ranges and tiles restart at the shard's first row), so
    nothing requires a coarse alignment: the default is 1."""
    units = -(-n_rows // align)
    base, extra = divmod(units, world)
    lo_u = rank * base + min(rank, extra)
    hi_u = lo_u + base + (1 if rank < extra else 0)
    return min(n_rows, lo_u * align), min(n_rows, hi_u * align)


@dataclass
class RecordLayout:
    """Byte layout of one rank's contribution: dense scores f32 | dense ids i32 | sparse scores | sparse ids i32."""
    n_queries: int
    k: int
    sparse_bytes: int      # 8 for BM25Okapi (float64), 4 for bm25s (float32)

    @property
    def sizes(self):
        n = self.n_queries * self.k
        return (n * 4, n * 4, n * self.sparse_bytes, n * 4)

    @property
    def offsets(self):
        o, out = 0, []
        for s in self.sizes:
            out.append(o)
            o += (s + 15) // 16 * 16
        return out, o

    @property
    def nbytes(self) -> int:
        return self.offsets[1]


def record_views(layout: RecordLayout, buf: torch.Tensor):
    """The four per-route arrays of one rank's record as typed [Q, k] views of the byte buffer ``buf``
    (dense scores f32, dense ids i32, sparse scores f64/f32, sparse ids i32): writing through them fills the
    message in place, reading them from a gathered buffer needs no unpacking."""
    offs, _ = layout.offsets
    q, k = layout.n_queries, layout.k
    sdt = torch.float64 if layout.sparse_bytes == 8 else torch.float32
    out = []
    for off, size, dt in zip(offs, layout.sizes, (torch.float32, torch.int32, sdt, torch.int32)):
        out.append(buf[off:off + size].view(dt).view(q, k))
    return out


class ShardedCoarseRanker:
    """dense + BM25 + RRF over a row-sharded corpus; every rank returns the full fused result.

    Per (batch size, k) one record buffer and one gather buffer are allocated once.  The two route kernels write
    their top-k straight into the record (typed views), ONE ``all_gather_into_tensor`` exchanges the records, and the
    merge kernel reads the gathered buffer in place (``ezr_merge_topk_parts``): no pack/unpack kernels, no per-step
    allocations.
    """

    def __init__(self, ranker, group=None):
        """``ranker``: :class:`easyrag_b200.batched.CoarseRanker` over this rank's shard (indexes built with
        ``row_lo`` / ``doc_lo`` = the shard's first global row)."""
        self.ranker = ranker
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._state = {}

    def _make(self, nq: int, k: int, sparse_dtype):
        from .batched import TopK
        dev = self.ranker.device
        layout = RecordLayout(nq, k, 8 if sparse_dtype == torch.float64 else 4)
        record = torch.zeros(layout.nbytes, dtype=torch.uint8, device=dev)
        gathered = torch.zeros(self.world * layout.nbytes, dtype=torch.uint8, device=dev)
        ds, di, ss, si = record_views(layout, record)
        cnt = lambda: torch.empty(nq, dtype=torch.int32, device=dev)
        mk = lambda dt: torch.empty(nq, k, dtype=dt, device=dev)
        return dict(
            layout=layout, record=record, gathered=gathered,
            d_local=TopK(ds, di, cnt()), s_local=TopK(ss, si, cnt()),
            views=record_views(layout, gathered[:layout.nbytes]),
            dense=TopK(mk(torch.float32), mk(torch.int32), cnt()),
            sparse=TopK(mk(sparse_dtype), mk(torch.int32), cnt()))

    def _buffers(self, nq: int, k: int, sparse_dtype):
        key = (nq, k, sparse_dtype)
        if key not in self._state:
            self._state[key] = self._make(nq, k, sparse_dtype)
        return self._state[key]

    def _join(self, st, k: int, k_out: int, K: int, canon, f_out, stream=None):
        """all-gather of the records, per-route merge of the world's lists, RRF -- on the current stream."""
        from . import batched
        dist.all_gather_into_tensor(st["gathered"], st["record"], group=self.group)   # the one collective
        g_ds, g_di, g_ss, g_si = st["views"]
        nbytes = st["layout"].nbytes
        dense = batched.merge_topk_parts(g_ds, g_di, self.world, nbytes, k, out=st["dense"], stream=stream)
        sparse = batched.merge_topk_parts(g_ss, g_si, self.world, nbytes, k, out=st["sparse"], stream=stream)
        fused = batched.rrf_fuse(sparse.ids, sparse.counts, dense.ids, dense.counts, k_out, K=K, canon=canon, out=f_out,
                                 stream=stream)
        return fused, sparse, dense

    def hybrid(self, queries, q_ptr, q_terms, k: int = 10, k_out: int = 10, K: int = 60, q_group=None,
               canon: Optional[torch.Tensor] = None):
        from . import batched
        r = self.ranker
        nq = queries.shape[0]
        if k > 32:
            raise ValueError("the sharded path merges per-shard lists of k <= 32")
        st = self._buffers(nq, k, r.sparse.score_dtype)
        _, _, f_out = r.routes(queries, q_ptr, q_terms, k, k_out, q_group=q_group, d_out=st["d_local"],
                               s_out=st["s_local"])
        return self._join(st, k, k_out, K, canon if canon is not None else r.canon, f_out)

    def submit(self, queries, q_ptr, q_terms, k: int = 10, k_out: int = 10, K: int = 60, q_group=None,
               canon: Optional[torch.Tensor] = None):
        """:meth:`hybrid` without the join to the caller's stream (see ``CoarseRanker.submit``): the all-gather, the
        merges and the RRF of this batch run on the ranker's join stream while the routes of the next submitted
        batch already occupy the SMs.  Every rank must submit the same sequence of batches (one collective each)."""
        from . import batched
        r = self.ranker
        nq = queries.shape[0]
        if k > 32:
            raise ValueError("the sharded path merges per-shard lists of k <= 32")

        def make():
            st = self._make(nq, k, r.sparse.score_dtype)
            st["f"] = batched.TopK(torch.empty(nq, k_out, dtype=torch.float64, device=r.device),
                                   torch.empty(nq, k_out, dtype=torch.int32, device=r.device),
                                   torch.empty(nq, dtype=torch.int32, device=r.device))
            return st
        slot = r._slot(("sharded", id(self), nq, k, k_out), make)
        r.launch_routes(slot, queries, q_ptr, q_terms, k, q_group, slot["d_local"], slot["s_local"])
        with torch.cuda.stream(r.s_tail):
            r.s_tail.wait_event(slot["ev_d"])
            r.s_tail.wait_event(slot["ev_s"])
            fused, sparse, dense = self._join(slot, k, k_out, K, canon if canon is not None else r.canon, slot["f"],
                                              stream=r.s_tail)
            slot["done"].record(r.s_tail)
        return batched.Ticket(fused, sparse, dense, slot)

    def join(self) -> None:
        self.ranker.join()
