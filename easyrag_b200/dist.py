"""Row-sharded coarse ranking across the GPUs of one box.

The reference is single-process; this is new.  Corpus rows (and the BM25 document axis) are cut
into contiguous shards, one per rank; queries are replicated.  Corpus-global BM25 statistics
(idf, avgdl) are shared, so every shard scores exactly as the unsharded index would.  Each rank
computes its local dense and BM25 top-k with *global* ids, packs both routes into one byte
record and a SINGLE all-gather (NCCL over NVLink on GPUs, gloo in the CPU tests) exchanges
them; every rank then merges G*k candidates per route under the canonical order -- the same
order the 1-GPU path uses, hence identical rank lists -- and runs RRF.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_rows: int, world: int, rank: int, align: int = 1) -> Tuple[int, int]:
    """Contiguous, ordered, exhaustive split of ``n_rows``; shard sizes are multiples of ``align`` except the last."""
    per = -(-n_rows // world)
    per = -(-per // align) * align
    lo = min(n_rows, rank * per)
    hi = min(n_rows, lo + per)
    return lo, hi


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


def pack_records(layout: RecordLayout, d_scores, d_ids, s_scores, s_ids, out: Optional[torch.Tensor] = None):
    """Copy the four per-route tensors into one contiguous uint8 buffer (one message per rank)."""
    offs, total = layout.offsets
    dev = d_scores.device
    if out is None:
        out = torch.zeros(total, dtype=torch.uint8, device=dev)
    for off, size, t in zip(offs, layout.sizes, (d_scores, d_ids, s_scores, s_ids)):
        out[off:off + size].copy_(t.contiguous().view(-1).view(torch.uint8))
    return out


def unpack_records(layout: RecordLayout, gathered: torch.Tensor, world: int):
    """gathered uint8 [world, nbytes] -> per-route candidate matrices [Q, world*k] (ids < 0 = empty)."""
    offs, total = layout.offsets
    q, k = layout.n_queries, layout.k
    sdt = torch.float64 if layout.sparse_bytes == 8 else torch.float32
    g = gathered.view(world, total)

    def route(off, size, dtype):
        x = g[:, off:off + size].contiguous().view(-1).view(dtype).view(world, q, k)
        return x.permute(1, 0, 2).contiguous().view(q, world * k)

    sizes = layout.sizes
    return (route(offs[0], sizes[0], torch.float32), route(offs[1], sizes[1], torch.int32),
            route(offs[2], sizes[2], sdt), route(offs[3], sizes[3], torch.int32))


def all_gather_bytes(local: torch.Tensor, group=None) -> torch.Tensor:
    """The one collective on the data path: all-gather of the packed per-shard top-k records."""
    world = dist.get_world_size(group)
    out = torch.empty(world * local.numel(), dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(out, local, group=group)
    return out.view(world, local.numel())


class ShardedCoarseRanker:
    """dense + BM25 + RRF over a row-sharded corpus; every rank returns the full fused result."""

    def __init__(self, ranker, group=None):
        """``ranker``: :class:`easyrag_b200.batched.CoarseRanker` over this rank's shard (indexes built with
        ``row_lo`` / ``doc_lo`` = the shard's first global row)."""
        self.ranker = ranker
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._pack = {}

    def hybrid(self, queries, q_ptr, q_terms, k: int = 10, k_out: int = 10, K: int = 60, q_group=None,
               canon: Optional[torch.Tensor] = None):
        from . import batched
        r = self.ranker
        nq = queries.shape[0]
        d_out, s_out, f_out = r.routes(queries, q_ptr, q_terms, k, k_out, q_group=q_group)
        layout = RecordLayout(nq, k, 8 if s_out.scores.dtype == torch.float64 else 4)
        key = (nq, k, layout.sparse_bytes)
        if key not in self._pack:
            self._pack[key] = torch.zeros(layout.nbytes, dtype=torch.uint8, device=r.device)
        local = pack_records(layout, d_out.scores, d_out.ids, s_out.scores, s_out.ids, out=self._pack[key])
        gathered = all_gather_bytes(local, self.group)
        ds, di, ss, si = unpack_records(layout, gathered, self.world)
        dense = batched.merge_topk(ds, di, k)
        sparse = batched.merge_topk(ss, si, k)
        cn = canon if canon is not None else r.canon
        fused = batched.rrf_fuse(sparse.ids, sparse.counts, dense.ids, dense.counts, k_out, K=K, canon=cn, out=f_out)
        return fused, sparse, dense
