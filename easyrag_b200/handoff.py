"""Hand-off of the coarse ranker's candidates to the fine-ranking stage.

The reference's ``LLMRerank._postprocess_nodes`` (rerankers.py:298-376) walks the coarse list in slices of
``embed_bs`` (32, yaml:28) and tokenises ``(query, get_node_content(node, embed_type))`` pairs per slice, then
overwrites ``node.score`` in place.  The reranker itself is outside this repository's scope; what belongs to
the coarse path is producing its input in exactly that layout, without re-deriving anything on the way:

* ``rerank_batches``   - the drop-in view: a fused ``List[NodeWithScore]`` -> the same slices of pairs.
* ``candidate_batches`` - the batched view: the ``[Q, k]`` id tensor of ``CoarseRanker`` -> per-query slices of
  document ids in rank order (one D2H copy for the whole batch; texts are looked up by the caller).

Host-side Python only (strings and lists), like the code it feeds.
"""
from __future__ import annotations

from typing import Iterator, List, Sequence, Tuple

from .retrievers import get_node_content

DEFAULT_RERANK_BATCH = 32      # yaml:28 (embed_bs of the reranker)


def rerank_batches(nodes: Sequence, query_str: str, embed_type: int = 0,
                   batch_size: int = DEFAULT_RERANK_BATCH) -> Iterator[Tuple[int, int, List[Tuple[str, str]]]]:
    """Yield ``(begin, end, [(query, text), ...])`` exactly as rerankers.py:309-322 slices the coarse list.

    ``nodes`` are ``NodeWithScore``-like objects (``.node``); the caller scores a slice and writes
    ``nodes[begin + i].score`` back, which is what the reference does (rerankers.py:365-370).
    """
    if batch_size < 1:
        raise ValueError("batch_size must be positive")
    n = len(nodes)
    for begin in range(0, n, batch_size):
        end = min(begin + batch_size, n)
        yield begin, end, [(query_str, get_node_content(nw.node, embed_type)) for nw in nodes[begin:end]]


def candidate_batches(ids, counts, batch_size: int = DEFAULT_RERANK_BATCH) -> Iterator[Tuple[int, int, List[int]]]:
    """Yield ``(query index, begin, [doc ids])`` for every slice of every query's candidate list, rank order kept.

    ``ids`` is ``[Q, k]`` (torch tensor on any device, or array-like) with ``-1`` padding, ``counts`` is ``[Q]``:
    the ``TopK`` a ``CoarseRanker`` / ``ShardedCoarseRanker`` returns.  One transfer for the whole batch.
    """
    if batch_size < 1:
        raise ValueError("batch_size must be positive")
    ids_h = ids.cpu().tolist() if hasattr(ids, "cpu") else [list(r) for r in ids]
    cnt_h = counts.cpu().tolist() if hasattr(counts, "cpu") else list(counts)
    for q, (row, c) in enumerate(zip(ids_h, cnt_h)):
        row = row[:c]
        for begin in range(0, c, batch_size):
            yield q, begin, row[begin:begin + batch_size]
