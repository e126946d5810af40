"""Oracle: retriever control flow of the reference.  TEST INFRASTRUCTURE ONLY.

Follows src/easyrag/custom/retrievers.py of the reference:
  tokenize_and_remove_stopwords  :72-76
  BM25Retriever.filter           :191-210
  HybridRetriever.fusion         :239-253
  HybridRetriever.reciprocal_rank_fusion :256-274
  QdrantRetriever (cosine top-k) :37-52 with Distance.COSINE (ingestion.py:180-182)

Nodes here are tiny stand-ins exposing exactly what those functions touch:
``.get_content()``, ``.score`` (mutable), ``.metadata``, ``.node``.
"""
from __future__ import annotations

from collections import defaultdict
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np


@dataclass
class ONode:
    text: str
    idx: int = -1
    metadata: Dict[str, object] = field(default_factory=dict)

    def get_content(self) -> str:
        return self.text


@dataclass
class OScored:
    node: ONode
    score: float

    def get_content(self) -> str:
        return self.node.get_content()

    @property
    def metadata(self):
        return self.node.metadata


def tokenize_and_remove_stopwords(tokenizer, text, stopwords):
    """retrievers.py:72-76."""
    return [w for w in tokenizer.cut(text) if w not in stopwords and w != ' ']


def canonical_order(scores: np.ndarray) -> np.ndarray:
    """(score desc, index desc) == ``scores.argsort(kind='stable')[::-1]``.

    numpy's default argsort used at retrievers.py:192 is not stable, so the
    reference's order inside a tie group is platform dependent (SURVEY.md
    8(c)); this is the declared canonical order every parity test uses.
    """
    return np.argsort(scores, kind="stable")[::-1]


def bm25_filter(scores: np.ndarray, nodes: Sequence[ONode], k: int,
                filter_dict: Optional[Dict[str, object]] = None,
                literal_argsort: bool = False) -> List[OScored]:
    """retrievers.py:191-210 (``literal_argsort=True`` reproduces :192 verbatim)."""
    top_n = scores.argsort()[::-1] if literal_argsort else canonical_order(scores)
    out: List[OScored] = []
    for ix in top_n:
        if scores[ix] <= 0:
            break
        flag = True
        if filter_dict is not None:
            for key, value in filter_dict.items():
                if nodes[ix].metadata[key] != value:
                    flag = False
                    break
        if flag:
            out.append(OScored(node=nodes[ix], score=float(scores[ix])))
        if len(out) == k:
            break
    out = sorted(out, key=lambda x: x.score, reverse=True)
    return out


def bm25_topk_ids(scores: np.ndarray, k: int, allowed: Optional[np.ndarray] = None):
    """Array form of :func:`bm25_filter` under the canonical order.

    Returns (ids int64[<=k], scores[<=k]).  ``allowed`` is a boolean mask (the
    ``filter_dict`` predicate evaluated per document).
    """
    order = canonical_order(scores)
    pos = scores[order] > 0
    if allowed is not None:
        pos &= allowed[order]
    # the reference stops at the first score <= 0; scores are sorted descending
    # so everything after it is <= 0 too and the two formulations coincide.
    ids = order[pos][:k]
    return ids.astype(np.int64), scores[ids]


def fusion(list_of_lists: Sequence[Sequence[OScored]], topk: int = 256) -> List[OScored]:
    """retrievers.py:239-253."""
    all_nodes, seen = [], set()
    for nodes in list_of_lists:
        for node in nodes:
            content = node.get_content()
            if content not in seen:
                all_nodes.append(node)
                seen.add(content)
    all_nodes = sorted(all_nodes, key=lambda n: n.score, reverse=True)
    return all_nodes[:min(len(all_nodes), topk)]


def reciprocal_rank_fusion(list_of_lists: Sequence[Sequence[OScored]], K: int = 60,
                           topk: int = 256) -> List[OScored]:
    """retrievers.py:256-274 (mutates ``.score`` of the returned items, as the reference does)."""
    rrf_map = defaultdict(float)
    text_to_node = {}
    for rank_list in list_of_lists:
        for rank, item in enumerate(rank_list, 1):
            content = item.get_content()
            text_to_node[content] = item
            rrf_map[content] += 1 / (rank + K)
    sorted_items = sorted(rrf_map.items(), key=lambda x: x[1], reverse=True)
    out = []
    for text, score in sorted_items:
        out.append(text_to_node[text])
        out[-1].score = score
    return out[:min(topk, len(out))]


def rrf_ids(lists: Sequence[Sequence[int]], canon: Optional[np.ndarray] = None, K: int = 60,
            topk: int = 256):
    """Integer-id form of :func:`reciprocal_rank_fusion`.

    ``canon[id]`` maps a document index to the smallest index carrying the same
    text (the dict key of retrievers.py:263-265).  Returns (representative ids,
    float64 scores): the representative is the *last writer* of
    ``text_to_node`` -- retrievers.py:264.
    """
    rrf: Dict[int, float] = {}
    rep: Dict[int, int] = {}
    for lst in lists:
        for rank, i in enumerate(lst, 1):
            key = int(canon[i]) if canon is not None else int(i)
            rep[key] = int(i)
            rrf[key] = rrf.get(key, 0.0) + 1 / (rank + K)
    items = sorted(rrf.items(), key=lambda x: x[1], reverse=True)[:topk]
    return (np.array([rep[k_] for k_, _ in items], dtype=np.int64),
            np.array([s for _, s in items], dtype=np.float64))


def fusion_ids(lists: Sequence[Sequence[int]], scores: Sequence[Sequence[float]],
               canon: Optional[np.ndarray] = None, topk: int = 256):
    """Integer-id form of :func:`fusion`: first occurrence of a text wins, stable sort by raw score."""
    seen, ids, sc = set(), [], []
    for lst, ss in zip(lists, scores):
        for i, s in zip(lst, ss):
            key = int(canon[i]) if canon is not None else int(i)
            if key not in seen:
                seen.add(key)
                ids.append(int(i))
                sc.append(float(s))
    order = sorted(range(len(ids)), key=lambda j: sc[j], reverse=True)[:topk]
    return (np.array([ids[j] for j in order], dtype=np.int64),
            np.array([sc[j] for j in order], dtype=np.float64))


def dense_topk(corpus: np.ndarray, queries: np.ndarray, k: int,
               allowed: Optional[np.ndarray] = None):
    """Exact cosine top-k as qdrant local mode does it: float32 dot of unit vectors.

    Reference: retrievers.py:44-47 -> QdrantVectorStore.aquery on a collection
    created with Distance.COSINE (ingestion.py:180-182).  Inputs are float32
    and already L2-normalised (the embedding classes normalise; qdrant
    re-normalises at insert, a no-op to 1e-7).  Ties resolve (score desc,
    id desc), the repository-wide canonical order.  ``allowed``: bool[Q, N] or
    bool[N] payload filter (ingestion.py:207-216).
    Returns (ids int64[Q,k], scores float32[Q,k]); rows padded with -1/-inf.
    """
    corpus = np.asarray(corpus, dtype=np.float32)
    queries = np.asarray(queries, dtype=np.float32)
    sims = queries @ corpus.T
    q, n = sims.shape
    ids = np.full((q, k), -1, dtype=np.int64)
    sc = np.full((q, k), -np.inf, dtype=np.float32)
    for i in range(q):
        s = sims[i]
        order = canonical_order(s)
        if allowed is not None:
            m = allowed[i] if allowed.ndim == 2 else allowed
            order = order[m[order]]
        order = order[:k]
        ids[i, :order.size] = order
        sc[i, :order.size] = s[order]
    return ids, sc


def rerank_inputs(query_ids: Sequence[int], passages: Sequence[Sequence[int]], sep: Sequence[int], prompt: Sequence[int],
                  bos: int, max_length: int = 1024):
    """``LLMRerank.get_inputs`` / ``get_inputs_v2_5`` (rerankers.py:196-293) on token ids, before padding.

    ``query_ids`` = tokenizer("A: " + query) and ``passages[i]`` = tokenizer("B: " + passage), add_special_tokens=False
    (what the reference computes at :221-226 / :266-275; their own truncations max_length*3//4 and max_length are
    applied here).  ``prepare_for_model(first, second, truncation='only_second', max_length=max_length)`` keeps
    ``first`` whole and cuts ``second`` from its end.  Returns (list of id lists, query_lengths, prompt_lengths).
    """
    q = list(query_ids)[: max_length * 3 // 4]
    items, qlens, plens = [], [], []
    for p in passages:
        first = [bos] + q
        second = list(sep) + list(p)[:max_length]
        second = second[: max(max_length - len(first), 0)]
        items.append(first + second + list(sep) + list(prompt))
        qlens.append(len(first) + len(sep))
        plens.append(len(sep) + len(prompt))
    return items, qlens, plens
