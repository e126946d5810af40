"""Seeded synthetic corpora / queries for the coarse-ranking path.

No corpus text ships with the reference (scripts/process.sh downloads it), so
tests and bench.py use this generator.  Everything is produced with torch ops
so the same code runs on ``cpu`` (tests, small configs) and on ``cuda``
(BASELINE.json configs 3-5, 3e8 tokens); a run is reproducible for a given
(seed, device type, torch version).

Model of a chunk after ``tokenize_and_remove_stopwords`` (retrievers.py:72-76):
raw tokens follow Zipf(s=1.07) over ``n_stop + vocab`` ranks; the ``n_stop`` most
frequent ranks are the stop-word list (hit_stopwords.txt has 749 entries and
removes roughly half of running text), so term ids 0..vocab-1 are the ranks
*after* the stop words, passed through a fixed random permutation so that
posting-list length is uncorrelated with term id.  Chunk length after
stop-word removal is a clipped lognormal with mean ~300 in [20, 800]
(chunk_size 1024 tiktoken tokens, easyrag.yaml:46).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch

ZIPF_S = 1.07
N_STOP = 100


@dataclass
class SparseCorpus:
    """Ragged term-id documents: tokens of doc i are ``tokens[doc_ptr[i]:doc_ptr[i+1]]``."""
    tokens: torch.Tensor      # int32 [T]
    doc_ptr: torch.Tensor     # int64 [N+1]
    vocab: int

    @property
    def n_docs(self) -> int:
        return self.doc_ptr.numel() - 1

    def doc_lists(self) -> List[np.ndarray]:
        t = self.tokens.cpu().numpy()
        p = self.doc_ptr.cpu().numpy()
        return [t[p[i]:p[i + 1]] for i in range(len(p) - 1)]


@dataclass
class QuerySet:
    term_ptr: torch.Tensor    # int32 [Q+1]
    terms: torch.Tensor       # int32 [sum]  (-1 = out-of-vocabulary token)
    vectors: Optional[torch.Tensor] = None    # bf16 [Q, D], unit norm
    group: Optional[torch.Tensor] = None      # int32 [Q], wanted ``dir`` id or -1 (no filter)

    @property
    def n(self) -> int:
        return self.term_ptr.numel() - 1

    def term_lists(self) -> List[np.ndarray]:
        t = self.terms.cpu().numpy()
        p = self.term_ptr.cpu().numpy()
        return [t[p[i]:p[i + 1]] for i in range(len(p) - 1)]


def _gen(seed: int, device) -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    return g


def zipf_cdf(vocab: int, n_stop: int = N_STOP, s: float = ZIPF_S, device="cpu") -> torch.Tensor:
    r = torch.arange(n_stop + 1, n_stop + vocab + 1, dtype=torch.float64, device=device)
    p = r.pow(-s)
    return torch.cumsum(p / p.sum(), 0)


def make_sparse_corpus(n_docs: int, vocab: int, seed: int, device="cpu", mean_len: float = 300.0,
                       min_len: int = 20, max_len: int = 800, chunk: int = 1 << 26) -> SparseCorpus:
    g = _gen(seed, device)
    sigma = 0.5
    mu = float(np.log(mean_len) - sigma * sigma / 2)
    ln = torch.randn(n_docs, generator=g, device=device, dtype=torch.float32) * sigma + mu
    lens = ln.exp().round().clamp_(min_len, max_len).to(torch.int64)
    doc_ptr = torch.zeros(n_docs + 1, dtype=torch.int64, device=device)
    torch.cumsum(lens, 0, out=doc_ptr[1:])
    total = int(doc_ptr[-1])
    cdf = zipf_cdf(vocab, device=device)
    perm = torch.randperm(vocab, generator=g, device=device).to(torch.int32)
    tokens = torch.empty(total, dtype=torch.int32, device=device)
    for s in range(0, total, chunk):
        e = min(total, s + chunk)
        u = torch.rand(e - s, generator=g, device=device, dtype=torch.float64)
        rank = torch.searchsorted(cdf, u).clamp_(max=vocab - 1)
        tokens[s:e] = perm[rank]
    return SparseCorpus(tokens=tokens, doc_ptr=doc_ptr, vocab=vocab)


def make_queries(corpus: SparseCorpus, n_queries: int, seed: int, min_terms: int = 4, max_terms: int = 12,
                 p_random: float = 0.08, p_oov: float = 0.02) -> QuerySet:
    """Each query samples 4-12 tokens (with replacement) from one random chunk; ~10% noise terms."""
    device = corpus.tokens.device
    g = _gen(seed, device)
    n = corpus.n_docs
    m = torch.randint(min_terms, max_terms + 1, (n_queries,), generator=g, device=device)
    term_ptr = torch.zeros(n_queries + 1, dtype=torch.int64, device=device)
    torch.cumsum(m, 0, out=term_ptr[1:])
    total = int(term_ptr[-1])
    qid = torch.repeat_interleave(torch.arange(n_queries, device=device), m)
    src_doc = torch.randint(0, n, (n_queries,), generator=g, device=device)[qid]
    lo = corpus.doc_ptr[src_doc]
    ln = corpus.doc_ptr[src_doc + 1] - lo
    off = (torch.rand(total, generator=g, device=device, dtype=torch.float64) * ln.to(torch.float64)).long()
    off = torch.minimum(off, ln - 1)
    terms = corpus.tokens[lo + off].clone()
    u = torch.rand(total, generator=g, device=device)
    rnd = torch.randint(0, corpus.vocab, (total,), generator=g, device=device, dtype=torch.int32)
    terms = torch.where(u < p_random, rnd, terms)
    terms = torch.where((u >= p_random) & (u < p_random + p_oov), torch.full_like(terms, -1), terms)
    return QuerySet(term_ptr=term_ptr.to(torch.int32), terms=terms.to(torch.int32))


def make_dense_corpus(n_rows: int, dim: int, seed: int, device="cpu", chunk: int = 1 << 18) -> torch.Tensor:
    """Unit-norm rows of N(0,1), stored bf16 (what GTEEmbedding emits: gte_embeddings.py:70-71)."""
    g = _gen(seed, device)
    out = torch.empty(n_rows, dim, dtype=torch.bfloat16, device=device)
    for s in range(0, n_rows, chunk):
        e = min(n_rows, s + chunk)
        x = torch.randn(e - s, dim, generator=g, device=device, dtype=torch.float32)
        out[s:e] = torch.nn.functional.normalize(x, dim=1).to(torch.bfloat16)
    return out


def make_dense_queries(corpus: torch.Tensor, n_queries: int, seed: int, noise: float = 0.7) -> torch.Tensor:
    """Queries near a random corpus row (so the top hit is meaningful), unit norm, bf16."""
    device = corpus.device
    g = _gen(seed, device)
    src = torch.randint(0, corpus.shape[0], (n_queries,), generator=g, device=device)
    x = corpus[src].float()
    x = x + noise * torch.nn.functional.normalize(
        torch.randn(n_queries, corpus.shape[1], generator=g, device=device), dim=1)
    return torch.nn.functional.normalize(x, dim=1).to(torch.bfloat16)


def make_groups(n_docs: int, n_groups: int, seed: int, device="cpu") -> torch.Tensor:
    """Per-document ``dir`` id (the metadata field ``filter_dict`` / qdrant filters match on)."""
    g = _gen(seed, device)
    return torch.randint(0, n_groups, (n_docs,), generator=g, device=device, dtype=torch.int32)


def make_duplicates(n_docs: int, frac: float, seed: int, device="cpu") -> torch.Tensor:
    """canon[i] = smallest index with the same text; ``frac`` of docs copy an earlier doc."""
    g = _gen(seed, device)
    canon = torch.arange(n_docs, device=device, dtype=torch.int32)
    if frac <= 0 or n_docs < 2:
        return canon
    is_dup = torch.rand(n_docs, generator=g, device=device) < frac
    is_dup[0] = False
    src = (torch.rand(n_docs, generator=g, device=device, dtype=torch.float64)
           * torch.arange(n_docs, device=device, dtype=torch.float64)).long()
    src = torch.minimum(src, torch.arange(n_docs, device=device) - 1).clamp_(min=0)
    # resolve chains: a duplicate of a duplicate points at the original
    c = torch.where(is_dup, src.to(torch.int32), canon)
    for _ in range(32):
        c2 = c[c.long()]
        if torch.equal(c2, c):
            break
        c = c2
    return c


class PseudoWordTokenizer:
    """Duck-typed stand-in for ``jieba.Tokenizer()`` (pipeline.py:177-178): ``.cut(text)``.

    Text is space-joined pseudo-words ``w<id>``; ``cut`` yields the words *and* the
    single-space separators, as jieba does (hence the ``word != ' '`` test at
    retrievers.py:75).
    """

    def cut(self, text: str):
        out = []
        for i, w in enumerate(text.split(' ')):
            if i:
                out.append(' ')
            if w:
                out.append(w)
        return out


def ids_to_text(ids) -> str:
    return ' '.join(f"w{int(i)}" if i >= 0 else "oov" for i in ids)
