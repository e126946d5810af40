"""Oracle: BM25 scoring as the reference computes it.  TEST INFRASTRUCTURE ONLY.

The reference (retrievers.py:103-118) delegates BM25 to two pip packages that
are NOT vendored under /root/reference and are not installed in this image:

  * rank-bm25==0.2.2 (requirements.txt:102) -> ``BM25Okapi(corpus, k1=1.5,
    b=0.75, epsilon=0.25)``, ``get_scores(tokens)``; float64.
  * bm25s==0.1.7 (requirements.txt:113) -> ``bm25s.BM25(k1, b).index(corpus)``,
    ``get_scores(tokens)``; method "lucene", float32.

Their published algorithms are restated below from the upstream documentation
(SURVEY.md section 8(c) records the evaluation order).  ``OkapiLiteral`` mirrors
the per-term / per-document loop the reference pays for (this is what the
published "17 s for 103 queries" measures); ``OkapiCSR`` is a vectorised
numpy formulation producing bit-identical float64 results, used at sizes where
the literal loop is too slow.  tests/test_oracle.py checks both against the
rank_bm25 README known answer and against each other.

bm25s: PARITY UNPINNED (no upstream vector available offline).
"""
from __future__ import annotations

import math
from typing import Dict, Hashable, List, Sequence

import numpy as np

K1 = 1.5       # retrievers.py:103
B = 0.75       # retrievers.py:104
EPSILON = 0.25  # retrievers.py:105


# --------------------------------------------------------------------------
# rank_bm25.BM25Okapi, literal
# --------------------------------------------------------------------------
class OkapiLiteral:
    """Literal restatement of rank_bm25 0.2.2 ``BM25Okapi`` (float64).

    Call sites in the reference: retrievers.py:113-118 (index build),
    retrievers.py:142-147 (throw-away index for ``docs``), retrievers.py:150
    (``get_scores``).
    """

    def __init__(self, corpus: Sequence[Sequence[Hashable]], k1: float = K1,
                 b: float = B, epsilon: float = EPSILON):
        self.k1, self.b, self.epsilon = k1, b, epsilon
        self.corpus_size = 0
        self.doc_freqs: List[Dict[Hashable, int]] = []
        self.doc_len: List[int] = []
        self.idf: Dict[Hashable, float] = {}
        nd: Dict[Hashable, int] = {}     # term -> number of docs containing it
        total = 0
        for doc in corpus:
            self.doc_len.append(len(doc))
            total += len(doc)
            tf: Dict[Hashable, int] = {}
            for w in doc:
                tf[w] = tf.get(w, 0) + 1
            self.doc_freqs.append(tf)
            for w in tf:                 # first-seen order inside the doc
                nd[w] = nd.get(w, 0) + 1
            self.corpus_size += 1
        self.avgdl = total / self.corpus_size      # ZeroDivisionError on empty corpus, as upstream
        self.nd = nd
        # idf with the epsilon floor: terms in > half the docs get eps * mean idf
        idf_sum = 0.0
        negative = []
        for w, n in nd.items():          # dict insertion order == first-seen order
            v = math.log(self.corpus_size - n + 0.5) - math.log(n + 0.5)
            self.idf[w] = v
            idf_sum += v
            if v < 0:
                negative.append(w)
        self.average_idf = idf_sum / len(self.idf)
        eps = self.epsilon * self.average_idf
        for w in negative:
            self.idf[w] = eps

    def get_scores(self, query: Sequence[Hashable]) -> np.ndarray:
        score = np.zeros(self.corpus_size)
        doc_len = np.array(self.doc_len)
        for q in query:                  # query order, duplicates repeat
            q_freq = np.array([(d.get(q) or 0) for d in self.doc_freqs])
            score += (self.idf.get(q) or 0) * (
                q_freq * (self.k1 + 1)
                / (q_freq + self.k1 * (1 - self.b + self.b * doc_len / self.avgdl)))
        return score


# --------------------------------------------------------------------------
# rank_bm25.BM25Okapi, CSR formulation (bit-identical, fast)
# --------------------------------------------------------------------------
class OkapiCSR:
    """Same numbers as :class:`OkapiLiteral`, built from integer term ids.

    ``docs`` is a list of int arrays (term ids, document order preserved).  The
    per-element evaluation order is the one numpy applies to the literal
    expression: t1=b*dl; t2=t1/avgdl; t3=(1-b)+t2; K=k1*t3; den=tf+K;
    num=tf*(k1+1); r=num/den; c=idf*r; score=score+c  (each IEEE-754 RN, no FMA).
    Documents that do not contain the term receive +/-0.0, which leaves the
    running float64 sum unchanged, so only postings are touched.
    """

    def __init__(self, docs: Sequence[np.ndarray], vocab_size: int, k1: float = K1,
                 b: float = B, epsilon: float = EPSILON):
        self.k1, self.b, self.epsilon = k1, b, epsilon
        n = len(docs)
        self.corpus_size = n
        self.doc_len = np.array([len(d) for d in docs], dtype=np.int64)
        self.avgdl = int(self.doc_len.sum()) / n
        # postings, document-major then transposed to term-major
        first_seen_order: List[int] = []
        seen = np.zeros(vocab_size, dtype=bool)
        p_doc, p_term, p_tf = [], [], []
        for i, d in enumerate(docs):
            d = np.asarray(d, dtype=np.int64)
            if d.size == 0:
                continue
            terms, first_pos, counts = np.unique(d, return_index=True, return_counts=True)
            order = np.argsort(first_pos, kind="stable")      # first-seen order inside doc
            for t in terms[order]:
                if not seen[t]:
                    seen[t] = True
                    first_seen_order.append(int(t))
            p_doc.append(np.full(terms.size, i, dtype=np.int64))
            p_term.append(terms)
            p_tf.append(counts)
        p_doc = np.concatenate(p_doc) if p_doc else np.zeros(0, np.int64)
        p_term = np.concatenate(p_term) if p_term else np.zeros(0, np.int64)
        p_tf = np.concatenate(p_tf) if p_tf else np.zeros(0, np.int64)
        o = np.lexsort((p_doc, p_term))
        self.post_doc = p_doc[o]
        self.post_tf = p_tf[o]
        df = np.bincount(p_term, minlength=vocab_size)
        self.df = df
        self.indptr = np.zeros(vocab_size + 1, dtype=np.int64)
        np.cumsum(df, out=self.indptr[1:])
        # idf, sequential float64 sum in first-seen order
        idf = np.zeros(vocab_size, dtype=np.float64)
        idf_sum = 0.0
        negative = []
        for t in first_seen_order:
            v = math.log(n - int(df[t]) + 0.5) - math.log(int(df[t]) + 0.5)
            idf[t] = v
            idf_sum += v
            if v < 0:
                negative.append(t)
        self.first_seen_order = np.array(first_seen_order, dtype=np.int64)
        self.average_idf = idf_sum / len(first_seen_order)
        for t in negative:
            idf[t] = self.epsilon * self.average_idf
        self.idf = idf
        self.K_d = self.k1 * ((1 - self.b) + (self.b * self.doc_len) / self.avgdl)

    def contributions(self, t: int) -> np.ndarray:
        """float64 ``idf * tf*(k1+1)/(tf+K_d)`` for every posting of term ``t``."""
        s, e = self.indptr[t], self.indptr[t + 1]
        tf = self.post_tf[s:e]
        return self.idf[t] * (tf * (self.k1 + 1) / (tf + self.K_d[self.post_doc[s:e]]))

    def get_scores(self, query_ids: Sequence[int]) -> np.ndarray:
        score = np.zeros(self.corpus_size)
        for t in query_ids:
            if t < 0 or t >= self.idf.shape[0] or self.idf[t] == 0.0:
                continue                        # unknown term / idf 0 -> contributes +/-0
            s, e = self.indptr[t], self.indptr[t + 1]
            score[self.post_doc[s:e]] += self.contributions(int(t))
        return score


# --------------------------------------------------------------------------
# bm25s.BM25 (method="lucene", float32)  -- PARITY UNPINNED
# --------------------------------------------------------------------------
class Bm25sLucene:
    """Restatement of bm25s 0.1.7 with defaults (reference: retrievers.py:107-111,136-140).

    idf = log(1 + (N - df + 0.5)/(df + 0.5)) stored as float32;
    tfc = tf / (k1*((1-b) + b*l_d/l_avg) + tf) in float64;
    stored weight = float32(float64(idf32) * tfc);
    get_scores: float32 zeros, then for each known query token in order
    ``np.add.at(scores, docs_of_token, weights_of_token)``.
    """

    def __init__(self, docs: Sequence[np.ndarray], vocab_size: int, k1: float = K1, b: float = B):
        n = len(docs)
        self.corpus_size = n
        doc_len = np.array([len(d) for d in docs], dtype=np.int64)
        l_avg = float(doc_len.mean()) if n else 0.0
        p_doc, p_term, p_w = [], [], []
        df = np.zeros(vocab_size, dtype=np.int64)
        per_doc = []
        for i, d in enumerate(docs):
            d = np.asarray(d, dtype=np.int64)
            terms, counts = np.unique(d, return_counts=True)
            df[terms] += 1
            per_doc.append((terms, counts))
        idf32 = np.zeros(vocab_size, dtype=np.float32)
        nz = df > 0
        idf32[nz] = np.array([math.log(1 + (n - int(x) + 0.5) / (int(x) + 0.5)) for x in df[nz]],
                             dtype=np.float64).astype(np.float32)
        self.idf32 = idf32
        for i, (terms, counts) in enumerate(per_doc):
            tfc = counts / (k1 * ((1 - b) + b * int(doc_len[i]) / l_avg) + counts)
            w = (idf32[terms] * tfc).astype(np.float32)
            p_doc.append(np.full(terms.size, i, dtype=np.int64))
            p_term.append(terms)
            p_w.append(w)
        p_doc = np.concatenate(p_doc) if p_doc else np.zeros(0, np.int64)
        p_term = np.concatenate(p_term) if p_term else np.zeros(0, np.int64)
        p_w = np.concatenate(p_w) if p_w else np.zeros(0, np.float32)
        o = np.lexsort((p_doc, p_term))
        self.post_doc, self.post_w = p_doc[o], p_w[o]
        self.df = df
        self.indptr = np.zeros(vocab_size + 1, dtype=np.int64)
        np.cumsum(np.bincount(p_term, minlength=vocab_size), out=self.indptr[1:])

    def get_scores(self, query_ids: Sequence[int]) -> np.ndarray:
        scores = np.zeros(self.corpus_size, dtype=np.float32)
        for t in query_ids:
            if t < 0 or t >= self.df.shape[0] or self.df[t] == 0:
                continue                        # token not in vocab: dropped
            s, e = self.indptr[t], self.indptr[t + 1]
            np.add.at(scores, self.post_doc[s:e], self.post_w[s:e])
        return scores
