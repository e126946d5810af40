"""CPU check of the inequalities the two-phase BM25 top-k rests on (easyrag_b200/csrc/bm25_pk.cuh header).

The GPU candidate pass accumulates integer upper-bound scores ``Q(d) = sum_j ceil(w_j * 2^e)`` and keeps a document
when ``Q(d) >= B - 1`` for a bound ``B = (k-th largest Q) - m - 1``; exact float64 scores are recomputed for the
survivors only.  That is exact iff the survivors are a superset of the exact top-k.  Restated here with numpy on
the oracle's contributions (same scale selection as ``ezr_bm25_pack``), so the argument is pinned without a GPU:

* the bracket ``Q - m - 1 <= 2^e * s <= Q + 1`` for every document and query,
* superset: canonical exact top-k  is contained in  {d : Q(d) >= B - 1},
* the same with the lowest-weight tokens skipped (``ezr_bm25_set_skipping``): ``Q_ess(d) >= B - 1 - NE``.
"""
import math

import numpy as np
import pytest

from oracle import bm25 as obm
from oracle import retrieve as ort
from easyrag_b200 import synth

WBITS = 19          # 32 - log2(8192): packed weight field of the default build


def _scale_log2(wmax: float) -> int:
    """ezr_bm25_pack: wmax < 2^ex (frexp)  ->  e = WBITS - 1 - ex, so ceil(w * 2^e) < 2^(WBITS-1)."""
    if wmax <= 0:
        return 0
    _, ex = math.frexp(wmax)
    return WBITS - 1 - ex


@pytest.fixture(scope="module")
def case():
    corpus = synth.make_sparse_corpus(6000, 3000, 99, mean_len=60, min_len=1, max_len=200)
    o = obm.OkapiCSR(corpus.doc_lists(), corpus.vocab)
    assert (o.idf >= 0).all()
    wmax = max(float(o.contributions(int(t)).max()) for t in np.nonzero(o.df)[0])
    e = _scale_log2(wmax)
    queries = synth.make_queries(corpus, 60, 100)
    lists = [[int(t) for t in terms] for terms in queries.term_lists()]
    present = np.nonzero(o.df)[0]
    rng = np.random.default_rng(3)
    lists += [[int(t) for t in rng.choice(present, 40)], [int(present[0])] * 5 + [int(present[7])]]
    return o, e, lists


def _packed(o, t, e):
    w = o.contributions(t)
    q = np.ceil(np.ldexp(w, e))
    assert (q < (1 << (WBITS - 1))).all() and (q[w > 0] >= 1).all()
    return q.astype(np.int64)


def _int_scores(o, e, tokens, skip=()):
    """(Q over the non-skipped tokens, m = number of valid tokens incl. skipped)."""
    q = np.zeros(o.corpus_size, dtype=np.int64)
    m = 0
    for j, t in enumerate(tokens):
        if t < 0 or t >= o.idf.shape[0] or o.df[t] == 0:
            continue
        m += 1
        if j in skip:
            continue
        s, en = o.indptr[t], o.indptr[t + 1]
        np.add.at(q, o.post_doc[s:en], _packed(o, t, e))
    return q, m


def test_integer_scores_bracket_the_float64_scores(case):
    o, e, lists = case
    for tokens in lists:
        s = o.get_scores(tokens)
        q, _ = _int_scores(o, e, tokens)
        m = len(tokens)                                   # the kernel uses the token count of the query
        scaled = np.ldexp(s, e)
        assert (q - m - 1 <= scaled).all() and (scaled <= q + 1).all()


@pytest.mark.parametrize("k", [1, 10, 32])
def test_candidates_are_a_superset_of_the_exact_topk(case, k):
    o, e, lists = case
    for tokens in lists:
        s = o.get_scores(tokens)
        ids, _ = ort.bm25_topk_ids(s, k, None)            # canonical exact top-k (positive scores only)
        q, _ = _int_scores(o, e, tokens)
        m = len(tokens)
        if (q > 0).sum() < k:
            continue                                      # fewer than k positives: no bound, everything is kept
        bound = int(np.sort(q)[-k]) - m - 1
        keep = q >= max(bound - 1, 1)
        assert keep[ids].all()


@pytest.mark.parametrize("k", [10])
def test_skipping_lowest_weight_tokens_keeps_the_superset(case, k):
    o, e, lists = case
    num, den = 3, 10                                      # kPkNeNum / kPkNeDen
    skipped_any = 0
    for tokens in lists:
        if len(tokens) > 32:
            continue                                      # the mask covers the first 32 tokens
        s = o.get_scores(tokens)
        ids, _ = ort.bm25_topk_ids(s, k, None)
        q, _ = _int_scores(o, e, tokens)
        m = len(tokens)
        if (q > 0).sum() < k:
            continue
        bound = int(np.sort(q)[-k]) - m - 1
        if bound <= 1:
            continue
        gm = [int(_packed(o, t, e).max()) if (0 <= t < o.idf.shape[0] and o.df[t] > 0) else 0 for t in tokens]
        order = sorted(range(len(tokens)), key=lambda j: (gm[j], j))
        budget = (bound - 1) * num // den
        skip, ne = set(), 0
        for j in order:
            if ne + gm[j] <= budget:
                ne += gm[j]
                skip.add(j)
            else:
                break
        skipped_any += bool(skip)
        q_ess, _ = _int_scores(o, e, tokens, skip=skip)
        assert (q - q_ess <= ne).all()                    # the skipped part never exceeds NE
        keep = q_ess >= max(bound - 1 - ne, 1)
        assert keep[ids].all()
    assert skipped_any > 0


def _chunks(n_ranges):
    """Range chunks of pk_launch (bm25.cu): 4, 4, 8, 16, ... with no tiny last chunk."""
    out, r0, span = [], 0, 4
    while r0 < n_ranges:
        ln = min(n_ranges - r0, span)
        if n_ranges - (r0 + ln) < span // 2:
            ln = n_ranges - r0
        out.append((r0, r0 + ln))
        r0 += ln
        if r0 > 4:
            span *= 2
    return out


@pytest.mark.parametrize("skipping", [False, True])
def test_chunked_bound_protocol_never_loses_a_topk_document(case, skipping):
    """Python model of bm25_cand_kernel + bm25_bound_kernel over many small ranges: running bound B, candidates with
    lower/upper bounds (L, U), pruning by U between chunks, non-essential tokens chosen from the updated bound."""
    o, e, lists = case
    k, R, num, den = 10, 256, 3, 10
    n_ranges = -(-o.corpus_size // R)
    assert len(_chunks(n_ranges)) >= 4 and sum(b - a for a, b in _chunks(n_ranges)) == n_ranges
    pruned_total = skipped_total = 0
    for tokens in lists:
        m = len(tokens)
        s = o.get_scores(tokens)
        ids, _ = ort.bm25_topk_ids(s, k, None)
        q_full, _ = _int_scores(o, e, tokens)
        gm = [int(_packed(o, t, e).max()) if (0 <= t < o.idf.shape[0] and o.df[t] > 0) else 0 for t in tokens]
        bound, skip, ne = 0, set(), 0
        cand = {}                                          # doc -> (L, U)
        for c0, c1 in _chunks(n_ranges):
            q_ess = _int_scores(o, e, tokens, skip=skip)[0] if skip else q_full
            for r in range(c0, c1):
                lo, hi = r * R, min((r + 1) * R, o.corpus_size)
                if bound > 0:                              # track mode: crossing threshold B - 1 - NE
                    thr = max(bound - 1 - ne, 1)
                    new = [d for d in range(lo, hi) if q_ess[d] >= thr]
                    vals = sorted((int(q_ess[d]) for d in new), reverse=True)
                    if len(vals) >= k:
                        bound = max(bound, vals[k - 1] - m - 1)
                else:                                      # no bound yet: this range's own k-th best, all tokens read
                    vals = np.sort(q_full[lo:hi])[::-1]
                    bl = int(vals[k - 1]) - m - 1 if (vals > 0).sum() >= k else 0
                    thr = max(bl - 1, 1)
                    new = [d for d in range(lo, hi) if q_full[d] >= thr]
                    bound = max(bound, bl)
                for d in new:
                    lval = int(q_ess[d]) if skip else int(q_full[d])
                    cand[d] = (lval, lval + (ne if skip else 0))
            if len(cand) >= k:                             # bm25_bound_kernel
                kth = sorted((lu[0] for lu in cand.values()), reverse=True)[k - 1]
                b = max(kth - (m + 1), bound)
                before = len(cand)
                cand = {d: lu for d, lu in cand.items() if lu[1] >= b - 1}
                pruned_total += before - len(cand)
                bound = b
                if skipping and b > 1 and m <= 32:
                    budget = (b - 1) * num // den
                    skip, ne = set(), 0
                    for j in sorted(range(m), key=lambda j: (gm[j], j)):
                        if ne + gm[j] <= budget:
                            ne += gm[j]
                            skip.add(j)
                        else:
                            break
                    skipped_total += len(skip)
        assert set(ids.tolist()) <= set(cand), (tokens, sorted(set(ids.tolist()) - set(cand)))
    assert pruned_total > 0 and (skipped_total > 0) == skipping
