"""CPU: the reference arm of bench.py (oracle port of the reference's CPU retrievers) is itself checked here."""
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

from _host_counts import stats_from_host_counts

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import bench                                   # noqa: E402
from oracle import bm25 as obm                 # noqa: E402
from oracle import retrieve as ort             # noqa: E402
from easyrag_b200 import synth                 # noqa: E402
from easyrag_b200.index import Bm25Stats       # noqa: E402


def test_cpu_reference_arm_matches_oracle():
    n, vocab, dim, nq, k = 3000, 800, 32, 12, 10
    corpus = synth.make_sparse_corpus(n, vocab, 3, mean_len=40, min_len=1, max_len=100)
    queries = synth.make_queries(corpus, nq, 4)
    vec = synth.make_dense_corpus(n, dim, 5)
    qvec = synth.make_dense_queries(vec, nq, 6)
    data = dict(stats=stats_from_host_counts(corpus.tokens, corpus.doc_ptr, vocab), queries=queries, vec=vec, qvec=qvec)
    ref = bench.CpuReference(data, SimpleNamespace(k=k))
    got = ref.run(0, nq)
    model = obm.OkapiCSR(corpus.doc_lists(), vocab)
    d_ids, _ = ort.dense_topk(vec.float().numpy(), qvec.float().numpy(), k)
    for i, terms in enumerate(queries.term_lists()):
        s_ids, _ = ort.bm25_topk_ids(model.get_scores([int(t) for t in terms]), k)
        want_ids, want_sc = ort.rrf_ids([s_ids, d_ids[i]], None, K=60, topk=k)
        ids, sc = got[i]
        # the reference's argsort()[::-1] is not stable: compare scores, and ids outside score ties
        assert np.array_equal(sc, want_sc)
        assert set(ids.tolist()) == set(want_ids.tolist()) or np.array_equal(ids, want_ids)


def test_bench_defaults_finish_fast_and_name_the_metric():
    sys.argv = ["bench.py"]
    a = bench.parse()
    assert a.gpus == 1 and a.steps <= 50 and a.warmup >= 3 and a.rows == 1_000_000 and a.dim == 768 and a.k == 10
    assert "queries/sec" in bench.METRIC and "1M" in bench.METRIC


def _fake_gpu_lists(ref, nq, k):
    """What a correct GPU path returns for the first nq queries: built from the oracle modules, not from bench.py."""
    from types import SimpleNamespace as NS
    model_scores = [ref.bm25_scores(i) for i in range(nq)]
    sims = (ref.qvec[:nq] @ ref.vec.T).numpy()
    s_ids = np.full((nq, k), -1, np.int32); s_sc = np.full((nq, k), -np.inf); s_cnt = np.zeros(nq, np.int32)
    d_ids = np.zeros((nq, k), np.int32); d_sc = np.zeros((nq, k), np.float32)
    f_ids = np.full((nq, k), -1, np.int32); f_sc = np.full((nq, k), -np.inf); f_cnt = np.zeros(nq, np.int32)
    for i in range(nq):
        ii, ss = ort.bm25_topk_ids(model_scores[i], k)
        s_ids[i, :ii.size], s_sc[i, :ii.size], s_cnt[i] = ii, ss, ii.size
        order = ort.canonical_order(sims[i])[:k]
        d_ids[i], d_sc[i] = order, sims[i][order]
        r_ids, r_sc = ort.rrf_ids([ii, order], None, K=60, topk=k)
        f_ids[i, :r_ids.size], f_sc[i, :r_ids.size], f_cnt[i] = r_ids, r_sc, r_ids.size
    t = torch.from_numpy
    return (NS(ids=t(f_ids), scores=t(f_sc), counts=t(f_cnt)), NS(ids=t(s_ids), scores=t(s_sc), counts=t(s_cnt)),
            NS(ids=t(d_ids), scores=t(d_sc), counts=torch.full((nq,), k, dtype=torch.int32)))


def test_parity_full_size_accepts_the_oracle_and_rejects_a_wrong_list():
    n, vocab, dim, nq, k = 4000, 300, 32, 16, 10
    corpus = synth.make_sparse_corpus(n, vocab, 13, mean_len=30, min_len=1, max_len=80)
    queries = synth.make_queries(corpus, nq, 14)
    vec = synth.make_dense_corpus(n, dim, 15)
    qvec = synth.make_dense_queries(vec, nq, 16)
    data = dict(stats=stats_from_host_counts(corpus.tokens, corpus.doc_ptr, vocab), queries=queries, vec=vec, qvec=qvec)
    ref = bench.CpuReference(data, SimpleNamespace(k=k))
    f, s, d = _fake_gpu_lists(ref, nq, k)
    res = bench.parity_full_size(ref, nq, f, s, d, k)
    assert res["ok"] and res["bm25_bit_exact"] and res["rrf_equal"] and res["dense_ids_equal"] == nq, res
    assert res["fused_lists_identical_to_oracle"] == nq and res["dense_max_abs"] == 0.0
    # a dense score off by more than the tolerance, a BM25 score off by one ulp, a swapped fused pair: all caught
    d2 = SimpleNamespace(ids=d.ids, scores=d.scores.clone(), counts=d.counts)
    d2.scores[3, 0] += 2e-3
    assert not bench.parity_full_size(ref, nq, f, s, d2, k)["dense_within_tol"]
    s2 = SimpleNamespace(ids=s.ids, scores=s.scores.clone(), counts=s.counts)
    s2.scores[5, 0] = np.nextafter(float(s2.scores[5, 0]), 1e9)
    assert not bench.parity_full_size(ref, nq, f, s2, d, k)["bm25_bit_exact"]
    f2 = SimpleNamespace(ids=f.ids.clone(), scores=f.scores, counts=f.counts)
    f2.ids[7, [0, 1]] = f2.ids[7, [1, 0]]
    assert not bench.parity_full_size(ref, nq, f2, s, d, k)["rrf_equal"]
    # a near-tie swap inside the tolerance is accepted and counted
    d3 = SimpleNamespace(ids=d.ids.clone(), scores=d.scores.clone(), counts=d.counts)
    sims = (ref.qvec[:nq] @ ref.vec.T).numpy()
    nxt = ort.canonical_order(sims[2])[k]                       # the 11th best replaces the 10th, scores barely differ?
    if sims[2][d.ids[2, k - 1]] - sims[2][nxt] < 1e-3:
        d3.ids[2, k - 1] = int(nxt); d3.scores[2, k - 1] = float(sims[2][nxt])
        f3, _, _ = f, None, None
        r = bench.parity_full_size(ref, nq, f, s, d3, k)
        assert r["dense_within_tol"] and r["dense_near_tie_swaps"] == 1


def test_canonical_topk_matches_stable_argsort_with_ties():
    rng = np.random.default_rng(0)
    for n, k in ((50, 10), (1000, 10), (7, 10), (10, 10)):
        score = rng.integers(0, 6, n).astype(np.float64)          # heavy ties, zeros included
        ids, sc, inside, straddle = bench.canonical_topk(score, k, positive_only=True)
        want_ids, want_sc = ort.bm25_topk_ids(score, k)
        assert np.array_equal(ids, want_ids) and np.array_equal(sc, want_sc)
        assert inside == bool(want_sc.size > 1 and np.any(want_sc[1:] == want_sc[:-1]))
        order = ort.canonical_order(score)
        if n > k and want_sc.size == k:
            assert straddle == bool(score[order[k]] == score[order[k - 1]])
