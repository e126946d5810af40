"""CPU: the reference arm of bench.py (oracle port of the reference's CPU retrievers) is itself checked here."""
import sys
from pathlib import Path
from types import SimpleNamespace

import numpy as np
import torch

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
    data = dict(stats=Bm25Stats.from_tokens(corpus.tokens, corpus.doc_ptr, vocab), queries=queries, vec=vec, qvec=qvec)
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
