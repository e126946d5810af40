"""GPU: the drop-in retriever classes against a literal replay of the reference's own control flow.

The "reference side" below is retrievers.py re-enacted with the oracle pieces: tokenise with the
same tokenizer object, rank_bm25 restatement (literal per-document loop), ``filter``, ``fusion``,
``reciprocal_rank_fusion`` on plain node objects.  The call pattern mirrors pipeline.py:190-237,
331-365,400-409.
"""
import asyncio

import numpy as np
import pytest
import torch

from oracle import bm25 as obm
from oracle import retrieve as ort
from easyrag_b200 import synth, _lib
from easyrag_b200.retrievers import B200VectorStore, BM25Retriever, HybridRetriever, QdrantRetriever
from easyrag_b200.schema import BaseEmbedding, NodeWithScore, QueryBundle, TextNode, build_qdrant_filters

pytestmark = pytest.mark.gpu
DIRS = ["director", "emsplus", "rcp", "umac"]       # the four `document` values of src/data/question.jsonl


@pytest.fixture(scope="module")
def world(lib_built):
    _lib.require_cuda()
    n, vocab, dim = 1200, 900, 64
    corpus = synth.make_sparse_corpus(n, vocab, 41, mean_len=30, min_len=1, max_len=80)
    docs = corpus.doc_lists()
    rng = np.random.default_rng(42)
    g = torch.Generator().manual_seed(43)
    emb = torch.nn.functional.normalize(torch.randn(n, dim, generator=g), dim=1)
    emb = emb.to(torch.bfloat16).float()                          # what GTEEmbedding emits: bf16 values as floats
    nodes = []
    for i, d in enumerate(docs):
        text = synth.ids_to_text(d)
        if i in (17, 400, 401):                                    # duplicate chunk texts (overlapping windows)
            text = synth.ids_to_text(docs[5])
        nodes.append(TextNode(text=text, id_=f"node-{i}",
                              metadata={"dir": DIRS[int(rng.integers(4))], "file_path": f"f{i % 7}",
                                        "know_path": synth.ids_to_text(d[:3])},
                              embedding=emb[i].tolist()))
    for i in (17, 400, 401):
        nodes[i].embedding = nodes[5].embedding
    tk = synth.PseudoWordTokenizer()
    stop = {f"w{i}" for i in range(5)}
    queries = [synth.ids_to_text(t) for t in synth.make_queries(corpus, 30, 44, min_terms=2, max_terms=8).term_lists()]
    return dict(nodes=nodes, tk=tk, stop=stop, queries=queries, emb=emb, dim=dim)


class _FakeEmbedding(BaseEmbedding):
    """Deterministic query embedding (the encoder has its own tests): hash of the text -> unit vector."""

    def __init__(self, dim):
        super().__init__(model_name="fake", embed_batch_size=8)
        self._dim = dim

    def _get_query_embedding(self, query):
        g = torch.Generator().manual_seed(abs(hash(query)) % (1 << 31))
        v = torch.nn.functional.normalize(torch.randn(self._dim, generator=g), dim=0)
        return v.to(torch.bfloat16).float().tolist()

    _get_text_embedding = _get_query_embedding


def _ref_sparse(world, query, k, embed_type=0, filter_dict=None):
    from easyrag_b200.retrievers import get_node_content
    nodes = world["nodes"]
    corpus = [ort.tokenize_and_remove_stopwords(world["tk"], get_node_content(n, embed_type), world["stop"])
              for n in nodes]
    model = obm.OkapiLiteral(corpus)
    toks = ort.tokenize_and_remove_stopwords(world["tk"], query, world["stop"])
    scores = model.get_scores(toks)
    onodes = [ort.ONode(n.get_content(), i, n.metadata) for i, n in enumerate(nodes)]
    return scores, ort.bm25_filter(scores, onodes, k, filter_dict)


def test_bm25_retriever_matches_reference_flow(world):
    r = BM25Retriever.from_defaults(nodes=world["nodes"], tokenizer=world["tk"], similarity_top_k=192,
                                    stopwords=world["stop"], embed_type=0, bm25_type=0)
    for qi, query in enumerate(world["queries"][:8]):
        r.filter_dict = {"dir": DIRS[qi % 4]} if qi % 2 else None
        got = asyncio.run(r.aretrieve(QueryBundle(query)))
        scores, ref = _ref_sparse(world, query, 192, filter_dict=r.filter_dict)
        assert [g.node.node_id for g in got] == [f"node-{o.node.idx}" for o in ref]
        assert [g.score for g in got] == [o.score for o in ref]           # float64, bit-exact
        assert all(isinstance(g, NodeWithScore) and g.node is world["nodes"][o.node.idx] for g, o in zip(got, ref))
        # get_scores / filter, the two-step form used by compressors.py and by _retrieve in the reference
        s = r.get_scores(query)
        assert s.tobytes() == scores.tobytes()
        again = r.filter(s)
        assert [g.node.node_id for g in again] == [g.node.node_id for g in got]


def test_bm25_retriever_path_route_and_small_k(world):
    # pipeline.py:201-208: second retriever over know_path strings, k = f_topk_3 = 6
    r = BM25Retriever.from_defaults(nodes=world["nodes"], tokenizer=world["tk"], similarity_top_k=6,
                                    stopwords=world["stop"], embed_type=5, bm25_type=0)
    for query in world["queries"][:6]:
        got = r.retrieve(query)
        _, ref = _ref_sparse(world, query, 6, embed_type=5)
        assert [g.node.node_id for g in got] == [f"node-{o.node.idx}" for o in ref]
        assert [g.score for g in got] == [o.score for o in ref]


def test_get_scores_with_adhoc_docs(world):
    # compressors.py:42 -> retrievers.py:131-147: throw-away index over the given sentences
    r = BM25Retriever.from_defaults(nodes=world["nodes"][:50], tokenizer=world["tk"], similarity_top_k=4,
                                    stopwords=world["stop"])
    docs = [n.get_content() for n in world["nodes"][100:140]]
    query = world["queries"][3]
    got = r.get_scores(query, docs)
    corpus = [ort.tokenize_and_remove_stopwords(world["tk"], d, world["stop"]) for d in docs]
    ref = obm.OkapiLiteral(corpus).get_scores(ort.tokenize_and_remove_stopwords(world["tk"], query, world["stop"]))
    assert got.tobytes() == ref.tobytes()
    assert np.array_equal(got.argsort(), ref.argsort())


def test_from_defaults_argument_check(world):
    with pytest.raises(ValueError):
        BM25Retriever.from_defaults(tokenizer=world["tk"])                  # retrievers.py:167-168


def test_dense_retriever_and_filters(world):
    store = B200VectorStore(world["nodes"])
    dense = QdrantRetriever(store, _FakeEmbedding(world["dim"]), similarity_top_k=20)
    emb = world["emb"].numpy().copy()
    for i in (17, 400, 401):
        emb[i] = emb[5]
    emb = emb / np.linalg.norm(emb, axis=1, keepdims=True)
    for qi, query in enumerate(world["queries"][:6]):
        dense.filters = build_qdrant_filters(DIRS[qi % 4]) if qi % 2 else None
        got = asyncio.run(dense.aretrieve(QueryBundle(query)))
        qv = np.array(_FakeEmbedding(world["dim"])._get_query_embedding(query), dtype=np.float32)
        qv = qv / np.linalg.norm(qv)
        allowed = None
        if dense.filters is not None:
            allowed = np.array([n.metadata["dir"] == DIRS[qi % 4] for n in world["nodes"]])
        ref_i, ref_s = ort.dense_topk(emb, qv[None], 20, allowed)
        sims = emb @ qv
        assert len(got) == int((ref_i[0] >= 0).sum())
        got_idx = [int(g.node.node_id.split("-")[1]) for g in got]
        assert np.abs(np.array([g.score for g in got]) - sims[got_idx]).max() <= 1e-3
        assert (sims[got_idx] >= ref_s[0, len(got) - 1] - 1e-3).all()
        if allowed is not None:
            assert allowed[got_idx].all()
        assert [g.score for g in got] == sorted((g.score for g in got), reverse=True)


def _dense_lists(store, world, k=15, filters=None):
    out = []
    for query in world["queries"][:8]:
        r = QdrantRetriever(store, _FakeEmbedding(world["dim"]), similarity_top_k=k)
        r.filters = filters
        got = asyncio.run(r.aretrieve(QueryBundle(query)))
        out.append([(g.node.node_id, g.score) for g in got])
    return out


def test_vector_store_append_equals_bulk_insert(world):
    """``add`` appends (amortised O(new nodes)) and ``add_embedded`` takes a tensor: same collection either way."""
    nodes = world["nodes"]
    bulk = B200VectorStore(nodes)
    piecewise = B200VectorStore(nodes[:400])
    piecewise.add(nodes[400:900])
    emb = torch.tensor([n.embedding for n in nodes[900:]], dtype=torch.float32)
    piecewise.add_embedded(nodes[900:], emb.to("cuda"))
    assert piecewise.index.n_rows == bulk.index.n_rows == len(nodes)
    assert torch.equal(piecewise.index.vectors, bulk.index.vectors)
    assert _dense_lists(piecewise, world) == _dense_lists(bulk, world)
    f = build_qdrant_filters("rcp")
    assert _dense_lists(piecewise, world, filters=f) == _dense_lists(bulk, world, filters=f)
    # the per-key class table is uploaded once per key set, not per query
    assert list(piecewise._group_cache) == [("dir",)]
    before = piecewise._group_cache[("dir",)].data_ptr()
    _dense_lists(piecewise, world, filters=build_qdrant_filters("umac"))
    assert piecewise._group_cache[("dir",)].data_ptr() == before


def test_vector_store_node_without_the_filter_key_matches_no_filter(world):
    nodes = list(world["nodes"][:300])
    bare = TextNode(text="no dir here", id_="node-bare", metadata={}, embedding=nodes[0].embedding)
    store = B200VectorStore(nodes + [bare])
    r = QdrantRetriever(store, _FakeEmbedding(world["dim"]), similarity_top_k=301)
    got_all = asyncio.run(r.aretrieve(QueryBundle(world["queries"][0])))
    assert "node-bare" in [g.node.node_id for g in got_all]
    for d in DIRS:
        r.filters = build_qdrant_filters(d)
        got = asyncio.run(r.aretrieve(QueryBundle(world["queries"][0])))          # a qdrant payload filter: no error
        assert "node-bare" not in [g.node.node_id for g in got]
        assert all(g.node.metadata["dir"] == d for g in got)


class _TensorEmbedding(_FakeEmbedding):
    """An embedding model with the ``embed_tensor`` fast path of GTEEmbedding / HuggingFaceEmbedding."""

    def __init__(self, dim, table):
        super().__init__(dim)
        self._table = table
        self.embed_batch_size = 7
        self.calls = 0

    def embed_tensor(self, texts):
        self.calls += 1
        rows = torch.stack([self._table[t] for t in texts])
        return rows.to(torch.bfloat16).to("cuda"), rows.to("cuda")


def test_vector_store_from_embed_model_writes_encoder_rows_in_place(world):
    nodes = world["nodes"][:100]
    g = torch.Generator().manual_seed(5)
    table = {n.get_content(): torch.randn(world["dim"], generator=g) * 3 for n in nodes}
    model = _TensorEmbedding(world["dim"], table)
    store = B200VectorStore.from_embed_model(nodes, model)
    assert model.calls == -(-100 // 7) and store.index.n_rows == 100 and len(store.nodes) == 100
    want = torch.stack([table[n.get_content()] for n in nodes]).to(torch.bfloat16).float()
    want = (want / want.norm(dim=1, keepdim=True)).to(torch.bfloat16)
    got = store.index.vectors.cpu()
    assert (got.float() - want.float()).abs().max().item() <= 2 ** -8          # one bf16 ulp of a unit-vector entry
    assert (got.float().norm(dim=1) - 1).abs().max().item() < 5e-3


def test_hybrid_retriever_rrf_matches_reference_flow(world):
    sparse = BM25Retriever.from_defaults(nodes=world["nodes"], tokenizer=world["tk"], similarity_top_k=24,
                                         stopwords=world["stop"])
    dense = QdrantRetriever(B200VectorStore(world["nodes"]), _FakeEmbedding(world["dim"]), similarity_top_k=24)
    hybrid = HybridRetriever(dense_retriever=dense, sparse_retriever=sparse, retrieval_type=3, topk=16)
    for query in world["queries"][:6]:
        got = asyncio.run(hybrid.aretrieve(QueryBundle(query)))
        s_nodes = asyncio.run(sparse.aretrieve(QueryBundle(query)))
        d_nodes = asyncio.run(dense.aretrieve(QueryBundle(query)))
        # reference fusion on the same two lists (text-keyed, sparse first, last writer wins)
        to_o = lambda lst: [ort.OScored(ort.ONode(x.get_content(), int(x.node.node_id.split("-")[1])), x.score)
                            for x in lst]
        ref = ort.reciprocal_rank_fusion([to_o(s_nodes), to_o(d_nodes)], topk=16)
        assert [int(g.node.node_id.split("-")[1]) for g in got] == [o.node.idx for o in ref]
        assert [g.score for g in got] == [o.score for o in ref]
    hybrid.retrieval_type = 2
    only_sparse = asyncio.run(hybrid.aretrieve(QueryBundle(world["queries"][0])))
    assert [g.node.node_id for g in only_sparse] == \
        [g.node.node_id for g in asyncio.run(sparse.aretrieve(QueryBundle(world["queries"][0])))]


def test_fusion_classmethods_on_the_class(world):
    # pipeline.py:362,408 call them on the class, with lists from two different BM25 retrievers
    chunk = BM25Retriever.from_defaults(nodes=world["nodes"], tokenizer=world["tk"], similarity_top_k=30,
                                        stopwords=world["stop"])
    path = BM25Retriever.from_defaults(nodes=world["nodes"], tokenizer=world["tk"], similarity_top_k=6,
                                       stopwords=world["stop"], embed_type=5)
    for query in world["queries"][:5]:
        a, b = chunk.retrieve(query), path.retrieve(query)
        to_o = lambda lst: [ort.OScored(ort.ONode(x.get_content(), int(x.node.node_id.split("-")[1])), x.score)
                            for x in lst]
        ref = ort.fusion([to_o(a), to_o(b)], topk=20)
        got = HybridRetriever.fusion([a, b], topk=20)
        assert [int(g.node.node_id.split("-")[1]) for g in got] == [o.node.idx for o in ref]
        assert [g.score for g in got] == [o.score for o in ref]
        ref = ort.reciprocal_rank_fusion([to_o(a), to_o(b)], topk=6)
        got = HybridRetriever.reciprocal_rank_fusion([a, b], topk=6)
        assert [int(g.node.node_id.split("-")[1]) for g in got] == [o.node.idx for o in ref]
        assert [g.score for g in got] == [o.score for o in ref]
    assert HybridRetriever.fusion([[], []]) == []
    assert HybridRetriever.reciprocal_rank_fusion([[], []]) == []
    # the reference accepts any number of lists (retrievers.py:243,261): one and three
    q = world["queries"][0]
    a, b = chunk.retrieve(q), path.retrieve(q)
    c = list(reversed(chunk.retrieve(world["queries"][1])))[:7]
    for lists in ([a], [a, b, c], [b, [], c, a]):
        ref = ort.reciprocal_rank_fusion([to_o(l) for l in lists], topk=9)
        got = HybridRetriever.reciprocal_rank_fusion([list(l) for l in lists], topk=9)
        assert [int(g.node.node_id.split("-")[1]) for g in got] == [o.node.idx for o in ref]
        assert [g.score for g in got] == [o.score for o in ref]
    a, b = chunk.retrieve(q), path.retrieve(q)             # fresh scores (RRF overwrote them, as in the reference)
    ref = ort.fusion([to_o(l) for l in (a, b, c)], topk=15)
    got = HybridRetriever.fusion([a, b, c], topk=15)
    assert [int(g.node.node_id.split("-")[1]) for g in got] == [o.node.idx for o in ref]


def test_rerank_packer_builds_the_reference_inputs_on_the_device(lib_built):
    """rerankers.py:196-293 (get_inputs / get_inputs_v2_5) + the 32-pair slices of :309-322, from fused device ids."""
    from easyrag_b200 import handoff
    rng = np.random.default_rng(9)
    n_docs, nq, k, max_length = 400, 7, 40, 64
    passages = [rng.integers(5, 1000, rng.integers(0, 120)).tolist() for _ in range(n_docs)]      # some longer than max_length
    sep, prompt, bos = [13], rng.integers(5, 1000, 9).tolist(), 1
    packer = handoff.RerankPacker(passages, sep, prompt, bos, max_length=max_length)
    queries = [rng.integers(5, 1000, rng.integers(1, 70)).tolist() for _ in range(nq)]            # some beyond 3/4 max_length
    q_ptr = torch.tensor(np.cumsum([0] + [len(q) for q in queries]), dtype=torch.int32)
    q_tok = torch.tensor([t for q in queries for t in q], dtype=torch.int32)
    cnt = rng.integers(0, k + 1, nq).astype(np.int32)
    cnt[0], cnt[1] = k, 0
    ids = np.full((nq, k), -1, np.int32)
    for q in range(nq):
        ids[q, :cnt[q]] = rng.choice(n_docs, cnt[q], replace=False)
    out = packer.pack(torch.from_numpy(ids).cuda(), torch.from_numpy(cnt).cuda(), q_ptr, q_tok)
    cu, tok, qlen = out.cu.cpu().numpy(), out.ids.cpu().numpy(), out.query_len.cpu().numpy()
    assert cu[0] == 0 and cu[-1] == tok.size and out.prompt_len == len(sep) + len(prompt)
    for q in range(nq):
        want, want_ql, want_pl = ort.rerank_inputs(queries[q], [passages[d] for d in ids[q, :cnt[q]]], sep, prompt, bos,
                                                   max_length)
        assert all(p == out.prompt_len for p in want_pl)
        for r in range(k):
            p = q * k + r
            got = tok[cu[p]:cu[p + 1]].tolist()
            if r < cnt[q]:
                assert got == want[r], (q, r)
                assert qlen[p] == want_ql[r] and len(got) <= max_length + len(sep) + len(prompt)
            else:
                assert got == [] and qlen[p] == 0
    # slices in the reference's order: per query, 32 pairs at a time
    sl = list(out.slices(32))
    assert sl[:2] == [(0, 0, 32), (0, 32, 40)] and len(sl) == nq * 2 and sl[-1] == (nq - 1, (nq - 1) * k + 32, nq * k)
