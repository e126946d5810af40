"""CPU oracle for the EasyRAG coarse-ranking path.  TEST INFRASTRUCTURE ONLY.

This package restates, in plain Python / numpy, the arithmetic and control flow
of the reference's coarse-ranking path (src/easyrag/custom/retrievers.py and
the un-vendored pip dependencies it calls: rank-bm25==0.2.2, bm25s==0.1.7,
qdrant local-mode cosine search).  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.
The product package ``easyrag_b200`` never imports anything from here.

Pinning status (see DESIGN.md "Oracle"):
  * BM25Okapi restatement: pinned against the rank_bm25 README known answer
    (tests/golden/kat.json) -- the only published vector for this path.
  * RRF / fusion / filter: pinned against hand-computed values derived from
    retrievers.py:239-274,191-210.
  * Qwen2 (GTE) encoder: pinned against outputs of the *reference's own*
    vendored model (src/easyrag/utils/modeling_qwen.py) generated in the
    authoring container by tests/golden/make_encoder_golden.py.
  * bm25s (bm25_type=1) and the BERT-shaped encoder: the upstream packages are
    absent from /root/reference and from this image -> PARITY UNPINNED for
    those two (restated from the published algorithm only).
"""
