"""Test infrastructure: the integer counting of the BM25 index build restated with numpy (no GPU).

The product counts on the GPU only (csrc/bm25_build.cu behind ``Bm25Stats.from_tokens``).  CPU tests that exercise
the host-exact arithmetic (``Bm25Stats.from_counts``: math.log idf, first-seen-order float64 sum) and the GPU tests
that check the kernels both use this independent restatement of retrievers.py:98-118 -> rank_bm25 ``_initialize``.
"""
import numpy as np
import torch

from easyrag_b200.index import Bm25Stats


def host_counts(tokens, doc_ptr, vocab):
    """-> dict(df, indptr, post_doc, post_tf, first_pos, doc_len): term-major postings, documents ascending."""
    tok = np.asarray(tokens.cpu() if hasattr(tokens, "cpu") else tokens).astype(np.int64)
    ptr = np.asarray(doc_ptr.cpu() if hasattr(doc_ptr, "cpu") else doc_ptr).astype(np.int64)
    n = ptr.size - 1
    lens = ptr[1:] - ptr[:-1]
    if tok.size and (tok.min() < 0 or tok.max() >= vocab):
        raise ValueError("token id out of range [0, vocab)")
    doc_of = np.repeat(np.arange(n, dtype=np.int64), lens)
    key = tok * max(n, 1) + doc_of
    ukey, counts = np.unique(key, return_counts=True)               # sorted: term-major, doc ascending
    post_term = ukey // max(n, 1)
    post_doc = (ukey - post_term * max(n, 1)).astype(np.int32)
    df = np.bincount(post_term, minlength=vocab).astype(np.int64)
    indptr = np.zeros(vocab + 1, dtype=np.int64)
    np.cumsum(df, out=indptr[1:])
    first_pos = np.full(vocab, np.iinfo(np.uint64).max, dtype=np.uint64)
    if tok.size:
        np.minimum.at(first_pos, tok, np.arange(tok.size, dtype=np.uint64))
    return dict(df=df, indptr=indptr, post_doc=post_doc, post_tf=counts.astype(np.int32), first_pos=first_pos,
                doc_len=lens.astype(np.int32))


def stats_from_host_counts(tokens, doc_ptr, vocab, bm25_type=0) -> Bm25Stats:
    n = int(doc_ptr.numel()) - 1
    if n == 0:
        raise ZeroDivisionError("division by zero")                  # rank_bm25: avgdl = num_doc / corpus_size
    c = host_counts(tokens, doc_ptr, vocab)
    t = torch.from_numpy
    return Bm25Stats.from_counts(n, vocab, int(c["doc_len"].sum()), t(c["doc_len"]), t(c["df"]), t(c["indptr"]),
                                 t(c["post_doc"]), t(c["post_tf"]), c["first_pos"], bm25_type=bm25_type)
