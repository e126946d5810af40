"""CPU: host-side logic of the product + the C-ABI library loads and exports every declared symbol."""
import ctypes
import os
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from _host_counts import stats_from_host_counts

from oracle import bm25 as obm
from easyrag_b200 import synth, _lib
from easyrag_b200.index import Bm25Stats
from easyrag_b200 import dist as ezdist
from easyrag_b200 import schema

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_header_symbol(lib_built):
    header = (ROOT / "include" / "easyrag_b200.h").read_text()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(ezr_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 15
    handle = ctypes.CDLL(str(lib_built))
    missing = [s for s in sorted(declared) if not hasattr(handle, s)]
    assert not missing, f"symbols declared in include/easyrag_b200.h but not exported: {missing}"
    # and the ctypes table covers the same set
    assert declared == set(_lib.SIGNATURES) or declared <= set(_lib.SIGNATURES)


def test_library_loads_without_gpu(lib_built):
    L = _lib.lib()
    assert L.ezr_version() >= 100
    assert isinstance(L.ezr_last_error(), bytes)
    # pure argument validation needs no device
    assert L.ezr_dense_set_kernel(7) == -1
    assert b"dense_set_kernel" in L.ezr_last_error()
    assert L.ezr_dense_set_kernel(0) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_fails_loudly_without_cuda(lib_built):
    with pytest.raises(_lib.EzrError):
        _lib.require_cuda()
    from easyrag_b200.retrievers import HybridRetriever
    from easyrag_b200.schema import NodeWithScore, TextNode
    a = [NodeWithScore(TextNode("x"), 1.0)]
    with pytest.raises(_lib.EzrError):
        HybridRetriever.reciprocal_rank_fusion([a, a])
    # the index build counts on the GPU only: no CPU implementation to fall back to
    c = synth.make_sparse_corpus(20, 30, 1, mean_len=5, min_len=1, max_len=9)
    with pytest.raises(_lib.EzrError):
        Bm25Stats.from_tokens(c.tokens, c.doc_ptr, c.vocab)


@pytest.mark.parametrize("bm25_type", [0, 1])
def test_bm25_stats_match_oracle(bm25_type):
    c = synth.make_sparse_corpus(500, 400, 11, mean_len=30, min_len=0, max_len=90)
    st = stats_from_host_counts(c.tokens, c.doc_ptr, c.vocab, bm25_type=bm25_type)
    docs = c.doc_lists()
    if bm25_type == 0:
        o = obm.OkapiCSR(docs, c.vocab)
        assert st.avgdl == o.avgdl
        assert st.average_idf == o.average_idf          # sequential float64 sum, first-seen order
        assert st.idf.tobytes() == o.idf.tobytes()
        assert np.array_equal(st.post_tf.numpy(), o.post_tf)
    else:
        o = obm.Bm25sLucene(docs, c.vocab)
        assert np.array_equal(st.idf.astype(np.float32), o.idf32)
    assert np.array_equal(st.indptr.numpy(), o.indptr)
    assert np.array_equal(st.post_doc.numpy(), o.post_doc)
    assert np.array_equal(st.df.numpy(), o.df)


def test_bm25_stats_empty_corpus_raises_like_reference():
    with pytest.raises(ZeroDivisionError):
        stats_from_host_counts(torch.zeros(0, dtype=torch.int32), torch.zeros(1, dtype=torch.int64), 4)


def test_shard_bounds_cover_exactly():
    for n in (0, 1, 63, 64, 1000, 125000 * 8 + 3):
        for world in (1, 2, 4, 8):
            prev = 0
            for r in range(world):
                lo, hi = ezdist.shard_bounds(n, world, r, align=64)
                assert lo == prev and lo <= hi <= n
                prev = hi
            assert prev == n


def test_shard_bounds_are_balanced():
    for world in (2, 4, 8):
        sizes = [b - a for a, b in (ezdist.shard_bounds(1_000_000, world, r, align=64) for r in range(world))]
        assert max(sizes) - min(sizes) <= 64 and sum(sizes) == 1_000_000


def test_record_views_alias_the_record_bytes():
    """The typed views the kernels write through and the merge reads from (in place, part p at p * nbytes)."""
    q, k, world = 5, 3, 4
    for sparse_bytes, sdt in ((8, torch.float64), (4, torch.float32)):
        lay = ezdist.RecordLayout(q, k, sparse_bytes)
        offs, total = lay.offsets
        assert total == lay.nbytes and total % 16 == 0 and all(o % 16 == 0 for o in offs)
        g = torch.Generator().manual_seed(0)
        gathered = torch.zeros(world * lay.nbytes, dtype=torch.uint8)
        parts = []
        for r in range(world):
            rec = gathered[r * lay.nbytes:(r + 1) * lay.nbytes]
            ds, di, ss, si = ezdist.record_views(lay, rec)
            assert ds.dtype == torch.float32 and ss.dtype == sdt and di.dtype == si.dtype == torch.int32
            assert ds.shape == (q, k) and ds.data_ptr() == rec.data_ptr() + offs[0] and si.data_ptr() == rec.data_ptr() + offs[3]
            ds.copy_(torch.rand(q, k, generator=g))
            di.copy_(torch.randint(0, 100, (q, k), generator=g, dtype=torch.int32))
            ss.copy_(torch.rand(q, k, generator=g, dtype=sdt))
            si.copy_(torch.randint(0, 100, (q, k), generator=g, dtype=torch.int32))
            parts.append([t.clone() for t in (ds, di, ss, si)])
        v0 = ezdist.record_views(lay, gathered[:lay.nbytes])
        for r in range(world):
            for j in range(4):
                # what ezr_merge_topk_parts addresses: part r of array j = view 0 of array j + r * nbytes
                raw = gathered[r * lay.nbytes + offs[j]: r * lay.nbytes + offs[j] + lay.sizes[j]]
                assert torch.equal(raw.view(v0[j].dtype).view(q, k), parts[r][j])


def test_schema_surface():
    n = schema.TextNode("hello", metadata={"dir": "a"})
    s = schema.NodeWithScore(n, 1.5)
    assert s.get_content() == "hello" and s.metadata["dir"] == "a" and s.node is n
    s.score = 2.0                                        # rerankers overwrite it (rerankers.py:92)
    f = schema.build_qdrant_filters("emsplus")
    assert schema.filter_conditions(f) == {"dir": "emsplus"}
    assert schema.filter_conditions(None) is None


def test_get_node_content_views():
    from easyrag_b200.retrievers import get_node_content
    n = schema.TextNode("body", metadata={"file_path": "fp", "know_path": "kp"})
    assert get_node_content(n, 0) == "body"
    assert get_node_content(n, 1) == "###\nfp\n\nbody"
    assert get_node_content(n, 2) == "###\nkp\n\nbody"
    assert get_node_content(n, 4) == "fp" and get_node_content(n, 5) == "kp"
    assert get_node_content(schema.TextNode("b"), 5) == ""


def test_rerank_handoff_slices_like_the_reference():
    # rerankers.py:309-322: slices of embed_bs over the coarse list, pairs (query, get_node_content(node, embed_type))
    import torch
    from easyrag_b200 import handoff
    from easyrag_b200.schema import TextNode, NodeWithScore
    nodes = [NodeWithScore(node=TextNode(text=f"chunk {i}", metadata={"file_path": f"f{i}.txt"}), score=1.0 / (i + 1))
             for i in range(70)]
    got = list(handoff.rerank_batches(nodes, "问题", embed_type=1, batch_size=32))
    assert [(b, e) for b, e, _ in got] == [(0, 32), (32, 64), (64, 70)]
    for b, e, pairs in got:
        assert pairs == [("问题", "###\nf%d.txt\n\nchunk %d" % (i, i)) for i in range(b, e)]
    assert list(handoff.rerank_batches([], "q")) == []
    ids = torch.tensor([[5, 3, 9, -1], [7, -1, -1, -1], [-1, -1, -1, -1]], dtype=torch.int32)
    cnt = torch.tensor([3, 1, 0], dtype=torch.int32)
    assert list(handoff.candidate_batches(ids, cnt, batch_size=2)) == [(0, 0, [5, 3]), (0, 2, [9]), (1, 0, [7])]
    import pytest
    with pytest.raises(ValueError):
        list(handoff.rerank_batches(nodes, "q", batch_size=0))
    # the oracle restatement of get_inputs (rerankers.py:253-293) that the device packer is tested against
    from oracle import retrieve as ort
    items, ql, pl = ort.rerank_inputs([7, 8, 9], [[20, 21], list(range(100, 200))], sep=[13], prompt=[50, 51], bos=1,
                                      max_length=16)
    assert items[0] == [1, 7, 8, 9, 13, 20, 21, 13, 50, 51] and ql == [5, 5] and pl == [3, 3]
    assert items[1] == [1, 7, 8, 9, 13] + list(range(100, 111)) + [13, 50, 51]       # pair cut to 16, passage gives way


def test_integration_md_struct_matches_the_binding():
    """The ctypes sample a maintainer would copy from INTEGRATION.md lists exactly the fields of the real binding
    (a struct that is 8 bytes short makes the library read a garbage pointer: ADVICE round 1)."""
    text = (ROOT / "INTEGRATION.md").read_text()
    block = text[text.index("class ezr_bm25_index(C.Structure):"):]
    block = block[:block.index("lib.ezr_bm25_topk.restype")]
    doc_fields = re.findall(r'\("([a-z_0-9]+)",\s*C\.(c_[a-z0-9_]+)\)', block)
    real = _lib.Bm25IndexStruct._fields_
    assert [n for n, _ in doc_fields] == [n for n, _ in real]
    assert all(getattr(ctypes, t) is rt for (_, t), (_, rt) in zip(doc_fields, real))      # c_int32 is an alias of c_int


def test_ctypes_struct_layout_matches_the_c_header(tmp_path):
    # ezr_bm25_index crosses the boundary by pointer: the ctypes mirror must have the C compiler's layout.
    import ctypes
    import shutil
    import subprocess
    from easyrag_b200 import _lib
    if shutil.which("gcc") is None:
        import pytest
        pytest.skip("gcc not available")
    fields = [name for name, _ in _lib.Bm25IndexStruct._fields_]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "layout.c"
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{root}/include/easyrag_b200.h"', 'int main(void) {',
             '  printf("%zu\\n", sizeof(ezr_bm25_index));']
    lines += [f'  printf("%zu\\n", offsetof(ezr_bm25_index, {f}));' for f in fields]
    lines += ['  return 0;', '}']
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-o", str(exe), str(src)], check=True)
    out = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert out[0] == ctypes.sizeof(_lib.Bm25IndexStruct)
    assert out[1:] == [getattr(_lib.Bm25IndexStruct, f).offset for f in fields]
