"""CPU, world_size 2, gloo: the N>1 plumbing (shard bounds, the in-place record layout, the single all-gather).

The merge itself is a CUDA kernel (tests/test_gpu_dist.py covers it); here the gathered records
are merged by the oracle and must reproduce the unsharded oracle result, which proves that the
exchange carries exactly what a canonical-order merge needs.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import bm25 as obm
from oracle import retrieve as ort
from easyrag_b200 import synth
from easyrag_b200 import dist as ezdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_topk(rows, k, base):
    ids = np.full((len(rows), k), -1, dtype=np.int32)
    sc = np.full((len(rows), k), -np.inf, dtype=rows[0].dtype)
    for i, s in enumerate(rows):
        ii, ss = ort.bm25_topk_ids(s, k)
        ids[i, :ii.size] = ii + base
        sc[i, :ii.size] = ss
    return ids, sc


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, vocab, nq, k, dim = 700, 200, 9, 5, 16
        c = synth.make_sparse_corpus(n, vocab, 21, mean_len=25, min_len=1, max_len=60)
        qs = synth.make_queries(c, nq, 22, min_terms=1, max_terms=6)
        model = obm.OkapiCSR(c.doc_lists(), vocab)                 # global statistics on every rank
        rows = [model.get_scores([int(t) for t in terms]) for terms in qs.term_lists()]
        dense = synth.make_dense_corpus(n, dim, 23).float().numpy()
        qv = synth.make_dense_queries(torch.from_numpy(dense), nq, 24).float().numpy()
        lo, hi = ezdist.shard_bounds(n, world, rank, align=64)
        s_ids, s_sc = _oracle_topk([r[lo:hi] for r in rows], k, lo)
        d_ids, d_sc = ort.dense_topk(dense[lo:hi], qv, k)
        d_ids = np.where(d_ids >= 0, d_ids + lo, -1).astype(np.int32)
        lay = ezdist.RecordLayout(nq, k, 8)
        # the product path (ShardedCoarseRanker): results are written through typed views of ONE byte record,
        # a single all_gather_into_tensor exchanges the records, the merge reads part p at p * nbytes
        record = torch.zeros(lay.nbytes, dtype=torch.uint8)
        for view, arr in zip(ezdist.record_views(lay, record), (d_sc, d_ids, s_sc, s_ids)):
            view.copy_(torch.from_numpy(arr))
        gathered = torch.zeros(world * lay.nbytes, dtype=torch.uint8)
        dist.all_gather_into_tensor(gathered, record)
        parts = [ezdist.record_views(lay, gathered[p * lay.nbytes:(p + 1) * lay.nbytes]) for p in range(world)]
        ds, di, ss, si = (torch.cat([parts[p][j] for p in range(world)], dim=1) for j in range(4))
        # canonical merge by the oracle: (score desc, id desc) over the gathered candidates
        ok = True
        full_s_ids, full_s_sc = _oracle_topk(rows, k, 0)
        full_d_ids, full_d_sc = ort.dense_topk(dense, qv, k)
        for q in range(nq):
            for cand_s, cand_i, ref_i, ref_s in ((ss[q].numpy(), si[q].numpy(), full_s_ids[q], full_s_sc[q]),
                                                 (ds[q].numpy(), di[q].numpy(), full_d_ids[q], full_d_sc[q])):
                valid = cand_i >= 0
                order = np.lexsort((-cand_i[valid], -cand_s[valid]))[:k]
                got_i = cand_i[valid][order]
                got_s = cand_s[valid][order]
                m = ref_i >= 0
                ok &= np.array_equal(got_i, ref_i[m]) and np.array_equal(got_s, ref_s[m])
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_two_rank_shard_exchange_reproduces_unsharded_result():
    world = 2
    port = _free_port()
    mgr = mp.get_context("spawn").Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
