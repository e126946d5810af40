#!/usr/bin/env python
"""Benchmark of the coarse-ranking hot path: queries/sec, dense + BM25 + RRF top-10 over 1M x 768 chunks.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload = BASELINE.json configs[2] ("dense+BM25 dual-route + RRF fusion top-10, 1M chunks, 10k queries",
the configuration the metric is quoted on); at N > 1 the SAME corpus is row-sharded (configs[3]) -> strong
scaling.  One step = one pass of the hot path over the whole batch of 10k synthetic queries.

  value : whole-job queries/s with query vectors + term ids resident in HBM when the timed region starts.
  e2e   : same metric through the public batched API from pinned HOST buffers, H2D of the queries and D2H of
          the fused (id, score) lists inside the timed region.
  roofline : dominant kernel, algorithmic bytes / CUDA-event duration on its launch stream.
  cpu_baseline : the oracle port of the reference's CPU retrievers (numpy BM25Okapi restatement + full argsort,
          fp32 BLAS cosine, Python RRF) on a bounded query sample over the same corpus, on this box's cores.

``--impl reference`` prints the CPU arm as its own line (rank 0 only under torchrun).
Nothing here reads /root/reference.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SEED = 20240922 + 3
METRIC = "queries/sec dense+BM25+RRF top-10 over 1M x 768 chunks"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--vocab", type=int, default=200_000)
    ap.add_argument("--queries", type=int, default=10_000)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--cpu-queries", type=int, default=256, help="bounded sample for the CPU baseline (~15 s)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--dense-kernel", type=int, default=0, help="0 auto, 1 simt, 2 tcgen05 SS, 3 tcgen05 TS")
    ap.add_argument("--overlap", type=int, default=0, help="1: dense and BM25 routes on two streams")
    ap.add_argument("--self-check", type=int, default=64,
                    help="after timing: first N queries through both BM25 kernel paths at full size, compared bit for bit")
    ap.add_argument("--bm25-skip", type=int, default=0, help="1: candidate pass skips non-essential terms (A/B)")
    ap.add_argument("--dense-probe", type=int, default=0, help="measurement probe of the dense kernel (results invalid)")
    ap.add_argument("--dense-stages", type=int, default=0, help="cap of the dense kernel's TMA ring (0 = all smem)")
    return ap.parse_args()


# ----------------------------------------------------------------------------------- data
def make_data(args, dev):
    """Synthetic config-3 data, generated on the GPU (3e8 tokens), identical on every rank (same seed)."""
    from easyrag_b200 import synth
    from easyrag_b200.index import Bm25Stats
    t0 = time.time()
    corpus = synth.make_sparse_corpus(args.rows, args.vocab, SEED, device=dev)
    queries = synth.make_queries(corpus, args.queries, SEED + 1)
    stats = Bm25Stats.from_tokens(corpus.tokens, corpus.doc_ptr, args.vocab, bm25_type=0)
    n_tokens = int(corpus.tokens.numel())
    del corpus
    vec = synth.make_dense_corpus(args.rows, args.dim, SEED + 2, device=dev)
    qvec = synth.make_dense_queries(vec, args.queries, SEED + 3)
    torch.cuda.synchronize()
    return dict(stats=stats, queries=queries, vec=vec, qvec=qvec, n_tokens=n_tokens, gen_s=time.time() - t0)


# --------------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.lines = []          # (arrival time, csv line)
        self.windows = []        # [t_begin, t_end] of the timed regions
        self.proc = None

    def begin(self):
        self.windows.append([time.perf_counter(), None])

    def end(self):
        self.windows[-1][1] = time.perf_counter()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        # nvidia-smi is started before the warm-up (its start-up takes longer than a short timed region); only
        # samples that arrived while a timed region was running count (a line arrives a few ms after its sample)
        inside = [ln for t, ln in self.lines if any(a <= t <= (b or t) + 0.03 for a, b in self.windows)]
        for ln in inside:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# -------------------------------------------------------------------------- CPU reference
class CpuReference:
    """Oracle port of the reference's CPU retrievers over the full corpus (kind = "port").

    BM25: rank_bm25.BM25Okapi.get_scores restated on CSR postings with numpy (oracle/bm25.py evaluation
    order; the literal reference loops over all N documents in Python per query term and is far slower),
    then BM25Retriever.filter's full ``argsort()[::-1]`` (retrievers.py:192).  Dense: exact fp32 cosine with
    the host BLAS on all cores + argpartition (qdrant local mode).  Fusion: Python RRF (retrievers.py:256-274).
    """

    def __init__(self, data, args):
        from oracle import retrieve as ort
        self.ort = ort
        st = data["stats"]
        self.indptr = st.indptr.cpu().numpy()
        self.post_doc = st.post_doc.cpu().numpy()
        self.post_tf = st.post_tf.cpu().numpy().astype(np.float64)
        self.idf = st.idf
        dl = st.doc_len.cpu().numpy()
        self.K_d = 1.5 * ((1 - 0.75) + (0.75 * dl) / st.avgdl)
        self.n = st.n_docs
        self.vec = data["vec"].float().cpu()
        self.qvec = data["qvec"].float().cpu()
        self.term_lists = data["queries"].term_lists()
        self.k = args.k
        self.cores = os.cpu_count()
        torch.set_num_threads(self.cores)

    def run(self, lo: int, hi: int):
        k = self.k
        sims = (self.qvec[lo:hi] @ self.vec.T).numpy()
        out = []
        for j, qi in enumerate(range(lo, hi)):
            score = np.zeros(self.n)
            for t in self.term_lists[qi]:
                if t < 0 or self.idf[t] == 0.0:
                    continue
                s, e = self.indptr[t], self.indptr[t + 1]
                tf = self.post_tf[s:e]
                d = self.post_doc[s:e]
                score[d] += self.idf[t] * (tf * 2.5 / (tf + self.K_d[d]))
            order = score.argsort()[::-1]                                   # retrievers.py:192
            sparse = [int(i) for i in order[:k] if score[i] > 0]
            part = np.argpartition(-sims[j], k)[:k]
            dense = [int(i) for i in part[np.argsort(-sims[j][part], kind="stable")]]
            out.append(self.ort.rrf_ids([sparse, dense], None, K=60, topk=k))
        return out

    def measure(self, n_queries: int, steps: int = 1, warmup: int = 0):
        n_queries = min(n_queries, len(self.term_lists))
        for _ in range(warmup):
            self.run(0, n_queries)
        t0 = time.perf_counter()
        for _ in range(steps):
            self.run(0, n_queries)
        dt = time.perf_counter() - t0
        return steps * n_queries / dt, dt / steps


# ------------------------------------------------------------------------------ our arm
def algorithmic_bytes(args, data, index, n_rows_local, n_slices_q):
    """Per-step algorithmic bytes of the two dominant kernels (DESIGN.md section 'Measurement')."""
    q = data["queries"]
    terms = q.terms.to(index.device).long()
    valid = (terms >= 0) & (terms < index.vocab)
    df = (index.indptr[1:] - index.indptr[:-1])
    postings = int(df[terms[valid]].sum())
    bm25 = postings * (4 + index.post_w.element_size()) + args.queries * args.k * 12
    # two-phase path: the candidate pass reads one packed 4-byte word per posting (+ the candidate ids it emits)
    bm25_cand = postings * 4 + args.queries * args.k * 4
    passes = -(-args.queries // 128)
    dense = passes * n_rows_local * args.dim * 2 + args.queries * args.dim * 2 + args.queries * args.k * 8
    return {"bm25_score": bm25, "bm25_cand": bm25_cand, "dense_tc": dense, "postings_per_step": postings, "dense_passes": passes}


def run_ours(args):
    import torch.distributed as dist
    from easyrag_b200 import _lib, batched
    from easyrag_b200 import dist as ezdist
    from easyrag_b200.index import Bm25Index, DenseIndex

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    _lib.require_cuda()
    L = _lib.lib()
    _lib.check(L.ezr_dense_set_kernel(args.dense_kernel))
    _lib.check(L.ezr_dense_set_stage_cap(args.dense_stages))
    _lib.check(L.ezr_dense_set_probe(args.dense_probe))
    _lib.check(L.ezr_bm25_set_skipping(args.bm25_skip))

    data = make_data(args, dev)
    lo, hi = ezdist.shard_bounds(args.rows, world, rank, align=8192)
    t0 = time.time()
    sparse = Bm25Index(data["stats"], device=dev, doc_lo=lo, doc_hi=hi)
    dense = DenseIndex(data["vec"][lo:hi], device=dev, row_lo=lo)
    torch.cuda.synchronize()
    build_s = time.time() - t0
    ranker = batched.CoarseRanker(dense, sparse, canon=None, overlap=bool(args.overlap))
    sharded = ezdist.ShardedCoarseRanker(ranker) if world > 1 else None
    k = args.k
    q = data["queries"]
    d_qvec = data["qvec"].contiguous()
    d_ptr, d_terms = q.term_ptr.to(dev), q.terms.to(dev)
    h_qvec = d_qvec.cpu().pin_memory()
    h_ptr, h_terms = q.term_ptr.cpu().pin_memory(), q.terms.cpu().pin_memory()
    h_ids = torch.empty(args.queries, k, dtype=torch.int32).pin_memory()
    h_sc = torch.empty(args.queries, k, dtype=torch.float64).pin_memory()
    e_qvec, e_ptr, e_terms = torch.empty_like(d_qvec), torch.empty_like(d_ptr), torch.empty_like(d_terms)

    def step_device():
        if sharded is not None:
            return sharded.hybrid(d_qvec, d_ptr, d_terms, k=k, k_out=k)[0]
        return ranker.hybrid(d_qvec, d_ptr, d_terms, k, k, k)[0]

    def step_e2e():
        e_qvec.copy_(h_qvec, non_blocking=True)
        e_ptr.copy_(h_ptr, non_blocking=True)
        e_terms.copy_(h_terms, non_blocking=True)
        if sharded is not None:
            f = sharded.hybrid(e_qvec, e_ptr, e_terms, k=k, k_out=k)[0]
        else:
            f = ranker.hybrid(e_qvec, e_ptr, e_terms, k, k, k)[0]
        h_ids.copy_(f.ids, non_blocking=True)
        h_sc.copy_(f.scores, non_blocking=True)

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        torch.cuda.synchronize()
        ms = torch.tensor([a.elapsed_time(b)], device=dev)
        if world > 1:
            dist.barrier()
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        step_device()
    torch.cuda.synchronize()
    _lib.check(L.ezr_profile_reset())
    _lib.check(L.ezr_profile_enable(1))
    launches0 = L.ezr_launch_count()
    sampler.begin()
    ms = timed(step_device, args.steps)
    sampler.end()
    launches_timed = L.ezr_launch_count() - launches0
    _lib.check(L.ezr_profile_enable(0))
    prof = {name: _lib.profile_read(name) for name in ("bm25_cand", "bm25_rescore", "bm25_score", "dense_tc",
                                                       "dense_simt", "merge", "fuse")}
    for _ in range(2):
        step_e2e()
    sampler.begin()
    ms_e2e = timed(step_e2e, args.steps)
    sampler.end()
    clocks = sampler.stop() if rank == 0 else None

    # Full-size property check outside the timed regions: the two-phase path (integer candidates + exact
    # rescoring) and the ordered float64 kernel are independent implementations; on this shard they must return
    # the same ids, scores and counts bit for bit.
    self_check = None
    if args.self_check > 0 and sparse.post_pk is not None and args.dense_probe == 0:
        nq = min(args.self_check, args.queries)
        qp = data["queries"].term_ptr[:nq + 1].to(dev)
        qt = data["queries"].terms.to(dev)
        a = batched.bm25_topk(sparse, qp, qt, k)
        b = batched.bm25_topk(sparse.ordered_view(), qp, qt, k)
        torch.cuda.synchronize()
        same = bool(torch.equal(a.ids, b.ids) and torch.equal(a.counts, b.counts)
                    and torch.equal(a.scores.view(torch.int64), b.scores.view(torch.int64)))
        self_check = {"bm25_two_phase_equals_ordered": same, "queries": nq, "postings_local": sparse.n_postings}
        if not same:
            raise SystemExit(f"bench.py self-check FAILED on rank {rank}: BM25 kernel paths disagree")
        # dense route: the tcgen05 kernel against the generic fp32 SIMT kernel + row selection over the full shard;
        # different accumulation orders, so the bar is the north star's cosine tolerance on the sorted score lists
        if k <= 16:
            r_tc = batched.dense_topk(dense, d_qvec[:nq], k)
            tc_name = L.ezr_dense_last_kernel().decode()
            _lib.check(L.ezr_dense_set_kernel(1))
            try:
                r_simt = batched.dense_topk(dense, d_qvec[:nq], k)
            finally:
                _lib.check(L.ezr_dense_set_kernel(args.dense_kernel))
            torch.cuda.synchronize()
            diff = float((r_tc.scores - r_simt.scores).abs().max())
            agree = float((r_tc.ids == r_simt.ids).float().mean())
            self_check.update({"dense_kernel": tc_name, "dense_vs_simt_max_abs_score_diff": diff,
                               "dense_vs_simt_id_agreement": agree, "dense_tol": 1e-3})
            if not diff <= 1e-3:
                raise SystemExit(f"bench.py self-check FAILED on rank {rank}: dense kernels differ by {diff}")
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = {}
    pk_file = ROOT / "MEASURED_PEAKS.json"
    if pk_file.exists():
        peaks = json.loads(pk_file.read_text())
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    alg = algorithmic_bytes(args, data, sparse, dense.n_rows, None)
    tf_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    tf_src = ("measured (MEASURED_PEAKS.json bf16_tflops_sustained: kernel timed inside a long step)"
              if "bf16_tflops_sustained" in peaks else "fallback 1400 TFLOP/s")
    dense_flops = 2.0 * dense.n_rows * args.dim * args.queries          # per launch: every query x every local row
    kernels = {}
    two_phase = prof["bm25_cand"][1] > 0
    for name in ("bm25_cand", "bm25_score", "dense_tc"):
        tot, n = prof[name]
        if n and not (name == "bm25_score" and two_phase):     # two-phase: bm25_score only sees overflowed queries
            avg_ms = tot / n
            kernels[name] = {"launches": n, "avg_ms": avg_ms, "alg_bytes_per_launch": alg[name],
                             "GBps": alg[name] / (avg_ms * 1e-3) / 1e9}
    others = {name: {"launches": prof[name][1], "avg_ms": prof[name][0] / prof[name][1]}
              for name in ("bm25_rescore", "bm25_score", "merge", "fuse") if prof[name][1] and name not in kernels}
    if "dense_tc" in kernels:
        # The persistent kernel shares each corpus pass between all resident query blocks through L2, so HBM is not
        # its bound (ncu: DRAM traffic ~ a few corpus passes per launch); it is a [Q x D] . [D x N] contraction on the
        # tensor pipe.  GBps above is kept as "bytes if every 128-query block streamed the corpus from HBM".
        kernels["dense_tc"]["flops_per_launch"] = dense_flops
        kernels["dense_tc"]["TFLOPs"] = dense_flops / (kernels["dense_tc"]["avg_ms"] * 1e-3) / 1e12
    dom = max(kernels, key=lambda n_: kernels[n_]["avg_ms"]) if kernels else None
    traffic = None
    tfile = ROOT / "profiles" / "traffic.json"
    if tfile.exists() and dom:
        traffic = json.loads(tfile.read_text()).get(dom)
    roofline = None
    if dom == "dense_tc":
        roofline = {"bound": "tensor", "kernel": dom, "achieved": kernels[dom]["TFLOPs"], "peak": tf_peak,
                    "unit": "TFLOP/s", "frac": kernels[dom]["TFLOPs"] / tf_peak, "traffic": traffic,
                    "peak_source": tf_src, "kernels": kernels}
    elif dom:
        roofline = {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["GBps"], "peak": hbm_peak, "unit": "GB/s",
                    "frac": kernels[dom]["GBps"] / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                    "kernels": kernels,
                    "note": ("algorithmic bytes = postings touched (4 B packed word each) + candidate ids"
                             if dom == "bm25_cand" else "algorithmic bytes = postings touched (12 B each) + outputs")
                            + "; ncu shows most of them are served from L2 (range-major grid), DRAM traffic per "
                              "launch is in `traffic`"}
    if roofline:
        roofline["other_kernels"] = others
    value = args.steps * args.queries / (ms * 1e-3)
    e2e_v = args.steps * args.queries / (ms_e2e * 1e-3)
    h2d = h_qvec.numel() * 2 + h_ptr.numel() * 4 + h_terms.numel() * 4
    d2h = h_ids.numel() * 4 + h_sc.numel() * 8
    cpu = None
    if world == 1 and not args.no_cpu:
        ref = CpuReference(data, args)
        v, dt = ref.measure(args.cpu_queries)
        cpu = {"value": v, "unit": "queries/s", "cores": ref.cores, "kind": "port",
               "sample": f"first {min(args.cpu_queries, args.queries)} of the {args.queries} queries over the full "
                         f"{args.rows} x {args.dim} corpus, {dt:.1f} s; numpy BM25Okapi restatement + full argsort, "
                         f"fp32 BLAS cosine, Python RRF"}
    line = {
        "metric": METRIC, "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "bf16 dense / f64 bm25", "data": "synthetic",
        "config": {"workload": "configs[2]: dense+BM25 dual-route + RRF top-10, 1M x 768 chunks, 10k queries/step"
                               + (f", row-sharded over {world} GPUs (configs[3])" if world > 1 else ""),
                   "rows": args.rows, "dim": args.dim, "vocab": args.vocab, "queries_per_step": args.queries,
                   "k": k, "rrf_K": 60, "tokens": data["n_tokens"], "postings_local": sparse.n_postings,
                   "queries_per_corpus_pass": 128, "timed_region_starts_from": "query vectors + term ids",
                   "l2": "inputs larger than L2 (corpus shard and postings >> 126 MB), no explicit flush",
                   "parallelism": f"rows{world}"},
        "e2e": {"value": e2e_v, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches_timed),
        "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "setup": {"generate_s": round(data["gen_s"], 1), "index_build_s": round(build_s, 1),
                  "index_bytes": sparse.index_bytes(), "dense_kernel": L.ezr_dense_last_kernel().decode(),
                  "self_check": self_check},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))     # data generation only
    data = make_data(args, dev)
    ref = CpuReference(data, args)
    n = min(args.cpu_queries, args.queries)
    v, dt = ref.measure(n, steps=args.steps, warmup=min(args.warmup, 1))
    sample = (f"each step = first {n} of the {args.queries} queries over the full {args.rows} x {args.dim} corpus; "
              f"numpy BM25Okapi restatement + full argsort, fp32 BLAS cosine ({ref.cores} threads), Python RRF")
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32 dense / f64 bm25", "data": "synthetic",
        "config": {"workload": "configs[2]: dense+BM25 dual-route + RRF top-10, 1M x 768 chunks (bounded query sample)",
                   "rows": args.rows, "dim": args.dim, "vocab": args.vocab, "queries_per_step": n, "k": args.k},
        "cpu_baseline": {"value": v, "unit": "queries/s", "cores": ref.cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
