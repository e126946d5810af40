#!/bin/bash
# round 2, session 23: BM25 document ranges of 16384 (18-bit packed weights) instead of 8192: fewer (query, range) units
mkdir -p gpurun_out
S=gpurun_out/r2s23_summary.txt; : > $S
VA=easyrag_b200/_lib/variant_a010162b/libeasyrag_b200.so     # -DEZR_BM25_RANGE=16384 (256 threads)
VB=$(ls -d easyrag_b200/_lib/variant_*/ | grep -v a010162b | head -1)libeasyrag_b200.so     # ... + 512 threads, 3 CTAs per SM
B="--steps 10 --warmup 3 --no-cpu --enc-chunks 0 --parity-queries 64 --self-check 64"
timeout 300 python bench.py $B > gpurun_out/r2s23_base.json 2> gpurun_out/r2s23_base.err; echo "base exit $?" >> $S
EASYRAG_B200_LIB=$VA timeout 300 python bench.py $B > gpurun_out/r2s23_r16k.json 2> gpurun_out/r2s23_r16k.err; echo "r16k exit $?" >> $S
EASYRAG_B200_LIB=$VB timeout 300 python bench.py $B > gpurun_out/r2s23_r16k_t512.json 2> gpurun_out/r2s23_r16k_t512.err; echo "r16k_t512 exit $?" >> $S
EASYRAG_B200_LIB=$VA timeout 300 python -m pytest tests/test_gpu_retrieval.py -m gpu -q -k "bm25 or hybrid or index" > gpurun_out/r2s23_tests_r16k.log 2>&1; echo "tests r16k exit $?" >> $S
cat $S
tail -n 6 gpurun_out/r2s23_tests_r16k.log
python - <<'PY'
import json
for v in ("base", "r16k", "r16k_t512"):
    f = f"gpurun_out/r2s23_{v}.json"
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(v, round(d["value"]), "ms", round(d["ms_per_step"], 3),
              {k: (round(x["avg_ms"], 3), round(x.get("avg_ms_in_timed_region", 0), 3)) for k, x in r["kernels"].items()},
              {k: round(x["avg_ms"], 4) for k, x in r["other_kernels"].items()},
              "parity", (d.get("parity_full_size") or {}).get("ok"), d["digest"].get("matches_committed_n1"), d["setup"]["self_check"]["bm25_two_phase_equals_ordered"], d["clocks"]["sm_mhz"])
    except Exception as e:
        print(v, "ERR", e); print(open(f.replace(".json", ".err")).read()[-1500:])
PY
