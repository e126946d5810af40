#!/bin/bash
# round 2, session 1: state check of the new bench (parity_full_size, overlap default, HostPipeline, small-batch regime)
mkdir -p gpurun_out
S=gpurun_out/r2s01_summary.txt; : > $S
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2s01_tests.log 2>&1; echo "tests exit $?" >> $S
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2s01_bench_n1.json 2> gpurun_out/r2s01_bench_n1.err; echo "bench exit $?" >> $S
timeout 900 python bench.py --steps 20 --warmup 5 --overlap 0 --no-cpu --parity-queries 0 > gpurun_out/r2s01_bench_n1_seq.json 2> gpurun_out/r2s01_bench_n1_seq.err; echo "bench-seq exit $?" >> $S
timeout 900 python bench.py --steps 30 --warmup 5 --queries 64 --no-cpu --parity-queries 64 > gpurun_out/r2s01_bench_q64.json 2> gpurun_out/r2s01_bench_q64.err; echo "bench-q64 exit $?" >> $S
timeout 900 python bench.py --steps 30 --warmup 5 --queries 1 --no-cpu --parity-queries 1 --self-check 1 > gpurun_out/r2s01_bench_q1.json 2> gpurun_out/r2s01_bench_q1.err; echo "bench-q1 exit $?" >> $S
cat $S
tail -n 4 gpurun_out/r2s01_tests.log
python - <<'PY'
import json
for tag in ("n1", "n1_seq", "q64", "q1"):
    f = f"gpurun_out/r2s01_bench_{tag}.json"
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(tag, round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), r["bound"], r["kernel"], round(r["achieved"]), round(r["frac"], 3),
              {k: (round(v["avg_ms"], 3), round(v.get("avg_ms_in_timed_region", 0), 3)) for k, v in r["kernels"].items()},
              {k: round(v["avg_ms"], 3) for k, v in r["other_kernels"].items()})
        print("   parity", d.get("parity_full_size"))
        print("   digest", d.get("digest"), "lat", d.get("latency_ms"), "cpu", d.get("cpu_baseline") and round(d["cpu_baseline"]["value"], 1))
        print("   terms", d.get("scaling_terms"))
    except Exception as e:
        print(tag, "ERR", e)
        print(open(f.replace(".json", ".err")).read()[-3000:])
PY
