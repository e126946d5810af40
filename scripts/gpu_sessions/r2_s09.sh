#!/bin/bash
# round 2, session 9: cta_group::2 GEMM (tests, per-shape microbench, encode), BM25 plan table A/B (interleaved, 2 reps)
mkdir -p gpurun_out
S=gpurun_out/r2s09_summary.txt; : > $S
timeout 900 python -m pytest tests/test_gpu_encoder.py -m gpu -q -x > gpurun_out/r2s09_enc_tests.log 2>&1; echo "enc tests exit $?" >> $S
timeout 600 python scripts/bench_gemm.py > gpurun_out/r2s09_gemm_shapes.jsonl 2> gpurun_out/r2s09_gemm_shapes.err; echo "gemm shapes exit $?" >> $S
timeout 600 python bench_encode.py --arch bert --chunks 40000 > gpurun_out/r2s09_enc_bert.json 2> gpurun_out/r2s09_enc_bert.err; echo "enc-bert exit $?" >> $S
timeout 600 python bench_encode.py --arch qwen2 --chunks 40000 > gpurun_out/r2s09_enc_qwen2.json 2> gpurun_out/r2s09_enc_qwen2.err; echo "enc-qwen2 exit $?" >> $S
timeout 1200 python -m pytest tests/test_gpu_retrieval.py -m gpu -q -x > gpurun_out/r2s09_tests.log 2>&1; echo "tests exit $?" >> $S
for rep in 1 2; do for plan in 1 0; do
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --enc-chunks 0 --parity-queries 0 --self-check 0 --bm25-plan $plan > gpurun_out/r2s09_bench_plan${plan}_$rep.json 2> gpurun_out/r2s09_bench_plan${plan}_$rep.err; echo "bench plan$plan rep$rep exit $?" >> $S
done; done
cat $S
tail -n 12 gpurun_out/r2s09_enc_tests.log
tail -n 5 gpurun_out/r2s09_tests.log
cat gpurun_out/r2s09_gemm_shapes.jsonl; tail -5 gpurun_out/r2s09_gemm_shapes.err
python - <<'PY'
import json
for t in ("enc_bert", "enc_qwen2"):
    try:
        d = json.loads(open(f"gpurun_out/r2s09_{t}.json").read().strip().splitlines()[-1])
        print(t, "chunks/s", round(d["chunks_per_s"]), "gemm", round(d["gemm"]["tflops"]), "attn", round(d["attention"]["tflops"]), "parity", d["parity"])
    except Exception as e:
        print(t, "ERR", e); print(open(f"gpurun_out/r2s09_{t}.err").read()[-2000:])
for rep in (1, 2):
  for plan in (1, 0):
    f = f"gpurun_out/r2s09_bench_plan{plan}_{rep}.json"
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d["roofline"]
        print("plan", plan, rep, round(d["value"]), "ms", round(d["ms_per_step"], 3),
              {k: (round(v["avg_ms"], 3), round(v.get("avg_ms_in_timed_region", 0), 3)) for k, v in r["kernels"].items()}, d["digest"].get("matches_committed_n1"), d["clocks"]["sm_mhz"])
    except Exception as e:
        print("plan", plan, rep, "ERR", e)
        print(open(f.replace(".json", ".err")).read()[-2000:])
PY
