#!/bin/bash
# round 2, session 7: cluster-pair GEMM (multicast W), fast GELU, BM25 candidate-pass rework, dense cluster-pair kernel A/B
mkdir -p gpurun_out
S=gpurun_out/r2s07_summary.txt; : > $S
timeout 900 python -m pytest tests/test_gpu_encoder.py -m gpu -q -x > gpurun_out/r2s07_enc_tests.log 2>&1; echo "enc tests exit $?" >> $S
timeout 1200 python -m pytest tests/test_gpu_retrieval.py tests/test_gpu_dropin.py -m gpu -q -x > gpurun_out/r2s07_tests.log 2>&1; echo "tests exit $?" >> $S
timeout 600 python bench_encode.py --arch bert --chunks 40000 > gpurun_out/r2s07_enc_bert.json 2> gpurun_out/r2s07_enc_bert.err; echo "enc-bert exit $?" >> $S
timeout 600 python bench_encode.py --arch qwen2 --chunks 40000 > gpurun_out/r2s07_enc_qwen2.json 2> gpurun_out/r2s07_enc_qwen2.err; echo "enc-qwen2 exit $?" >> $S
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --enc-chunks 0 --parity-queries 64 > gpurun_out/r2s07_bench_k4.json 2> gpurun_out/r2s07_bench_k4.err; echo "bench k4 exit $?" >> $S
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --enc-chunks 0 --parity-queries 64 --dense-kernel 5 > gpurun_out/r2s07_bench_k5.json 2> gpurun_out/r2s07_bench_k5.err; echo "bench k5 exit $?" >> $S
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --enc-chunks 0 --parity-queries 0 --dense-kernel 5 --overlap 0 > gpurun_out/r2s07_bench_k5_seq.json 2> gpurun_out/r2s07_bench_k5_seq.err; echo "bench k5 seq exit $?" >> $S
cat $S
tail -n 12 gpurun_out/r2s07_enc_tests.log
tail -n 12 gpurun_out/r2s07_tests.log
python - <<'PY'
import json
for t in ("enc_bert", "enc_qwen2"):
    try:
        d = json.loads(open(f"gpurun_out/r2s07_{t}.json").read().strip().splitlines()[-1])
        print(t, "chunks/s", round(d["chunks_per_s"]), "gemm", round(d["gemm"]["tflops"]), "attn", round(d["attention"]["tflops"]), "parity", d["parity"])
    except Exception as e:
        print(t, "ERR", e); print(open(f"gpurun_out/r2s07_{t}.err").read()[-2000:])
for tag in ("k4", "k5", "k5_seq"):
    f = f"gpurun_out/r2s07_bench_{tag}.json"
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(tag, round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), r["bound"], r["kernel"], round(r["achieved"]), round(r["frac"], 3),
              {k: (round(v["avg_ms"], 3), round(v.get("avg_ms_in_timed_region", 0), 3)) for k, v in r["kernels"].items()},
              {k: round(v["avg_ms"], 3) for k, v in r["other_kernels"].items()}, d["setup"]["dense_kernel"], d["clocks"])
        p = d.get("parity_full_size") or {}
        print("   parity ok", p.get("ok"), "digest", d["digest"].get("matches_committed_n1"))
    except Exception as e:
        print(tag, "ERR", e)
        print(open(f.replace(".json", ".err")).read()[-2500:])
PY
