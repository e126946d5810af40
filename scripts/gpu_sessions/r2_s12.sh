#!/bin/bash
# round 2, session 12: GEMM with per-epilogue warp counts + residual L2 prefetch; BM25 candidate-pass CTA size A/B
mkdir -p gpurun_out
S=gpurun_out/r2s12_summary.txt; : > $S
T128=easyrag_b200/_lib/variant_aa063dbb/libeasyrag_b200.so     # -DEZR_BM25_PK_THREADS=128
T512=easyrag_b200/_lib/variant_ec167540/libeasyrag_b200.so     # -DEZR_BM25_PK_THREADS=512 -DEZR_BM25_PK_MINB=3
timeout 900 python -m pytest tests/test_gpu_encoder.py -m gpu -q -x > gpurun_out/r2s12_enc_tests.log 2>&1; echo "enc tests exit $?" >> $S
timeout 600 python scripts/bench_gemm.py > gpurun_out/r2s12_gemm.jsonl 2> gpurun_out/r2s12_gemm.err; echo "gemm exit $?" >> $S
timeout 600 python bench_encode.py --arch bert --chunks 40000 > gpurun_out/r2s12_enc_bert.json 2> gpurun_out/r2s12_enc_bert.err; echo "enc-bert exit $?" >> $S
timeout 600 python bench_encode.py --arch qwen2 --chunks 40000 > gpurun_out/r2s12_enc_qwen2.json 2> gpurun_out/r2s12_enc_qwen2.err; echo "enc-qwen2 exit $?" >> $S
for rep in 1 2; do for v in base t128 t512; do
case $v in base) LIBV=easyrag_b200/_lib/libeasyrag_b200.so;; t128) LIBV=$T128;; t512) LIBV=$T512;; esac
EASYRAG_B200_LIB=$LIBV timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --enc-chunks 0 --parity-queries 0 --self-check 64 > gpurun_out/r2s12_bench_${v}_$rep.json 2> gpurun_out/r2s12_bench_${v}_$rep.err; echo "bench $v rep$rep exit $?" >> $S
done; done
cat $S
tail -n 8 gpurun_out/r2s12_enc_tests.log
python - <<'PY'
import json
for l in open("gpurun_out/r2s12_gemm.jsonl"):
    d = json.loads(l); print(d["gemm"], d["N"], d["K"], round(d["ms"], 4), round(d["tflops"]))
for t in ("enc_bert", "enc_qwen2"):
    try:
        d = json.loads(open(f"gpurun_out/r2s12_{t}.json").read().strip().splitlines()[-1])
        print(t, "chunks/s", round(d["chunks_per_s"]), "gemm", round(d["gemm"]["tflops"]), "attn", round(d["attention"]["tflops"]), "other_ms", round(d["other_ms"]), "parity", d["parity"]["ok"])
    except Exception as e:
        print(t, "ERR", e); print(open(f"gpurun_out/r2s12_{t}.err").read()[-2000:])
for rep in (1, 2):
  for v in ("base", "t128", "t512"):
    f = f"gpurun_out/r2s12_bench_{v}_{rep}.json"
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(v, rep, round(d["value"]), "ms", round(d["ms_per_step"], 3),
              {k: (round(x["avg_ms"], 3), round(x.get("avg_ms_in_timed_region", 0), 3)) for k, x in r["kernels"].items()},
              d["digest"].get("matches_committed_n1"), d["setup"]["self_check"]["bm25_two_phase_equals_ordered"], d["clocks"]["sm_mhz"])
    except Exception as e:
        print(v, rep, "ERR", e)
        print(open(f.replace(".json", ".err")).read()[-1500:])
PY
tail -3 gpurun_out/r2s12_gemm.err
