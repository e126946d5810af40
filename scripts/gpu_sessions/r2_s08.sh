#!/bin/bash
# round 2, session 8: BM25 exact skipping A/B, GEMM per-shape microbench, attention v3 at L=512, ncu of dense / bm25 / gemm
mkdir -p gpurun_out
S=gpurun_out/r2s08_summary.txt; : > $S
timeout 1200 python -m pytest tests/test_gpu_retrieval.py tests/test_gpu_dropin.py -m gpu -q -x > gpurun_out/r2s08_tests.log 2>&1; echo "tests exit $?" >> $S
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --enc-chunks 0 --parity-queries 64 --bm25-skip 1 > gpurun_out/r2s08_bench_skip1.json 2> gpurun_out/r2s08_bench_skip1.err; echo "bench skip1 exit $?" >> $S
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --enc-chunks 0 --parity-queries 64 --bm25-skip 0 > gpurun_out/r2s08_bench_skip0.json 2> gpurun_out/r2s08_bench_skip0.err; echo "bench skip0 exit $?" >> $S
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --enc-chunks 0 --parity-queries 64 --bm25-skip 1 --dense-kernel 5 > gpurun_out/r2s08_bench_skip1_k5.json 2> gpurun_out/r2s08_bench_skip1_k5.err; echo "bench skip1 k5 exit $?" >> $S
timeout 600 python scripts/bench_gemm.py > gpurun_out/r2s08_gemm_shapes.jsonl 2> gpurun_out/r2s08_gemm_shapes.err; echo "gemm shapes exit $?" >> $S
timeout 600 python bench_encode.py --arch bert --chunks 20000 --len-min 512 --len-max 512 > gpurun_out/r2s08_enc_bert_L512.json 2> gpurun_out/r2s08_enc_bert_L512.err; echo "enc-L512 exit $?" >> $S
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|attn_tc_kernel" -s 40 -c 6 -o gpurun_out/r2s08_prof_enc python bench_encode.py --arch bert --chunks 2048 --enc-queries 128 > gpurun_out/r2s08_ncu_enc.log 2>&1; echo "ncu-enc exit $?" >> $S
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"dense_ts_kernel|bm25_cand_kernel|bm25_rescore_kernel" -s 40 -c 8 -o gpurun_out/r2s08_prof_retr python bench.py --steps 1 --warmup 3 --cal-steps 1 --no-cpu --enc-chunks 0 --parity-queries 0 --self-check 0 --overlap 0 --bm25-skip 1 > gpurun_out/r2s08_ncu_retr.log 2>&1; echo "ncu-retr exit $?" >> $S
cat $S
tail -n 8 gpurun_out/r2s08_tests.log
cat gpurun_out/r2s08_gemm_shapes.jsonl
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/r2s08_enc_bert_L512.json").read().strip().splitlines()[-1])
    print("L512 chunks/s", round(d["chunks_per_s"]), "gemm", round(d["gemm"]["tflops"]), "attn", round(d["attention"]["tflops"]), d["parity"])
except Exception as e:
    print("L512 ERR", e); print(open("gpurun_out/r2s08_enc_bert_L512.err").read()[-2000:])
for tag in ("skip0", "skip1", "skip1_k5"):
    f = f"gpurun_out/r2s08_bench_{tag}.json"
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(tag, round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), r["bound"], r["kernel"], round(r["achieved"]), round(r["frac"], 3),
              {k: (round(v["avg_ms"], 3), round(v.get("avg_ms_in_timed_region", 0), 3)) for k, v in r["kernels"].items()},
              {k: round(v["avg_ms"], 3) for k, v in r["other_kernels"].items()}, d["setup"]["dense_kernel"])
        p = d.get("parity_full_size") or {}
        print("   parity ok", p.get("ok"), "digest", d["digest"].get("matches_committed_n1"), "self", d["setup"]["self_check"])
    except Exception as e:
        print(tag, "ERR", e)
        print(open(f.replace(".json", ".err")).read()[-2500:])
PY
