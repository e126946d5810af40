#!/bin/bash
# round 2, session 11: TMA-store GEMM epilogue (8 and 16 epilogue warps), 4-way rescore, submitted (pipelined) steps
mkdir -p gpurun_out
find gpurun_out -name '*.ncu-rep' -size +20M -delete 2>/dev/null
S=gpurun_out/r2s11_summary.txt; : > $S
V16=easyrag_b200/_lib/variant_7928f0ea/libeasyrag_b200.so     # -DEZR_GEMM_EPI_WARPS=16 -DEZR_GEMM_STAGES=5
timeout 900 python -m pytest tests/test_gpu_encoder.py -m gpu -q -x > gpurun_out/r2s11_enc_tests.log 2>&1; echo "enc tests exit $?" >> $S
EASYRAG_B200_LIB=$V16 timeout 900 python -m pytest tests/test_gpu_encoder.py -m gpu -q -x -k gemm > gpurun_out/r2s11_enc_tests_v16.log 2>&1; echo "enc tests v16 exit $?" >> $S
timeout 600 python scripts/bench_gemm.py > gpurun_out/r2s11_gemm_w8.jsonl 2> gpurun_out/r2s11_gemm_w8.err; echo "gemm w8 exit $?" >> $S
EASYRAG_B200_LIB=$V16 timeout 600 python scripts/bench_gemm.py > gpurun_out/r2s11_gemm_w16.jsonl 2> gpurun_out/r2s11_gemm_w16.err; echo "gemm w16 exit $?" >> $S
timeout 600 python bench_encode.py --arch bert --chunks 40000 > gpurun_out/r2s11_enc_bert.json 2> gpurun_out/r2s11_enc_bert.err; echo "enc-bert exit $?" >> $S
timeout 1500 python -m pytest tests/test_gpu_retrieval.py -m gpu -q -x > gpurun_out/r2s11_tests.log 2>&1; echo "tests exit $?" >> $S
for rep in 1 2; do for pl in 1 0; do
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --enc-chunks 0 --parity-queries 0 --self-check 0 --pipeline $pl > gpurun_out/r2s11_bench_pl${pl}_$rep.json 2> gpurun_out/r2s11_bench_pl${pl}_$rep.err; echo "bench pipeline$pl rep$rep exit $?" >> $S
done; done
cat $S
tail -n 15 gpurun_out/r2s11_enc_tests.log
tail -n 5 gpurun_out/r2s11_enc_tests_v16.log
tail -n 15 gpurun_out/r2s11_tests.log
for t in w8 w16; do echo "== $t"; python - <<PY
import json
for l in open("gpurun_out/r2s11_gemm_$t.jsonl"):
    d = json.loads(l); print(d["gemm"], d["N"], d["K"], round(d["ms"], 4), round(d["tflops"]))
PY
tail -3 gpurun_out/r2s11_gemm_$t.err; done
python - <<'PY'
import json
for t in ("enc_bert",):
    try:
        d = json.loads(open(f"gpurun_out/r2s11_{t}.json").read().strip().splitlines()[-1])
        print(t, "chunks/s", round(d["chunks_per_s"]), "gemm", round(d["gemm"]["tflops"]), "attn", round(d["attention"]["tflops"]), "parity", d["parity"])
    except Exception as e:
        print(t, "ERR", e); print(open(f"gpurun_out/r2s11_{t}.err").read()[-2000:])
for rep in (1, 2):
  for pl in (1, 0):
    f = f"gpurun_out/r2s11_bench_pl{pl}_{rep}.json"
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d["roofline"]
        print("pipeline", pl, rep, round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3),
              {k: (round(v["avg_ms"], 3), round(v.get("avg_ms_in_timed_region", 0), 3)) for k, v in r["kernels"].items()},
              {k: round(v["avg_ms"], 4) for k, v in r["other_kernels"].items()}, d["digest"].get("matches_committed_n1"), d["clocks"]["sm_mhz"])
    except Exception as e:
        print("pipeline", pl, rep, "ERR", e)
        print(open(f.replace(".json", ".err")).read()[-2000:])
PY
