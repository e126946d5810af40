#!/bin/bash
# round 2, session 10: where the K=768 GEMM shapes lose time -- epilogue probes (variant builds) and one ncu capture
mkdir -p gpurun_out
S=gpurun_out/r2s10_summary.txt; : > $S
V1=easyrag_b200/_lib/variant_a818f068/libeasyrag_b200.so     # -DEZR_GEMM_PROBE=1: epilogue math, no stores
V2=easyrag_b200/_lib/variant_25fd07f3/libeasyrag_b200.so     # -DEZR_GEMM_PROBE=2: no epilogue
timeout 600 python scripts/bench_gemm.py > gpurun_out/r2s10_gemm_base.jsonl 2> gpurun_out/r2s10_gemm_base.err; echo "base exit $?" >> $S
EASYRAG_B200_LIB=$V1 timeout 600 python scripts/bench_gemm.py > gpurun_out/r2s10_gemm_probe1.jsonl 2> gpurun_out/r2s10_gemm_probe1.err; echo "probe1 exit $?" >> $S
EASYRAG_B200_LIB=$V2 timeout 600 python scripts/bench_gemm.py > gpurun_out/r2s10_gemm_probe2.jsonl 2> gpurun_out/r2s10_gemm_probe2.err; echo "probe2 exit $?" >> $S
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm -c 28 \
   -o gpurun_out/r2s10_gemm python scripts/bench_gemm.py --iters 1 > gpurun_out/r2s10_ncu.log 2>&1; echo "ncu exit $?" >> $S
cat $S
for t in base probe1 probe2; do echo "== $t"; python - <<PY
import json
for l in open("gpurun_out/r2s10_gemm_$t.jsonl"):
    d = json.loads(l); print(d["gemm"], d["N"], d["K"], round(d["ms"], 4), round(d["tflops"]))
PY
tail -3 gpurun_out/r2s10_gemm_$t.err; done
tail -5 gpurun_out/r2s10_ncu.log
