#!/bin/bash
# round 2, session 20: attention A/B on one box -- two Q buffers + 4-deep rings + cross-item QK^T (default) vs one Q buffer + 2-deep rings
mkdir -p gpurun_out
S=gpurun_out/r2s20_summary.txt; : > $S
V=$(ls -d easyrag_b200/_lib/variant_*/ | head -1)libeasyrag_b200.so
for rep in 1 2; do
timeout 300 python bench_encode.py --arch bert --chunks 20000 --len-min 512 --len-max 512 > gpurun_out/r2s20_L512_deep_$rep.json 2> gpurun_out/r2s20_L512_deep_$rep.err; echo "deep L512 $rep exit $?" >> $S
EASYRAG_B200_LIB=$V timeout 300 python bench_encode.py --arch bert --chunks 20000 --len-min 512 --len-max 512 > gpurun_out/r2s20_L512_flat_$rep.json 2> gpurun_out/r2s20_L512_flat_$rep.err; echo "flat L512 $rep exit $?" >> $S
timeout 300 python bench_encode.py --arch bert --chunks 20000 > gpurun_out/r2s20_rag_deep_$rep.json 2> gpurun_out/r2s20_rag_deep_$rep.err; echo "deep ragged $rep exit $?" >> $S
EASYRAG_B200_LIB=$V timeout 300 python bench_encode.py --arch bert --chunks 20000 > gpurun_out/r2s20_rag_flat_$rep.json 2> gpurun_out/r2s20_rag_flat_$rep.err; echo "flat ragged $rep exit $?" >> $S
done
cat $S
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2s20_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("r2s20_")[1], "chunks/s", round(d["chunks_per_s"]), "gemm", round(d["gemm"]["tflops"]), "attn", round(d["attention"]["tflops"]), "attn ms", round(d["attention"]["ms"], 1))
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json", ".err")).read()[-1000:])
PY
