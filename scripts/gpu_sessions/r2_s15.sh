#!/bin/bash
# round 2, session 15: GEMM epilogue in 64-column blocks (TMA store per block, residual through the staging buffer):
# 6 ring stages + single staging buffers vs 5 stages + double buffers
mkdir -p gpurun_out
S=gpurun_out/r2s15_summary.txt; : > $S
V5=easyrag_b200/_lib/variant_5a165916/libeasyrag_b200.so     # -DEZR_GEMM_PLAIN_STAGES=5
timeout 900 python -m pytest tests/test_gpu_encoder.py -m gpu -q > gpurun_out/r2s15_enc_tests.log 2>&1; echo "enc tests exit $?" >> $S
EASYRAG_B200_LIB=$V5 timeout 900 python -m pytest tests/test_gpu_encoder.py -m gpu -q -k gemm > gpurun_out/r2s15_enc_tests_v5.log 2>&1; echo "enc tests v5 exit $?" >> $S
for rep in 1 2; do
timeout 600 python scripts/bench_gemm.py > gpurun_out/r2s15_gemm_s6_$rep.jsonl 2> gpurun_out/r2s15_gemm_s6_$rep.err; echo "gemm s6 $rep exit $?" >> $S
EASYRAG_B200_LIB=$V5 timeout 600 python scripts/bench_gemm.py > gpurun_out/r2s15_gemm_s5_$rep.jsonl 2> gpurun_out/r2s15_gemm_s5_$rep.err; echo "gemm s5 $rep exit $?" >> $S
done
timeout 600 python bench_encode.py --arch bert --chunks 40000 > gpurun_out/r2s15_enc_bert.json 2> gpurun_out/r2s15_enc_bert.err; echo "enc-bert exit $?" >> $S
EASYRAG_B200_LIB=$V5 timeout 600 python bench_encode.py --arch bert --chunks 40000 > gpurun_out/r2s15_enc_bert_v5.json 2> gpurun_out/r2s15_enc_bert_v5.err; echo "enc-bert v5 exit $?" >> $S
timeout 600 python bench_encode.py --arch qwen2 --chunks 40000 > gpurun_out/r2s15_enc_qwen2.json 2> gpurun_out/r2s15_enc_qwen2.err; echo "enc-qwen2 exit $?" >> $S
cat $S
tail -n 12 gpurun_out/r2s15_enc_tests.log
tail -n 5 gpurun_out/r2s15_enc_tests_v5.log
for t in s6_1 s5_1 s6_2 s5_2; do echo "== $t"; python - <<PY
import json
print(" ".join(f'{json.loads(l)["gemm"]}:{round(json.loads(l)["tflops"])}' for l in open("gpurun_out/r2s15_gemm_$t.jsonl")))
PY
tail -2 gpurun_out/r2s15_gemm_$t.err; done
python - <<'PY'
import json
for t in ("enc_bert", "enc_bert_v5", "enc_qwen2"):
    try:
        d = json.loads(open(f"gpurun_out/r2s15_{t}.json").read().strip().splitlines()[-1])
        print(t, "chunks/s", round(d["chunks_per_s"]), "gemm", round(d["gemm"]["tflops"]), "attn", round(d["attention"]["tflops"]), "ms", round(d["gemm"]["ms"]), round(d["attention"]["ms"]), round(d["other_ms"]), "parity", d["parity"]["ok"])
    except Exception as e:
        print(t, "ERR", e); print(open(f"gpurun_out/r2s15_{t}.err").read()[-2000:])
PY
