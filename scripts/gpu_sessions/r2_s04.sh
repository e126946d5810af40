#!/bin/bash
# round 2, session 4: attention v2 (persistent CTAs, single-pass softmax): tests, ragged + L=512 encode bench, ncu
mkdir -p gpurun_out
S=gpurun_out/r2s04_summary.txt; : > $S
timeout 600 python -m pytest tests/test_gpu_encoder.py -m gpu -q -x > gpurun_out/r2s04_enc_tests.log 2>&1; echo "enc tests exit $?" >> $S
timeout 900 python -m pytest tests/test_gpu_retrieval.py tests/test_gpu_dropin.py -m gpu -q -x > gpurun_out/r2s04_tests.log 2>&1; echo "tests exit $?" >> $S
timeout 600 python bench_encode.py --arch bert --chunks 40000 > gpurun_out/r2s04_enc_bert.json 2> gpurun_out/r2s04_enc_bert.err; echo "enc-bert exit $?" >> $S
timeout 600 python bench_encode.py --arch bert --chunks 20000 --len-min 512 --len-max 512 > gpurun_out/r2s04_enc_bert_L512.json 2> gpurun_out/r2s04_enc_bert_L512.err; echo "enc-L512 exit $?" >> $S
timeout 600 python bench_encode.py --arch qwen2 --chunks 40000 > gpurun_out/r2s04_enc_qwen2.json 2> gpurun_out/r2s04_enc_qwen2.err; echo "enc-qwen2 exit $?" >> $S
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attn_tc_kernel" -s 30 -c 2 -o gpurun_out/r2s04_prof_attn python bench_encode.py --arch bert --chunks 2048 --enc-queries 128 > gpurun_out/r2s04_ncu_attn.log 2>&1; echo "ncu-attn exit $?" >> $S
cat $S
tail -n 12 gpurun_out/r2s04_enc_tests.log
tail -n 5 gpurun_out/r2s04_tests.log
python - <<'PY'
import json
for t in ("enc_bert", "enc_bert_L512", "enc_qwen2"):
    try:
        d = json.loads(open(f"gpurun_out/r2s04_{t}.json").read().strip().splitlines()[-1])
        print(t, "chunks/s", round(d["chunks_per_s"]), "gemm", round(d["gemm"]["tflops"]), "attn", round(d["attention"]["tflops"]), d["attention"]["kernel"], "attn ms", round(d["attention"]["ms"]), "parity", d["parity"])
    except Exception as e:
        print(t, "ERR", e); print(open(f"gpurun_out/r2s04_{t}.err").read()[-2000:])
PY
