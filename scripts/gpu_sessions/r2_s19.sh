#!/bin/bash
# round 2, session 19: attention v4 (resolved plan one item ahead, two Q buffers + 4-deep K/V rings at hd 64, the next
# item's first QK^T issued under the last tile's softmax)
mkdir -p gpurun_out
S=gpurun_out/r2s19_summary.txt; : > $S
timeout 900 python -m pytest tests/test_gpu_encoder.py -m gpu -q > gpurun_out/r2s19_enc_tests.log 2>&1; echo "enc tests exit $?" >> $S
timeout 600 python bench_encode.py --arch bert --chunks 40000 > gpurun_out/r2s19_enc_bert.json 2> gpurun_out/r2s19_enc_bert.err; echo "enc-bert exit $?" >> $S
timeout 600 python bench_encode.py --arch bert --chunks 20000 --len-min 512 --len-max 512 > gpurun_out/r2s19_enc_bert_L512.json 2> gpurun_out/r2s19_enc_bert_L512.err; echo "enc-bert-L512 exit $?" >> $S
timeout 600 python bench_encode.py --arch qwen2 --chunks 40000 > gpurun_out/r2s19_enc_qwen2.json 2> gpurun_out/r2s19_enc_qwen2.err; echo "enc-qwen2 exit $?" >> $S
cat $S
tail -n 8 gpurun_out/r2s19_enc_tests.log
python - <<'PY'
import json
for t in ("enc_bert", "enc_bert_L512", "enc_qwen2"):
    try:
        d = json.loads(open(f"gpurun_out/r2s19_{t}.json").read().strip().splitlines()[-1])
        print(t, "chunks/s", round(d["chunks_per_s"]), "gemm", round(d["gemm"]["tflops"]), "attn", round(d["attention"]["tflops"]), "ms", round(d["gemm"]["ms"]), round(d["attention"]["ms"]), round(d["other_ms"]), "parity", d["parity"]["ok"])
    except Exception as e:
        print(t, "ERR", e); print(open(f"gpurun_out/r2s19_{t}.err").read()[-2000:])
PY
