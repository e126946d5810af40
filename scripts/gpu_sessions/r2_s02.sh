#!/bin/bash
# round 2, session 2: tcgen05 attention kernel -- parity tests, then the encoder bench (was 145 TFLOP/s with mma.sync)
mkdir -p gpurun_out
S=gpurun_out/r2s02_summary.txt; : > $S
timeout 600 python -m pytest tests/test_gpu_encoder.py -m gpu -q -x -k "attention" > gpurun_out/r2s02_attn_tests.log 2>&1; echo "attn tests exit $?" >> $S
timeout 900 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_retrieval.py tests/test_gpu_dropin.py -m gpu -q > gpurun_out/r2s02_tests.log 2>&1; echo "tests exit $?" >> $S
timeout 600 python bench_encode.py --arch bert > gpurun_out/r2s02_enc_bert.json 2> gpurun_out/r2s02_enc_bert.err; echo "enc-bert exit $?" >> $S
timeout 600 python bench_encode.py --arch qwen2 > gpurun_out/r2s02_enc_qwen2.json 2> gpurun_out/r2s02_enc_qwen2.err; echo "enc-qwen2 exit $?" >> $S
cat $S
tail -n 30 gpurun_out/r2s02_attn_tests.log
tail -n 6 gpurun_out/r2s02_tests.log
cat gpurun_out/r2s02_enc_bert.json gpurun_out/r2s02_enc_qwen2.json
tail -n 5 gpurun_out/r2s02_enc_bert.err
