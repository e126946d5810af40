#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_encoder.py -m gpu -q > gpurun_out/s11_enc.log 2>&1; echo "enc exit $?" >> gpurun_out/summary11.txt
timeout 900 python bench_encode.py --arch bert > gpurun_out/enc_bert_j.json 2> gpurun_out/enc_bert_j.err; echo "enc-bert exit $?" >> gpurun_out/summary11.txt
timeout 900 python bench_encode.py --arch qwen2 > gpurun_out/enc_qwen2_j.json 2> gpurun_out/enc_qwen2_j.err; echo "enc-qwen2 exit $?" >> gpurun_out/summary11.txt
timeout 900 python __graft_entry__.py --smoke > gpurun_out/s11_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/summary11.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel" -s 40 -c 8 -o gpurun_out/prof_r1j_enc python bench_encode.py --arch bert --chunks 1024 --queries 128 > gpurun_out/ncu_j2.log 2>&1; echo "ncu-enc exit $?" >> gpurun_out/summary11.txt
cat gpurun_out/summary11.txt
tail -n 5 gpurun_out/s11_*.log
cat gpurun_out/enc_bert_j.json gpurun_out/enc_qwen2_j.json
tail -n 3 gpurun_out/enc_bert_j.err
