#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_encoder.py -m gpu -q > gpurun_out/s12_enc.log 2>&1; echo "enc exit $?" >> gpurun_out/summary12.txt
timeout 900 python -m pytest tests/test_gpu_retrieval.py tests/test_gpu_dropin.py -m gpu -q > gpurun_out/s12_retr.log 2>&1; echo "retr exit $?" >> gpurun_out/summary12.txt
timeout 900 python bench_encode.py --arch bert > gpurun_out/enc_bert_k.json 2> gpurun_out/enc_bert_k.err; echo "enc-bert exit $?" >> gpurun_out/summary12.txt
timeout 900 python bench_encode.py --arch qwen2 > gpurun_out/enc_qwen2_k.json 2> gpurun_out/enc_qwen2_k.err; echo "enc-qwen2 exit $?" >> gpurun_out/summary12.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1k.json 2> gpurun_out/bench_r1k.err; echo "bench exit $?" >> gpurun_out/summary12.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bm25_|dense_|merge_|select_|fuse_" -c 60 --csv --log-file gpurun_out/launches_r1k.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_k_launch.log 2>&1; echo "ncu-list exit $?" >> gpurun_out/summary12.txt
cat gpurun_out/summary12.txt
tail -n 4 gpurun_out/s12_*.log
cat gpurun_out/enc_bert_k.json gpurun_out/enc_qwen2_k.json gpurun_out/bench_r1k.json
tail -n 3 gpurun_out/bench_r1k.err
