#!/bin/bash
# two-phase BM25 (packed postings -> integer candidates -> exact rescoring): parity tests, then A/B bench + ncu of the new kernels
mkdir -p gpurun_out; rm -f gpurun_out/summary20.txt
timeout 900 python -m pytest tests/test_gpu_retrieval.py tests/test_gpu_dropin.py -m gpu -q -x > gpurun_out/s20_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/summary20.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_r1t_packed.json 2> gpurun_out/bench_r1t_packed.err; echo "bench packed exit $?" >> gpurun_out/summary20.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"bm25_cand_kernel|bm25_rescore_kernel|bm25_bound_kernel" -s 42 -c 14 -o gpurun_out/prof_r1t python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_t.log 2>&1; echo "ncu-full exit $?" >> gpurun_out/summary20.txt
cat gpurun_out/summary20.txt
tail -n 15 gpurun_out/s20_tests.log
cat gpurun_out/bench_r1t_packed.json
tail -n 3 gpurun_out/bench_r1t_packed.err
