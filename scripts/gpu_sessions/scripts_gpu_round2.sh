#!/bin/bash
# GPU session 2: first bench numbers, ncu launch list, ncu full capture of the two dominant kernels, sanitizer
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_retrieval.py -m gpu -q -k "rrf or fusion" > gpurun_out/s2_fuse.log 2>&1; echo "fuse exit $?" >> gpurun_out/summary2.txt
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err; echo "bench exit $?" >> gpurun_out/summary2.txt
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r1_ref.json 2> gpurun_out/bench_r1_ref.err; echo "ref exit $?" >> gpurun_out/summary2.txt
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bm25_|dense_|merge_|select_|fuse_" -c 60 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1; echo "ncu-list exit $?" >> gpurun_out/summary2.txt
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"bm25_score_kernel|dense_tc_kernel" -s 6 -c 2 -o gpurun_out/prof_r1 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "ncu-full exit $?" >> gpurun_out/summary2.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py --smoke > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/summary2.txt
cat gpurun_out/summary2.txt
cat gpurun_out/bench_r1_n1.json
tail -n 3 gpurun_out/bench_r1_n1.err
cat gpurun_out/bench_r1_ref.json
