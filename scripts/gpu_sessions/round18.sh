#!/bin/bash
# final round-1 evidence run: all GPU tests, bench (ours + reference arm), launch list, --set full of the two top kernels,
# compute-sanitizer on the smoke path
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/s18_all_gpu_tests.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/summary18.txt
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1r_reference.json 2> gpurun_out/bench_r1r_reference.err; echo "ref exit $?" >> gpurun_out/summary18.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1r_n1.json 2> gpurun_out/bench_r1r_n1.err; echo "bench exit $?" >> gpurun_out/summary18.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bm25_|dense_|merge_|select_|fuse_" -c 60 --csv --log-file gpurun_out/launches_r1r.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_r_launch.log 2>&1; echo "ncu-list exit $?" >> gpurun_out/summary18.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"bm25_score_kernel|dense_ts_kernel" -s 6 -c 2 -o gpurun_out/prof_r1r python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_r.log 2>&1; echo "ncu-full exit $?" >> gpurun_out/summary18.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py --smoke > gpurun_out/sanitizer_memcheck_r1r.log 2>&1; echo "memcheck exit $?" >> gpurun_out/summary18.txt
timeout 900 python bench_encode.py --arch bert > gpurun_out/enc_bert_r.json 2> gpurun_out/enc_bert_r.err; echo "enc-bert exit $?" >> gpurun_out/summary18.txt
timeout 900 python bench_encode.py --arch qwen2 > gpurun_out/enc_qwen2_r.json 2> gpurun_out/enc_qwen2_r.err; echo "enc-qwen2 exit $?" >> gpurun_out/summary18.txt
cat gpurun_out/summary18.txt
tail -n 4 gpurun_out/s18_all_gpu_tests.log
tail -n 2 gpurun_out/sanitizer_memcheck_r1r.log
cat gpurun_out/bench_r1r_n1.json
