#!/bin/bash
# evidence run after the two-phase BM25 + TS128 dense changes: all GPU tests, both bench arms, launch list, --set full, memcheck
mkdir -p gpurun_out; rm -f gpurun_out/summary28.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/s28_all_gpu_tests.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/summary28.txt
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2b_reference.json 2> gpurun_out/bench_r2b_reference.err; echo "ref exit $?" >> gpurun_out/summary28.txt
timeout 900 python bench.py > gpurun_out/bench_r2b_n1.json 2> gpurun_out/bench_r2b_n1.err; echo "bench exit $?" >> gpurun_out/summary28.txt
timeout 600 python bench.py --overlap 1 --dense-stages 3 --no-cpu > gpurun_out/bench_r2b_n1_overlap.json 2> gpurun_out/bench_r2b_n1_overlap.err; echo "bench overlap exit $?" >> gpurun_out/summary28.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bm25_|dense_|merge_|select_|fuse_" -c 120 --csv --log-file gpurun_out/launches_r2b.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_b_launch.log 2>&1; echo "ncu-list exit $?" >> gpurun_out/summary28.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"bm25_cand_kernel|bm25_rescore_kernel|bm25_bound_kernel|dense_ts_kernel" -s 39 -c 13 -o gpurun_out/prof_r2b python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_b.log 2>&1; echo "ncu-full exit $?" >> gpurun_out/summary28.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py --smoke > gpurun_out/sanitizer_memcheck_r2b.log 2>&1; echo "memcheck exit $?" >> gpurun_out/summary28.txt
cat gpurun_out/summary28.txt
tail -n 3 gpurun_out/s28_all_gpu_tests.log
tail -n 2 gpurun_out/sanitizer_memcheck_r2b.log
cat gpurun_out/bench_r2b_n1.json
python - <<'PY'
import json
for f in ('gpurun_out/bench_r2b_n1_overlap.json','gpurun_out/bench_r2b_reference.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['value'],1), d.get('ms_per_step'), d.get('clocks'))
PY
