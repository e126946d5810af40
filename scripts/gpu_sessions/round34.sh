#!/bin/bash
# racecheck + memcheck of the hot path (smoke), all GPU tests, default bench with the BM25 and dense full-size self-checks
mkdir -p gpurun_out; rm -f gpurun_out/summary34.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python __graft_entry__.py --smoke > gpurun_out/sanitizer_racecheck_r2h.log 2>&1; echo "racecheck exit $? $(tail -n 1 gpurun_out/sanitizer_racecheck_r2h.log)" >> gpurun_out/summary34.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py --smoke > gpurun_out/sanitizer_memcheck_r2h.log 2>&1; echo "memcheck exit $? $(tail -n 1 gpurun_out/sanitizer_memcheck_r2h.log)" >> gpurun_out/summary34.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/s34_all_gpu_tests.log 2>&1; echo "gpu tests exit $? $(tail -n 1 gpurun_out/s34_all_gpu_tests.log)" >> gpurun_out/summary34.txt
timeout 900 python bench.py > gpurun_out/bench_r2h_n1.json 2> gpurun_out/bench_r2h_n1.err; echo "bench exit $?" >> gpurun_out/summary34.txt
cat gpurun_out/summary34.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2h_n1.json').read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['gpu_launches'], d['clocks'])
print(d['setup']['self_check'])
print(d['roofline']['kernel'], round(d['roofline']['frac'],3), {n:round(v['avg_ms'],2) for n,v in d['roofline']['kernels'].items()})
PY
tail -n 3 gpurun_out/bench_r2h_n1.err
