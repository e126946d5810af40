#!/bin/bash
# final state of the round: all GPU tests, smoke, default bench
mkdir -p gpurun_out; rm -f gpurun_out/summary36.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/s36_all_gpu_tests.log 2>&1; echo "gpu tests exit $? $(tail -n 1 gpurun_out/s36_all_gpu_tests.log)" >> gpurun_out/summary36.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/s36_smoke.log 2>&1; echo "smoke exit $? $(tail -n 1 gpurun_out/s36_smoke.log)" >> gpurun_out/summary36.txt
timeout 900 python bench.py > gpurun_out/bench_r2j_n1.json 2> gpurun_out/bench_r2j_n1.err; echo "bench exit $?" >> gpurun_out/summary36.txt
cat gpurun_out/summary36.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2j_n1.json').read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['gpu_launches'], d['clocks'])
print(d['setup']['self_check'])
print(d['roofline']['kernel'], round(d['roofline']['frac'],3), {n:round(v['avg_ms'],2) for n,v in d['roofline']['kernels'].items()}, {n:round(v['avg_ms'],3) for n,v in d['roofline']['other_kernels'].items()})
PY
