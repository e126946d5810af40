#!/bin/bash
# confirmation of the committed state: all GPU tests (incl. the 4200-token hand-over case), default bench with the
# real launch counter and the 256-query CPU sample, reference arm, smoke
mkdir -p gpurun_out; rm -f gpurun_out/summary31.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/s31_all_gpu_tests.log 2>&1; echo "gpu tests exit $? $(tail -n 1 gpurun_out/s31_all_gpu_tests.log)" >> gpurun_out/summary31.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/s31_smoke.log 2>&1; echo "smoke exit $? $(tail -n 1 gpurun_out/s31_smoke.log)" >> gpurun_out/summary31.txt
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2e_reference.json 2> gpurun_out/bench_r2e_reference.err; echo "ref exit $?" >> gpurun_out/summary31.txt
timeout 900 python bench.py > gpurun_out/bench_r2e_n1.json 2> gpurun_out/bench_r2e_n1.err; echo "bench exit $?" >> gpurun_out/summary31.txt
cat gpurun_out/summary31.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2e_n1.json').read().strip().splitlines()[-1])
print(round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'], d['steps'], d['clocks'])
print(d['roofline']['kernel'], d['roofline']['frac'], {n:round(v['avg_ms'],2) for n,v in d['roofline']['kernels'].items()})
print(d['cpu_baseline'])
r=json.loads(open('gpurun_out/bench_r2e_reference.json').read().strip().splitlines()[-1]); print('ref', r['value'], r['cpu_baseline']['sample'][:80])
PY
