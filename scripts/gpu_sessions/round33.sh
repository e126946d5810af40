#!/bin/bash
# after the int-overflow fix in the rescoring binary search: configs[4] shape on one GPU (1.14e9 postings) with the
# full-size self-check (two-phase vs ordered kernel, 256 queries), then the default bench with its self-check
mkdir -p gpurun_out; rm -f gpurun_out/summary33.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --rows 4000000 --dim 1024 --queries 4096 --self-check 256 > gpurun_out/bench_r2g_c5.json 2> gpurun_out/bench_r2g_c5.err; echo "bench-c5shape exit $?" >> gpurun_out/summary33.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --self-check 256 > gpurun_out/bench_r2g_n1.json 2> gpurun_out/bench_r2g_n1.err; echo "bench default exit $?" >> gpurun_out/summary33.txt
cat gpurun_out/summary33.txt
python - <<'PY'
import json
for f in ('gpurun_out/bench_r2g_c5.json','gpurun_out/bench_r2g_n1.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), round(d['ms_per_step'],2), d['setup'], {n:round(v['avg_ms'],2) for n,v in d['roofline']['kernels'].items()}, d['roofline'].get('other_kernels'), d['roofline']['kernels'].get('dense_tc',{}).get('TFLOPs'), d['clocks'])
    except Exception as e: print(f,'ERR',e)
PY
tail -n 3 gpurun_out/bench_r2g_c5.err
