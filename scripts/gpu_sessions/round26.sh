#!/bin/bash
# dense TS kernel with 128-row corpus tiles: parity tests, A/B against 64-row tiles, probes
mkdir -p gpurun_out; rm -f gpurun_out/summary26.txt
timeout 900 python -m pytest tests/test_gpu_retrieval.py -m gpu -q -x -k "dense or hybrid" > gpurun_out/s26_tests.log 2>&1; echo "tests exit $? $(tail -n 1 gpurun_out/s26_tests.log)" >> gpurun_out/summary26.txt
run() { tag=$1; shift; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu "$@" > gpurun_out/bench_r1z_$tag.json 2> gpurun_out/bench_r1z_$tag.err; echo "bench $tag exit $?" >> gpurun_out/summary26.txt; }
run ts128
run ts64 --dense-kernel 3
run ts128_noTMA --dense-probe 1
run ts128_fewMMA --dense-probe 2
run ts128_noTMA_fewMMA --dense-probe 3
cat gpurun_out/summary26.txt
tail -n 30 gpurun_out/s26_tests.log | head -60
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r1z_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['roofline']['kernels']; o=d['roofline'].get('other_kernels',{})
        print(f.split('r1z_')[1][:-5], round(d['value']), 'ms', round(d['ms_per_step'],2), {n:round(v['avg_ms'],2) for n,v in k.items()}, d['setup']['dense_kernel'])
    except Exception as e: print(f, 'ERR', e)
PY
