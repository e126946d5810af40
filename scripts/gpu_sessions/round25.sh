#!/bin/bash
# BM25 branch-free apply check + dense kernel probes (which resource paces the dense kernel?)
mkdir -p gpurun_out; rm -f gpurun_out/summary25.txt
timeout 900 python -m pytest tests/test_gpu_retrieval.py -m gpu -q -x -k "bm25 or hybrid" > gpurun_out/s25_tests.log 2>&1; echo "tests exit $? $(tail -n 1 gpurun_out/s25_tests.log)" >> gpurun_out/summary25.txt
run() { tag=$1; shift; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu "$@" > gpurun_out/bench_r1y_$tag.json 2> gpurun_out/bench_r1y_$tag.err; echo "bench $tag exit $?" >> gpurun_out/summary25.txt; }
run base
run probe1_no_tma --dense-probe 1
run probe2_few_mma --dense-probe 2
run probe3_no_insert --dense-probe 3
cat gpurun_out/summary25.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r1y_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['roofline']['kernels']; o=d['roofline'].get('other_kernels',{})
        print(f.split('r1y_')[1][:-5], round(d['value']), 'ms', round(d['ms_per_step'],2), {n:round(v['avg_ms'],2) for n,v in k.items()}, {n:round(v['avg_ms'],3) for n,v in o.items()}, d['clocks'])
    except Exception as e: print(f, 'ERR', e)
PY
