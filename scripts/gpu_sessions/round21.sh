#!/bin/bash
# BM25 candidate kernel with 128-posting pieces; dense/BM25 overlap experiment with a capped dense TMA ring
mkdir -p gpurun_out; rm -f gpurun_out/summary21.txt
timeout 900 python -m pytest tests/test_gpu_retrieval.py -m gpu -q -x -k "bm25 or hybrid" > gpurun_out/s21_tests.log 2>&1; echo "tests exit $?" >> gpurun_out/summary21.txt
run() { tag=$1; shift; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu "$@" > gpurun_out/bench_r1u_$tag.json 2> gpurun_out/bench_r1u_$tag.err; echo "bench $tag exit $?" >> gpurun_out/summary21.txt; }
run base
run st5 --dense-stages 5
run st6 --dense-stages 6
run ov7 --overlap 1
run ov6 --overlap 1 --dense-stages 6
run ov5 --overlap 1 --dense-stages 5
run ov4 --overlap 1 --dense-stages 4
cat gpurun_out/summary21.txt
tail -n 3 gpurun_out/s21_tests.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r1u_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['roofline']['kernels']; o=d['roofline'].get('other_kernels',{})
        print(f.split('r1u_')[1][:-5], round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), {n:round(v['avg_ms'],2) for n,v in k.items()}, {n:round(v['avg_ms'],3) for n,v in o.items()})
    except Exception as e: print(f, 'ERR', e)
PY
