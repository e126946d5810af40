#!/bin/bash
# dense/BM25 overlap with max shared-memory carveout on both kernels
mkdir -p gpurun_out; rm -f gpurun_out/summary22.txt
run() { tag=$1; shift; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu "$@" > gpurun_out/bench_r1v_$tag.json 2> gpurun_out/bench_r1v_$tag.err; echo "bench $tag exit $?" >> gpurun_out/summary22.txt; }
run base
run ov6 --overlap 1 --dense-stages 6
run ov5 --overlap 1 --dense-stages 5
run ov4 --overlap 1 --dense-stages 4
run ov3 --overlap 1 --dense-stages 3
cat gpurun_out/summary22.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r1v_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['roofline']['kernels']; o=d['roofline'].get('other_kernels',{})
        print(f.split('r1v_')[1][:-5], round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), {n:round(v['avg_ms'],2) for n,v in k.items()}, {n:round(v['avg_ms'],3) for n,v in o.items()})
    except Exception as e: print(f, 'ERR', e)
PY
