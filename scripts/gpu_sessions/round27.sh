#!/bin/bash
# dense TS128 with unrolled k-chunk issue: tests, bench, valid pacing probes (insertion path off), overlap with a 3-stage ring
mkdir -p gpurun_out; rm -f gpurun_out/summary27.txt
timeout 900 python -m pytest tests/test_gpu_retrieval.py -m gpu -q -x -k "dense or hybrid" > gpurun_out/s27_tests.log 2>&1; echo "tests exit $? $(tail -n 1 gpurun_out/s27_tests.log)" >> gpurun_out/summary27.txt
run() { tag=$1; shift; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu "$@" > gpurun_out/bench_r2a_$tag.json 2> gpurun_out/bench_r2a_$tag.err; echo "bench $tag exit $?" >> gpurun_out/summary27.txt; }
run ts128
run p4_noInsert --dense-probe 4
run p5_noTMA_noInsert --dense-probe 5
run p6_fewMMA_noInsert --dense-probe 6
run ts64_p5 --dense-kernel 3 --dense-probe 5
run ov3 --overlap 1 --dense-stages 3
cat gpurun_out/summary27.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r2a_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['roofline']['kernels']
        print(f.split('r2a_')[1][:-5], round(d['value']), 'ms', round(d['ms_per_step'],2), {n:round(v['avg_ms'],2) for n,v in k.items()}, d['setup']['dense_kernel'])
    except Exception as e: print(f, 'ERR', e)
PY
