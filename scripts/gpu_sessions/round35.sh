#!/bin/bash
# MaxScore-style skipping of non-essential terms in the BM25 candidate pass: parity tests, then A/B bench (skip on / off)
mkdir -p gpurun_out; rm -f gpurun_out/summary35.txt
timeout 900 python -m pytest tests/test_gpu_retrieval.py tests/test_gpu_dropin.py -m gpu -q -x -k "bm25 or hybrid or sparse or retriever" > gpurun_out/s35_tests.log 2>&1; echo "tests exit $? $(tail -n 1 gpurun_out/s35_tests.log)" >> gpurun_out/summary35.txt
run() { tag=$1; shift; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --self-check 256 "$@" > gpurun_out/bench_r2i_$tag.json 2> gpurun_out/bench_r2i_$tag.err; echo "bench $tag exit $?" >> gpurun_out/summary35.txt; }
run skip1
run skip0 --bm25-skip 0
run skip1_b
cat gpurun_out/summary35.txt
tail -n 5 gpurun_out/s35_tests.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r2i_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); k=d['roofline']['kernels']; o=d['roofline']['other_kernels']
        print(f.split('r2i_')[1][:-5], round(d['value']), 'ms', round(d['ms_per_step'],2), {n:round(v['avg_ms'],2) for n,v in k.items()}, 'rescore', round(o['bm25_rescore']['avg_ms'],3), 'ordered', round(o['bm25_score']['avg_ms'],3), d['setup']['self_check']['bm25_two_phase_equals_ordered'], d['clocks']['sm_mhz'])
    except Exception as e: print(f,'ERR',e)
PY
tail -n 3 gpurun_out/bench_r2i_skip1.err
