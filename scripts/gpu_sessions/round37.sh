#!/bin/bash
# last session: one warp vote per work item (default) vs per posting slot (variant -DEZR_BM25_PK_VOTE_EACH=1); parity + A/B
mkdir -p gpurun_out; rm -f gpurun_out/summary37.txt
timeout 600 python -m pytest tests/test_gpu_retrieval.py -m gpu -q -x -k "bm25 or hybrid" > gpurun_out/s37_tests.log 2>&1; echo "tests exit $? $(tail -n 1 gpurun_out/s37_tests.log)" >> gpurun_out/summary37.txt
V=$(ls -d easyrag_b200/_lib/variant_*/ | head -n 1)
for rep in 1 2; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --self-check 256 > gpurun_out/bench_r2k_item_$rep.json 2> gpurun_out/bench_r2k_item_$rep.err; echo "item $rep exit $?" >> gpurun_out/summary37.txt
  EASYRAG_B200_LIB=$PWD/${V}libeasyrag_b200.so timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --self-check 256 > gpurun_out/bench_r2k_each_$rep.json 2> gpurun_out/bench_r2k_each_$rep.err; echo "each $rep exit $?" >> gpurun_out/summary37.txt
done
cat gpurun_out/summary37.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r2k_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); k=d['roofline']['kernels']
        print(f.split('r2k_')[1][:-5], round(d['value']), 'ms', round(d['ms_per_step'],2), {n:round(v['avg_ms'],3) for n,v in k.items()}, d['setup']['self_check']['bm25_two_phase_equals_ordered'], d['clocks']['sm_mhz'])
    except Exception as e: print(f,'ERR',e)
PY
