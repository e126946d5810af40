#!/bin/bash
# two-phase BM25 launch-shape sweep (variants listed in round23_variants.txt)
mkdir -p gpurun_out; rm -f gpurun_out/summary23.txt
for lib in easyrag_b200/_lib/libeasyrag_b200.so easyrag_b200/_lib/variant_*/libeasyrag_b200.so; do
  tag=$(basename $(dirname $lib))
  export EASYRAG_B200_LIB=$PWD/$lib
  timeout 600 python -m pytest tests/test_gpu_retrieval.py -m gpu -q -x -k "bm25" > gpurun_out/s23_$tag.log 2>&1; echo "$tag tests exit $? $(tail -n 1 gpurun_out/s23_$tag.log)" >> gpurun_out/summary23.txt
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_r1w_$tag.json 2> gpurun_out/bench_r1w_$tag.err; echo "$tag bench exit $?" >> gpurun_out/summary23.txt
done
cat gpurun_out/summary23.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r1w_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['roofline']['kernels']; o=d['roofline'].get('other_kernels',{})
        print(f.split('r1w_')[1][:-5], round(d['value']), 'ms', round(d['ms_per_step'],2), {n:round(v['avg_ms'],2) for n,v in k.items()}, {n:round(v['avg_ms'],3) for n,v in o.items()})
    except Exception as e: print(f, 'ERR', e)
PY
