#!/bin/bash
# A/B: branch-free atomics with spare slots (default) vs predicated atomics (variant a4e9daad = -DEZR_BM25_PK_BRANCHY=1)
mkdir -p gpurun_out; rm -f gpurun_out/summary29.txt
for rep in 1 2; do
for lib in easyrag_b200/_lib/libeasyrag_b200.so easyrag_b200/_lib/variant_a4e9daad/libeasyrag_b200.so; do
  tag=$(basename $(dirname $lib))_$rep
  EASYRAG_B200_LIB=$PWD/$lib timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_r2c_$tag.json 2> gpurun_out/bench_r2c_$tag.err; echo "$tag exit $?" >> gpurun_out/summary29.txt
done; done
EASYRAG_B200_LIB=$PWD/easyrag_b200/_lib/variant_a4e9daad/libeasyrag_b200.so timeout 600 python -m pytest tests/test_gpu_retrieval.py -m gpu -q -x -k "bm25" > gpurun_out/s29_tests.log 2>&1; echo "variant tests exit $? $(tail -n 1 gpurun_out/s29_tests.log)" >> gpurun_out/summary29.txt
cat gpurun_out/summary29.txt
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r2c_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['roofline']['kernels']
        print(f.split('r2c_')[1][:-5], round(d['value']), 'ms', round(d['ms_per_step'],2), {n:round(v['avg_ms'],3) for n,v in k.items()}, d['gpu_launches'], d['clocks'])
    except Exception as e: print(f, 'ERR', e)
PY
