#!/bin/bash
# negative-idf index through the ordered kernel (new test) + BM25 tests; configs[4] shape on one GPU with the current kernels
mkdir -p gpurun_out; rm -f gpurun_out/summary32.txt
timeout 900 python -m pytest tests/test_gpu_retrieval.py -m gpu -q -k "bm25" > gpurun_out/s32_tests.log 2>&1; echo "bm25 tests exit $? $(tail -n 1 gpurun_out/s32_tests.log)" >> gpurun_out/summary32.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu --rows 4000000 --dim 1024 --queries 4096 > gpurun_out/bench_r2f_c5.json 2> gpurun_out/bench_r2f_c5.err; echo "bench-c5shape exit $?" >> gpurun_out/summary32.txt
cat gpurun_out/summary32.txt
tail -n 6 gpurun_out/s32_tests.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2f_c5.json').read().strip().splitlines()[-1])
print(round(d['value']), round(d['ms_per_step'],2), d['setup'], {n:round(v['avg_ms'],2) for n,v in d['roofline']['kernels'].items()}, d['roofline'].get('other_kernels'), d['roofline']['kernels'].get('dense_tc',{}).get('TFLOPs'), d['clocks'])
PY
tail -n 3 gpurun_out/bench_r2f_c5.err
