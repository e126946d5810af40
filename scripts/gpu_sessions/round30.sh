#!/bin/bash
# 2 GPUs: sharded path (row shards, one all_gather of packed records) with the two-phase BM25 and TS128 dense kernels
mkdir -p gpurun_out; rm -f gpurun_out/summary30.txt
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q > gpurun_out/s30_dist_tests.log 2>&1; echo "dist tests exit $? $(tail -n 1 gpurun_out/s30_dist_tests.log)" >> gpurun_out/summary30.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2 --steps 30 --warmup 3 > gpurun_out/bench_r2d_n2.json 2> gpurun_out/bench_r2d_n2.err; echo "bench n2 exit $?" >> gpurun_out/summary30.txt
timeout 600 python bench.py --gpus 1 --steps 30 --warmup 3 --no-cpu > gpurun_out/bench_r2d_n1.json 2> gpurun_out/bench_r2d_n1.err; echo "bench n1 exit $?" >> gpurun_out/summary30.txt
cat gpurun_out/summary30.txt
python - <<'PY'
import json
for f in ('gpurun_out/bench_r2d_n1.json','gpurun_out/bench_r2d_n2.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); k=d['roofline']['kernels']
        print(f, d['n_gpus'], round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), {n:round(v['avg_ms'],2) for n,v in k.items()}, d['gpu_launches'], d['clocks'])
    except Exception as e: print(f,'ERR',e)
PY
tail -n 5 gpurun_out/bench_r2d_n2.err
