#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_retrieval.py -m gpu -q -k "dense" > gpurun_out/s14_dense.log 2>&1; echo "dense exit $?" >> gpurun_out/summary14.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_r1n.json 2> gpurun_out/bench_r1n.err; echo "bench exit $?" >> gpurun_out/summary14.txt
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu --rows 4000000 --dim 1024 --queries 4096 > gpurun_out/bench_r1n_c5.json 2> gpurun_out/bench_r1n_c5.err; echo "bench-c5shape exit $?" >> gpurun_out/summary14.txt
cat gpurun_out/summary14.txt
tail -n 4 gpurun_out/s14_*.log
python - <<'PY'
import json
for f in ("gpurun_out/bench_r1n.json","gpurun_out/bench_r1n_c5.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["e2e"]["value"]), {k:(round(v["avg_ms"],2), round(v.get("TFLOPs",0))) for k,v in d["roofline"]["kernels"].items()}, d["setup"])
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace('.json','.err')).read()[-2000:])
PY
