#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_retrieval.py -m gpu -q -k "bm25 or hybrid or select" > gpurun_out/s16_retr.log 2>&1; echo "retr exit $?" >> gpurun_out/summary16.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_r1p.json 2> gpurun_out/bench_r1p.err; echo "bench exit $?" >> gpurun_out/summary16.txt
cat gpurun_out/summary16.txt
tail -n 3 gpurun_out/s16_*.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/bench_r1p.json").read().strip().splitlines()[-1])
print(round(d["value"]), round(d["e2e"]["value"]), {k:(round(v["avg_ms"],2)) for k,v in d["roofline"]["kernels"].items()})
PY
