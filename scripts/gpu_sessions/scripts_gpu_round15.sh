#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_retrieval.py tests/test_gpu_dropin.py -m gpu -q > gpurun_out/s15_retr.log 2>&1; echo "retr exit $?" >> gpurun_out/summary15.txt
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_r1o.json 2> gpurun_out/bench_r1o.err; echo "bench exit $?" >> gpurun_out/summary15.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"bm25_score_kernel" -s 3 -c 1 -o gpurun_out/prof_r1o python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_o.log 2>&1; echo "ncu exit $?" >> gpurun_out/summary15.txt
cat gpurun_out/summary15.txt
tail -n 4 gpurun_out/s15_*.log
python - <<'PY'
import json
for f in ("gpurun_out/bench_r1o.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["e2e"]["value"]), {k:(round(v["avg_ms"],2), round(v.get("TFLOPs",0))) for k,v in d["roofline"]["kernels"].items()})
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace('.json','.err')).read()[-2000:])
PY
