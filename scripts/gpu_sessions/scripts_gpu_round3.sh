#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_retrieval.py -m gpu -q -x -k "not dense and not hybrid" > gpurun_out/s3_bm25.log 2>&1; echo "bm25 exit $?" >> gpurun_out/summary3.txt
timeout 600 python -m pytest tests/test_gpu_retrieval.py -m gpu -q -k "dense and not 3]" > gpurun_out/s3_dense.log 2>&1; echo "dense exit $?" >> gpurun_out/summary3.txt
timeout 600 python -m pytest tests/test_gpu_retrieval.py -m gpu -q -k "dense and 3]" > gpurun_out/s3_dense_ts.log 2>&1; echo "dense-ts exit $?" >> gpurun_out/summary3.txt
timeout 600 python -m pytest tests/test_gpu_retrieval.py tests/test_gpu_dropin.py -m gpu -q -k "hybrid or dropin or retriever or fusion" > gpurun_out/s3_rest.log 2>&1; echo "rest exit $?" >> gpurun_out/summary3.txt
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_r1b_ss.json 2> gpurun_out/bench_r1b_ss.err; echo "bench-ss exit $?" >> gpurun_out/summary3.txt
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu --dense-kernel 3 > gpurun_out/bench_r1b_ts.json 2> gpurun_out/bench_r1b_ts.err; echo "bench-ts exit $?" >> gpurun_out/summary3.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"bm25_score_kernel" -s 3 -c 1 -o gpurun_out/prof_r1b_bm25 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_b1.log 2>&1; echo "ncu-bm25 exit $?" >> gpurun_out/summary3.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"dense_t" -s 3 -c 1 -o gpurun_out/prof_r1b_dense python bench.py --steps 1 --warmup 3 --no-cpu --dense-kernel 3 > gpurun_out/ncu_b2.log 2>&1; echo "ncu-dense exit $?" >> gpurun_out/summary3.txt
cat gpurun_out/summary3.txt
tail -n 4 gpurun_out/s3_bm25.log gpurun_out/s3_dense.log gpurun_out/s3_dense_ts.log gpurun_out/s3_rest.log
python - <<'PY'
import json
for f in ("gpurun_out/bench_r1b_ss.json","gpurun_out/bench_r1b_ts.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d["value"], d["e2e"]["value"], {k:(round(v["avg_ms"],2), round(v["GBps"])) for k,v in d["roofline"]["kernels"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
