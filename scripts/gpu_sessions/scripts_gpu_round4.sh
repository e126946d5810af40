#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_retrieval.py -m gpu -q -x -k "not dense and not hybrid" > gpurun_out/s4_bm25.log 2>&1; echo "bm25 exit $?" >> gpurun_out/summary4.txt
timeout 600 python -m pytest tests/test_gpu_retrieval.py -m gpu -q -k "dense or hybrid" > gpurun_out/s4_dense.log 2>&1; echo "dense exit $?" >> gpurun_out/summary4.txt
timeout 600 python -m pytest tests/test_gpu_dropin.py -m gpu -q > gpurun_out/s4_dropin.log 2>&1; echo "dropin exit $?" >> gpurun_out/summary4.txt
timeout 900 python -m pytest tests/test_gpu_encoder.py -m gpu -q -k "gemm" > gpurun_out/s4_enc_gemm.log 2>&1; echo "enc-gemm exit $?" >> gpurun_out/summary4.txt
timeout 900 python -m pytest tests/test_gpu_encoder.py -m gpu -q -k "attention or norm or pool" > gpurun_out/s4_enc_ops.log 2>&1; echo "enc-ops exit $?" >> gpurun_out/summary4.txt
timeout 900 python -m pytest tests/test_gpu_encoder.py -m gpu -q -k "encoder" > gpurun_out/s4_enc_model.log 2>&1; echo "enc-model exit $?" >> gpurun_out/summary4.txt
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; echo "bench exit $?" >> gpurun_out/summary4.txt
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu --dense-kernel 2 > gpurun_out/bench_r1c_ss.json 2> gpurun_out/bench_r1c_ss.err; echo "bench-ss exit $?" >> gpurun_out/summary4.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"bm25_score_kernel|dense_ts_kernel" -s 6 -c 2 -o gpurun_out/prof_r1c python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_c.log 2>&1; echo "ncu exit $?" >> gpurun_out/summary4.txt
cat gpurun_out/summary4.txt
tail -n 6 gpurun_out/s4_*.log
python - <<'PY'
import json
for f in ("gpurun_out/bench_r1c.json","gpurun_out/bench_r1c_ss.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["e2e"]["value"]), {k:(round(v["avg_ms"],2), round(v["GBps"])) for k,v in d["roofline"]["kernels"].items()}, d["setup"])
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace('.json','.err')).read()[-1500:])
PY
