#!/bin/bash
# 2-GPU session: NCCL shard path parity + scaling bench + retests of the changed kernels
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus5.txt
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q > gpurun_out/s5_dist.log 2>&1; echo "dist exit $?" >> gpurun_out/summary5.txt
timeout 600 python -m pytest tests/test_gpu_retrieval.py -m gpu -q -k "dense or hybrid" > gpurun_out/s5_dense.log 2>&1; echo "dense exit $?" >> gpurun_out/summary5.txt
timeout 600 python -m pytest tests/test_gpu_encoder.py -m gpu -q -k "encoder" > gpurun_out/s5_enc.log 2>&1; echo "enc exit $?" >> gpurun_out/summary5.txt
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_r1d_n1.json 2> gpurun_out/bench_r1d_n1.err; echo "bench1 exit $?" >> gpurun_out/summary5.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r1d_n2.json 2> gpurun_out/bench_r1d_n2.err; echo "bench2 exit $?" >> gpurun_out/summary5.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_r1d_ref2.json 2> gpurun_out/bench_r1d_ref2.err; echo "ref2 exit $?" >> gpurun_out/summary5.txt
cat gpurun_out/summary5.txt
tail -n 5 gpurun_out/s5_*.log
python - <<'PY'
import json
for f in ("gpurun_out/bench_r1d_n1.json","gpurun_out/bench_r1d_n2.json"):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["e2e"]["value"]), {k:(round(v["avg_ms"],2), round(v["GBps"])) for k,v in d["roofline"]["kernels"].items()})
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace('.json','.err')).read()[-2500:])
PY
tail -c 600 gpurun_out/bench_r1d_ref2.json
