#!/bin/bash
# 8-GPU session: scaling bench exactly as the driver launches it, plus a quick parity check of the changed BM25 kernel
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus13.txt
timeout 600 python -m pytest tests/test_gpu_retrieval.py -m gpu -q -x -k "bm25 or hybrid" > gpurun_out/s13_bm25.log 2>&1; echo "bm25 exit $?" >> gpurun_out/summary13.txt
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q > gpurun_out/s13_dist.log 2>&1; echo "dist exit $?" >> gpurun_out/summary13.txt
for N in 1 2 4 8; do
  if [ $N -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
  fi
  echo "scale $N exit $?" >> gpurun_out/summary13.txt
done
cat gpurun_out/summary13.txt
tail -n 3 gpurun_out/s13_*.log
python - <<'PY'
import json
for n in (1,2,4,8):
    f=f"gpurun_out/scale_n{n}.json"
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(n, round(d["value"]), round(d["e2e"]["value"]), round(d["ms_per_step"],2), {k:(round(v["avg_ms"],2)) for k,v in d["roofline"]["kernels"].items()})
    except Exception as e:
        print(n, "ERR", e); print(open(f.replace('.json','.err')).read()[-2000:])
PY
