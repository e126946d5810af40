#!/bin/bash
# BM25 launch-shape sweep: the same kernel built with different (range, threads, min blocks)
mkdir -p gpurun_out
for lib in easyrag_b200/_lib/libeasyrag_b200.so easyrag_b200/_lib/variant_*/libeasyrag_b200.so; do
  tag=$(basename $(dirname $lib))
  export EASYRAG_B200_LIB=$PWD/$lib
  timeout 600 python -m pytest tests/test_gpu_retrieval.py -m gpu -q -x -k "bm25" > gpurun_out/s17_$tag.log 2>&1; echo "$tag tests exit $?" >> gpurun_out/summary17.txt
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_r1q_$tag.json 2> gpurun_out/bench_r1q_$tag.err; echo "$tag bench exit $?" >> gpurun_out/summary17.txt
  python - <<PY >> gpurun_out/summary17.txt
import json,ctypes
try:
    d=json.loads(open("gpurun_out/bench_r1q_$tag.json").read().strip().splitlines()[-1])
    print("   $tag range", ctypes.CDLL("$lib").ezr_bm25_range_size(), round(d["value"]), {k:(round(v["avg_ms"],2)) for k,v in d["roofline"]["kernels"].items()})
except Exception as e:
    print("   $tag ERR", e)
PY
done
unset EASYRAG_B200_LIB
timeout 900 python -m pytest tests/test_gpu_encoder.py -m gpu -q -k "dropin" > gpurun_out/s17_enc_dropin.log 2>&1; echo "enc dropin exit $?" >> gpurun_out/summary17.txt
cat gpurun_out/summary17.txt
tail -n 5 gpurun_out/s17_enc_dropin.log
