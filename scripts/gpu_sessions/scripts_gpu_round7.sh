#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_retrieval.py -m gpu -q -x > gpurun_out/s7_retr.log 2>&1; echo "retr exit $?" >> gpurun_out/summary7.txt
timeout 900 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_dropin.py -m gpu -q > gpurun_out/s7_enc.log 2>&1; echo "enc exit $?" >> gpurun_out/summary7.txt
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; echo "bench exit $?" >> gpurun_out/summary7.txt
timeout 900 python bench_encode.py --arch bert > gpurun_out/enc_bert_f.json 2> gpurun_out/enc_bert_f.err; echo "enc-bert exit $?" >> gpurun_out/summary7.txt
timeout 900 python bench_encode.py --arch qwen2 > gpurun_out/enc_qwen2_f.json 2> gpurun_out/enc_qwen2_f.err; echo "enc-qwen2 exit $?" >> gpurun_out/summary7.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"bm25_score_kernel|dense_ts_kernel" -s 6 -c 2 -o gpurun_out/prof_r1f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_f.log 2>&1; echo "ncu exit $?" >> gpurun_out/summary7.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|attn_bidir" -s 40 -c 8 -o gpurun_out/prof_r1f_enc python bench_encode.py --arch bert --chunks 1024 --queries 128 > gpurun_out/ncu_f2.log 2>&1; echo "ncu-enc exit $?" >> gpurun_out/summary7.txt
cat gpurun_out/summary7.txt
tail -n 5 gpurun_out/s7_*.log
cat gpurun_out/enc_bert_f.json gpurun_out/enc_qwen2_f.json
python - <<'PY'
import json
for f in ("gpurun_out/bench_r1f.json",):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d["value"]), round(d["e2e"]["value"]), {k:(round(v["avg_ms"],2), round(v["GBps"])) for k,v in d["roofline"]["kernels"].items()})
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace('.json','.err')).read()[-2500:])
PY
