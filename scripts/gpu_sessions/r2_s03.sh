#!/bin/bash
# round 2, session 3: validate index-build kernels / dense seeding / vector store; default bench with the encode block;
# ncu --set full of the tcgen05 attention kernel (why 152 TFLOP/s?) and L=512 uniform attention throughput
mkdir -p gpurun_out
S=gpurun_out/r2s03_summary.txt; : > $S
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2s03_tests.log 2>&1; echo "tests exit $?" >> $S
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2s03_bench_n1.json 2> gpurun_out/r2s03_bench_n1.err; echo "bench exit $?" >> $S
timeout 600 python bench_encode.py --arch bert --chunks 20000 --len-min 512 --len-max 512 > gpurun_out/r2s03_enc_bert_L512.json 2> gpurun_out/r2s03_enc_bert_L512.err; echo "enc-L512 exit $?" >> $S
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attn_tc_kernel" -s 30 -c 2 -o gpurun_out/r2s03_prof_attn python bench_encode.py --arch bert --chunks 2048 --enc-queries 128 > gpurun_out/r2s03_ncu_attn.log 2>&1; echo "ncu-attn exit $?" >> $S
cat $S
tail -n 6 gpurun_out/r2s03_tests.log
cat gpurun_out/r2s03_enc_bert_L512.json
python - <<'PY'
import json
f = "gpurun_out/r2s03_bench_n1.json"
try:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    r = d["roofline"]
    print("n1", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), r["bound"], r["kernel"], round(r["achieved"]), round(r["frac"], 3),
          {k: (round(v["avg_ms"], 3), round(v.get("avg_ms_in_timed_region", 0), 3)) for k, v in r["kernels"].items()})
    print("   parity", d.get("parity_full_size"))
    print("   digest", d.get("digest"))
    print("   setup", d.get("setup"))
    print("   encode", json.dumps(d.get("encode"))[:1500])
except Exception as e:
    print("ERR", e)
    print(open(f.replace(".json", ".err")).read()[-3000:])
PY
