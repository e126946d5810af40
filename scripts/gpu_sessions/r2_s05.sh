#!/bin/bash
# round 2, session 5 (--gpus 2): NCCL parity test of the sharded ranker, BM25 candidate-pass rework, N=1 / N=2 bench
mkdir -p gpurun_out
S=gpurun_out/r2s05_summary.txt; : > $S
timeout 900 python -m pytest tests/test_gpu_dist.py tests/test_gpu_retrieval.py -m gpu -q -x > gpurun_out/r2s05_tests.log 2>&1; echo "tests exit $?" >> $S
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --enc-chunks 0 > gpurun_out/r2s05_bench_n1.json 2> gpurun_out/r2s05_bench_n1.err; echo "bench n1 exit $?" >> $S
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2s05_bench_n2.json 2> gpurun_out/r2s05_bench_n2.err; echo "bench n2 exit $?" >> $S
cat $S
tail -n 6 gpurun_out/r2s05_tests.log
python - <<'PY'
import json
for tag in ("n1", "n2"):
    f = f"gpurun_out/r2s05_bench_{tag}.json"
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(tag, round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), r["bound"], r["kernel"], round(r["achieved"]), round(r["frac"], 3),
              {k: (round(v["avg_ms"], 3), round(v.get("avg_ms_in_timed_region", 0), 3)) for k, v in r["kernels"].items()},
              {k: round(v["avg_ms"], 3) for k, v in r["other_kernels"].items()})
        print("   parity", d.get("parity_full_size"))
        print("   digest", d.get("digest"))
        print("   terms", d.get("scaling_terms"))
        if d.get("encode"): print("   encode", {k: d["encode"][k] for k in ("chunks_per_s", "n_gpus")}, d["encode"]["gemm"]["tflops"], d["encode"]["attention"]["tflops"])
    except Exception as e:
        print(tag, "ERR", e)
        print(open(f.replace(".json", ".err")).read()[-3000:])
PY
