#!/bin/bash
# round 2, session 6 (--gpus 8): strong scaling N=1,2,4,8 of configs[2]/[3] on one box, configs[4] (4M x 1024, batch 64) on 8 GPUs
mkdir -p gpurun_out
S=gpurun_out/r2s06_summary.txt; : > $S
run() { # N extra-args tag
  N=$1; TAG=$2; shift 2
  if [ "$N" = "1" ]; then
    timeout 900 python bench.py --gpus 1 "$@" > gpurun_out/r2s06_$TAG.json 2> gpurun_out/r2s06_$TAG.err
  else
    timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$N$N bench.py --gpus $N "$@" > gpurun_out/r2s06_$TAG.json 2> gpurun_out/r2s06_$TAG.err
  fi
  echo "$TAG exit $?" >> $S
}
run 8 scale_n8 --steps 20 --warmup 5 --no-cpu --enc-chunks 0
run 4 scale_n4 --steps 20 --warmup 5 --no-cpu --enc-chunks 0 --parity-queries 64
run 2 scale_n2 --steps 20 --warmup 5 --no-cpu --enc-chunks 0 --parity-queries 64
run 1 scale_n1 --steps 20 --warmup 5 --no-cpu --enc-chunks 0 --parity-queries 64
run 8 c5_n8_b64 --rows 4000000 --dim 1024 --queries 64 --steps 30 --warmup 5 --no-cpu --enc-chunks 0 --parity-queries 64
run 8 c5_n8_b10k --rows 4000000 --dim 1024 --queries 10000 --steps 10 --warmup 3 --no-cpu --enc-chunks 0 --parity-queries 32
run 8 encode_n8 --workload encode
cat $S
python - <<'PY'
import json
for tag in ("scale_n1", "scale_n2", "scale_n4", "scale_n8", "c5_n8_b64", "c5_n8_b10k", "encode_n8"):
    f = f"gpurun_out/r2s06_{tag}.json"
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        if tag.startswith("encode"):
            e = d["encode"]; print(tag, round(d["value"]), "chunks/s", e["n_gpus"], "gemm", e["gemm"]["tflops"], "attn", e["attention"]["tflops"]); continue
        r = d["roofline"]
        print(tag, round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3), r["bound"], r["kernel"], round(r["achieved"]), round(r["frac"], 3),
              {k: (round(v["avg_ms"], 3), round(v.get("avg_ms_in_timed_region", 0), 3)) for k, v in r["kernels"].items()},
              {k: round(v["avg_ms"], 3) for k, v in r["other_kernels"].items()}, "lat", d.get("latency_ms"))
        p = d.get("parity_full_size") or {}
        print("   parity ok", p.get("ok"), p.get("queries"), "digest", d["digest"].get("matches_committed_n1"), d["digest"]["fused_sha256"][:16], "terms", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in d["scaling_terms"].items() if k != "note"})
    except Exception as e:
        print(tag, "ERR", e)
        print(open(f.replace(".json", ".err")).read()[-2500:])
PY
