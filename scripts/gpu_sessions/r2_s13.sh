#!/bin/bash
# round 2, session 13 (2 GPUs): submitted vs joined steps at the per-GPU load of the 8-GPU run (125k rows per rank)
mkdir -p gpurun_out
S=gpurun_out/r2s13_summary.txt; : > $S
run() {  # tag, extra args
  tag=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 40 --warmup 5 --no-cpu --enc-chunks 0 --parity-queries 0 --self-check 0 "$@" \
    > gpurun_out/r2s13_$tag.json 2> gpurun_out/r2s13_$tag.err; echo "$tag exit $?" >> $S
}
for rep in 1 2; do
run small_pl1_$rep --rows 250000 --pipeline 1
run small_pl0_$rep --rows 250000 --pipeline 0
done
run full_pl1 --pipeline 1
run full_pl0 --pipeline 0
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q -x > gpurun_out/r2s13_dist_tests.log 2>&1; echo "dist tests exit $?" >> $S
cat $S; tail -4 gpurun_out/r2s13_dist_tests.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2s13_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        st = d.get("scaling_terms", {})
        print(f.split("r2s13_")[1], round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3),
              "seq", round(st.get("sequential_step_ms", 0), 3), "fixed", round(st.get("fixed_ms", 0), 3), d["digest"].get("matches_committed_n1"))
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json", ".err")).read()[-1500:])
PY
