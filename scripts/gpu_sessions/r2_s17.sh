#!/bin/bash
# round 2, session 17 (2 GPUs): routes side by side (3-stage dense ring) vs one after the other on a side stream (full
# ring, cluster-pair dense kernel), both with submitted steps; at the 8-GPU per-rank load (125k rows) and at 500k rows
mkdir -p gpurun_out
S=gpurun_out/r2s17_summary.txt; : > $S
run() {  # tag, extra args
  tag=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 40 --warmup 5 --no-cpu --enc-chunks 0 --parity-queries 0 --self-check 0 "$@" \
    > gpurun_out/r2s17_$tag.json 2> gpurun_out/r2s17_$tag.err; echo "$tag exit $?" >> $S
}
for rep in 1 2; do
run small_serial_$rep --rows 250000 --serial-routes 1
run small_overlap_$rep --rows 250000 --serial-routes 0
done
run small_serial_span8 --rows 250000 --serial-routes 1 --bm25-span 8
run small_serial_span16 --rows 250000 --serial-routes 1 --bm25-span 16
run small_overlap_span8 --rows 250000 --serial-routes 0 --bm25-span 8
run small_overlap_span16 --rows 250000 --serial-routes 0 --bm25-span 16
run full_serial --serial-routes 1
run full_overlap_span8 --serial-routes 0 --bm25-span 8
run full_overlap --serial-routes 0
run x2_serial --rows 2000000 --serial-routes 1
run x2_overlap --rows 2000000 --serial-routes 0
cat $S
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2s17_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        st = d.get("scaling_terms", {})
        print(f.split("r2s17_")[1], round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms", round(d["ms_per_step"], 3),
              "seq", round(st.get("sequential_step_ms", 0), 3), "dense", round(st.get("dense_ms", 0), 3), "cand", round(st.get("bm25_cand_ms", 0), 3),
              d["setup"]["dense_kernel"], d["digest"].get("matches_committed_n1"))
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json", ".err")).read()[-1500:])
PY
