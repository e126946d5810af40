#!/bin/bash
# round 2, session 18 (8 GPUs): the scaling curve with submitted steps on one box: N = 8, 4, 1 and the encode workload at N = 8
mkdir -p gpurun_out
S=gpurun_out/r2s18_summary.txt; : > $S
run() {  # n, tag, extra args
  n=$1; tag=$2; shift 2
  if [ "$n" = 1 ]; then
    timeout 900 python bench.py --gpus 1 "$@" > gpurun_out/r2s18_$tag.json 2> gpurun_out/r2s18_$tag.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $n "$@" > gpurun_out/r2s18_$tag.json 2> gpurun_out/r2s18_$tag.err
  fi
  echo "run $n $tag exit $?" >> $S
}
run 8 scale_n8 --steps 40 --warmup 5 --no-cpu --enc-chunks 0 --parity-queries 64
run 8 scale_n8_joined --steps 40 --warmup 5 --no-cpu --enc-chunks 0 --parity-queries 0 --self-check 0 --pipeline 0
run 4 scale_n4 --steps 40 --warmup 5 --no-cpu --enc-chunks 0 --parity-queries 0 --self-check 0
run 1 scale_n1 --steps 20 --warmup 5 --no-cpu --enc-chunks 0 --parity-queries 0 --self-check 0
run 8 encode_n8 --workload encode --steps 3 --warmup 1
cat $S
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2s18_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        st = d.get("scaling_terms") or {}
        print(f.split("r2s18_")[1], round(d["value"]), d["unit"], "e2e", round(d["e2e"]["value"]) if d.get("e2e") else None, "ms", round(d["ms_per_step"], 3),
              "seq", round(st.get("sequential_step_ms") or 0, 3), "dense", round(st.get("dense_ms") or 0, 3), "cand", round(st.get("bm25_cand_ms") or 0, 3),
              "fixed", round(st.get("fixed_ms") or 0, 3), (d.get("digest") or {}).get("matches_committed_n1"), (d.get("parity_full_size") or {}).get("ok"), (d.get("clocks") or {}).get("sm_mhz"))
    except Exception as e:
        print(f, "ERR", e); print(open(f.replace(".json", ".err")).read()[-1500:])
PY
