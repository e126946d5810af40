#!/bin/bash
# round 2, session 21: validation of the final tree -- every GPU test, smoke(), the default bench line, the reference arm
mkdir -p gpurun_out
S=gpurun_out/r2s21_summary.txt; : > $S
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2s21_gpu_tests.log 2>&1; echo "gpu tests exit $?" >> $S
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2s21_smoke.log 2>&1; echo "smoke exit $?" >> $S
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2s21_bench_default.json 2> gpurun_out/r2s21_bench_default.err; echo "bench default exit $?" >> $S
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2s21_bench_reference.json 2> gpurun_out/r2s21_bench_reference.err; echo "bench reference exit $?" >> $S
cat $S
tail -n 5 gpurun_out/r2s21_gpu_tests.log
tail -n 2 gpurun_out/r2s21_smoke.log
python - <<'PY'
import json
for tag in ("default", "reference"):
    f = f"gpurun_out/r2s21_bench_{tag}.json"
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(tag, {k: d.get(k) for k in ("value", "unit", "n_gpus", "ms_per_step", "gpu_launches", "impl")}, "e2e", (d.get("e2e") or {}).get("value"))
        if tag == "default":
            r = d["roofline"]
            print("   roofline", r["kernel"], round(r["achieved"]), round(r["frac"], 3), "parity", (d.get("parity_full_size") or {}).get("ok"), "digest", d["digest"].get("matches_committed_n1"), "clocks", d["clocks"]["sm_mhz"])
            e = d.get("encode") or {}
            print("   encode", e.get("chunks_per_s"), (e.get("gemm") or {}).get("tflops"), (e.get("attention") or {}).get("tflops"), (e.get("parity") or {}).get("ok"))
    except Exception as ex:
        print(tag, "ERR", ex); print(open(f.replace(".json", ".err")).read()[-2000:])
PY
