#!/bin/bash
# round 2, session 16: validation of the tree as committed -- every GPU test, smoke(), the default bench line (both
# arms), the ncu launch list of the default command and one full capture of the encoder kernels
mkdir -p gpurun_out
S=gpurun_out/r2s16_summary.txt; : > $S
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2s16_gpu_tests.log 2>&1; echo "gpu tests exit $?" >> $S
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2s16_smoke.log 2>&1; echo "smoke exit $?" >> $S
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/r2s16_bench_default.json 2> gpurun_out/r2s16_bench_default.err; echo "bench default exit $?" >> $S
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2s16_bench_reference.json 2> gpurun_out/r2s16_bench_reference.err; echo "bench reference exit $?" >> $S
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2s16_launches.csv python bench.py --steps 2 --warmup 3 --cal-steps 1 --no-cpu --enc-chunks 2048 --parity-queries 0 --self-check 0 > gpurun_out/r2s16_ncu_launches.log 2>&1; echo "ncu launch list exit $?" >> $S
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_tc_kernel|attn_tc_kernel" -s 40 -c 6 -o gpurun_out/r2s16_prof_enc python bench_encode.py --arch bert --chunks 2048 --enc-queries 128 > gpurun_out/r2s16_ncu_enc.log 2>&1; echo "ncu-enc exit $?" >> $S
cat $S
tail -n 6 gpurun_out/r2s16_gpu_tests.log
tail -n 3 gpurun_out/r2s16_smoke.log
python - <<'PY'
import json
for tag in ("default", "reference"):
    f = f"gpurun_out/r2s16_bench_{tag}.json"
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(tag, {k: d.get(k) for k in ("metric", "value", "unit", "n_gpus", "ms_per_step", "gpu_launches", "impl")})
        print("   e2e", d.get("e2e"), "cpu", d.get("cpu_baseline"))
        if tag == "default":
            r = d["roofline"]
            print("   roofline", r["bound"], r["kernel"], round(r["achieved"]), r["peak"], round(r["frac"], 3), r.get("traffic"),
                  {k: (round(v["avg_ms"], 3), round(v.get("avg_ms_in_timed_region", 0), 3)) for k, v in r["kernels"].items()})
            p = d.get("parity_full_size") or {}
            print("   parity", p.get("ok"), {k: p.get(k) for k in ("queries", "bm25", "dense", "rrf")} if p else None)
            print("   digest", d["digest"], "clocks", d["clocks"])
            e = d.get("encode") or {}
            print("   encode", {k: e.get(k) for k in ("chunks_per_s", "queries_per_s")}, (e.get("gemm") or {}).get("tflops"), (e.get("attention") or {}).get("tflops"), (e.get("parity") or {}).get("ok"))
    except Exception as ex:
        print(tag, "ERR", ex)
        print(open(f.replace(".json", ".err")).read()[-2500:])
PY
head -c 600 gpurun_out/r2s16_launches.csv | tail -c 300; wc -l gpurun_out/r2s16_launches.csv
