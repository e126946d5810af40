#!/bin/bash
# round 2, session 22: one full ncu capture of the retrieval kernels of the final tree (routes on one stream)
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"dense_ts_kernel|bm25_cand_kernel|bm25_rescore_kernel|bm25_plan_kernel|bm25_bound_kernel" -s 66 -c 22 -o gpurun_out/r2s22_prof_retr python bench.py --steps 1 --warmup 3 --cal-steps 1 --no-cpu --enc-chunks 0 --parity-queries 0 --self-check 0 --overlap 0 > gpurun_out/r2s22_ncu_retr.log 2>&1; echo "ncu-retr exit $?"
tail -4 gpurun_out/r2s22_ncu_retr.log; ls -la gpurun_out/r2s22_prof_retr.ncu-rep
