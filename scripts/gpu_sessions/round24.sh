#!/bin/bash
# new default BM25 shape (8192-doc ranges): all GPU tests, bench base vs overlap, launch list + --set full, memcheck on smoke
mkdir -p gpurun_out; rm -f gpurun_out/summary24.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/s24_all_gpu_tests.log 2>&1; echo "gpu tests exit $?" >> gpurun_out/summary24.txt
run() { tag=$1; shift; timeout 600 python bench.py --steps 10 --warmup 3 "$@" > gpurun_out/bench_r1x_$tag.json 2> gpurun_out/bench_r1x_$tag.err; echo "bench $tag exit $?" >> gpurun_out/summary24.txt; }
run base
run ov4 --overlap 1 --dense-stages 4 --no-cpu
run ov6 --overlap 1 --dense-stages 6 --no-cpu
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"bm25_|dense_|merge_|select_|fuse_" -c 120 --csv --log-file gpurun_out/launches_r1x.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_x_launch.log 2>&1; echo "ncu-list exit $?" >> gpurun_out/summary24.txt
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"bm25_cand_kernel|bm25_rescore_kernel|bm25_bound_kernel|dense_ts_kernel" -s 42 -c 14 -o gpurun_out/prof_r1x python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_x.log 2>&1; echo "ncu-full exit $?" >> gpurun_out/summary24.txt
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python __graft_entry__.py --smoke > gpurun_out/sanitizer_memcheck_r1x.log 2>&1; echo "memcheck exit $?" >> gpurun_out/summary24.txt
cat gpurun_out/summary24.txt
tail -n 4 gpurun_out/s24_all_gpu_tests.log
tail -n 2 gpurun_out/sanitizer_memcheck_r1x.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r1x_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        k=d['roofline']['kernels']; o=d['roofline'].get('other_kernels',{})
        print(f.split('r1x_')[1][:-5], round(d['value']), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), {n:round(v['avg_ms'],2) for n,v in k.items()}, {n:round(v['avg_ms'],3) for n,v in o.items()}, d['clocks'])
    except Exception as e: print(f, 'ERR', e)
PY
