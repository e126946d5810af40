#!/bin/bash
# first GPU session: parity tests, split so a trap in the tcgen05 kernel cannot poison the other results
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import os; print('cpus', os.cpu_count())" >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests/test_gpu_retrieval.py -m gpu -q -x -k "not dense and not hybrid" > gpurun_out/t1_bm25_fusion.log 2>&1; echo "t1 exit $?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_dropin.py -m gpu -q > gpurun_out/t2_dropin.log 2>&1; echo "t2 exit $?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_retrieval.py -m gpu -q -k "dense and not 2]" > gpurun_out/t3_dense_simt.log 2>&1; echo "t3 exit $?" >> gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_gpu_retrieval.py -m gpu -q -k "(dense and 2]) or hybrid or fewer" > gpurun_out/t4_dense_tc.log 2>&1; echo "t4 exit $?" >> gpurun_out/summary.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/t5_smoke.log 2>&1; echo "t5 exit $?" >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
tail -5 gpurun_out/t1_bm25_fusion.log gpurun_out/t2_dropin.log gpurun_out/t3_dense_simt.log gpurun_out/t4_dense_tc.log gpurun_out/t5_smoke.log
