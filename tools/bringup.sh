#!/bin/bash
# First-contact check on a GPU box: one plain conv through the tcgen05 kernel under a watchdog,
# then the whole igemm parity file.  Output -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 300 python -m pytest tests/test_igemm_gpu.py -m gpu -x -q -k "res3x3_reflect and forward" > gpurun_out/first.log 2>&1
echo "first rc=$?" >> gpurun_out/first.log
tail -5 gpurun_out/first.log
timeout 900 python -m pytest tests/test_igemm_gpu.py -m gpu -v > gpurun_out/igemm.log 2>&1
echo "igemm rc=$?" >> gpurun_out/igemm.log
grep -E "PASSED|FAILED|ERROR|passed|failed" gpurun_out/igemm.log | tail -60
