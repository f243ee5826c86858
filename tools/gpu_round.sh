#!/bin/bash
# One GPU visit: all GPU parity tests, conv micro-bench, bench with the per-plan igemm table. Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader > gpurun_out/gpu.txt
PCGAN_SKIP_TRAJ=1 timeout 1500 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/tests.log 2>&1
tail -15 gpurun_out/tests.log | cut -c1-300
timeout 300 python tools/bench_conv.py > gpurun_out/bench_conv.log 2>&1; tail -12 gpurun_out/bench_conv.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --dump-igemm gpurun_out/igemm_table.txt > gpurun_out/bench.log 2> gpurun_out/bench.err
tail -2 gpurun_out/bench.err; cut -c1-1500 gpurun_out/bench.log
