#!/bin/bash
# Short GPU visit for kernel iterations: igemm parity, conv micro-bench, one bench line with the igemm table.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_igemm_gpu.py -m gpu -q --tb=short -x > gpurun_out/q_tests.log 2>&1
tail -5 gpurun_out/q_tests.log | cut -c1-300
timeout 300 python tools/bench_conv.py > gpurun_out/bench_conv.log 2>&1; tail -18 gpurun_out/bench_conv.log
if [ -n "$QUICK_EXTRA_TESTS" ]; then
  timeout 900 python -m pytest $QUICK_EXTRA_TESTS -m gpu -q --tb=short -x > gpurun_out/q_tests2.log 2>&1
  tail -5 gpurun_out/q_tests2.log | cut -c1-300
fi
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-igemm gpurun_out/igemm_table.txt > gpurun_out/bench.log 2> gpurun_out/bench.err
tail -2 gpurun_out/bench.err; cut -c1-600 gpurun_out/bench.log
