#!/bin/bash
# Round-2 GPU visit: whole GPU suite (with the B=64 trajectories), the bench lines of every workload, sanitizer logs.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader > gpurun_out/gpu.txt
timeout 2400 python -m pytest tests -m gpu -q --tb=short -s > gpurun_out/tests.log 2>&1
echo "tests rc=$?"; grep -E "passed|failed|^FAILED|Error|loss_[GD] " gpurun_out/tests.log | cut -c1-300 | tail -15
timeout 900 python bench.py --dump-igemm gpurun_out/igemm_table.txt > gpurun_out/bench.log 2> gpurun_out/bench.err
echo "bench rc=$?"; tail -2 gpurun_out/bench.err; cut -c1-2500 gpurun_out/bench.log
for w in c256 bayesian siamese; do
  timeout 600 python bench.py --workload $w --steps 40 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$w.log 2> gpurun_out/bench_$w.err
  echo "$w rc=$?"; tail -2 gpurun_out/bench_$w.err | cut -c1-300; cut -c1-700 gpurun_out/bench_$w.log
done
if [ -n "$SANITIZE" ]; then
  for tool in synccheck racecheck memcheck; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/one_step.py > gpurun_out/sanitizer_$tool.log 2>&1
    echo "$tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|losses|hazard" gpurun_out/sanitizer_$tool.log | head -5
  done
fi
