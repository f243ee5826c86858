#!/bin/bash
# One GPU visit: module parity, smoke, short bench. Logs -> gpurun_out/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_networks_gpu.py -m gpu -q --tb=line -s > gpurun_out/networks.log 2>&1
grep -E "passed|failed|^G |^D |^E |^Basic|Error|^/root" gpurun_out/networks.log | cut -c1-700 | head -40
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -4 gpurun_out/smoke.log
PCGAN_SKIP_TRAJ=1 timeout 600 python -m pytest tests/test_step_gpu.py -m gpu -q --tb=short -s > gpurun_out/step.log 2>&1; grep -E "passed|failed|step losses|Error|error" gpurun_out/step.log | cut -c1-600 | head
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.log | cut -c1-3000
