#!/bin/bash
# The parity-proper GPU tests (chains, every plan of the step, steps, modes); all failures reported (no -x).
mkdir -p gpurun_out
PCGAN_SKIP_TRAJ=${PCGAN_SKIP_TRAJ:-1} timeout 1500 python -m pytest ${PARITY_TESTS:-tests/test_chain_gpu.py tests/test_bench_geometry_gpu.py tests/test_step_gpu.py tests/test_encoder_modes_gpu.py tests/test_networks_gpu.py} -m gpu -q --tb=short -s > gpurun_out/parity.log 2>&1
echo "rc=$?"
grep -E "chain N=|convolutions, |passed|failed|^FAILED|rel-L2|Error" gpurun_out/parity.log | cut -c1-400 | tail -60
