#!/bin/bash
# The tests added last (netIP chain + step, siamese branches / graph, opcheck), then the siamese bench line with the captured trainer.
mkdir -p gpurun_out
timeout 1200 python -m pytest ${NEW_TESTS:-tests/test_chain_gpu.py::test_alexnet_chain_teacher_forced tests/test_step_gpu.py::test_identity_preserving_step_losses tests/test_encoder_modes_gpu.py tests/test_opcheck_gpu.py tests/test_elementwise_gpu.py} -m gpu -q --tb=short -s > gpurun_out/new_tests.log 2>&1
echo "rc=$?"; grep -E "chain N=|IP step|siamese|passed|failed|^FAILED|rel-L2|Error|^E  " gpurun_out/new_tests.log | cut -c1-330 | tail -40
timeout 300 python bench.py --workload siamese --steps 100 --warmup 3 --no-cpu-baseline > gpurun_out/bench_siamese.log 2> gpurun_out/bench_siamese.err
echo "siamese rc=$?"; tail -2 gpurun_out/bench_siamese.err | cut -c1-200; cut -c1-400 gpurun_out/bench_siamese.log
