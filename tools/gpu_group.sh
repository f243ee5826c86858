#!/bin/bash
mkdir -p gpurun_out
PCGAN_SKIP_TRAJ=1 timeout 1500 python -m pytest tests/test_networks_gpu.py tests/test_step_gpu.py tests/test_encoder_modes_gpu.py tests/test_chain_gpu.py -m gpu -q --tb=short -s > gpurun_out/group_tests.log 2>&1
echo "rc=$?"; grep -E "grouped|passed|failed|^FAILED|Error|^E  " gpurun_out/group_tests.log | cut -c1-300 | tail -25
for v in 1 0; do
  PCGAN_GROUP=$v timeout 400 python bench.py --steps 60 --warmup 3 --no-cpu-baseline --no-gpu-reference --dump-igemm gpurun_out/igemm_group$v.txt > gpurun_out/bench_group$v.log 2> gpurun_out/bench_group$v.err
  echo "group=$v rc=$?"; tail -2 gpurun_out/bench_group$v.err | cut -c1-200
  python - $v <<'PY'
import json,sys
for l in open('gpurun_out/bench_group%s.log'%sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print('ms/step %.3f e2e %.3f igemm %.3f launches %s'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['gpu_launches_per_step']))
PY
done
