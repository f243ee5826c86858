#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_elementwise_gpu.py tests/test_chain_gpu.py -m gpu -q --tb=short -s > gpurun_out/fused_tests.log 2>&1
echo "tests rc=$?"; grep -E "one-pass|chain N=|passed|failed|^FAILED|Error" gpurun_out/fused_tests.log | cut -c1-300 | tail -20
for v in 1 0; do
  PCGAN_NORM_FUSED=$v timeout 400 python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-gpu-reference --dump-igemm gpurun_out/igemm_fused$v.txt > gpurun_out/bench_fused$v.log 2> gpurun_out/bench_fused$v.err
  echo "fused=$v rc=$?"; tail -2 gpurun_out/bench_fused$v.err | cut -c1-200
  python - $v <<'PY'
import json,sys
for l in open('gpurun_out/bench_fused%s.log'%sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print('ms/step %.3f e2e %.3f igemm %.3f hbm %s'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms_per_step'], json.dumps(d['roofline_hbm']['per_kernel'])[:600]))
PY
done
cat gpurun_out/igemm_fused1.txt.ops
