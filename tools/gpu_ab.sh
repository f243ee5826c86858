#!/bin/bash
# A/B of the kernel switches (each env var alone off against all on) + the GPU suite on the all-on build.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader > gpurun_out/gpu.txt
PCGAN_SKIP_TRAJ=1 timeout 900 python -m pytest tests -m gpu -q --tb=short -x > gpurun_out/tests.log 2>&1
echo "tests rc=$?"; grep -E "passed|failed|Error|error" gpurun_out/tests.log | cut -c1-300 | tail -8
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --dump-igemm gpurun_out/igemm_$name.txt > gpurun_out/bench_$name.log 2> gpurun_out/bench_$name.err
  echo "$name rc=$?"
  python - "$name" <<'PY'
import json,sys
n=sys.argv[1]
for l in open('gpurun_out/bench_%s.log'%n):
    if l.startswith('{'):
        d=json.loads(l); print(n,'ms/step %.3f e2e %.3f igemm %.3f frac %.3f mhz %s'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['roofline']['frac'], d['clocks']['sm_mhz']))
PY
}
run all X=1
run nokps PCGAN_KPS=0
run nosplit PCGAN_SPLITP=0
run nophase PCGAN_PHASE_STREAMS=0
run none PCGAN_KPS=0 PCGAN_SPLITP=0 PCGAN_PHASE_STREAMS=0
