#!/bin/bash
# One GPU visit: all GPU parity tests, conv micro-bench, bench with the per-plan igemm table. Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader > gpurun_out/gpu.txt
PCGAN_SKIP_TRAJ=${PCGAN_SKIP_TRAJ:-1} timeout 1800 python -m pytest tests -m gpu -q --tb=short -x -s > gpurun_out/tests.log 2>&1
grep -E "passed|failed|Error|error|loss_[GD] " gpurun_out/tests.log | cut -c1-300 | tail -15
timeout 300 python tools/bench_conv.py > gpurun_out/bench_conv.log 2>&1; tail -12 gpurun_out/bench_conv.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --dump-igemm gpurun_out/igemm_table.txt > gpurun_out/bench.log 2> gpurun_out/bench.err
tail -2 gpurun_out/bench.err; cut -c1-1800 gpurun_out/bench.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_nograph.log 2> gpurun_out/bench_nograph.err
python -c "
import json
for f in ('bench','bench_nograph'):
    for l in open('gpurun_out/%s.log'%f):
        if l.startswith('{'):
            d=json.loads(l); print(f, 'ms/step %.2f e2e %.2f host %.2f igemm %.2f frac %.3f'%(d['ms_per_step'], d['e2e']['ms_per_step'], d['host_enqueue_ms_per_step'], d['roofline']['kernel_ms_per_step'], d['roofline']['frac']))
"
