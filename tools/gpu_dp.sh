#!/bin/bash
# 2-GPU visit: data-parallel consistency check, then the bench at N=2 (one graph with NCCL inside) and at N=1 for the ratio.
mkdir -p gpurun_out
N=${NGPU:-2}
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/dp_check.py > gpurun_out/dp_check.log 2>&1
echo "dp_check rc=$?"; grep -E "OK:|buckets|Error|error|assert" gpurun_out/dp_check.log | head -8
BENCH_WATCHDOG_S=200 timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/bench_dp$N.log 2> gpurun_out/bench_dp$N.err
echo "bench N=$N rc=$?"; tail -3 gpurun_out/bench_dp$N.err | cut -c1-300
[ -n "$SKIP_N1" ] || timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/bench_dp1.log 2> gpurun_out/bench_dp1.err
python - $N <<'PY'
import json,sys
n=sys.argv[1]
for f in ('bench_dp1','bench_dp'+n):
    for l in open('gpurun_out/%s.log'%f):
        if l.startswith('{'):
            d=json.loads(l); print(f,'value %.1f e2e %.1f ms %.3f e2e_ms %.3f'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['e2e']['ms_per_step']))
PY
