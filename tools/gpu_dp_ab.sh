#!/bin/bash
# N-GPU A/B of the all-reduce placement: buckets inside the sweep vs all at its end, default NCCL channels vs 4.
mkdir -p gpurun_out
N=${NGPU:-2}
run() { name=$1; shift
  env "$@" BENCH_WATCHDOG_S=150 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29613 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/dpab_$name.log 2> gpurun_out/dpab_$name.err
  echo "$name rc=$?"
  python - $name <<'PY'
import json,sys
for l in open('gpurun_out/dpab_%s.log'%sys.argv[1]):
    if l.startswith('{'):
        d=json.loads(l); print(sys.argv[1],'value %.1f e2e %.1f ms %.3f e2e_ms %.3f'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['e2e']['ms_per_step']))
PY
}
run overlap X=1
run atend PCGAN_BUCKET_OVERLAP=0
run overlap_ch4 NCCL_MAX_NCHANNELS=4
run atend_ch4 PCGAN_BUCKET_OVERLAP=0 NCCL_MAX_NCHANNELS=4
timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-gpu-reference > gpurun_out/dpab_n1.log 2> gpurun_out/dpab_n1.err
python - <<'PY'
import json
for l in open('gpurun_out/dpab_n1.log'):
    if l.startswith('{'):
        d=json.loads(l); print('n1 value %.1f e2e %.1f ms %.3f e2e_ms %.3f'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['e2e']['ms_per_step']))
PY
