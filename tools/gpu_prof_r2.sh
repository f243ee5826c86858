#!/bin/bash
# Round-2 ncu evidence: (1) launch list of a bench run (eager launches), (2) full capture of representative igemm launches
# (resblock fwd / wgrad, narrow layers), (3) full capture of the norm kernels incl. the one-pass cluster kernel.
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-graph > gpurun_out/plain.log 2>&1 &&
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-graph > gpurun_out/ncu.log 2>&1
echo "ncu launches rc=$?"; wc -l gpurun_out/launches.csv
WHICH=1,2,3,5,6 python tools/prof_layers.py > gpurun_out/pl_plain.log 2>&1 &&
WHICH=1,2,3,5,6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm -f -o gpurun_out/prof_igemm python tools/prof_layers.py > gpurun_out/pl_ncu.log 2>&1
echo "ncu igemm rc=$?"
ITERS=1 python tools/bench_norm.py > gpurun_out/norm_plain.log 2>&1 &&
ITERS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:norm_ -s 3 -c 8 -f -o gpurun_out/prof_norm python tools/bench_norm.py > gpurun_out/norm_ncu.log 2>&1
echo "ncu norm rc=$?"; ls -la gpurun_out/*.ncu-rep; tail -12 gpurun_out/norm_plain.log
