#!/bin/bash
# ncu passes for profiles/ (run after gpu_round.sh succeeded on the same tree):
#  1. launch list of a bench run (per-launch device time of every kernel of the step; eager launches so names are visible)
#  2. full capture (--set full, source) of representative igemm launches and of the norm kernels
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/plain.log 2>&1 &&
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/ncu.log 2>&1
echo "ncu launches rc=$?"; wc -l gpurun_out/launches.csv
WHICH=2,4,6 python tools/prof_layers.py > gpurun_out/pl_plain.log 2>&1 &&
WHICH=2,4,6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm -f -o gpurun_out/prof_igemm python tools/prof_layers.py > gpurun_out/pl_ncu.log 2>&1
echo "ncu igemm rc=$?"
ITERS=1 python tools/bench_norm.py > gpurun_out/norm_plain.log 2>&1 &&
ITERS=1 timeout 600 ncu --set full --clock-control none -k regex:norm_ -s 3 -c 6 -f -o gpurun_out/prof_norm python tools/bench_norm.py > gpurun_out/norm_ncu.log 2>&1
echo "ncu norm rc=$?"; ls -la gpurun_out/*.ncu-rep
