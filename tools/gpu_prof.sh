#!/bin/bash
# ncu passes (run after gpu_round.sh succeeded on the same tree): launch list of a bench run, full capture of the flagship conv.
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu.log 2>&1
echo "ncu launches rc=$?"; wc -l gpurun_out/launches.csv
python tools/prof_two.py > gpurun_out/prof_plain.log 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm -c 6 -f -o gpurun_out/prof_igemm python tools/prof_two.py > gpurun_out/prof_ncu.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/*.ncu-rep
