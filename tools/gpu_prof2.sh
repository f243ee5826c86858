#!/bin/bash
# ncu --set full (with source) of representative narrow igemm launches; plain run first.
mkdir -p gpurun_out; rm -f gpurun_out/prof_narrow.ncu-rep
WHICH=${WHICH:-1,3,5} REPS=2 python tools/prof_layers.py > gpurun_out/pl_plain.log 2>&1 || { tail -5 gpurun_out/pl_plain.log; exit 1; }
WHICH=${WHICH:-1,3,5} REPS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:igemm -f -o gpurun_out/prof_narrow python tools/prof_layers.py > gpurun_out/pl_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/*.ncu-rep
