#!/bin/bash
# instruction counts of the norm-family kernels at one shape, for the in-tree library and (if present) build/base
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size
export NB_ONCE=1 NB_SHAPES=${NB_SHAPES:-32x256}
timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/nb_ncu_new.csv -k regex:'norm_|halo_fold' python tools/norm_bench.py > gpurun_out/nb_ncu_new.log 2>&1
if [ -f build/base/libpcgan_kernels.so ]; then
PCGAN_KERNELS_LIB=/root/repo/build/base/libpcgan_kernels.so timeout 300 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/nb_ncu_base.csv -k regex:'norm_|halo_fold' python tools/norm_bench.py > gpurun_out/nb_ncu_base.log 2>&1
fi
