#!/bin/bash
# Final round-2 evidence on one box: full GPU suite, smoke, the four bench workloads, per-plan tables, the ncu launch list of
# a bench run, one full capture of the rewritten norm kernels, the norm micro-benchmark.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -15 > gpurun_out/r2_final_gpu_tests.log; tail -3 gpurun_out/r2_final_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --dump-igemm gpurun_out/r2_igemm_table.txt > gpurun_out/r2_bench_default.log 2> gpurun_out/r2_bench_default.err; echo "bench default rc=$?"
for w in c256 bayesian siamese; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline --no-gpu-reference > gpurun_out/r2_bench_$w.log 2> gpurun_out/r2_bench_$w.err; echo "bench $w rc=$?"
done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_reference.log 2>/dev/null; echo "reference arm rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-graph > gpurun_out/r2_ncu.log 2>&1
echo "ncu launches rc=$?"; wc -l gpurun_out/r2_launches.csv
NB_SHAPES=32x256,64x128,128x64 timeout 300 python tools/norm_bench.py > gpurun_out/r2_norm_bench.txt 2>&1; tail -1 gpurun_out/r2_norm_bench.txt
NB_SHAPES=32x256 bash tools/gpu_nb_ncu.sh; echo "nb ncu rc=$?"
