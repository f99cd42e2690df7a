#!/bin/bash
mkdir -p gpurun_out
# launch list (shares) of one default bench run
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_v3.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
grep -c "query_tc3" gpurun_out/launches_v3.csv
# full capture of the dominant kernel at the bench size (257^3), plus the per-texel GEMM
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"query_tc3|g0_kernel" -c 2 -o gpurun_out/prof_tc3 python bench.py --steps 1 --warmup 1 --no-recon --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out/*.ncu-rep
