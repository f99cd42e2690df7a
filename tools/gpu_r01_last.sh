#!/bin/bash
# Last GPU call of round 1 (12 GPU-minutes left): parity suite, default bench line, ncu evidence of the final kernels.
# Every step writes into gpurun_out/ as soon as it ends; most important first.
mkdir -p gpurun_out
T0=$SECONDS
timeout 330 python -m pytest tests -x -q -m gpu --timeout 200 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$? t=$((SECONDS-T0))s" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 240 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$? t=$((SECONDS-T0))s"
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/bench_default.json'))
    print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'], d['e2e']['value'], d['clocks'])
    print(d.get('recon')); print(d.get('cpu_baseline'))
except Exception as e:
    print("bench line unreadable", e)
PY
# ncu --set full of the dominant kernel + the per-frame G0 GEMM at 257^3 (first call: both launch)
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"query_tc3|g0_tc" -c 2 -f -o gpurun_out/r01_final_tc \
  python tools/tc_prof.py 257 > gpurun_out/ncu_tc.log 2>&1; echo "ncu tc rc=$? t=$((SECONDS-T0))s"
ncu -i gpurun_out/r01_final_tc.ncu-rep --page raw --csv > gpurun_out/r01_final_tc_raw.csv 2>/dev/null
# launch list of the default bench's dense step (shares only)
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r01_final_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-recon --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list rc=$? t=$((SECONDS-T0))s"
# ncu --set full of one coarse-to-fine frame (octree + visible surface + marching cubes kernels): HBM GB/s evidence
timeout 200 ncu --set full --clock-control none --profile-from-start off -c 60 -f -o gpurun_out/r01_final_frame \
  python tools/recon_trace.py --no-profiler --frames 1 --mc > gpurun_out/ncu_frame.log 2>&1; echo "ncu frame rc=$? t=$((SECONDS-T0))s"
ncu -i gpurun_out/r01_final_frame.ncu-rep --page raw --csv > gpurun_out/r01_final_frame_raw.csv 2>/dev/null
timeout 120 python tools/recon_trace.py --mc > gpurun_out/recon_trace_mc.txt 2>&1; echo "trace rc=$? t=$((SECONDS-T0))s"
ls -la gpurun_out | head -30
