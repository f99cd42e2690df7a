#!/bin/bash
# Round-2 final 1-GPU evidence: parity suite, sanitizer target, default bench line + reference arm, ncu of the final kernels.
# Every step writes into gpurun_out/ as soon as it ends; most important first.
mkdir -p gpurun_out
T0=$SECONDS
timeout 400 python -m pytest tests -x -q -m gpu --timeout 300 > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc=$? t=$((SECONDS-T0))s"; tail -3 gpurun_out/r02f_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02f_smoke.log 2>&1; echo "smoke rc=$? t=$((SECONDS-T0))s"; tail -2 gpurun_out/r02f_smoke.log
timeout 400 python bench.py > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench rc=$? t=$((SECONDS-T0))s"
python - <<'PY'
import json
try:
    d = json.load(open('gpurun_out/r02f_bench.json'))
    print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['clocks'], d['parity_max_abs'])
    print(json.dumps(d.get('recon'))[:1200]); print(d.get('configs4_dense513')); print(d.get('cpu_baseline'))
except Exception as e:
    print("bench line unreadable", e)
PY
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02f_bench_reference_arm.json 2>/dev/null; echo "reference arm rc=$? t=$((SECONDS-T0))s"
# ncu --set full of the dominant kernel + the per-frame G0 GEMM at 257^3
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"query_tc3|g0_tc" -c 2 -f -o gpurun_out/r02_final_tc \
  python tools/tc_prof.py 257 > gpurun_out/r02f_ncu_tc.log 2>&1; echo "ncu tc rc=$? t=$((SECONDS-T0))s"
ncu -i gpurun_out/r02_final_tc.ncu-rep --page raw --csv > gpurun_out/r02_final_tc_raw.csv 2>/dev/null
# launch list of the default bench's dense step (shares only)
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_final_launches.csv \
  python bench.py --steps 3 --warmup 3 --no-recon --no-cpu-baseline > gpurun_out/r02f_bench_under_ncu.log 2>&1; echo "launch list rc=$? t=$((SECONDS-T0))s"
# ncu --set full of one coarse-to-fine frame (octree + visible surface + marching cubes kernels): HBM GB/s evidence
timeout 300 ncu --set full --clock-control none --profile-from-start off -c 70 -f -o gpurun_out/r02_final_frame \
  python tools/recon_trace.py --no-profiler --frames 1 --mc > gpurun_out/r02f_ncu_frame.log 2>&1; echo "ncu frame rc=$? t=$((SECONDS-T0))s"
ncu -i gpurun_out/r02_final_frame.ncu-rep --page raw --csv > gpurun_out/r02_final_frame_raw.csv 2>/dev/null
timeout 120 python tools/recon_trace.py --mc 2>&1 | grep -v Warn > gpurun_out/r02_final_recon_trace_fv_mc.txt; echo "trace rc=$? t=$((SECONDS-T0))s"
timeout 120 python tools/recon_trace.py --color 2>&1 | grep -v Warn | head -14 > gpurun_out/r02_final_recon_trace_color.txt
# precision sweep of the tensor-core program on the GPU (12 seeds x 20 000 points)
timeout 200 python tools/precision_sweep.py > gpurun_out/r02_final_precision_sweep.log 2>&1; tail -4 gpurun_out/r02_final_precision_sweep.log
# compute-sanitizer memcheck over every kernel family (opt-in pytest target)
MONOPORT_B200_RUN_SANITIZER=1 MONOPORT_B200_SANITIZER_LOG=gpurun_out/r02_sanitizer_memcheck.log timeout 1500 python -m pytest tests/test_sanitizer_gpu.py -q -m gpu > gpurun_out/r02f_pytest_sanitizer.log 2>&1; echo "sanitizer rc=$? t=$((SECONDS-T0))s"; tail -3 gpurun_out/r02f_pytest_sanitizer.log
ls -la gpurun_out | grep r02_final | head
