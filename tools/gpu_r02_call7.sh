#!/bin/bash
# Round-2 call 7 (1 GPU): CTA-pair (cta_group::2) variant of the geometry program: parity, then A/B against the one-CTA kernel
mkdir -p gpurun_out
T0=$SECONDS
timeout 300 python -m pytest tests/test_query_gpu.py tests/test_engine_gpu.py -x -q -m gpu --timeout 120 > gpurun_out/r02c7_pytest.log 2>&1; echo "pytest rc=$? t=$((SECONDS-T0))s"; tail -5 gpurun_out/r02c7_pytest.log
timeout 200 python bench.py --no-cpu-baseline --no-recon > gpurun_out/r02c7_bench_cg2.json 2> gpurun_out/r02c7_bench_cg2.err; echo "bench cg2 rc=$? t=$((SECONDS-T0))s"
MONOPORT_B200_TC_CG=1 timeout 200 python bench.py --no-cpu-baseline --no-recon > gpurun_out/r02c7_bench_cg1.json 2> gpurun_out/r02c7_bench_cg1.err; echo "bench cg1 rc=$? t=$((SECONDS-T0))s"
python - <<'PY'
import json
for f in ("cg2", "cg1"):
    try:
        d = json.load(open("gpurun_out/r02c7_bench_%s.json" % f))
        print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"])
    except Exception as e:
        print(f, "unreadable", e)
PY
MONOPORT_B200_TC_PROF=1 timeout 120 python tools/tc_prof.py 257 2>&1 | grep "tc prof" | head -22 > gpurun_out/r02c7_tc_inkernel_cycles_cg2.txt; cat gpurun_out/r02c7_tc_inkernel_cycles_cg2.txt
timeout 120 python tools/recon_trace.py 2>&1 | grep -v Warn | head -8
