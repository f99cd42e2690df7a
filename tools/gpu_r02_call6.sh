#!/bin/bash
# Round-2 call 6 (1 GPU): warp-uniform MMA issue in every tensor-core kernel + range guard + colour W3 pair + marching cubes
mkdir -p gpurun_out
T0=$SECONDS
timeout 500 python -m pytest tests -x -q -m gpu --timeout 200 > gpurun_out/r02c6_pytest.log 2>&1; echo "pytest rc=$? t=$((SECONDS-T0))s"; tail -5 gpurun_out/r02c6_pytest.log
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02c6_bench.json 2> gpurun_out/r02c6_bench.err; echo "bench rc=$? t=$((SECONDS-T0))s"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02c6_bench.json"))
r = d.get("recon") or {}
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"], {k: v for k, v in r.items() if k.startswith("frames")})
PY
MONOPORT_B200_TC_PROF=1 timeout 120 python tools/tc_prof.py 257 2>&1 | grep "tc prof" > gpurun_out/r02c6_tc_inkernel_cycles.txt; cat gpurun_out/r02c6_tc_inkernel_cycles.txt
timeout 120 python tools/recon_trace.py --mc 2>&1 | grep -v Warn > gpurun_out/r02c6_recon_trace_mc.txt; head -24 gpurun_out/r02c6_recon_trace_mc.txt
