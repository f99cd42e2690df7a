#!/bin/bash
# Round-2 call 9 (1 GPU): full parity suite (frame graph, range guard, bench head) + the restructured bench line
mkdir -p gpurun_out
T0=$SECONDS
timeout 600 python -m pytest tests -x -q -m gpu --timeout 300 > gpurun_out/r02c9_pytest.log 2>&1; echo "pytest rc=$? t=$((SECONDS-T0))s"; tail -15 gpurun_out/r02c9_pytest.log
timeout 500 python bench.py > gpurun_out/r02c9_bench.json 2> gpurun_out/r02c9_bench.err; echo "bench rc=$? t=$((SECONDS-T0))s"; tail -5 gpurun_out/r02c9_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02c9_bench.json"))
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"], "parity", d["parity_max_abs"], "vol ok", d["volume_matches_single_gpu"], "e2e", d["e2e"]["value"])
print(json.dumps(d.get("recon"), indent=1)[:2500])
print(d.get("configs4_dense513"))
print(d.get("cpu_baseline"))
PY
