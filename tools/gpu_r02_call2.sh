#!/bin/bash
# Round-2 call 2 (1 GPU): parity suite with the colour head on the tensor cores by default + the new ADVICE regression tests;
# rate probe of the layer-1 pipeline (tools/tc_rate.cu): cta_group::1 vs cta_group::2, bulk copies vs tensor-map TMA.
mkdir -p gpurun_out
T0=$SECONDS
timeout 400 python -m pytest tests -x -q -m gpu --timeout 200 > gpurun_out/r02c2_pytest.log 2>&1; echo "pytest rc=$? t=$((SECONDS-T0))s"; tail -5 gpurun_out/r02c2_pytest.log
for cfg in "1 0" "1 1" "1 9" "1 3" "1 11" "1 5" "2 8" "2 9" "2 11" "2 13" "2 10"; do
  timeout 60 tools/bin/tc_rate $cfg 4096 2>&1 | tail -1
done | tee gpurun_out/r02c2_tc_rate.txt
echo "t=$((SECONDS-T0))s"
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02c2_bench.json 2> gpurun_out/r02c2_bench.err; echo "bench rc=$? t=$((SECONDS-T0))s"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02c2_bench.json"))
r = d.get("recon") or {}
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], {k: v for k, v in r.items() if k.startswith("frames")})
PY
