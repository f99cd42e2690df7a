#!/bin/bash
# Round-2 call 21 (1 GPU): all tensor-core kernels with one elected region per chunk / phase in their issuers, weight multicast
# by default for multi-wave launches: the whole GPU suite, then the default bench line.
mkdir -p gpurun_out
T0=$SECONDS
timeout 500 python -m pytest tests -x -q -m gpu --timeout 300 > gpurun_out/r02c21_pytest.log 2>&1; echo "pytest rc=$? t=$((SECONDS-T0))s"; tail -3 gpurun_out/r02c21_pytest.log
timeout 400 python bench.py > gpurun_out/r02c21_bench.json 2> gpurun_out/r02c21_bench.err; echo "bench rc=$? t=$((SECONDS-T0))s"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02c21_bench.json'))
print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['clocks'], d['parity_max_abs'])
print(json.dumps(d.get('recon'))[:1400]); print(d.get('configs4_dense513'))
PY
timeout 120 python tools/recon_trace.py --color 2>&1 | grep -v Warn | head -16 | cut -c1-120
