#!/bin/bash
# Round-2 call 15 (1 GPU): marching cubes with the queue-based emission and classify fused with the chunk totals (parity +
# per-kernel timings), the default bench's dense line (e2e now reads the volume back in overlapped chunks).
mkdir -p gpurun_out
T0=$SECONDS
timeout 300 python -m pytest tests/test_engine_gpu.py tests/test_query_gpu.py -x -q -m gpu --timeout 200 > gpurun_out/r02c15_pytest.log 2>&1; echo "pytest rc=$? t=$((SECONDS-T0))s"; tail -2 gpurun_out/r02c15_pytest.log
timeout 120 python tools/recon_trace.py --mc 2>&1 | grep -v Warn > gpurun_out/r02c15_recon_trace_fv_mc.txt; grep -E "mcubes|first_hit|HitF|per frame" gpurun_out/r02c15_recon_trace_fv_mc.txt | head -16 | cut -c1-120
timeout 300 python bench.py --no-recon --no-cpu-baseline > gpurun_out/r02c15_bench_dense.json 2> gpurun_out/r02c15_bench.err; echo "bench rc=$? t=$((SECONDS-T0))s"
python - <<'PY'
import json
d = json.load(open('gpurun_out/r02c15_bench_dense.json'))
print(d['value'], d['ms_per_step'], d['e2e'], d['clocks'])
PY
