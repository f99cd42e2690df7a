#!/bin/bash
# Round-2 call 19 (1 GPU): the geometry kernel with ONE elected region per chunk / phase in its MMA issuer (and no profiling
# code in the shipped instantiation) -- parity, then a same-box A/B against the previous commit's library.
mkdir -p gpurun_out
T0=$SECONDS
timeout 300 python -m pytest tests/test_query_gpu.py tests/test_engine_gpu.py -x -q -m gpu --timeout 200 > gpurun_out/r02c19_pytest.log 2>&1; echo "pytest rc=$? t=$((SECONDS-T0))s"; tail -2 gpurun_out/r02c19_pytest.log
L=monoport_b200/lib
cp $L/libmonoport_b200.so $L/_cur.so
{
for rep in 1 2; do
  echo "== new issuer"; cp $L/_cur.so $L/libmonoport_b200.so
  timeout -k 5 120 python tools/tc_prof.py 257 2>&1 | grep -E "ms per volume"
  echo "== previous commit"; cp $L/libmonoport_b200_prev.so $L/libmonoport_b200.so
  timeout -k 5 120 python tools/tc_prof.py 257 2>&1 | grep -E "ms per volume"
done
cp $L/_cur.so $L/libmonoport_b200.so
echo "== new issuer, in-kernel attribution (PROF instantiation)"
MONOPORT_B200_TC_PROF=1 timeout -k 5 120 python tools/tc_prof.py 257 2>&1 | grep -E "tc prof\]" | head -21
} 2>&1 | tee gpurun_out/r02c19_issuer_ab.txt
timeout 300 python bench.py --no-recon --no-cpu-baseline > gpurun_out/r02c19_bench_dense.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02c19_bench_dense.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['clocks'])"
