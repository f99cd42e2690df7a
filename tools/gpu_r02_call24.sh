#!/bin/bash
# Round-2 call 24 (1 GPU): the fp32 tail split over both warpgroups (TMEM released after the loads) -- parity, then a same-box
# A/B against the previous commit's library (dense 257^3 and one coarse-to-fine frame).
mkdir -p gpurun_out
T0=$SECONDS
timeout 300 python -m pytest tests/test_query_gpu.py tests/test_engine_gpu.py -x -q -m gpu --timeout 200 > gpurun_out/r02c24_pytest.log 2>&1; echo "pytest rc=$? t=$((SECONDS-T0))s"; tail -2 gpurun_out/r02c24_pytest.log
L=monoport_b200/lib
cp $L/libmonoport_b200.so $L/_cur.so
{
for rep in 1 2; do
  echo "== tail over both warpgroups"; cp $L/_cur.so $L/libmonoport_b200.so
  timeout -k 5 120 python tools/tc_prof.py 257 2>&1 | grep -E "ms per volume"
  echo "== previous commit"; cp $L/libmonoport_b200_prev.so $L/libmonoport_b200.so
  timeout -k 5 120 python tools/tc_prof.py 257 2>&1 | grep -E "ms per volume"
done
echo "== previous commit: frame"; timeout 120 python tools/recon_trace.py 2>&1 | grep -E "GPU busy|query_tc3_kernel" | head -3 | cut -c1-120
cp $L/_cur.so $L/libmonoport_b200.so
echo "== tail over both warpgroups: frame"; timeout 120 python tools/recon_trace.py 2>&1 | grep -E "GPU busy|query_tc3_kernel" | head -3 | cut -c1-120
echo "== in-kernel attribution (PROF instantiation)"
MONOPORT_B200_TC_PROF=1 timeout -k 5 120 python tools/tc_prof.py 257 2>&1 | grep -E "tc prof\]" | head -21
} 2>&1 | tee gpurun_out/r02c24_tail_split_ab.txt
