#!/bin/bash
# First GPU call of round 2 (1 GPU): validate what round 1 wrote against the CPU model only.
#   1. default parity suite (must stay green)
#   2. the same suite with the tcgen05 program of the colour head switched on (MONOPORT_B200_TC_NETC=1)
#   3. default bench line, then the configs[2] colour frame rate with the tensor-core colour head
#   4. per-kernel timeline of a frame (marching-cubes rewrite, scan tail)
# Second call (gpurun --gpus 2): MONOPORT_B200_TEST_FUSED=1 python -m pytest tests/test_shard_multigpu.py -q ;
#   torchrun ... bench.py --gpus 2 [--fused-gather]  (A/B of the fused slab exchange against the NCCL all-gather)
mkdir -p gpurun_out
T0=$SECONDS
timeout 300 python -m pytest tests -x -q -m gpu --timeout 200 > gpurun_out/r02_pytest_default.log 2>&1; echo "pytest default rc=$? t=$((SECONDS-T0))s"; tail -2 gpurun_out/r02_pytest_default.log
MONOPORT_B200_TC_NETC=1 timeout 300 python -m pytest tests -q -m gpu --timeout 200 > gpurun_out/r02_pytest_tc_netc.log 2>&1; echo "pytest tc netC rc=$? t=$((SECONDS-T0))s"; tail -5 gpurun_out/r02_pytest_tc_netc.log
timeout 300 python bench.py > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err; echo "bench rc=$? t=$((SECONDS-T0))s"
MONOPORT_B200_TC_NETC=1 timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r02_bench_tc_netc.json 2> gpurun_out/r02_bench_tc_netc.err; echo "bench tc netC rc=$? t=$((SECONDS-T0))s"
# A/B: brick-ordered dense grid (8x4x4 bricks per tile: layer-0 taps hit L1) against the row order, same box
MONOPORT_B200_GRID_BRICK=1 timeout 200 python bench.py --no-recon --no-cpu-baseline > gpurun_out/r02_bench_brick.json 2> gpurun_out/r02_bench_brick.err; echo "bench brick rc=$? t=$((SECONDS-T0))s"
MONOPORT_B200_GRID_BRICK=1 timeout 200 python -m pytest tests/test_query_gpu.py tests/test_engine_gpu.py -q -m gpu -k "grid or full_size or reconstruction or sharded" > gpurun_out/r02_pytest_brick.log 2>&1; tail -2 gpurun_out/r02_pytest_brick.log
python - <<'PY'
import json
for f in ("r02_bench_default", "r02_bench_tc_netc", "r02_bench_brick"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        r = d.get("recon") or {}
        print(f, d["value"], d["ms_per_step"], d["roofline"]["frac"], "fv fps", r.get("frames_per_s_with_forward_vertices"), "mc fps",
              r.get("frames_per_s_with_marching_cubes"), "colour fps", r.get("frames_per_s_geometry_plus_netC_colour"))
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 120 python tools/recon_trace.py --mc 2>&1 | grep -v Warn > gpurun_out/r02_recon_trace_mc.txt; head -26 gpurun_out/r02_recon_trace_mc.txt
# A/B: classify with the uniform-chunk fast path (MONOPORT_B200_MC_FAST=1), parity first
MONOPORT_B200_MC_FAST=1 timeout 120 python -m pytest tests/test_engine_gpu.py -q -m gpu -k "marching or reconstruction" > gpurun_out/r02_pytest_mc_fast.log 2>&1; tail -2 gpurun_out/r02_pytest_mc_fast.log
MONOPORT_B200_MC_FAST=1 timeout 120 python tools/recon_trace.py --mc 2>&1 | grep -v Warn | grep -E "classify|wall" > gpurun_out/r02_recon_trace_mc_fast.txt; cat gpurun_out/r02_recon_trace_mc_fast.txt
MONOPORT_B200_TC_NETC=1 timeout 120 python tools/recon_trace.py --color 2>&1 | grep -v Warn > gpurun_out/r02_recon_trace_color_tc.txt; head -12 gpurun_out/r02_recon_trace_color_tc.txt
# where a dense tile's 66 k cycles go in the FINAL program v3 (in-kernel clock64 attribution + one traced tile): decides which of
# the ideas of DESIGN.md §5 to try first
MONOPORT_B200_TC_PROF=1 timeout 120 python tools/tc_prof.py 257 2>&1 | grep "tc prof" > gpurun_out/r02_tc_v3_final_inkernel_cycles.txt; cat gpurun_out/r02_tc_v3_final_inkernel_cycles.txt
MONOPORT_B200_TC_TRACE=1 timeout 120 python tools/tc_prof.py 257 2>&1 | grep "tc trace" > gpurun_out/r02_tc_v3_final_trace_tile8.txt; wc -l gpurun_out/r02_tc_v3_final_trace_tile8.txt
MONOPORT_B200_GRID_BRICK=1 MONOPORT_B200_TC_PROF=1 timeout 120 python tools/tc_prof.py 257 2>&1 | grep "tc prof" > gpurun_out/r02_tc_v3_brick_inkernel_cycles.txt; grep -E "total|h0ready|wfull" gpurun_out/r02_tc_v3_brick_inkernel_cycles.txt
