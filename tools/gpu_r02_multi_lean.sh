#!/bin/bash
# Round 2, lean multi-GPU call (gpurun --gpus N) after the issuer rewrite: sharded == single-GPU checks (dense, fused exchange,
# list-sharded octree engines), then the dense bench line with both exchanges (no recon section: N x box time).
# usage: bash tools/gpu_r02_multi_lean.sh N
N=${1:-2}
mkdir -p gpurun_out
T0=$SECONDS
if [ "$N" -le 2 ]; then
  timeout 400 python -m pytest tests/test_shard_multigpu.py -q -m gpu > gpurun_out/r02b_pytest_multigpu_n${N}.log 2>&1; echo "pytest multi-gpu rc=$? t=$((SECONDS-T0))s"; tail -3 gpurun_out/r02b_pytest_multigpu_n${N}.log
fi
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/shard_check.py --fused --octree 2>&1 | grep -E "rank 0|OK|Error|error" | tee gpurun_out/r02b_shard_check_n${N}.txt
for flag in "--exchange nccl" "--exchange fused"; do
  tag=$(echo $flag | awk '{print $2}')
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 --no-recon --no-cpu-baseline $flag \
    > gpurun_out/r02b_bench_n${N}_${tag}.raw 2> gpurun_out/r02b_bench_n${N}_${tag}.err; echo "bench N=$N $tag rc=$? t=$((SECONDS-T0))s"
  grep '^{' gpurun_out/r02b_bench_n${N}_${tag}.raw | tail -1 > gpurun_out/r02b_bench_n${N}_${tag}.json
  python -c "
import json; d=json.load(open('gpurun_out/r02b_bench_n${N}_${tag}.json'))
print('$tag', round(d['value'],1), round(d['ms_per_step'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3), 'e2e', round(d['e2e']['value'],1), 'vol ok', d['volume_matches_single_gpu'], 'parity', d['parity_max_abs'])
print(d.get('configs4_dense513'))" 2>/dev/null || tail -5 gpurun_out/r02b_bench_n${N}_${tag}.err
done
