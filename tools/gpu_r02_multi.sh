#!/bin/bash
# Round 2, multi-GPU call (gpurun --gpus 2, then --gpus 8): the fused slab exchange against the NCCL all-gather.
#   usage: bash tools/gpu_r02_multi.sh N
N=${1:-2}
mkdir -p gpurun_out
T0=$SECONDS
MONOPORT_B200_TEST_FUSED=1 timeout 400 python -m pytest tests/test_shard_multigpu.py -q -m gpu > gpurun_out/r02_pytest_multigpu.log 2>&1; echo "pytest multi-gpu rc=$? t=$((SECONDS-T0))s"; tail -3 gpurun_out/r02_pytest_multigpu.log
for flag in "" "--fused-gather"; do
  tag=$([ -z "$flag" ] && echo nccl || echo fused)
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 10 --warmup 3 $flag \
    > gpurun_out/r02_bench_n${N}_${tag}.json 2> gpurun_out/r02_bench_n${N}_${tag}.err; echo "bench N=$N $tag rc=$? t=$((SECONDS-T0))s"
  python -c "import json; d=json.load(open('gpurun_out/r02_bench_n${N}_${tag}.json')); print('$tag', d['value'], d['ms_per_step'], d['e2e']['value'])" 2>/dev/null || tail -3 gpurun_out/r02_bench_n${N}_${tag}.err
done
