#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/shard_check.py 2>&1 | grep -E "shard_check|Error|rank 0" | head -5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "rc=$?"; cut -c1-420 gpurun_out/bench_n8.json; tail -2 gpurun_out/bench_n8.err | cut -c1-200
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 2>/dev/null | cut -c1-200
