#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 tools/shard_check.py > gpurun_out/shard_check4.log 2>&1; echo "rc=$?" >> gpurun_out/shard_check4.log; tail -4 gpurun_out/shard_check4.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "rc=$?"; cut -c1-700 gpurun_out/bench_n4.json; tail -2 gpurun_out/bench_n4.err
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'], d['recon'])"
