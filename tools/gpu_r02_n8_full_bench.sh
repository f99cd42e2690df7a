#!/bin/bash
# Round 2, 8 GPUs, final code: the driver-style scaling bench line at N = 8 with every section (dense + recon + list-sharded engine
# + frame-parallel stream + 513^3).
mkdir -p gpurun_out
T0=$SECONDS
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 \
  > gpurun_out/r02_final_bench_n8.raw 2> gpurun_out/r02_final_bench_n8.err; echo "bench N=8 rc=$? t=$((SECONDS-T0))s"
grep '^{' gpurun_out/r02_final_bench_n8.raw | tail -1 > gpurun_out/r02_final_bench_n8.json
python -c "
import json; d=json.load(open('gpurun_out/r02_final_bench_n8.json'))
print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), 'vol ok', d['volume_matches_single_gpu'], d['config'].get('exchange'))
print(json.dumps(d.get('recon'))[:1800]); print(d.get('configs4_dense513'))" || tail -5 gpurun_out/r02_final_bench_n8.err
