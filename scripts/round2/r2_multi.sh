#!/bin/bash
# usage: r2_multi.sh N   (under gpurun --gpus N)
N=$1
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_halo_gpu.py -q -p no:cacheprovider 2>&1 | tail -4
for H in exec sum; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --halo $H 2>gpurun_out/r2_scale_${N}_$H.err | tail -1 > gpurun_out/r2_scale_${N}_$H.json
python -c "
import json; d=json.load(open('gpurun_out/r2_scale_${N}_$H.json')); print($N, '$H', 'ms/step', d['ms_per_step'], 'value', d['value'], 'kernel_ms', d['roofline']['kernel_ms'], 'e2e ms', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'], 'parity', d.get('parity',{}).get('rel_err'))" || tail -5 gpurun_out/r2_scale_${N}_$H.err
done
for H in exec sum; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 benchmarks/cg_multi.py --size 128 --degree 5 --halo $H 2>gpurun_out/r2_cg_${N}_$H.err | tail -1 > gpurun_out/r2_cg_${N}_$H.json
cut -c1-330 gpurun_out/r2_cg_${N}_$H.json; tail -2 gpurun_out/r2_cg_${N}_$H.err
done
