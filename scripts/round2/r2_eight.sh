#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_halo_gpu.py -q -p no:cacheprovider -k "4" 2>&1 | tail -3
N=8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --halo exec 2>gpurun_out/r2_scale_${N}_exec.err | tail -1 > gpurun_out/r2_scale_${N}_exec.json
python -c "
import json; d=json.load(open('gpurun_out/r2_scale_${N}_exec.json')); print($N, 'exec ms/step', d['ms_per_step'], 'value', d['value'], 'kernel_ms', d['roofline']['kernel_ms'], 'e2e ms', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'], 'parity', d.get('parity',{}).get('rel_err'))" || tail -5 gpurun_out/r2_scale_${N}_exec.err
for M in 8 4; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $M --master-addr 127.0.0.1 --master-port 29631 benchmarks/cg_multi.py --size 128 --degree 5 --halo exec 2>gpurun_out/r2_cg_${M}_exec.err | tail -1 > gpurun_out/r2_cg_${M}_exec.json
cut -c1-330 gpurun_out/r2_cg_${M}_exec.json; tail -2 gpurun_out/r2_cg_${M}_exec.err | cut -c1-300
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --no-parity --no-e2e --halo sum 2>/dev/null | tail -1 > gpurun_out/r2_scale_${N}_sum.json
python -c "
import json; d=json.load(open('gpurun_out/r2_scale_${N}_sum.json')); print($N, 'sum ms/step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'])"
