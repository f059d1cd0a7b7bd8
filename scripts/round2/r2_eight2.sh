#!/bin/bash
mkdir -p gpurun_out
N=8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --halo exec 2>gpurun_out/r2b_scale_${N}_exec.err | tail -1 > gpurun_out/r2b_scale_${N}_exec.json
python -c "
import json; d=json.load(open('gpurun_out/r2b_scale_${N}_exec.json')); print($N, 'exec ms/step', d['ms_per_step'], 'value', d['value'], 'kernel_ms', d['roofline']['kernel_ms'], 'e2e ms', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'], 'parity', d.get('parity',{}).get('rel_err'))" || tail -15 gpurun_out/r2b_scale_${N}_exec.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29622 bench.py --impl reference --gpus $N --steps 20 --warmup 5 2>/dev/null | tail -1 | cut -c1-300
