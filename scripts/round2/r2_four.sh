#!/bin/bash
N=4
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_halo_gpu.py -q -p no:cacheprovider -k "action" 2>&1 | tail -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --halo exec 2>gpurun_out/r2b_scale_${N}_exec.err | tail -1 > gpurun_out/r2b_scale_${N}_exec.json
python -c "
import json; d=json.load(open('gpurun_out/r2b_scale_${N}_exec.json')); print($N, 'exec ms/step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'e2e ms', d['e2e']['ms_per_step'], 'parity', d.get('parity',{}).get('rel_err'))" || tail -15 gpurun_out/r2b_scale_${N}_exec.err
