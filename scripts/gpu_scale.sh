#!/bin/bash
# usage: gpu_scale.sh N   (run under gpurun --gpus N)
N=$1
mkdir -p gpurun_out
nvidia-smi -L | head -8
if [ "$N" -ge 2 ]; then python -m pytest tests/test_halo_gpu.py -x -q 2>&1 | tail -3; fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu 2>gpurun_out/scale_$N.err | tail -1 > gpurun_out/scale_$N.json
python -c "
import json; d=json.load(open('gpurun_out/scale_$N.json')); print($N, 'ms/step', d['ms_per_step'], 'value', d['value'], 'kernel_ms', d['roofline']['kernel_ms'], 'e2e ms', d['e2e']['ms_per_step'], 'e2e', d['e2e']['value'], 'launches', d['gpu_launches'], d['clocks'])"
tail -2 gpurun_out/scale_$N.err
