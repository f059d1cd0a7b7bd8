#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_halo_gpu.py -x -q 2>&1 | tail -15
for N in 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu 2>gpurun_out/multi_$N.err | tail -1 > gpurun_out/multi_$N.json
python -c "
import json; d=json.load(open('gpurun_out/multi_$N.json')); print($N, d['ms_per_step'], d['value'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
tail -3 gpurun_out/multi_$N.err
done
python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/multi_1.json; python -c "
import json; d=json.load(open('gpurun_out/multi_1.json')); print(1, d['ms_per_step'], d['value'], d['roofline']['kernel_ms'], d['e2e']['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
