#!/bin/bash
N=$1
mkdir -p gpurun_out
python -m pytest tests/test_halo_gpu.py -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 benchmarks/cg_multi.py --n 128 --degree 5 2>/dev/null | tail -1 | tee gpurun_out/cg_multi_$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29642 benchmarks/cg_multi.py --n 256 --degree 3 2>/dev/null | tail -1 | tee -a gpurun_out/cg_multi_$N.json
