#!/bin/bash
N=$1
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 benchmarks/cg_multi.py --size 128 --degree 5 > gpurun_out/cg_multi_$N.json 2> gpurun_out/cg_multi_$N.err
tail -1 gpurun_out/cg_multi_$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29643 benchmarks/cg_multi.py --size 128 --degree 4 2>/dev/null | tail -1 | tee -a gpurun_out/cg_multi_$N.json
python benchmarks/cg_multi.py --size 128 --degree 5 2>/dev/null | tail -1 | tee -a gpurun_out/cg_multi_$N.json
