#!/bin/bash
N=$1
for H in exec sum; do
FDB_PHASE_TIMING=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 benchmarks/cg_multi.py --size 128 --degree 5 --halo $H --iters 10 2>&1 | grep -E "phase timing|config5" | cut -c1-420
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 benchmarks/cg_multi.py --size 128 --degree 5 --halo $H 2>/dev/null | tail -1 > gpurun_out/r2_cg_${N}_$H.json; cut -c1-300 gpurun_out/r2_cg_${N}_$H.json
done
