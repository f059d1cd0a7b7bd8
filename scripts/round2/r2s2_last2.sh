#!/bin/bash
# session 2, last GPU seconds: 16 warps per SM (128 registers + slim staging + parked coefficients)
mkdir -p gpurun_out
FDB_MINB=4 FDB_SLIM4=1 timeout 80 python tools/time_action.py --check 2>&1 | tail -1 | tee gpurun_out/s2_last2_slim4.json
