#!/bin/bash
# session 2: warp-specialised kernel -- which role limits?  (probes: 1 = no arithmetic, 2 = no gather/scatter traffic)
mkdir -p gpurun_out; : > gpurun_out/s2_ws2.jsonl
run() { env "$@" timeout 200 python tools/time_action.py 2>&1 | tail -1 | tee -a gpurun_out/s2_ws2.jsonl; }
run FDB_WS=3
run FDB_WS=4
run FDB_WS=3 FDB_WS_PROBE=1
run FDB_WS=3 FDB_WS_PROBE=2
run FDB_WS=3 FDB_WS_PROBE=3
run FDB_WS=1 FDB_WS_PROBE=2
FDB_WS=3 timeout 300 ncu --set full --import-source on --clock-control none -k regex:helmholtz_action_ws -s 2 -c 1 \
  -o gpurun_out/s2_ws3 -f python tools/time_action.py --n 96 --steps 2 2>&1 | tail -3
ls -la gpurun_out/*.ncu-rep
