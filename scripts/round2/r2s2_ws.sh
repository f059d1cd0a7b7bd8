#!/bin/bash
# session 2: warp-specialised action kernel -- parity, sanitizer, timing
mkdir -p gpurun_out; : > gpurun_out/s2_ws.jsonl
for ws in 1 2; do
  FDB_WS=$ws timeout 120 python tools/time_action.py --check-only 2>&1 | tail -2 | tee -a gpurun_out/s2_ws.jsonl
done
FDB_WS=1 timeout 200 compute-sanitizer --tool memcheck python tools/time_action.py --check-only 2>&1 | tail -6 | tee gpurun_out/s2_ws_memcheck.txt
FDB_WS=1 timeout 240 compute-sanitizer --tool racecheck python tools/time_action.py --check-only 2>&1 | tail -6 | tee gpurun_out/s2_ws_racecheck.txt
for ws in 0 1 2; do
  FDB_WS=$ws timeout 200 python tools/time_action.py 2>&1 | tail -1 | tee -a gpurun_out/s2_ws.jsonl
done
FDB_WS=1 timeout 300 python -m pytest tests/test_action_gpu.py tests/test_full_size_gpu.py -x -q 2>&1 | tail -3 | tee gpurun_out/s2_ws_tests.txt
