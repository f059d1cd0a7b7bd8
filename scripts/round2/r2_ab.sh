#!/bin/bash
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-parity > gpurun_out/r2_ab_$name.json 2>gpurun_out/r2_ab_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_ab_$name.json").read().strip().splitlines()[-1])
    print("$name ms_per_step %.3f kernel_ms %.3f" % (d["ms_per_step"], d["roofline"]["kernel_ms"]))
except Exception as e:
    print("$name failed", e); print(open("gpurun_out/r2_ab_$name.err").read()[-500:])
PY
}
run base FDB_CD1=0
run cd1 FDB_CD1=1
run norot1 FDB_NOROT=1
run norot2 FDB_NOROT=2
timeout 300 python -m pytest tests/test_action_gpu.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -3
FDB_NOROT=1 timeout 300 python -m pytest tests/test_action_gpu.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -3
