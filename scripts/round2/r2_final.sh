#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu -p no:cacheprovider 2>&1 | tee gpurun_out/r2_final_suite.log | tail -5
python -c "import __graft_entry__ as g; g.smoke()"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_final_bench.json 2>gpurun_out/r2_final_bench.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_final_bench.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["roofline"]["frac"], d["roofline"].get("fp64_pipe_frac"), d["e2e"]["ms_per_step"], d["parity"]["rel_err"], d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"], d["clocks"])
PY
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 | cut -c1-400
