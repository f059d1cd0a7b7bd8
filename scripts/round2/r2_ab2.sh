#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift
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
run norot3 FDB_CD1=0 FDB_NOROT=3
run norot4 FDB_CD1=0 FDB_NOROT=4
run norot5 FDB_CD1=0 FDB_NOROT=5
FDB_NOROT=3 timeout 300 python -m pytest tests/test_action_gpu.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -2
timeout 600 python -m pytest tests/test_jit_gpu.py tests/test_matrix_gpu.py -q -m gpu -p no:cacheprovider -x -k "matrix" 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_matrix_launches_sym.csv python benchmarks/run_configs.py --only "blocked_matrix_case" > gpurun_out/r2_matrix_sym.jsonl 2>&1
python - <<'PY'
import csv,collections
t=collections.defaultdict(lambda:[0,0.0])
rows=[r for r in csv.reader(open("gpurun_out/r2_matrix_launches_sym.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
for r in rows[1:]:
    try: v=float(r[vi].replace(",",""))
    except: continue
    t[r[ki][:70]][0]+=1; t[r[ki][:70]][1]+=v
for k,(n,v) in sorted(t.items(), key=lambda x:-x[1][1])[:5]: print("  %-70s n=%d total=%.3f ms"%(k,n,v/1e6))
PY
