#!/bin/bash
timeout 900 python -m pytest tests/test_matrix_gpu.py tests/test_jit_gpu.py -q -m gpu -p no:cacheprovider -k "matrix" 2>&1 | tail -3
timeout 300 python benchmarks/run_configs.py --only "blocked_matrix_case" 2>&1 | cut -c1-300
python - <<'PY'
import sys, json
sys.path.insert(0, ".")
from firedrake_b200 import _lib
from benchmarks.run_configs import matrix_case
L = _lib.init(0)
for n, p in ((64, 3), (24, 4)):
    r = matrix_case(f"Poisson CG{p} matrix", n, p)
    print(json.dumps({k: r[k] for k in ("case", "n", "assemble_ms", "nnz")}), flush=True)
    L.fdb_mirror_drop_all()
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_matrix_launches_sym2.csv python benchmarks/run_configs.py --only "blocked_matrix_case" > /dev/null 2>&1
python - <<'PY'
import csv,collections
t=collections.defaultdict(lambda:[0,0.0])
rows=[r for r in csv.reader(open("gpurun_out/r2_matrix_launches_sym2.csv")) if len(r)>10]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
for r in rows[1:]:
    try: v=float(r[vi].replace(",",""))
    except: continue
    t[r[ki][:70]][0]+=1; t[r[ki][:70]][1]+=v
for k,(n,v) in sorted(t.items(), key=lambda x:-x[1][1])[:5]: print("  %-70s n=%d total=%.3f ms"%(k,n,v/1e6))
PY
