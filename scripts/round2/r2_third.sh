#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_action_gpu.py tests/test_full_size_gpu.py tests/test_assemble_gpu.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -5
for v in 1 0; do
FDB_NOROT=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-parity > gpurun_out/r2_bench_norot$v.json 2>/dev/null
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_norot$v.json").read().strip().splitlines()[-1])
print("NOROT=$v ms_per_step", d["ms_per_step"], "kernel_ms", d["roofline"]["kernel_ms"])
PY
done
timeout 300 python benchmarks/run_configs.py --only "action_case" 2>&1 | cut -c1-260
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_matrix_launches_dmma.csv python benchmarks/run_configs.py --only "blocked_matrix_case" > /dev/null 2>&1
FDB_MATRIX_DMMA=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_matrix_launches_old.csv python benchmarks/run_configs.py --only "blocked_matrix_case" > /dev/null 2>&1
python - <<'PY'
import csv,collections
for f in ("dmma","old"):
    t=collections.defaultdict(lambda:[0,0.0])
    try:
        rows=[r for r in csv.reader(open(f"gpurun_out/r2_matrix_launches_{f}.csv")) if len(r)>10]
        hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
        for r in rows[1:]:
            try: v=float(r[vi].replace(",",""))
            except: continue
            t[r[ki][:70]][0]+=1; t[r[ki][:70]][1]+=v
        print(f)
        for k,(n,v) in sorted(t.items(), key=lambda x:-x[1][1])[:6]: print("  %-70s n=%d total=%.3f ms"%(k,n,v/1e6))
    except Exception as e: print(f,"failed",e)
PY
