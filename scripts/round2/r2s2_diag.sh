#!/bin/bash
# session 2: where does the headline kernel's time go?  timing-only experiment builds (tools/build_variant.sh)
mkdir -p gpurun_out; : > gpurun_out/s2_diag.jsonl
V=firedrake_b200/lib/variants
for v in base noscatter nogather nocompute; do
  for mb in 3 2; do
    FDB_MINB=$mb FDB200_LIB=$V/libfdb200_$v.so timeout 300 python tools/time_action.py 2>&1 | tail -1 | tee -a gpurun_out/s2_diag.jsonl
  done
done
