#!/bin/bash
# session 2: register-relief variants of the fused headline kernel (tools/build_variant.sh), parity + timing
mkdir -p gpurun_out; : > gpurun_out/s2_variants.jsonl
V=firedrake_b200/lib/variants
timeout 200 python tools/time_action.py --check 2>&1 | tail -1 | tee -a gpurun_out/s2_variants.jsonl
for v in rcp3 stash stash_rcp3 stash_rcp3_cd1 all; do
  FDB200_LIB=$V/libfdb200_$v.so timeout 200 python tools/time_action.py --check 2>&1 | tail -1 | tee -a gpurun_out/s2_variants.jsonl
done
