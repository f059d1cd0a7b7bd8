#!/bin/bash
for v in "" _w4r168 _w2r168 _w3r152; do
  lib=/root/repo/firedrake_b200/lib/libfdb200$v.so
  echo "== $v"
  FDB200_LIB=$lib python bench.py --n 256 --steps 5 --warmup 3 --no-e2e --no-cpu 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['value'])"
done
