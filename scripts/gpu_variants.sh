#!/bin/bash
for v in "" _w3r224 _w2r200 _w1r184 _w3r168 _w1r152; do
  lib=/root/repo/firedrake_b200/lib/libfdb200$v.so
  echo "== $v"
  FDB200_LIB=$lib python bench.py --n 256 --steps 5 --warmup 3 --no-e2e --no-cpu 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['kernel_ms'], d['value'])"
done
