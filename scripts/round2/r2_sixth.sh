#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_patch_asm_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -8
for c in 0 4 8 12 16; do
FDB_CHUNK=$c timeout 200 python bench.py --n 128 --steps 20 --warmup 5 --no-cpu --no-e2e --no-parity 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('n128 chunk $c ms_per_step %.4f kernel_ms %.4f' % (d['ms_per_step'], d['roofline']['kernel_ms']))"
done
for c in 0 8 16; do
FDB_CHUNK=$c timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e --no-parity 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('n256 chunk $c ms_per_step %.4f kernel_ms %.4f' % (d['ms_per_step'], d['roofline']['kernel_ms']))"
done
