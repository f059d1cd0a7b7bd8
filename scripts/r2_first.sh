#!/bin/bash
# Round-2 first GPU call: run the tests that were gated in round 1 (no -x), then the validated suite.
mkdir -p gpurun_out
export FDB_RUN_UNVALIDATED=1
timeout 900 python -m pytest tests/test_jit_gpu.py -q -m gpu -rA -p no:cacheprovider 2>&1 | tee gpurun_out/r2_first_jit.log | tail -60
timeout 600 python -m pytest tests/ -q -m gpu -p no:cacheprovider --deselect tests/test_jit_gpu.py 2>&1 | tee gpurun_out/r2_first_suite.log | tail -8
timeout 200 python bench.py --warp 0 --steps 5 --warmup 3 --no-cpu 2>/dev/null | tail -1 > gpurun_out/bench_unwarped_general.json
FDB_AFFINE=1 timeout 200 python bench.py --warp 0 --steps 5 --warmup 3 --no-cpu 2>/dev/null | tail -1 > gpurun_out/bench_unwarped_affine.json
cat gpurun_out/bench_unwarped_general.json | cut -c1-400
cat gpurun_out/bench_unwarped_affine.json | cut -c1-400
timeout 300 python benchmarks/mg_solve.py --coarse 16 --levels 3 --degree 3 > gpurun_out/mg_solve.json 2>&1; tail -3 gpurun_out/mg_solve.json
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv
