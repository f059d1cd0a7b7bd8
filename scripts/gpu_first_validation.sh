#!/bin/bash
# First GPU call of the next round: run the GPU tests that were written after round 1's GPU
# budget was spent (generic NVRTC wrapper builder, blocked matrices, interpolation, 0-forms,
# multigrid), without -x so that one failure does not hide the others,
# then one case under compute-sanitizer (races in the generated atomics / shuffles show here).
#   gpurun --timeout 1800 -- 'bash scripts/gpu_first_validation.sh'   (about 15-20 GPU-minutes)
mkdir -p gpurun_out
export FDB_RUN_UNVALIDATED=1
python -m pytest tests/test_jit_gpu.py -q -m gpu -rA 2>&1 | tee gpurun_out/first_validation.log | tail -40
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_jit_gpu.py -q -m gpu \
    -k "access_modes or golden" > gpurun_out/first_validation_memcheck.log 2>&1
tail -5 gpurun_out/first_validation_memcheck.log
timeout 300 compute-sanitizer --tool racecheck python -m pytest tests/test_jit_gpu.py -q -m gpu \
    -k "access_modes" > gpurun_out/first_validation_racecheck.log 2>&1
tail -5 gpurun_out/first_validation_racecheck.log
# the validated suite must be unaffected
python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
# affine-cell kernel variant (opt-in): un-warped 256^3 CG3 with and without it
python bench.py --warp 0 --steps 5 --warmup 3 --no-cpu 2>/dev/null | tail -1 > gpurun_out/bench_unwarped_general.json
FDB_AFFINE=1 python bench.py --warp 0 --steps 5 --warmup 3 --no-cpu 2>/dev/null | tail -1 > gpurun_out/bench_unwarped_affine.json
python - <<'PY'
import json
for f in ("general", "affine"):
    try:
        d = json.loads(open(f"gpurun_out/bench_unwarped_{f}.json").read())
        print(f, d["ms_per_step"], d["config"].get("kernel_variant"))
    except Exception as e:
        print(f, "failed", e)
PY
FDB_RUN_UNVALIDATED=1 python benchmarks/run_configs.py --unvalidated > gpurun_out/run_configs_unvalidated.jsonl 2>&1; tail -4 gpurun_out/run_configs_unvalidated.jsonl
python benchmarks/mg_solve.py --coarse 16 --levels 3 --degree 3 > gpurun_out/mg_solve.json 2>&1; tail -1 gpurun_out/mg_solve.json
