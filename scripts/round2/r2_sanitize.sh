#!/bin/bash
mkdir -p gpurun_out
export FDB_SANITIZE=1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_action_gpu.py tests/test_dg_advection_gpu.py tests/test_patch_asm_gpu.py -q -m gpu -p no:cacheprovider -x -k "thread_per_cell or dg or patch or matches_oracle" > gpurun_out/r2_memcheck1.log 2>&1; echo "memcheck1 rc=$?"; tail -4 gpurun_out/r2_memcheck1.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_matrix_gpu.py -q -m gpu -p no:cacheprovider -x -k "matrix_matches_oracle or bc_lgmaps" > gpurun_out/r2_memcheck2.log 2>&1; echo "memcheck2 rc=$?"; tail -4 gpurun_out/r2_memcheck2.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_matrix_gpu.py tests/test_patch_asm_gpu.py -q -m gpu -p no:cacheprovider -x -k "dmma-4 or dmma-3 or patch_inverses" > gpurun_out/r2_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/r2_racecheck.log
