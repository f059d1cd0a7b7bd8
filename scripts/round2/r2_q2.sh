#!/bin/bash
timeout 600 python -m pytest tests/test_action_gpu.py tests/test_assemble_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -4
timeout 300 python benchmarks/run_configs.py --only "Poisson CG2 action" 2>&1 | cut -c1-260
FDB_Q2_THREAD=0 timeout 300 python benchmarks/run_configs.py --only "Poisson CG2 action" 2>&1 | cut -c1-260
