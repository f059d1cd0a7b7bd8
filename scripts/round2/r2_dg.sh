#!/bin/bash
timeout 600 python -m pytest tests/test_dg_advection_gpu.py tests/test_full_size_gpu.py -q -m gpu -p no:cacheprovider -k "dg" 2>&1 | tail -3
timeout 300 python benchmarks/run_configs.py --only "dg_case" 2>&1 | cut -c1-300
