#!/bin/bash
timeout 600 python -m pytest tests/test_halo_gpu.py -q -p no:cacheprovider -k "action" 2>&1 | tail -12
timeout 600 python -m pytest tests/test_action_gpu.py -q -m gpu -p no:cacheprovider -k "generation or host_pointer" 2>&1 | tail -5
