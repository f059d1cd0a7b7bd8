#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python benchmarks/run_configs.py 2>/dev/null | grep -E "DG" | cut -c1-300
