#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python benchmarks/run_configs.py 2>/dev/null | grep -E "CG5|CG4|CG3 action, lex" | cut -c1-300
FDB_NO_SLIM=1 python benchmarks/run_configs.py 2>/dev/null | grep -E "CG5 action|CG4" | cut -c1-300
