#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -q -m gpu -p no:cacheprovider 2>&1 | tee gpurun_out/r2_final2_suite.log | tail -5
python -c "import __graft_entry__ as g; g.smoke()"
