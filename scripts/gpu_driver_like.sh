#!/bin/bash
# what the driver runs at round end, in order
mkdir -p gpurun_out
python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python bench.py --impl reference --gpus 1 --steps 10 --warmup 3 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 900 gpurun_out/bench_ref.json; tail -4 gpurun_out/bench_ref.err
( time python bench.py --gpus 1 --steps 10 --warmup 3 ) > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -c 3500 gpurun_out/bench_ours.json; tail -4 gpurun_out/bench_ours.err
