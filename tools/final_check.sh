#!/bin/bash
# end-of-session validation on one GPU: full GPU test suite, smoke, [short ncu pass <tag>], default bench (tools/README.md)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=5 > gpurun_out/pytest_gpu_final.log 2>&1; tail -12 gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
if [ -n "$1" ]; then bash tools/prof_min.sh $1 | tail -8; fi
python bench.py > gpurun_out/bench_1gpu_final.json 2> gpurun_out/bench_1gpu_final.err; tail -c 600 gpurun_out/bench_1gpu_final.json; tail -2 gpurun_out/bench_1gpu_final.err
