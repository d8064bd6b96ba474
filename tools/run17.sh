#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu_r02b.log 2>&1; tail -14 gpurun_out/pytest_gpu_r02b.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_1gpu_r02b.json 2> gpurun_out/bench_1gpu_r02b.err; tail -c 3000 gpurun_out/bench_1gpu_r02b.json; tail -3 gpurun_out/bench_1gpu_r02b.err
