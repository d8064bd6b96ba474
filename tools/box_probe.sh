#!/bin/bash
# first call of the round: what the box has (cores, RAM), the GPU test suite, oracle timings at the BASELINE sizes
mkdir -p gpurun_out
{ nproc; free -g; nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv; python -c "import importlib.util as u; print('gpytorch', u.find_spec('gpytorch'))"; } > gpurun_out/box.txt 2>&1
python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/box.txt
python tools/oracle_time.py 4096 ag >> gpurun_out/box.txt 2>&1
python tools/oracle_time.py 8192 ag >> gpurun_out/box.txt 2>&1
python tools/oracle_time.py 8192 cf >> gpurun_out/box.txt 2>&1
python tools/oracle_time.py 16384 ag >> gpurun_out/box.txt 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/box.txt
