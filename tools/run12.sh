#!/bin/bash
mkdir -p gpurun_out
python tools/gemm_check.py 2>&1 | tail -8
python tools/tile_probe.py 2>&1
echo "== 1 CTA / SM"; DGP_SMEM_PAD=20000 python tools/tile_probe.py 2>&1 | grep "K= 2048\|K=  512"
REPS=4 python tools/perf_probe.py 2048 4096 8192 16384 2>&1 | grep -v "nlml only" | cut -c1-200
python -m pytest tests/test_batch.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_roll.log 2>&1; tail -3 gpurun_out/pytest_roll.log
python tools/batch_probe.py uniform 2>&1 | cut -c1-200
