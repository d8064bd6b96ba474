#!/bin/bash
REPS=4 python tools/perf_probe.py 2048 16384 2>&1 | grep -v "nlml only" | cut -c1-140
python tools/batch_probe.py uniform 2>&1 | cut -c1-190 | grep "16 x 2048\|16 x 7680"
python -m pytest tests/test_batch.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
