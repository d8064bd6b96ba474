#!/bin/bash
# round-2 session b: column strips on two streams (DGP_STRIP_BLOCKS) against the single trailing stream
mkdir -p gpurun_out
python -m pytest tests/test_batch.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_strips.log 2>&1; tail -3 gpurun_out/pytest_strips.log
for sb in 0 16 8 32; do
  echo "== DGP_STRIP_BLOCKS=$sb"
  DGP_STRIP_BLOCKS=$sb REPS=4 python tools/perf_probe.py 4096 8192 16384 2>&1 | grep -v "nlml only"
done
for sb in 0 16; do
  echo "== batch DGP_STRIP_BLOCKS=$sb"
  DGP_STRIP_BLOCKS=$sb python tools/batch_probe.py uniform 2>&1
done
