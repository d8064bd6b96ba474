#!/bin/bash
# round-2 session b: inverse merges launched as the factorisation passes them (DGP_EAGER_INV / DGP_EAGER_LAG)
mkdir -p gpurun_out
python -m pytest tests/test_batch.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_eager.log 2>&1; tail -3 gpurun_out/pytest_eager.log
for cfg in "0 0" "1 0" "1 8" "1 16" "1 32"; do
  set -- $cfg
  echo "== DGP_EAGER_INV=$1 DGP_EAGER_LAG=$2"
  DGP_EAGER_INV=$1 DGP_EAGER_LAG=$2 REPS=4 python tools/perf_probe.py 2048 4096 8192 16384 2>&1 | grep -v "nlml only" | cut -c1-100
done
for cfg in "0 0" "1 0" "1 16"; do
  set -- $cfg
  echo "== batch DGP_EAGER_INV=$1 DGP_EAGER_LAG=$2"
  DGP_EAGER_INV=$1 DGP_EAGER_LAG=$2 python tools/batch_probe.py uniform 2>&1 | cut -c1-110
done
