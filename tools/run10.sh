#!/bin/bash
for st in 0 8000 16000 32000; do
  echo "== DGP_STAGGER_NS=$st"
  DGP_STAGGER_NS=$st python tools/tile_probe.py 2>&1
done
for st in 0 12000 24000; do
  echo "== DGP_STAGGER_NS=$st DGP_EAGER_INV=0"
  DGP_EAGER_INV=0 DGP_STAGGER_NS=$st REPS=4 python tools/perf_probe.py 8192 16384 2>&1 | grep -v "nlml only" | cut -c1-120
done
