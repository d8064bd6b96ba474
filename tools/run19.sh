#!/bin/bash
for pb in 4 2 3 6; do
  echo "== DGP_PANEL_BLOCKS=$pb"
  DGP_PANEL_BLOCKS=$pb REPS=4 python tools/perf_probe.py 1024 2048 4096 8192 16384 2>&1 | grep -v "nlml only" | cut -c1-140
  DGP_PANEL_BLOCKS=$pb python tools/batch_probe.py uniform 2>&1 | cut -c1-190
done
