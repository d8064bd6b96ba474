#!/bin/bash
for cfg in "1 1" "0 1" "1 0" "0 0"; do
  set -- $cfg
  echo "== DGP_INPANEL_LEFT=$1 DGP_EAGER_INV=$2"
  DGP_INPANEL_LEFT=$1 DGP_EAGER_INV=$2 REPS=5 python tools/perf_probe.py 1024 2048 4096 8192 2>&1 | grep -v "nlml only" | cut -c1-150
done
