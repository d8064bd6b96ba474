#!/bin/bash
# timing experiment: what the panel chain costs the trailing updates (results of DGP_SKIP runs are garbage by construction)
for sk in 0 7 1 2 4 6; do
  for sb in 0 16; do
    echo "== DGP_SKIP=$sk DGP_STRIP_BLOCKS=$sb"
    DGP_SKIP=$sk DGP_STRIP_BLOCKS=$sb REPS=3 python tools/perf_probe.py 16384 2>&1 | grep -v "nlml only" | cut -c1-160
  done
done
echo "== inpanel_left=1"; DGP_INPANEL_LEFT=1 REPS=3 python tools/perf_probe.py 8192 16384 2>&1 | grep -v "nlml only" | cut -c1-160
