#!/bin/bash
for il in 1 0; do
  echo "== batch DGP_INPANEL_LEFT=$il"
  DGP_INPANEL_LEFT=$il python tools/batch_probe.py uniform config4 2>&1 | cut -c1-200
done
