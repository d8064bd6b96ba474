#!/bin/bash
# per-launch list of one evaluation (single stream semantics irrelevant: ncu serialises)
DGP_STRIP_BLOCKS=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_one_r02b.csv python tools/one_nlml.py 16384 grad > gpurun_out/ncu_one.log 2>&1
tail -2 gpurun_out/ncu_one.log
