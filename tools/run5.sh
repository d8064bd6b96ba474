#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_batch.py tests/test_multisite.py -m gpu -x -q --timeout 300 2>&1 | tail -3
python tools/batch_probe.py all > gpurun_out/probe_left1.log 2>&1; tail -17 gpurun_out/probe_left1.log
for v in 0 1; do DGP_INPANEL_LEFT=$v python bench.py --no-extra --no-cpu > gpurun_out/bench_left$v.json 2>gpurun_out/bench_left$v.err; python -c "
import json; d=json.load(open('gpurun_out/bench_left$v.json')); print('single-site inpanel_left=$v', d['ms_per_step'], d['roofline']['phases_ms'])"; done
DGP_INPANEL_LEFT=1 python tools/batch_probe.py single 2>&1 | tail -5
DGP_INPANEL_LEFT=0 python tools/batch_probe.py single 2>&1 | tail -5
DGP_INPANEL_LEFT=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 300 2>&1 | tail -3
