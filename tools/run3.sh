#!/bin/bash
mkdir -p gpurun_out
python tools/batch_probe.py all > gpurun_out/probe_raster1.log 2>&1; tail -25 gpurun_out/probe_raster1.log
DGP_RASTER=0 python tools/batch_probe.py uniform > gpurun_out/probe_raster0.log 2>&1; tail -8 gpurun_out/probe_raster0.log
python bench.py --no-extra --no-cpu > gpurun_out/bench_raster1.json 2>gpurun_out/bench_raster1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_raster1.json')); print('raster1', d['ms_per_step'], d['roofline']['phases_ms'])"
DGP_RASTER=0 python bench.py --no-extra --no-cpu > gpurun_out/bench_raster0.json 2>gpurun_out/bench_raster0.err; python -c "
import json; d=json.load(open('gpurun_out/bench_raster0.json')); print('raster0', d['ms_per_step'], d['roofline']['phases_ms'])"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_batch.py -m gpu -x -q --timeout 300 2>&1 | tail -3
