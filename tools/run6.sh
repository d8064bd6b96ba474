#!/bin/bash
# round-2 ncu session: single-site evidence (prof.sh) + per-launch list of one batched evaluation (config-4 group 0)
mkdir -p gpurun_out
bash tools/prof.sh r02a > gpurun_out/prof_r02a.log 2>&1; tail -3 gpurun_out/prof_r02a.log
DGP_RASTER=0 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file gpurun_out/dram_n16384_r02a_raster0.csv python tools/one_nlml.py 16384 grad > gpurun_out/ncu_dram_r0.log 2>&1
python tools/one_batch.py 0 2 > gpurun_out/one_batch_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --launch-skip 240 -c 300 --csv --log-file gpurun_out/batch_g0_launches_r02a.csv python tools/one_batch.py 0 2 > gpurun_out/ncu_batch.log 2>&1
tail -2 gpurun_out/ncu_batch.log; ls -la gpurun_out | tail -30
