#!/bin/bash
# short ncu pass after a schedule change (run under gpurun, one GPU):  tools/prof_min.sh <tag>
#   --set full capture of the standalone covariance generator (largest strip), DRAM bytes and the launch list of one evaluation
TAG=${1:-rXX}
OUT=gpurun_out
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_cov_lower --launch-skip 1 --launch-count 1 \
    -o $OUT/prof_gen_$TAG -f python tools/one_nlml.py 16384 grad > $OUT/ncu_gen_$TAG.log 2>&1
timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file $OUT/dram_n16384_$TAG.csv \
    python tools/one_nlml.py 16384 grad > $OUT/ncu_dram_$TAG.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches_n16384_$TAG.csv \
    python bench.py --steps 1 --warmup 3 --no-extra --no-cpu > $OUT/ncu_list_$TAG.log 2>&1
ls -la $OUT/*_$TAG*
