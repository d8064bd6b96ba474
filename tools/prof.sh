#!/bin/bash
# ncu evidence for one round (run under gpurun, one GPU):  tools/prof.sh <tag>
#   1. launch list of the bench command (per-launch gpu__time_duration, cold-cache / serialised)
#   2. DRAM bytes of every launch of one evaluation (roofline.traffic)
#   3. --set full captures of the dominant launches: K=512 trailing update, LAUUM+grad, one inverse-merge product, potf2
TAG=${1:-rXX}
OUT=gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --no-extra --no-cpu"
$BENCH > $OUT/bench_plain_$TAG.json 2> $OUT/bench_plain_$TAG.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2700 --csv --log-file $OUT/launches_n16384_$TAG.csv $BENCH > $OUT/ncu_list_$TAG.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1340 --csv --log-file $OUT/dram_n16384_$TAG.csv \
    python tools/one_nlml.py 16384 grad > $OUT/ncu_dram_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_gemm --launch-skip 17 --launch-count 1 \
    -o $OUT/prof_trail_$TAG -f python tools/one_nlml.py 16384 grad > $OUT/ncu_t_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_gemm<\(int\)0, \(int\)1>' --launch-count 1 \
    -o $OUT/prof_lauum_$TAG -f python tools/one_nlml.py 16384 grad > $OUT/ncu_l_$TAG.log 2>&1
# last merge level of the recursive inverse: the two products are the 2nd- and 3rd-to-last k_gemm<0,0> launches before LAUUM
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_gemm<\(int\)0, \(int\)0>' --launch-skip 139 --launch-count 1 \
    -o $OUT/prof_invm_$TAG -f python tools/one_nlml.py 16384 grad > $OUT/ncu_i_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_potf2 --launch-skip 5 --launch-count 1 \
    -o $OUT/prof_potf2_$TAG -f python tools/one_nlml.py 16384 grad > $OUT/ncu_p_$TAG.log 2>&1
ls -la $OUT/*_$TAG*
