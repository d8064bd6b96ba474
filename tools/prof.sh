#!/bin/bash
# ncu evidence for one round (run under gpurun, one GPU):  tools/prof.sh <tag>
#   1. launch list of the bench command (per-launch gpu__time_duration, cold-cache / serialised)
#   2. DRAM bytes of every launch of one evaluation (roofline.traffic)
#   3. --set full captures of the dominant launches: K=512 trailing update, LAUUM, one inverse-merge product, the
#      gradient contraction, the diagonal-block kernel
# Launch indices (n = 16384, panel width 4, half tiles below 296 tiles): among k_gemm<1,0,8> the 6th launch is the K = 512
# update of the column strip [16, 24) by panel 1 (1792 tiles; 3 in-panel updates, the next panel's other columns and the rest of its strip come before it); among k_gemm<0,0,8>: 53 panel solves, 14 inverse products, LAUUM last.
TAG=${1:-rXX}
OUT=gpurun_out
BENCH="python bench.py --steps 1 --warmup 3 --no-extra --no-cpu"
$BENCH > $OUT/bench_plain_$TAG.json 2> $OUT/bench_plain_$TAG.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches_n16384_$TAG.csv $BENCH > $OUT/ncu_list_$TAG.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1500 --csv --log-file $OUT/dram_n16384_$TAG.csv \
    python tools/one_nlml.py 16384 grad > $OUT/ncu_dram_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_gemm<\(int\)1, \(int\)0, \(int\)8>' --launch-skip 5 --launch-count 1 \
    -o $OUT/prof_trail_$TAG -f python tools/one_nlml.py 16384 grad > $OUT/ncu_t_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_gemm<\(int\)0, \(int\)0, \(int\)8>' --launch-skip 67 --launch-count 1 \
    -o $OUT/prof_lauum_$TAG -f python tools/one_nlml.py 16384 grad > $OUT/ncu_l_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_gemm<\(int\)0, \(int\)0, \(int\)8>' --launch-skip 66 --launch-count 1 \
    -o $OUT/prof_invm_$TAG -f python tools/one_nlml.py 16384 grad > $OUT/ncu_i_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_grad_contract --launch-count 1 \
    -o $OUT/prof_gradc_$TAG -f python tools/one_nlml.py 16384 grad > $OUT/ncu_g_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_potf2_v2 --launch-skip 5 --launch-count 1 \
    -o $OUT/prof_potf2_$TAG -f python tools/one_nlml.py 16384 grad > $OUT/ncu_p_$TAG.log 2>&1
ls -la $OUT/*_$TAG*
