#!/bin/bash
# ncu captures of the dominant kernels of one NLML+grad evaluation at n=16384 (run under gpurun, one GPU).
# Usage: tools/prof.sh <tag>
TAG=${1:-rXX}
OUT=gpurun_out
python tools/one_nlml.py 16384 grad > $OUT/plain_$TAG.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_potf2 --launch-skip 5 --launch-count 1 \
    -o $OUT/prof_potf2_$TAG -f python tools/one_nlml.py 16384 grad > $OUT/ncu_p_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_gemm --launch-skip 17 --launch-count 1 \
    -o $OUT/prof_trail_$TAG -f python tools/one_nlml.py 16384 grad > $OUT/ncu_t_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:k_gemm<\(int\)0, \(int\)1>' --launch-count 1 \
    -o $OUT/prof_lauum_$TAG -f python tools/one_nlml.py 16384 grad > $OUT/ncu_l_$TAG.log 2>&1
ls -la $OUT/*.ncu-rep
