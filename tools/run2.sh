#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multisite.py tests/test_batch.py -m gpu -x -q --timeout 300 > gpurun_out/t2.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/t2.log
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
( time python bench.py ) > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench1.err; head -c 6000 gpurun_out/bench1.json
