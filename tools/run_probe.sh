timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputests_v2.log 2>&1; tail -3 gpurun_out/gputests_v2.log
python tools/perf_probe.py 2048 4096 8192 16384 > gpurun_out/perf_l.log 2>&1
cat gpurun_out/perf_l.log
