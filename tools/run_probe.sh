timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/gputests_v2.log 2>&1; tail -3 gpurun_out/gputests_v2.log
python tools/perf_probe.py 1024 2048 4096 8192 16384 > gpurun_out/perf_k.log 2>&1
cat gpurun_out/perf_k.log
python tools/trace_probe.py 4096 > gpurun_out/trace_probe.log 2>&1
