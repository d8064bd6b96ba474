timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputests_r01e.log 2>&1; tail -3 gpurun_out/gputests_r01e.log
python tools/perf_probe.py 4096 16384 2>&1 | grep -v "nlml only" | cut -c1-150
cp build/libdgp_gc3.so discontinuum_b200/libdgp.so
echo GC3; python tools/perf_probe.py 4096 16384 2>&1 | grep -v "nlml only" | cut -c1-150
