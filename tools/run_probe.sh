REPS=6 python tools/perf_probe.py 512 1024 2048 2>&1 | grep -v "nlml only" | cut -c1-170
echo GRAPHS; DGP_GRAPHS=1 REPS=6 python tools/perf_probe.py 512 1024 2048 2>&1 | grep -v "nlml only" | cut -c1-170
