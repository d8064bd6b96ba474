timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputests_r01d.log 2>&1; tail -3 gpurun_out/gputests_r01d.log
python bench.py > gpurun_out/bench_r01d.json 2> gpurun_out/bench_r01d.err; tail -c 1500 gpurun_out/bench_r01d.json
bash tools/prof.sh r01d > gpurun_out/prof_r01d.log 2>&1; tail -12 gpurun_out/prof_r01d.log
