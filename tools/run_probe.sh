timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/gputests_r01f.log 2>&1; tail -3 gpurun_out/gputests_r01f.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/bench_r01f.json 2> gpurun_out/bench_r01f.err; python -c "
import json
d=json.loads(open('gpurun_out/bench_r01f.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['phases_ms'], d['extra']['sites']['value'], d['extra']['sites']['fit_tflops'], d['extra']['predict']['value'], d['clocks'])"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r01f.json 2>&1; tail -c 600 gpurun_out/bench_ref_r01f.json
