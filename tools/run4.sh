#!/bin/bash
mkdir -p gpurun_out
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench2 rc=$?"; tail -5 gpurun_out/bench2.err; head -c 300 gpurun_out/bench2.json; echo
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/bench2.json') if l.startswith('{')][-1])
print(d['value'], d['e2e'])
s = d['extra']['sites']; print('sites', s['value'], s['fit_tflops_per_gpu'], s['per_rank'])
c = d['extra']['config5']; print('c5', c['predict']['seconds'], c['predict']['value'], c['sample']['seconds'], c['sample']['collectives'], c['check_sharded_vs_single_gpu_draws_rel'], c['sample']['sd_over_sqrt_var_median'])
PY
