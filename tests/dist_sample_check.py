"""2+ GPU check of multisite.sample_sharded, launched with torchrun (not collected by pytest):
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tests/dist_sample_check.py
Every rank factorises the same site; the distributed draws must equal single-GPU dgp_sample with the same Philox
normals, on every rank."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist

import helpers as H
from discontinuum_b200 import capi, models, multisite, synthetic

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, m, S = (int(a) for a in (sys.argv[1:4] if len(sys.argv) >= 4 else (700, 2300, 16)))
X, y, noise = synthetic.loadest_site(n, 52)
Xs = synthetic.daily_grid(X, m) + np.array([0.0009, 0.0])
stream = torch.cuda.current_stream()
eng = capi.Engine(max_n=n, max_m=2048, device=local, stream=stream.cuda_stream)
eng.set_train(models.loadest_spec(2).to_c(), X, y, noise)
eng.factorize(H.loadest_theta1())
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
got, info = multisite.sample_sharded(eng, Xs, S, dist=dist, seed=7, jitter=1e-7)
torch.cuda.synchronize(); t1 = time.perf_counter()
ok = True
if m <= 20000:
    want, info1 = eng.sample_ex(Xs, S, Z=None, seed=7, jitter=1e-7)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    err = float(np.max(np.abs(got - want)) / np.max(np.abs(want)))
    ok = info == 0 and info1 == 0 and err <= 1e-9
    print(f"rank {rank}/{world}: n={n} m={m} S={S} sharded {t1-t0:.3f}s single {t2-t1:.3f}s info={info} rel err {err:.2e} {'OK' if ok else 'FAIL'}", flush=True)
else:
    mu, var = eng.predict(Xs[:4096])
    sd = (got[:, :4096] - mu).std(axis=0) / np.sqrt(np.maximum(var, 1e-12))
    ok = info == 0 and abs(float(np.median(sd)) - 1.0) < 0.1
    print(f"rank {rank}/{world}: n={n} m={m} S={S} sharded {t1-t0:.3f}s info={info} median sd/sqrt(var) {np.median(sd):.3f} {'OK' if ok else 'FAIL'}", flush=True)
eng.close()
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
