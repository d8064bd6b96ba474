"""Run-to-run variance of the site batch, completion-order polling against the fixed round (development aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from discontinuum_b200 import multisite, synthetic
rng = np.random.default_rng(42)
ns = (2000 + 6000 * rng.uniform(size=128)).astype(int)[:16]
sites = {s: synthetic.loadest_site(int(n), 1000 + s) for s, n in enumerate(ns)}
grids = {s: synthetic.daily_grid(sites[s][0], 10958) for s in sites}
flop = sum(float(n) ** 3 * 100 for n in ns)
multisite.fit_sites_local({0: sites[0]}, iterations=3, device=0, concurrency=1)
for rep in range(3):
    for mode in (True, False):
        for conc in (4, 6):
            multisite._COMPLETION_ORDER = mode
            t0 = time.perf_counter()
            multisite.fit_sites_local(sites, iterations=100, device=0, concurrency=conc, predict=grids)
            dt = time.perf_counter() - t0
            print(f"rep {rep} completion_order={mode} conc={conc}: {dt:.2f} s  {16/dt:.3f} sites/s  {flop/dt/1e12:.1f} TF", flush=True)
