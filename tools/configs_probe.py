"""BASELINE configs 1 and 2 end to end through the model classes (development aid / DESIGN table):
   config 1: loadest-gp, one synthetic site n = 1000, fit (100 iterations) + predict on a 10 958-point daily grid
   config 2: rating-gp, one synthetic gauge n = 2000, fit (100 iterations) + 7 rating curves of 250 points + 12 053 daily points
Host arrays in, host arrays out; wall-clock seconds.  With CPU=1 the oracle's restatement of the reference loop
(float64 torch on the host cores) is timed on a few iterations beside it."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from discontinuum_b200 import models

def loadest_arrays(n, seed):
    rng = np.random.default_rng(seed)
    days = np.sort(rng.uniform(0, 10958, n))
    t = np.datetime64("1990-01-01") + (days * 86400e9).astype("timedelta64[ns]")
    flow = np.exp(1.0 + 0.8 * np.sin(2 * np.pi * days / 365.25) + 0.5 * rng.standard_normal(n))
    conc = np.exp(0.3 * np.log(flow) + 0.2 * np.cos(2 * np.pi * days / 365.25) + 0.2 * rng.standard_normal(n))
    return {"time": t, "flow": flow}, conc

def daily(n_days, seed, key):
    rng = np.random.default_rng(seed)
    d = np.arange(n_days)
    t = np.datetime64("1990-01-01") + d.astype("timedelta64[D]")
    v = np.exp(1.0 + 0.8 * np.sin(2 * np.pi * d / 365.25) + 0.3 * rng.standard_normal(n_days))
    return {"time": t, key: v}

iters = int(os.environ.get("ITERS", "100"))
# ---- config 1
cov, conc = loadest_arrays(1000, 0)
grid = daily(10958, 1, "flow")
m = models.LoadestGP(); m.fit(cov, conc, iterations=3)          # warm-up: module load, workspace
torch.cuda.synchronize()
fits = []
for rep in range(int(os.environ.get("REPS", "7"))):
    t0 = time.perf_counter(); m = models.LoadestGP(); m.fit(cov, conc, iterations=iters); torch.cuda.synchronize(); t1 = time.perf_counter()
    fits.append(t1 - t0)
t1 = time.perf_counter()
target, se = m.predict(grid); torch.cuda.synchronize(); t2 = time.perf_counter()
fits.sort()
print(f"config 1 loadest n=1000: fit {iters} it median {fits[len(fits) // 2]:.3f} s (min {fits[0]:.3f}, max {fits[-1]:.3f}; "
      f"{fits[len(fits) // 2] / iters * 1e3:.2f} ms/it), predict 10958 pts {t2 - t1:.3f} s, "
      f"objective {m.history[0]:.4f} -> {m.history[-1]:.4f}", flush=True)
# ---- config 2
rng = np.random.default_rng(7)
n = 2000
days = np.sort(rng.uniform(0, 12053, n))
t = np.datetime64("1990-01-01") + (days * 86400e9).astype("timedelta64[ns]")
stage = rng.lognormal(1.0, 0.5, n)
q = 3.0 * (stage - 0.5 * stage.min()) ** 1.6 * np.exp(0.03 * rng.standard_normal(n))
gse = rng.choice(np.array([1.01, 1.025, 1.04, 1.06]), n)
torch.manual_seed(0)
r = models.RatingGP(); r.fit({"time": t, "stage": stage}, q, target_unc=gse, iterations=3)
torch.cuda.synchronize()
fits = []
for rep in range(int(os.environ.get("REPS", "7"))):
    torch.manual_seed(0)
    t0 = time.perf_counter(); r = models.RatingGP(); r.fit({"time": t, "stage": stage}, q, target_unc=gse, iterations=iters); torch.cuda.synchronize(); t1 = time.perf_counter()
    fits.append(t1 - t0)
fits.sort()
t1 = time.perf_counter()
for k in range(7):
    hs = np.linspace(stage.min(), stage.max(), 250)
    r.predict({"time": np.full(250, t[(k * n) // 7]), "stage": hs})
dg = daily(12053, 2, "stage")
r.predict(dg); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"config 2 rating n=2000: fit {iters} it median {fits[len(fits) // 2]:.3f} s (min {fits[0]:.3f}, max {fits[-1]:.3f}; "
      f"{fits[len(fits) // 2] / iters * 1e3:.2f} ms/it), 7 curves + 12053 daily pts {t2 - t1:.3f} s, "
      f"objective {r.history[0]:.4f} -> {r.history[-1]:.4f}", flush=True)
if os.environ.get("CPU"):
    from helpers import orc
    torch.set_num_threads(os.cpu_count() or 1)
    raw = orc.loadest_init_raw()
    t0 = time.perf_counter()
    orc.fit_adam("loadest", raw, torch.tensor(m.X), torch.tensor(m.y), torch.tensor(m.fixed_noise), iterations=5)
    dt = (time.perf_counter() - t0) / 5
    print(f"CPU oracle loop, config 1: {dt * 1e3:.1f} ms/it on {torch.get_num_threads()} threads", flush=True)
