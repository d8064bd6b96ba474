"""BASELINE config 5 at full size on one GPU (development aid / record run): n = 32768 training points, predictive mean +
variance on a 100k-point daily grid, S = 1000 joint posterior draws.  Checks size-independent properties only."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from discontinuum_b200 import capi, models, synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
m = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
S = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
X, y, noise = synthetic.loadest_site(n, 1000)
th = np.array([0.05, 0.7, 1.0, 1.0, 2.0, 1.3, 0.5, 0.2, 0.3, 0.4])
eng = capi.Engine(max_n=n, max_m=2048)
eng.set_train(models.loadest_spec(2).to_c(), X, y, noise)
t0 = time.perf_counter(); val, grad, info = eng.nlml_grad(th); t1 = time.perf_counter()
val, grad, info = eng.nlml_grad(th); t2 = time.perf_counter()
print(f"n={n}: nlml+grad {1e3*(t2-t1):.1f} ms ({n**3/(t2-t1)/1e12:.2f} TF/s) info={info} nlml={val:.6f}", flush=True)
t0 = time.perf_counter(); _, info = eng.factorize(th); t1 = time.perf_counter()
print(f"factorize {1e3*(t1-t0):.1f} ms info={info}", flush=True)
grid = synthetic.daily_grid(X, m)
t0 = time.perf_counter(); mu, var = eng.predict(grid); t1 = time.perf_counter()
print(f"predict m={m}: {t1-t0:.2f} s  {m/(t1-t0):.0f} points/s  ({m*float(n)**2/(t1-t0)/1e12:.2f} TF/s)  var range [{var.min():.3e}, {var.max():.3e}]", flush=True)
assert np.all(np.isfinite(mu)) and var.min() > -1e-8
if S > 0:
    rng = np.random.default_rng(0)
    Z = rng.standard_normal((S, m))
    t0 = time.perf_counter(); draws, info = eng.sample(grid, Z, jitter=1e-6); t1 = time.perf_counter()
    flop = float(m) * n * n + float(m) ** 2 * n + float(m) ** 3 / 3 + float(m) ** 2 * S
    print(f"sample m={m} S={S}: {t1-t0:.2f} s info={info} ({flop/(t1-t0)/1e12:.2f} TF/s)", flush=True)
    d = draws - mu
    sd = d.std(axis=0)
    ratio = sd / np.sqrt(np.maximum(var, 1e-12))
    print(f"draw mean error max {np.abs(draws.mean(0)-mu).max():.3e}; sd/sqrt(var): median {np.median(ratio):.3f}, 1%..99% [{np.quantile(ratio,0.01):.3f}, {np.quantile(ratio,0.99):.3f}]")
eng.close()
