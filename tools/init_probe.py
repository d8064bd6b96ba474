"""Where does first-call time go? (development aid)"""
import os, sys, time
t00 = time.perf_counter()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from discontinuum_b200 import capi, models, synthetic
t0 = time.perf_counter(); print(f"imports {t0-t00:.2f}s", flush=True)
X, y, noise = synthetic.loadest_site(2000, 1)
t1 = time.perf_counter(); eng = capi.Engine(max_n=2000, max_m=2048); t2 = time.perf_counter()
print(f"Engine() first {t2-t1:.3f}s", flush=True)
eng.set_train(models.loadest_spec(2).to_c(), X, y, noise); t3 = time.perf_counter(); print(f"set_train {t3-t2:.3f}s", flush=True)
th = np.array([0.05, 0.7, 1.0, 1.0, 2.0, 1.3, 0.5, 0.2, 0.3, 0.4])
for k in range(3):
    t = time.perf_counter(); eng.nlml_grad(th); print(f"nlml_grad #{k} {time.perf_counter()-t:.4f}s", flush=True)
t = time.perf_counter(); eng.factorize(th); print(f"factorize {time.perf_counter()-t:.4f}s", flush=True)
t = time.perf_counter(); eng.predict(synthetic.daily_grid(X, 4000)); print(f"predict {time.perf_counter()-t:.4f}s", flush=True)
t = time.perf_counter(); eng2 = capi.Engine(max_n=8000, max_m=2048); print(f"Engine() second (8000) {time.perf_counter()-t:.4f}s", flush=True)
t = time.perf_counter(); eng2.close(); print(f"close {time.perf_counter()-t:.4f}s", flush=True)
