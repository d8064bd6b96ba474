"""Phase timing of one NLML+grad evaluation at several n (development aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from discontinuum_b200 import capi, models, synthetic

def theta1():
    return np.array([0.05, 0.7, 1.0, 1.0, 2.0, 1.3, 0.5, 0.2, 0.3, 0.4])

sizes = [int(a) for a in sys.argv[1:]] or [2048, 4096, 8192, 16384]
reps = int(os.environ.get("REPS", "3"))
for n in sizes:
    X, y, noise = synthetic.loadest_site(n, 1000)
    eng = capi.Engine(max_n=n, max_m=256)
    eng.set_train(models.loadest_spec(2).to_c(), X, y, noise)
    eng.set_timing(True)
    th = theta1()
    for r in range(reps):
        t0 = time.perf_counter()
        val, grad, info = eng.nlml_grad(th)
        dt = (time.perf_counter() - t0) * 1e3
        ms = eng.last_timing()
    tot = sum(ms)
    print(f"n={n} info={info} nlml={val:.6f} wall={dt:.2f}ms dev={tot:.2f}ms potrf={ms[0]:.2f} trtri={ms[1]:.2f} lauum_grad={ms[2]:.2f} rest={ms[3]:.3f} "
          f"TF/s total={n**3/tot/1e9:.2f} potrf={n**3/3/ms[0]/1e9:.2f} trtri={n**3/3/ms[1]/1e9:.2f} lauum={n**3/3/ms[2]/1e9:.2f} launches={eng.launches}", flush=True)
    t0 = time.perf_counter(); val0, info0 = eng.nlml(th); dt0 = (time.perf_counter() - t0) * 1e3
    print(f"   nlml only: wall={dt0:.2f}ms diff={abs(val0-val):.3e}", flush=True)
    eng.close()
