import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from discontinuum_b200 import capi, models, synthetic
th = np.array([0.05, 0.7, 1.0, 1.0, 2.0, 1.3, 0.5, 0.2, 0.3, 0.4])
for n in [int(a) for a in sys.argv[1:]]:
    X, y, noise = synthetic.loadest_site(n, 1000)
    eng = capi.Engine(max_n=n, max_m=128)
    eng.set_train(models.loadest_spec(2).to_c(), X, y, noise)
    print(n, [eng.nlml(th) for _ in range(3)], flush=True)
    eng.close()
