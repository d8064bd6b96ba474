"""One batched evaluation of a BASELINE config-4 group under ncu (development aid): usage one_batch.py <group 0..7> [evals]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np
import helpers as H
from discontinuum_b200 import capi, models, synthetic

g = int(sys.argv[1]) if len(sys.argv) > 1 else 0
evals = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rng = np.random.default_rng(42)
ns_all = np.sort((2000 + 6000 * rng.uniform(size=128)).astype(int))[::-1]
ns = [int(v) for v in ns_all[16 * g:16 * g + 16]]
sites = [synthetic.loadest_site(n, 2000 + k) for k, n in enumerate(ns)]
b = capi.BatchEngine(max_sites=16, max_n=max(ns))
b.set_train(models.loadest_spec(2).to_c(), sites)
th = np.stack([H.loadest_theta1()] * 16)
for _ in range(evals):
    v, gr, info = b.nlml_grad(th)
print("sizes", ns, "launches", b.launches, "info", info.tolist())
b.close()
