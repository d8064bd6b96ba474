import sys, time, resource
import os; R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import torch, numpy as np
import helpers as H
from helpers import orc
from discontinuum_b200 import synthetic
n = int(sys.argv[1]); mode = sys.argv[2]
import psutil
need = 5.0 * (n / 4096) ** 2 * 1.3
if psutil.virtual_memory().available / 1e9 < need + 8:
    print(n, mode, 'skipped: needs', need, 'GB, available', psutil.virtual_memory().available / 1e9); sys.exit(0)
torch.set_num_threads(os.cpu_count() or 1)
X, y, noise = synthetic.loadest_site(n, 1000)
Xt, yt, nt = torch.tensor(X), torch.tensor(y), torch.tensor(noise)
nat = H.loadest_nat_from_theta(H.loadest_theta1())
t0 = time.perf_counter()
if mode == "cf":
    v, g, a, L = orc.nlml_grad_closed_form(orc.loadest_cov, orc.loadest_mean, nat, Xt, yt, nt)
elif mode == "ag":
    v, g = orc.nlml_grad_autograd(orc.loadest_cov, orc.loadest_mean, nat, Xt, yt, nt)
else:
    with torch.no_grad():
        K = orc.loadest_cov(Xt, Xt, nat); K.diagonal().add_(nt)
        v, L, a = orc.nlml_from_K(K, yt - orc.loadest_mean(Xt, nat))
print(n, mode, "sec", time.perf_counter() - t0, "maxrss GB", resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e6, float(v), torch.get_num_threads())
