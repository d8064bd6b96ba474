"""Stage-by-stage GPU-vs-oracle diagnostic (development aid; the judged parity tests are tests/test_gpu_*.py)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import helpers as H
from helpers import orc
from discontinuum_b200 import capi, models, synthetic


def check(model, n, theta_kind, eng=None, verbose=True):
    if model == "loadest":
        X, y, noise = synthetic.loadest_site(n, 1000)
        spec = models.loadest_spec(2)
        theta = H.loadest_theta0() if theta_kind == 0 else H.loadest_theta1()
        nat = H.loadest_nat_from_theta(theta)
        cov, mean, extra = orc.loadest_cov, orc.loadest_mean, None
    else:
        X, y, noise = synthetic.rating_gauge(n, 7)
        blo, bhi = models.stage_quantile_bounds(X[:, 1])
        spec = models.rating_spec(blo, bhi)
        theta = H.rating_theta0(blo, bhi) if theta_kind == 0 else H.rating_theta1(blo, bhi)
        nat = H.rating_nat_from_theta(theta)
        cov, mean, extra = orc.rating_cov, orc.rating_mean, "noise"
    Xt, yt, nt = torch.tensor(X), torch.tensor(y), torch.tensor(noise)
    eng = capi.Engine(max_n=n, max_m=512)
    eng.set_train(spec.to_c(), X, y, noise)
    out = {}
    K = eng.covmat(theta)
    Ko = cov(Xt, Xt, nat).numpy()
    out["cov_abs"] = float(np.abs(K - Ko).max())
    v, g, alpha_o, L_o = orc.nlml_grad_closed_form(cov, mean, nat, Xt, yt, nt, extra_key=extra)
    val, info = eng.nlml(theta)
    out["info"] = info
    out["nlml_rel"] = abs(val - float(v)) / abs(float(v))
    L = eng.chol()
    out["chol_abs"] = float(np.abs(np.tril(L) - L_o.numpy()).max())
    out["chol_upper_abs"] = float(np.abs(np.triu(L, 1)).max())
    eng.set_debug_kinv(True)
    val2, grad, info2 = eng.nlml_grad(theta)
    out["nlml2_rel"] = abs(val2 - float(v)) / abs(float(v))
    a = eng.alpha()
    out["alpha_rel"] = float(np.abs(a - alpha_o.numpy()).max() / np.abs(alpha_o.numpy()).max())
    Ki = eng.kinv()
    Kio = torch.cholesky_inverse(L_o).numpy()
    out["kinv_rel"] = float(np.abs(np.tril(Ki) - np.tril(Kio)).max() / np.abs(Kio).max())
    if model == "loadest":
        go = H.loadest_theta_from_nat({k: t.numpy() for k, t in g.items()})
    else:
        go = H.rating_theta_from_nat({k: t.numpy() for k, t in g.items()})
    out["grad_rel"] = float(np.max(np.abs(grad - go) / np.maximum(np.abs(go), 1e-12 * np.abs(go).max())))
    out["grad_rel_max"] = float(np.abs(grad - go).max() / np.abs(go).max())
    # predict
    m = 300
    Xs = synthetic.daily_grid(X, m)
    val3, info3 = eng.factorize(theta)
    mu, var = eng.predict(Xs)
    en = nat["noise"] if extra else None
    mu_o, _, var_o = orc.predict(cov, mean, nat, Xt, yt, nt, torch.tensor(Xs), extra_noise=en)
    out["mu_rel"] = float(np.abs(mu - mu_o.numpy()).max() / np.abs(mu_o.numpy()).max())
    out["var_rel"] = float(np.abs(var - var_o.numpy()).max() / np.abs(var_o.numpy()).max())
    if verbose:
        print(model, n, theta_kind, " ".join(f"{k}={v:.3g}" for k, v in out.items()), flush=True)
        if out["grad_rel"] > 1e-6:
            print("   grad gpu", grad); print("   grad ref", go)
    eng.close()
    return out


if __name__ == "__main__":
    t0 = time.time()
    for model in ("loadest", "rating"):
        for n in (1, 2, 100, 128, 129, 300, 1000):
            for tk in (0, 1):
                if model == "rating" and n < 20:
                    continue
                try:
                    check(model, n, tk)
                except Exception as e:  # noqa: BLE001
                    print("FAILED", model, n, tk, repr(e), flush=True)
    print("elapsed", time.time() - t0)
