"""Generate the known-answer vectors that pin the oracle (TEST INFRASTRUCTURE).

The reference's tests hold no numbers for this path and gpytorch is not installable here, so the pins are
self-generated: every quantity is evaluated with mpmath at 40 significant digits from the formulas of
SURVEY Appendix A (an implementation independent of oracle/gp_oracle.py: scalar loops, mp.cholesky,
numerical differentiation at high precision for the gradient) and rounded to float64.

    python oracle/make_golden.py        -> tests/golden/kat_loadest.json, tests/golden/kat_rating.json
"""
from __future__ import annotations

import json
import os

import mpmath as mp
import numpy as np

mp.mp.dps = 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")

SQ3, SQ5 = mp.sqrt(3), mp.sqrt(5)


def m32(r):
    a = SQ3 * r
    return (1 + a) * mp.e ** (-a)


def m52(r):
    a = SQ5 * r
    return (1 + a + a * a / 3) * mp.e ** (-a)


def per(d, p, lam):
    return mp.e ** (-2 * mp.sin(mp.pi * d / p) ** 2 / lam)


def loadest_k(x, z, th):
    c, s1, lam, p, l1, s2, l2, s3, l3t, l3q = th
    dt, dq = x[0] - z[0], x[1] - z[1]
    seasonal = s1 * per(dt, p, lam) * m52(abs(dt) / l1)
    flow = s2 * mp.e ** (-(dq / l2) ** 2 / 2)
    resid = s3 * m32(mp.sqrt((dt / l3t) ** 2 + (dq / l3q) ** 2))
    return seasonal + flow + resid


def loadest_mean(x, th):
    return th[0]


RATING_KEYS = ("pl_a", "pl_b", "pl_c", "noise", "gate_b", "shiftA_s", "shiftA_lh", "shiftA_lt", "shiftB_s", "shiftB_lh",
               "shiftB_lt", "bend_s", "bend_lh", "bend_lt", "base_s", "base_l", "per_s", "per_period", "per_lam", "per_l")


def rating_k(x, z, th):
    t = dict(zip(RATING_KEYS, th))
    dt = abs(x[0] - z[0])
    dh = abs(mp.log(x[1] + mp.mpf("1e-6")) - mp.log(z[1] + mp.mpf("1e-6")))
    g1 = 1 / (1 + mp.e ** (20 * (x[1] - t["gate_b"])))
    g2 = 1 / (1 + mp.e ** (20 * (z[1] - t["gate_b"])))
    shift = lambda n: t[n + "_s"] * m52(dh / t[n + "_lh"]) * m32(dt / t[n + "_lt"])
    bend = t["bend_s"] * m52(dh / t["bend_lh"]) * m52(dt / t["bend_lt"])
    base = t["base_s"] * m52(dh / t["base_l"])
    periodic = t["per_s"] * per(x[0] - z[0], t["per_period"], t["per_lam"]) * m52(dt / t["per_l"])
    return g1 * g2 * (shift("shiftA") + shift("shiftB")) + (1 - g1) * (1 - g2) * bend + base + periodic


def rating_mean(x, th):
    return th[0] + th[1] * mp.log(x[1] - th[2])


def nlml(kfun, mfun, th, X, y, noise, extra_idx=None):
    n = len(X)
    K = mp.matrix(n, n)
    for i in range(n):
        for j in range(i + 1):
            K[i, j] = K[j, i] = kfun(X[i], X[j], th)
        K[i, i] += noise[i] + (th[extra_idx] if extra_idx is not None else 0)
    L = mp.cholesky(K)
    r = mp.matrix([y[i] - mfun(X[i], th) for i in range(n)])
    z = mp.lu_solve(L, r)  # L is lower triangular; lu_solve is exact enough at 40 digits
    logdet = sum(mp.log(L[i, i]) for i in range(n))
    return (z.T * z)[0] / 2 + logdet + mp.mpf(n) / 2 * mp.log(2 * mp.pi), K, L, r


def predict(kfun, mfun, th, X, y, noise, Xs, extra_idx=None):
    val, K, L, r = nlml(kfun, mfun, th, X, y, noise, extra_idx)
    alpha = mp.lu_solve(K, r)
    mus, vs = [], []
    for xs in Xs:
        kx = mp.matrix([kfun(xs, X[j], th) for j in range(len(X))])
        mus.append(mfun(xs, th) + (kx.T * alpha)[0])
        v = mp.lu_solve(K, kx)
        vs.append(kfun(xs, xs, th) - (kx.T * v)[0])
    return mus, vs, alpha


def case(model, n, theta, seed):
    rng = np.random.default_rng(seed)
    if model == "loadest":
        X = np.stack([np.sort(rng.uniform(-3, 3, n)), rng.standard_normal(n)], 1)
        noise = np.full(n, 0.01)
        kfun, mfun, extra = loadest_k, loadest_mean, None
        Xs = np.stack([np.linspace(-3.2, 3.2, 5), rng.standard_normal(5)], 1)
    else:
        X = np.stack([np.sort(rng.uniform(-3, 3, n)), 1 + rng.uniform(0, 1, n)], 1)
        noise = rng.choice(np.array([1e-3, 4e-3, 9e-3]), n)
        kfun, mfun, extra = rating_k, rating_mean, 3
        Xs = np.stack([np.linspace(-3.2, 3.2, 5), 1 + rng.uniform(0, 1, 5)], 1)
    y = rng.standard_normal(n)
    th = [mp.mpf(float(t)) for t in theta]
    Xm = [[mp.mpf(float(v)) for v in row] for row in X]
    ym = [mp.mpf(float(v)) for v in y]
    nm = [mp.mpf(float(v)) for v in noise]
    Xsm = [[mp.mpf(float(v)) for v in row] for row in Xs]
    val, K, L, r = nlml(kfun, mfun, th, Xm, ym, nm, extra)

    def f(*args):
        return nlml(kfun, mfun, list(args), Xm, ym, nm, extra)[0]

    grad = []
    for i in range(len(th)):
        g = mp.diff(lambda t, i=i: f(*[t if k == i else th[k] for k in range(len(th))]), th[i], h=mp.mpf("1e-12"))
        grad.append(g)
    mus, vs, alpha = predict(kfun, mfun, th, Xm, ym, nm, Xsm, extra)
    Kplain = [[float(kfun(Xm[i], Xm[j], th)) for j in range(n)] for i in range(n)]
    return {"model": model, "n": n, "theta": [float(t) for t in theta], "X": X.tolist(), "y": y.tolist(),
            "noise": noise.tolist(), "Xs": Xs.tolist(), "nlml": float(val), "nlml_str": mp.nstr(val, 30),
            "grad": [float(g) for g in grad], "alpha": [float(a) for a in alpha], "K": Kplain,
            "L": [[float(L[i, j]) for j in range(n)] for i in range(n)],
            "mu": [float(m) for m in mus], "var_latent": [float(v) for v in vs]}


def main():
    os.makedirs(OUT, exist_ok=True)
    sp0 = float(mp.log(2))  # softplus(0): GPyTorch's initial value of every positive parameter
    lo_theta0 = [0.0] + [sp0] * 9
    lo_theta1 = [0.05, 0.7, 1.0, 1.0, 2.0, 1.3, 0.5, 0.2, 0.3, 0.4]
    cases = []
    for n in (1, 2, 3, 8, 24):
        cases.append(case("loadest", n, lo_theta0, 100 + n))
        cases.append(case("loadest", n, lo_theta1, 200 + n))
    with open(os.path.join(OUT, "kat_loadest.json"), "w") as fh:
        json.dump(cases, fh)
    ra_theta0 = [0.0, 1.3, 0.5, sp0 + 1e-4, 1.5] + [sp0] * 15
    ra_theta1 = [0.1, 1.6, 0.55, 0.002, 1.52, 0.4, 1.2, 2.5, 0.1, 2.0, 0.15, 0.3, 1.0, 3.0, 0.9, 0.8, 0.05, 1.0, 0.9, 4.0]
    cases = []
    for n in (1, 2, 3, 8, 24):
        cases.append(case("rating", n, ra_theta0, 300 + n))
        cases.append(case("rating", n, ra_theta1, 400 + n))
    with open(os.path.join(OUT, "kat_rating.json"), "w") as fh:
        json.dump(cases, fh)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
