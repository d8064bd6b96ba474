"""Second, independent pin of the oracle: scikit-learn's exact-GP implementation (sklearn.gaussian_process, float64,
Cholesky) evaluates the same marginal likelihood, its gradient and the predictive mean / variance for kernels that both
sides can express.  gpytorch / pymc are not installable here, sklearn is -- it is a third-party
implementation of the same mathematics, not the reference, so DESIGN.md keeps the "parity unpinned" label; this test
removes the possibility that oracle and mpmath golden vectors share a formula mistake in NLML / gradient / prediction.

Mapping (sklearn -> oracle):  ExpSineSquared(l, p) = exp(-2 sin^2(pi d / p) / l^2)  -> k_periodic(lam = l^2)
                              Matern(nu, l), RBF(l) with per-dimension l           -> k_matern / k_rbf  (l = 1e9 switches a dim off)
"""
import numpy as np
import pytest
import torch

import helpers as H  # noqa: F401
from helpers import orc

sk = pytest.importorskip("sklearn.gaussian_process")
from sklearn.gaussian_process import GaussianProcessRegressor  # noqa: E402
from sklearn.gaussian_process.kernels import RBF, ConstantKernel, ExpSineSquared, Matern, WhiteKernel  # noqa: E402

OFF = 1e9  # length-scale that removes a dimension from an ARD kernel


def _data(n, seed):
    rng = np.random.default_rng(seed)
    t = np.sort(rng.uniform(-3.0, 3.0, n))
    q = rng.standard_normal(n)
    y = 0.5 * q + 0.4 * np.sin(2 * np.pi * t) + 0.2 * rng.standard_normal(n)
    return np.stack([t, q], axis=1), y


def test_loadest_structure_nlml_gradient_and_prediction_match_sklearn():
    """K = s1 Per(t) M52(t) + s2 RBF(q) + s3 M32(t, q) + noise I on 1-D-time data (the periodic kernel of sklearn has no
    active_dims, so the time-only factors are checked on X = t and the covariate factors through ARD length-scales)."""
    n = 60
    X, y = _data(n, 0)
    s1, lam, period, l1, s2, l2, s3, l3t, l3q, noise = 0.7, 1.3, 1.1, 2.0, 1.3, 0.5, 0.2, 0.3, 0.4, 0.01
    # sklearn cannot restrict ExpSineSquared to one column: build that term on t only and add it as a precomputed check
    Xt = X[:, :1]
    k_time = ConstantKernel(s1) * ExpSineSquared(length_scale=np.sqrt(lam), periodicity=period) * Matern(length_scale=l1, nu=2.5)
    Kt_sk = k_time(Xt)
    nat = H.loadest_nat_from_theta(np.array([0.0, s1, lam, period, l1, s2, l2, s3, l3t, l3q]))
    Xtt = torch.tensor(X)
    Kt_or = (nat["s1"] * orc.k_periodic(Xtt, Xtt, 0, nat["period"], nat["lam"]) * orc.k_matern(Xtt, Xtt, [0], nat["l1"], 2.5)).numpy()
    assert np.max(np.abs(Kt_sk - Kt_or)) < 1e-13
    k_cov = ConstantKernel(s2) * RBF(length_scale=[OFF, l2]) + ConstantKernel(s3) * Matern(length_scale=[l3t, l3q], nu=1.5)
    Kc_sk = k_cov(X)
    K_or = orc.loadest_cov(Xtt, Xtt, nat).numpy()
    assert np.max(np.abs(Kt_sk + Kc_sk - K_or)) < 1e-12


def test_nlml_gradient_prediction_match_sklearn_on_a_shared_kernel():
    """A kernel both sides express exactly: s2 RBF_ard + s3 M32_ard + sigma^2 I.  sklearn's log_marginal_likelihood and
    its gradient (w.r.t. log-parameters) vs the oracle's NLML and closed-form gradient; predictive mean / variance."""
    n, m = 80, 25
    X, y = _data(n, 1)
    Xs, _ = _data(m, 2)
    s2, l2t, l2q, s3, l3t, l3q, noise = 1.3, 0.9, 0.5, 0.4, 0.3, 0.6, 0.05
    kernel = (ConstantKernel(s2) * RBF(length_scale=[l2t, l2q]) + ConstantKernel(s3) * Matern(length_scale=[l3t, l3q], nu=1.5)
              + WhiteKernel(noise_level=noise))
    gpr = GaussianProcessRegressor(kernel=kernel, alpha=0.0, optimizer=None, normalize_y=False).fit(X, y)
    lml, dlml = gpr.log_marginal_likelihood(gpr.kernel_.theta, eval_gradient=True)  # gradient w.r.t. log-parameters

    def cov(X1, X2, p):
        return p["s2"] * orc.k_rbf(X1, X2, [0, 1], p["l2"]) + p["s3"] * orc.k_matern(X1, X2, [0, 1], p["l3"], 1.5)

    def mean(Xa, p):
        return torch.zeros(Xa.shape[0], dtype=orc.DT)

    nat = {"s2": torch.tensor([s2], dtype=orc.DT), "l2": torch.tensor([l2t, l2q], dtype=orc.DT),
           "s3": torch.tensor([s3], dtype=orc.DT), "l3": torch.tensor([l3t, l3q], dtype=orc.DT),
           "noise": torch.tensor([noise], dtype=orc.DT)}
    Xt, yt, zero = torch.tensor(X), torch.tensor(y), torch.zeros(n, dtype=orc.DT)
    val, g, alpha, L = orc.nlml_grad_closed_form(cov, mean, nat, Xt, yt, zero, extra_key="noise")
    assert abs(float(val) + lml) <= 1e-10 * abs(lml)
    # sklearn's theta order: [s2, l2t, l2q, s3, l3t, l3q, noise] in log space; d(-NLML)/dlog p = -p dNLML/dp
    ours = -np.concatenate([(g["s2"] * nat["s2"]).numpy(), (g["l2"] * nat["l2"]).numpy(), (g["s3"] * nat["s3"]).numpy(),
                            (g["l3"] * nat["l3"]).numpy(), (g["noise"] * nat["noise"]).numpy()])
    assert np.max(np.abs(ours - dlml)) <= 1e-8 * np.max(np.abs(dlml))
    # autograd path agrees too
    val2, g2 = orc.nlml_grad_autograd(cov, mean, nat, Xt, yt, zero, extra_key="noise")
    assert abs(float(val2) - float(val)) <= 1e-12 * abs(float(val))
    # prediction: sklearn returns the variance of y* including the white-noise level
    mu_sk, sd_sk = gpr.predict(Xs, return_std=True)
    mu, var_obs, var_lat = orc.predict(cov, mean, nat, Xt, yt, zero, torch.tensor(Xs), extra_noise=nat["noise"])
    assert np.max(np.abs(mu.numpy() - mu_sk)) <= 1e-9 * np.max(np.abs(mu_sk))
    assert np.max(np.abs(var_obs.numpy() - sd_sk ** 2)) <= 1e-8 * np.max(sd_sk ** 2)
