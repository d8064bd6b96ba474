"""GPU parity at BASELINE.json's sizes: the CUDA path (through the C ABI) against the CPU oracle on the same seeded
inputs, at the largest sizes the oracle finishes in tens of seconds on the GPU box's host cores, and size-independent
properties above that.  Tolerance of the north star: <= 1e-6 relative in float64 (written per assertion).

  config 1 / 4 : predictive mean + variance at n = 4096 on the 10 958-point daily grid
  config 2     : rating-gp n = 2000 (theta0 and theta1 are in test_gpu_parity.py) -> 7 rating curves x 250 stages and
                 the 12 053-point daily grid; NLML + gradient at n = 4096
  config 3     : NLML + full gradient + alpha at n = 8192; NLML + alpha against torch.linalg.cholesky at n = 16 384;
                 properties (finite-difference gradient, mu(X) = y - noise alpha, 0 <= var <= noise) at n = 32 768
  config 5     : joint draws with supplied normals at n = 2048, m = 4096
"""
import time

import numpy as np
import pytest
import torch

import helpers as H
from helpers import orc
from discontinuum_b200 import capi, models, synthetic

pytestmark = pytest.mark.gpu
RTOL = 1e-6


def _engine(spec, X, y, noise, max_m=2048):
    eng = capi.Engine(max_n=X.shape[0], max_m=max_m)
    eng.set_train(spec.to_c(), X, y, noise)
    return eng


def _tt(*arrs):
    return [torch.tensor(a) for a in arrs]


def _close(got, want, rtol=RTOL):
    got, want = np.asarray(got), np.asarray(want)
    assert np.max(np.abs(got - want)) <= rtol * np.max(np.abs(want)), (np.max(np.abs(got - want)), np.max(np.abs(want)))


def test_nlml_grad_alpha_vs_oracle_n8192_loadest(cuda_device):
    n = 8192
    X, y, noise = synthetic.loadest_site(n, 1000)
    theta = H.loadest_theta1()
    eng = _engine(models.loadest_spec(2), X, y, noise)
    val, grad, info = eng.nlml_grad(theta)
    alpha = eng.alpha()
    eng.close()
    assert info == 0
    Xt, yt, nt = _tt(X, y, noise)
    t0 = time.time()
    v, g, a, _ = orc.nlml_grad_closed_form(orc.loadest_cov, orc.loadest_mean, H.loadest_nat_from_theta(theta), Xt, yt, nt)
    print(f"oracle n={n}: {time.time() - t0:.1f} s")
    assert abs(val - float(v)) <= RTOL * abs(float(v))
    _close(grad, H.loadest_theta_from_nat({k: t.numpy() for k, t in g.items()}))
    _close(alpha, a.numpy())


def test_nlml_grad_alpha_vs_oracle_n4096_rating(cuda_device):
    n = 4096
    X, y, noise = synthetic.rating_gauge(n, 7)
    b_lo, b_hi = models.stage_quantile_bounds(X[:, 1])
    theta = H.rating_theta1(b_lo, b_hi)
    eng = _engine(models.rating_spec(b_lo, b_hi), X, y, noise)
    val, grad, info = eng.nlml_grad(theta)
    alpha = eng.alpha()
    eng.close()
    assert info == 0
    Xt, yt, nt = _tt(X, y, noise)
    v, g, a, _ = orc.nlml_grad_closed_form(orc.rating_cov, orc.rating_mean, H.rating_nat_from_theta(theta), Xt, yt, nt,
                                           extra_key="noise")
    assert abs(val - float(v)) <= RTOL * abs(float(v))
    _close(grad, H.rating_theta_from_nat({k: t.numpy() for k, t in g.items()}))
    _close(alpha, a.numpy())


def test_nlml_alpha_vs_torch_cholesky_n16384(cuda_device):
    """The BASELINE metric's size: NLML and alpha against a dense float64 covariance build + torch.linalg.cholesky on
    the host (one shot, no autograd); the gradient at this size is covered by the finite-difference property test."""
    n = 16384
    X, y, noise = synthetic.loadest_site(n, 1000)
    theta = H.loadest_theta1()
    eng = _engine(models.loadest_spec(2), X, y, noise)
    val, grad, info = eng.nlml_grad(theta)
    alpha = eng.alpha()
    eng.close()
    assert info == 0
    Xt, yt, nt = _tt(X, y, noise)
    nat = H.loadest_nat_from_theta(theta)
    t0 = time.time()
    with torch.no_grad():
        K = orc.loadest_cov(Xt, Xt, nat)
        K.diagonal().add_(nt)
        v, L, a = orc.nlml_from_K(K, yt - orc.loadest_mean(Xt, nat))
    print(f"host cholesky n={n}: {time.time() - t0:.1f} s")
    assert abs(val - float(v)) <= RTOL * abs(float(v))
    _close(alpha, a.numpy())
    # d NLML / d mean constant = -sum(alpha): one gradient component the host factorisation gives for free
    assert abs(grad[0] + float(a.sum())) <= RTOL * max(1.0, abs(float(a.sum())))
    del K, L


def test_predict_vs_oracle_config1_grid(cuda_device):
    n, m = 4096, 10958
    X, y, noise = synthetic.loadest_site(n, 1003)
    theta = H.loadest_theta1()
    Xs = synthetic.daily_grid(X, m)
    eng = _engine(models.loadest_spec(2), X, y, noise)
    _, info = eng.factorize(theta)
    assert info == 0
    mu, var = eng.predict(Xs)
    eng.close()
    mu_o, _, var_o = orc.predict(orc.loadest_cov, orc.loadest_mean, H.loadest_nat_from_theta(theta), *_tt(X, y, noise, Xs))
    _close(mu, mu_o.numpy())
    _close(var, var_o.numpy())
    assert np.all(var > 0)


def test_predict_rating_config2_curves_and_daily_grid(cuda_device):
    n = 2000
    X, y, noise = synthetic.rating_gauge(n, 7)
    b_lo, b_hi = models.stage_quantile_bounds(X[:, 1])
    theta = H.rating_theta1(b_lo, b_hi)
    nat = H.rating_nat_from_theta(theta)
    # 7 rating curves: 250 stages over the observed range at 7 dates (rating_gp/plot.py plot_ratings_in_time), + daily grid
    stages = np.linspace(X[:, 1].min(), X[:, 1].max(), 250)
    curves = np.concatenate([np.stack([np.full(250, t), stages], axis=1) for t in np.linspace(-15.0, 15.0, 7)])
    Xs = np.ascontiguousarray(np.concatenate([curves, synthetic.daily_grid(X, 12053)]))
    eng = _engine(models.rating_spec(b_lo, b_hi), X, y, noise)
    _, info = eng.factorize(theta)
    assert info == 0
    mu, var = eng.predict(Xs)
    eng.close()
    mu_o, _, var_o = orc.predict(orc.rating_cov, orc.rating_mean, nat, *_tt(X, y, noise, Xs), extra_noise=nat["noise"])
    _close(mu, mu_o.numpy())
    _close(var, var_o.numpy())


def test_sample_vs_oracle_n2048_m4096(cuda_device):
    n, m, S = 2048, 4096, 16
    X, y, noise = synthetic.loadest_site(n, 1005)
    theta = H.loadest_theta1()
    Xs = synthetic.daily_grid(X, m) + np.array([0.0004, 0.01])
    Z = np.random.default_rng(5).standard_normal((S, m))
    eng = _engine(models.loadest_spec(2), X, y, noise)
    eng.factorize(theta)
    draws, info = eng.sample(Xs, Z, jitter=1e-6)
    eng.close()
    assert info == 0
    want, _ = orc.sample(orc.loadest_cov, orc.loadest_mean, H.loadest_nat_from_theta(theta), *_tt(X, y, noise, Xs, Z), jitter=1e-6)
    # draws = mu + Z Lpost': the Cholesky factor of a posterior covariance with condition ~1e8 amplifies rounding
    # differences in Sigma*, hence 1e-5 on the draws themselves (as in the small-size test) ...
    assert np.max(np.abs(draws - want.numpy())) <= 1e-5 * np.max(np.abs(want.numpy()))
    # ... and so does their second moment (what the annual-flux uncertainty is made of)
    _close(np.mean(draws * draws, axis=0), np.mean(want.numpy() ** 2, axis=0), 1e-5)


def test_full_size_properties_n32768(cuda_device):
    """Upper end of BASELINE config 3 / the training size of config 5 (the oracle needs minutes and > 100 GB here):
    (i) central-difference directional derivative of NLML against the analytic gradient; (ii) mu(X) = y - noise alpha;
    (iii) 0 <= latent variance at training inputs <= noise."""
    n = 32768
    X, y, noise = synthetic.loadest_site(n, 1000)
    eng = _engine(models.loadest_spec(2), X, y, noise, max_m=512)
    th = H.loadest_theta1()
    val, grad, info = eng.nlml_grad(th)
    assert info == 0 and np.isfinite(val)
    d = np.random.default_rng(0).standard_normal(th.shape[0]) * th * 1e-2
    d[0] = 1e-2
    eps = 1e-3
    vp, ip = eng.nlml(th + eps * d)
    vm, im = eng.nlml(th - eps * d)
    assert ip == 0 and im == 0
    fd = (vp - vm) / (2 * eps)
    assert abs(fd - grad @ d) <= 1e-5 * abs(grad @ d) + 1e-8 * abs(val), (fd, grad @ d)
    _, info = eng.factorize(th)
    assert info == 0
    alpha = eng.alpha()
    idx = np.arange(0, n, 71)[:400]
    mu, var = eng.predict(X[idx])
    assert np.max(np.abs(mu - (y[idx] - noise[idx] * alpha[idx]))) <= 1e-7 * max(1.0, np.max(np.abs(y)))
    assert np.all(var > -1e-9) and np.all(var <= noise[idx] + 1e-9)
    eng.close()
