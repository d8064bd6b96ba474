"""Edge cases of the C ABI on the GPU: smallest inputs, chunk boundaries, poisoned hyper-parameters."""
import numpy as np
import pytest
import torch

import helpers as H
from helpers import orc
from discontinuum_b200 import capi, models, synthetic

pytestmark = pytest.mark.gpu
RTOL = 1e-6


def _oracle_predict(theta, X, y, noise, Xs):
    nat = H.loadest_nat_from_theta(theta)
    mu, _, var = orc.predict(orc.loadest_cov, orc.loadest_mean, nat, torch.tensor(X), torch.tensor(y), torch.tensor(noise),
                             torch.tensor(Xs))
    return mu.numpy(), var.numpy()


@pytest.mark.parametrize("n", [1, 2, 5, 127])
def test_smallest_training_sets_predict_and_sample(cuda_device, n):
    """n below one 128-block (down to a single observation): NLML, gradient, prediction and one joint draw."""
    X, y, noise = synthetic.loadest_site(max(n, 2), 77)
    X, y, noise = np.ascontiguousarray(X[:n]), np.ascontiguousarray(y[:n]), np.ascontiguousarray(noise[:n])
    theta = H.loadest_theta1()
    eng = capi.Engine(max_n=n, max_m=256)
    eng.set_train(models.loadest_spec(2).to_c(), X, y, noise)
    nat = H.loadest_nat_from_theta(theta)
    v, g, _, _ = orc.nlml_grad_closed_form(orc.loadest_cov, orc.loadest_mean, nat, torch.tensor(X), torch.tensor(y), torch.tensor(noise))
    val, grad, info = eng.nlml_grad(theta)
    assert info == 0 and abs(val - float(v)) <= RTOL * max(1.0, abs(float(v)))
    go = H.loadest_theta_from_nat({k: t.numpy() for k, t in g.items()})
    assert np.max(np.abs(grad - go)) <= RTOL * max(1.0, np.max(np.abs(go)))
    eng.factorize(theta)
    for m in (1, 3):
        Xs = np.ascontiguousarray(synthetic.daily_grid(X if n > 1 else np.vstack([X, X + 1.0]), m))
        mu, var = eng.predict(Xs)
        mu_o, var_o = _oracle_predict(theta, X, y, noise, Xs)
        assert mu.shape == (m,) and np.max(np.abs(mu - mu_o)) <= RTOL * max(1.0, np.max(np.abs(mu_o)))
        assert np.max(np.abs(var - var_o)) <= RTOL * max(1e-6, np.max(np.abs(var_o)))
    Z = np.random.default_rng(0).standard_normal((1, 3))
    draws, info = eng.sample(Xs, Z, jitter=1e-8)
    assert info == 0 and draws.shape == (1, 3) and np.all(np.isfinite(draws))
    eng.close()


def test_prediction_chunk_boundaries(cuda_device):
    """m = chunk - 1, chunk, chunk + 1 and 1 (the engine predicts in chunks of max_m points): identical values per point."""
    n = 300
    X, y, noise = synthetic.loadest_site(n, 3)
    theta = H.loadest_theta1()
    eng = capi.Engine(max_n=n, max_m=256)
    eng.set_train(models.loadest_spec(2).to_c(), X, y, noise)
    eng.factorize(theta)
    Xs = synthetic.daily_grid(X, 513) + np.array([0.0007, 0.0])
    mu_all, var_all = eng.predict(Xs)
    mu_o, var_o = _oracle_predict(theta, X, y, noise, Xs)
    assert np.max(np.abs(mu_all - mu_o)) <= RTOL * np.max(np.abs(mu_o)) and np.max(np.abs(var_all - var_o)) <= RTOL * np.max(np.abs(var_o))
    for m in (1, 255, 256, 257, 512):
        mu, var = eng.predict(np.ascontiguousarray(Xs[:m]))
        assert np.array_equal(mu, mu_all[:m]) and np.array_equal(var, var_all[:m]), m
    eng.close()


def test_poisoned_theta_reports_and_engine_stays_usable(cuda_device):
    """NaN hyper-parameters: a positive `info` (or a DgpError), never a hang or a sticky device fault; the next
    evaluation with sane values on the same handle is exact again.  Same for one poisoned site of a batch."""
    n = 400
    X, y, noise = synthetic.loadest_site(n, 5)
    spec = models.loadest_spec(2)
    theta = H.loadest_theta1()
    eng = capi.Engine(max_n=n, max_m=256)
    eng.set_train(spec.to_c(), X, y, noise)
    good, g0, info = eng.nlml_grad(theta)
    assert info == 0
    for idx in (0, 1, 4):   # constant mean, an output scale, a length scale
        bad = theta.copy()
        bad[idx] = float("nan")
        try:
            val, grad, info = eng.nlml_grad(bad)
            assert info != 0 or not np.isfinite(val)
        except capi.DgpError:
            pass
        val, grad, info = eng.nlml_grad(theta)
        assert info == 0 and val == good and np.array_equal(grad, g0)
    eng.close()
    site2 = synthetic.loadest_site(250, 6)
    batch = capi.BatchEngine(max_sites=2, max_n=n)
    batch.set_train(spec.to_c(), [(X, y, noise), site2])
    thetas = np.stack([theta, theta])
    v0, gr0, i0 = batch.nlml_grad(thetas)
    assert not i0.any() and v0[0] == good
    thetas_bad = thetas.copy()
    thetas_bad[1, 0] = float("nan")
    v1, gr1, i1 = batch.nlml_grad(thetas_bad)
    assert i1[0] == 0 and v1[0] == good and np.array_equal(gr1[0], g0) and (i1[1] != 0 or not np.isfinite(v1[1]))
    v2, gr2, i2 = batch.nlml_grad(thetas)
    assert not i2.any() and np.array_equal(v2, v0) and np.array_equal(gr2, gr0)
    batch.close()
