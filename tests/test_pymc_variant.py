"""PyMC-parameterised loadest model on the B200 engine: covariance and MAP objective
against the oracle's restatement of PyMC's own formulas, then the engine surface."""
import numpy as np
import pytest
import torch

import helpers as H  # noqa: F401  (puts the repo root on sys.path)
from helpers import orc

from discontinuum_b200 import pymc_variant as pv


def test_pymc_variable_table_and_mapping():
    vars_ = pv.loadest_pymc_vars(3)
    assert sum(v.size for v in vars_) == 13 and pv.loadest_pymc_spec(3).ntheta == 13
    v = {x.name: torch.tensor(x.init) for x in vars_}
    nat = pv.pymc_to_natural(v)
    assert nat.shape == (13,) and float(nat[0]) == 1.0 and abs(float(nat[1]) - 4.0 * (4.0 / 3.0) ** 2) < 1e-15
    # log prior of pm.Exponential(scale=1.5) at its mean, pm.HalfNormal(1) at 1
    assert abs(float(vars_[4].logp(torch.tensor([1.5], dtype=torch.float64))) - (-np.log(1.5) - 1.0)) < 1e-15
    assert abs(float(vars_[0].logp(torch.tensor([1.0], dtype=torch.float64))) - (0.5 * np.log(2 / np.pi) - 0.5)) < 1e-15


def _arrays(n, seed):
    rng = np.random.default_rng(seed)
    days = np.sort(rng.uniform(0, 3650, n))
    time = np.datetime64("2000-01-01") + (days * 86400e9).astype("timedelta64[ns]")
    flow = np.exp(1.0 + 0.8 * np.sin(2 * np.pi * days / 365.25) + 0.5 * rng.standard_normal(n))
    conc = np.exp(0.3 * np.log(flow) + 0.2 * np.cos(2 * np.pi * days / 365.25) + 0.2 * rng.standard_normal(n))
    return {"time": time, "flow": flow}, conc


@pytest.mark.gpu
def test_pymc_variant_covariance_objective_and_surface(cuda_device):
    cov, conc = _arrays(180, 3)
    m = pv.LoadestGPMarginalPyMCB200()
    m.fit(cov, conc, maxiter=6)
    assert m.is_fitted and set(m.mp) == {v.name for v in m.vars}
    X, y = torch.tensor(m.X), torch.tensor(m.y)
    # covariance: engine tiles at the reparameterised theta == PyMC's own formulas
    v = {k: torch.tensor(val) for k, val in m.mp.items()}
    K = m._engine.covmat(m._theta())
    Kw = orc.pymc_loadest_cov(X, X, v).numpy()
    assert np.max(np.abs(K - Kw)) <= 1e-12 * np.max(np.abs(Kw))
    # MAP objective and its gradient w.r.t. the unconstrained vector
    u0 = np.concatenate([np.log(x.init) if x.positive else x.init for x in m.vars]) + 0.05
    val, grad = m._neg_logp(u0)
    u = torch.tensor(u0, requires_grad=True)
    want = orc.pymc_loadest_neg_logp(m._split(u), X, y)
    want.backward()
    assert abs(val - float(want.detach())) <= 1e-6 * abs(float(want.detach()))
    assert np.max(np.abs(grad - u.grad.numpy())) <= 1e-6 * np.max(np.abs(u.grad.numpy()))
    # BFGS made progress from PyMC's initial point
    f0, _ = m._neg_logp(u0 - 0.05)
    assert m.map_result.fun < f0
    # surface
    target, se = m.predict(cov)
    assert target.shape == (180,) and np.all(target > 0) and np.all(se >= 1.0)
    grid, index, covs = m.predict_grid("flow")
    assert grid.shape[1] == 18 and grid.shape[0] == index.shape[0]
    sim = m.sample({k: a[:40] for k, a in cov.items()}, n=16, seed=0)
    assert sim.shape == (16, 40) and np.all(np.isfinite(sim)) and np.all(sim > 0)
    with pytest.raises(RuntimeError, match="hasn't been fitted"):
        pv.LoadestGPMarginalPyMCB200().predict(cov)
