"""CPU: pin the oracle against the committed mpmath known-answer vectors (tests/golden, made by
oracle/make_golden.py) and against analytic closed forms.  No GPU, no engine."""
import json
import math
import os

import numpy as np
import pytest
import torch

import helpers as H
from helpers import orc

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _cases(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


def _setup(case):
    X, y, noise = (torch.tensor(case[k], dtype=H.DT) for k in ("X", "y", "noise"))
    if case["model"] == "loadest":
        nat = H.loadest_nat_from_theta(case["theta"])
        return X, y, noise, nat, orc.loadest_cov, orc.loadest_mean, None
    nat = H.rating_nat_from_theta(case["theta"])
    return X, y, noise, nat, orc.rating_cov, orc.rating_mean, "noise"


@pytest.mark.parametrize("fname", ["kat_loadest.json", "kat_rating.json"])
def test_oracle_matches_mpmath(fname):
    for case in _cases(fname):
        X, y, noise, nat, cov, mean, extra = _setup(case)
        K = cov(X, X, nat).numpy()
        assert np.max(np.abs(K - np.array(case["K"]))) < 5e-15
        val, g, alpha, L = orc.nlml_grad_closed_form(cov, mean, nat, X, y, noise, extra_key=extra)
        assert abs(float(val) - case["nlml"]) <= 1e-12 * max(1.0, abs(case["nlml"]))
        assert np.max(np.abs(L.numpy() - np.array(case["L"]))) < 1e-12
        assert np.max(np.abs(alpha.numpy() - np.array(case["alpha"]))) <= 1e-9 * np.max(np.abs(case["alpha"]))
        to_theta = H.loadest_theta_from_nat if case["model"] == "loadest" else H.rating_theta_from_nat
        gv = to_theta({k: t.numpy() for k, t in g.items()})
        gold = np.array(case["grad"])
        assert np.max(np.abs(gv - gold)) <= 1e-8 * max(1.0, np.max(np.abs(gold))), (case["n"], gv, gold)
        _, g2 = orc.nlml_grad_autograd(cov, mean, nat, X, y, noise, extra_key=extra)
        gv2 = to_theta({k: t.numpy() for k, t in g2.items()})
        assert np.max(np.abs(gv2 - gold)) <= 1e-8 * max(1.0, np.max(np.abs(gold)))
        en = nat["noise"] if extra else None
        mu, _, var = orc.predict(cov, mean, nat, X, y, noise, torch.tensor(case["Xs"], dtype=H.DT), extra_noise=en)
        assert np.max(np.abs(mu.numpy() - np.array(case["mu"]))) <= 1e-9 * max(1.0, np.max(np.abs(case["mu"])))
        assert np.max(np.abs(var.numpy() - np.array(case["var_latent"]))) <= 1e-9


def test_analytic_n1():
    """n = 1: NLML = r^2 / (2 v) + log(v) / 2 + log(2 pi) / 2 with v = k(x, x) + noise."""
    X = torch.tensor([[0.3, -0.2]], dtype=H.DT)
    y = torch.tensor([0.7], dtype=H.DT)
    th = H.loadest_theta1()
    nat = H.loadest_nat_from_theta(th)
    v = th[1] + th[5] + th[7] + 0.01
    r = 0.7 - th[0]
    want = r * r / (2 * v) + 0.5 * math.log(v) + 0.5 * math.log(2 * math.pi)
    got = orc.nlml(orc.loadest_cov, orc.loadest_mean, nat, X, y, torch.tensor([0.01], dtype=H.DT))
    assert abs(float(got) - want) < 1e-14


def test_analytic_n2_predict():
    """n = 2, posterior at a training input with tiny noise reproduces the observation and ~zero variance."""
    X = torch.tensor([[0.0, 0.0], [1.3, 0.4]], dtype=H.DT)
    y = torch.tensor([0.5, -0.25], dtype=H.DT)
    nat = H.loadest_nat_from_theta(H.loadest_theta1())
    noise = torch.full((2,), 1e-10, dtype=H.DT)
    mu, _, var = orc.predict(orc.loadest_cov, orc.loadest_mean, nat, X, y, noise, X.clone())
    assert torch.allclose(mu, y, atol=1e-8) and float(var.abs().max()) < 1e-8


def test_kernel_invariants():
    X = torch.tensor(np.random.default_rng(0).normal(size=(40, 2)))
    nat = H.loadest_nat_from_theta(H.loadest_theta1())
    K = orc.loadest_cov(X, X, nat)
    assert torch.allclose(K, K.T, atol=1e-15)
    assert torch.allclose(torch.diagonal(K), torch.full((40,), 0.7 + 1.3 + 0.2, dtype=H.DT), atol=1e-14)
    perm = torch.randperm(40)
    y = torch.tensor(np.random.default_rng(1).normal(size=40))
    n0 = orc.nlml(orc.loadest_cov, orc.loadest_mean, nat, X, y, orc.loadest_noise(40))
    n1 = orc.nlml(orc.loadest_cov, orc.loadest_mean, nat, X[perm], y[perm], orc.loadest_noise(40))
    assert abs(float(n0 - n1)) < 1e-10


def test_objective_prior_scale():
    """The training objective is -(logN + sum log prior) / n; at GPyTorch's init the period prior dominates."""
    X, y = torch.zeros(3, 2, dtype=H.DT), torch.zeros(3, dtype=H.DT)
    X[:, 0] = torch.tensor([0.0, 0.4, 0.9])
    raw = orc.loadest_init_raw()
    obj = orc.objective("loadest", raw, X, y, orc.loadest_noise(3))
    nat = orc.loadest_natural(raw)
    lp_period = float(orc.lp_normal(nat["period"], 1.0, 0.01))
    assert lp_period < -400
    manual = (orc.nlml(orc.loadest_cov, orc.loadest_mean, nat, X, y, orc.loadest_noise(3)) - orc.loadest_log_prior(nat)) / 3
    assert abs(float(obj - manual)) < 1e-12
