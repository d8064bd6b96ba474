"""Opportunistic pin against the third-party library that owns the reference's arithmetic (SURVEY 8c pin 5).

gpytorch is not installed in this image or on the GPU box, so these tests report "skipped" today; on a box that has
it they compare the oracle (CPU) and the CUDA engine (GPU) with what the reference's engine computes:
  objective : -ExactMarginalLogLikelihood(likelihood, model)(model(x), y)   discontinuum/engines/gpytorch.py:318,353
  gradient  : objective.backward() w.r.t. the raw parameters                 discontinuum/engines/gpytorch.py:384
  predict   : likelihood(model(x)) in eval mode, .mean / .variance           discontinuum/engines/gpytorch.py:618-624
in float64 with the Cholesky path forced (gpytorch.settings.max_cholesky_size(10**9), no fast-pred-var low-rank
approximation).  The model is assembled exactly as loadest_gp/models/gpytorch.py:48-128 does.
"""
import numpy as np
import pytest
import torch

import helpers as H
from helpers import orc
from discontinuum_b200 import synthetic

gpytorch = pytest.importorskip("gpytorch", reason="gpytorch is not installed: parity stays pinned by mpmath / scikit-learn only")

RTOL = 1e-6


def _reference_loadest_model(X, y):
    """loadest_gp/models/gpytorch.py:48-128, float64."""
    from gpytorch.kernels import MaternKernel, PeriodicKernel, RBFKernel, ScaleKernel
    from gpytorch.priors import GammaPrior, HalfNormalPrior, NormalPrior

    noise = 0.1 ** 2 * torch.ones(y.shape[0], dtype=torch.float64).reshape(1, -1)
    likelihood = gpytorch.likelihoods.FixedNoiseGaussianLikelihood(noise=noise, learn_additional_noise=False)

    class ExactGPModel(gpytorch.models.ExactGP):
        def __init__(self, train_x, train_y, lik):
            super().__init__(train_x, train_y, lik)
            n_d = train_x.shape[1]
            dims = np.arange(n_d)
            time_dim, cov_dims = [dims[0]], dims[1:]
            self.mean_module = gpytorch.means.ConstantMean()
            seasonal = ScaleKernel(PeriodicKernel(period_length_prior=NormalPrior(loc=1, scale=0.01), active_dims=time_dim)
                                   * MaternKernel(nu=2.5, active_dims=time_dim), outputscale_prior=HalfNormalPrior(scale=1))
            covariates = ScaleKernel(RBFKernel(ard_num_dims=cov_dims.shape[0], lengthscale_prior=GammaPrior(concentration=2, rate=3),
                                               active_dims=cov_dims), outputscale_prior=HalfNormalPrior(scale=2))
            residual = ScaleKernel(MaternKernel(ard_num_dims=dims.shape[0], nu=1.5, active_dims=dims,
                                                lengthscale_prior=GammaPrior(concentration=2, rate=10)),
                                   outputscale_prior=HalfNormalPrior(scale=0.2))
            self.covar_module = seasonal + covariates + residual

        def forward(self, x):
            return gpytorch.distributions.MultivariateNormal(self.mean_module(x), self.covar_module(x))

    model = ExactGPModel(X, y, likelihood).double()
    likelihood = likelihood.double()
    return model, likelihood


def _set_raw(model, raw):
    """Write the oracle's raw parameter dict into the gpytorch module tree (same softplus constraints)."""
    k = model.covar_module.kernels
    with torch.no_grad():
        model.mean_module.raw_constant.copy_(raw["mean_c"].reshape(model.mean_module.raw_constant.shape))
        k[0].raw_outputscale.copy_(raw["s1"].reshape(k[0].raw_outputscale.shape))
        per, mat = k[0].base_kernel.kernels
        per.raw_lengthscale.copy_(raw["lam"].reshape(per.raw_lengthscale.shape))
        per.raw_period_length.copy_(raw["period"].reshape(per.raw_period_length.shape))
        mat.raw_lengthscale.copy_(raw["l1"].reshape(mat.raw_lengthscale.shape))
        k[1].raw_outputscale.copy_(raw["s2"].reshape(k[1].raw_outputscale.shape))
        k[1].base_kernel.raw_lengthscale.copy_(raw["l2"].reshape(k[1].base_kernel.raw_lengthscale.shape))
        k[2].raw_outputscale.copy_(raw["s3"].reshape(k[2].raw_outputscale.shape))
        k[2].base_kernel.raw_lengthscale.copy_(raw["l3"].reshape(k[2].base_kernel.raw_lengthscale.shape))


def _gpytorch_objective_and_grad(model, likelihood, X, y):
    model.train()
    likelihood.train()
    mll = gpytorch.mlls.ExactMarginalLogLikelihood(likelihood, model)
    with gpytorch.settings.max_cholesky_size(10 ** 9):
        obj = -mll(model(X), y)
        obj = obj.sum()
        obj.backward()
    k = model.covar_module.kernels
    per, mat = k[0].base_kernel.kernels
    g = {"mean_c": model.mean_module.raw_constant.grad, "s1": k[0].raw_outputscale.grad, "lam": per.raw_lengthscale.grad,
         "period": per.raw_period_length.grad, "l1": mat.raw_lengthscale.grad, "s2": k[1].raw_outputscale.grad,
         "l2": k[1].base_kernel.raw_lengthscale.grad, "s3": k[2].raw_outputscale.grad, "l3": k[2].base_kernel.raw_lengthscale.grad}
    return float(obj.detach()), {n: v.detach().reshape(-1).clone() for n, v in g.items()}


def _gpytorch_predict(model, likelihood, Xs):
    model.eval()
    likelihood.eval()
    with torch.no_grad(), gpytorch.settings.max_cholesky_size(10 ** 9), gpytorch.settings.fast_pred_var(False):
        pred = likelihood(model(Xs))
        return pred.mean.reshape(-1).numpy(), pred.variance.reshape(-1).numpy()


def _case(n=600, m=300, fitted_like=True):
    X, y, noise = synthetic.loadest_site(n, 1000)
    theta = H.loadest_theta1() if fitted_like else H.loadest_theta0()
    nat = H.loadest_nat_from_theta(theta)
    raw = orc.loadest_raw_from_natural(nat)
    Xs = synthetic.daily_grid(X, m) + np.array([0.0007, 0.0])
    return X, y, noise, theta, nat, raw, Xs


@pytest.mark.parametrize("fitted_like", [False, True])
def test_oracle_matches_gpytorch_objective_gradient_prediction(fitted_like):
    X, y, noise, theta, nat, raw, Xs = _case(fitted_like=fitted_like)
    Xt, yt, nt = torch.tensor(X), torch.tensor(y), torch.tensor(noise)
    model, lik = _reference_loadest_model(Xt, yt)
    _set_raw(model, raw)
    obj_g, grad_g = _gpytorch_objective_and_grad(model, lik, Xt, yt)
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in raw.items()}
    obj_o = orc.objective("loadest", leaves, Xt, yt, nt)
    obj_o.backward()
    assert abs(float(obj_o) - obj_g) <= RTOL * abs(obj_g)
    for k in leaves:
        assert float((leaves[k].grad.reshape(-1) - grad_g[k]).abs().max()) <= RTOL * max(1.0, float(grad_g[k].abs().max())), k
    mu_g, var_g = _gpytorch_predict(model, lik, torch.tensor(Xs))
    mu_o, var_o, _ = orc.predict(orc.loadest_cov, orc.loadest_mean, nat, Xt, yt, nt, torch.tensor(Xs))
    assert np.max(np.abs(mu_o.numpy() - mu_g)) <= RTOL * np.max(np.abs(mu_g))
    assert np.max(np.abs(var_o.numpy() - var_g)) <= RTOL * np.max(np.abs(var_g))


@pytest.mark.gpu
def test_engine_matches_gpytorch_objective_gradient_prediction(cuda_device):
    from discontinuum_b200 import capi, models
    from discontinuum_b200.spec import GPModule

    X, y, noise, theta, nat, raw, Xs = _case(n=2000, m=1500)
    Xt, yt = torch.tensor(X), torch.tensor(y)
    model, lik = _reference_loadest_model(Xt, yt)
    _set_raw(model, raw)
    obj_g, grad_g = _gpytorch_objective_and_grad(model, lik, Xt, yt)
    module = GPModule(models.loadest_spec(2))
    for name, value in zip([p.name for p in module.spec.params], theta):
        module.set_natural(name, float(value))
    natv, dnat, lp, dlp = module.host_chain()
    eng = capi.Engine(max_n=X.shape[0], max_m=2048)
    eng.set_train(module.spec.to_c(), X, y, noise)
    val, grad, info = eng.nlml_grad(np.ascontiguousarray(natv))
    assert info == 0
    n = X.shape[0]
    obj = (val - lp) / n
    graw = (grad - dlp) * dnat / n
    assert abs(obj - obj_g) <= RTOL * abs(obj_g)
    want = H.loadest_theta_from_nat({k: v.numpy() for k, v in grad_g.items()})
    assert np.max(np.abs(graw - want)) <= RTOL * max(1.0, np.max(np.abs(want)))
    eng.factorize(np.ascontiguousarray(natv))
    mu, var = eng.predict(Xs)
    eng.close()
    mu_g, var_g = _gpytorch_predict(model, lik, torch.tensor(Xs))
    assert np.max(np.abs(mu - mu_g)) <= RTOL * np.max(np.abs(mu_g))
    assert np.max(np.abs(np.maximum(var, 1e-10) - var_g)) <= RTOL * np.max(np.abs(var_g))
