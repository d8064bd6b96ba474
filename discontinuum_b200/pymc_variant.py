"""The reference's PyMC engine surface on the B200 engine.

  MarginalPyMCB200.fit / predict / predict_grid / sample  <- src/discontinuum/engines/pymc.py:24-169
  loadest_pymc_spec / LoadestGPMarginalPyMCB200            <- src/loadest_gp/models/pymc.py:30-90

PyMC's kernels are the same family in a different parameterisation, so the model is a host-side
reparameterisation onto the same dgp_spec tree (nothing new on the GPU):
  eta**2 * Periodic(period, ls) * Matern52(ls')   ->  scale = eta^2, periodic "lam" = 4 ls^2   (PyMC: exp(-sin^2(pi d/T) / (2 ls^2)))
  eta**2 * ExpQuad(ls)                            ->  scale = eta^2, RBF lengthscale ls
  eta**2 * Matern32(ls)                           ->  scale = eta^2, Matern-3/2 lengthscales ls
  WhiteNoise(sigma=0.1) + pm.gp.Marginal's 1e-6 jitter  ->  fixed noise 0.01 + 1e-6 on the diagonal, zero mean.
`pm.find_MAP(method="BFGS")` maximises log p(y | theta) + sum log prior(theta) over the transformed (log) variables
without the Jacobian term; here scipy's BFGS does the same with the objective and its gradient from dgp_nlml_grad.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np
import torch

from . import capi
from .engine import JITTERS, MIN_VARIANCE, DataMixin, ModelConfig, NotPSDError, _assign_coords, is_fitted
from .data import LogStandardPipeline, TimePipeline
from .spec import CovSpec, Factor

PYMC_SIGMA = 0.1        # cov_noise = WhiteNoise(sigma=0.1), loadest_gp/models/pymc.py:80-83
PYMC_JITTER = 1e-6      # pm.gp.Marginal JITTER_DEFAULT


def loadest_pymc_spec(ndim: int = 2) -> CovSpec:
    """K = eta_per^2 Per(t) M52(t) + eta_trend^2 RBF(t) + eta_cov^2 RBF(q) + eta_res^2 M32(t, q); zero mean."""
    if ndim < 2:
        raise ValueError("loadest-gp needs time + at least one covariate")
    s = CovSpec(ndim=ndim)
    cols = [s.col_copy(d) for d in range(ndim)]
    s1 = s.param("seasonal.scale"); lam = s.param("seasonal.periodic.lam"); per = s.param("seasonal.periodic.period")
    l1 = s.param("seasonal.matern52.ls")
    st = s.param("trend.scale"); lt = s.param("trend.rbf.ls")
    s2 = s.param("covariates.scale"); l2 = [s.param(f"covariates.rbf.ls.{d}") for d in range(ndim - 1)]
    s3 = s.param("residual.scale"); l3 = [s.param(f"residual.matern32.ls.{d}") for d in range(ndim)]
    s.term(s1, [Factor(capi.PERIODIC, [cols[0]], [lam], per), Factor(capi.MATERN52, [cols[0]], [l1])])
    s.term(st, [Factor(capi.RBF, [cols[0]], [lt])])
    s.term(s2, [Factor(capi.RBF, cols[1:], l2)])
    s.term(s3, [Factor(capi.MATERN32, cols, l3)])
    s.mean_kind = capi.MEAN_ZERO
    return s


class _Var:
    """One PyMC random variable: prior, initial value, log transform for positive supports."""

    def __init__(self, name, prior, init, size=1, positive=True):
        self.name, self.prior, self.size, self.positive = name, prior, size, positive
        self.init = np.full(size, float(init))

    def logp(self, x: torch.Tensor) -> torch.Tensor:
        p = self.prior
        if p[0] == "halfnormal":      # pm.HalfNormal(sigma)
            return (0.5 * math.log(2.0 / math.pi) - math.log(p[1]) - 0.5 * (x / p[1]) ** 2).sum()
        if p[0] == "gamma":           # pm.Gamma(alpha, beta)
            a, b = p[1], p[2]
            return (a * math.log(b) - math.lgamma(a) + (a - 1.0) * torch.log(x) - b * x).sum()
        if p[0] == "normal":          # pm.Normal(mu, sigma)
            return (-0.5 * math.log(2.0 * math.pi) - math.log(p[2]) - 0.5 * ((x - p[1]) / p[2]) ** 2).sum()
        if p[0] == "exponential":     # pm.Exponential(scale) -> lam = 1 / scale
            return (-math.log(p[1]) - x / p[1]).sum()
        raise ValueError(p)


def loadest_pymc_vars(ndim: int = 2) -> List[_Var]:
    """Priors and initial values of loadest_gp/models/pymc.py:37-78 (PyMC's default initial point of a variable
    without initval is its prior's moment)."""
    return [
        _Var("eta_per", ("halfnormal", 1.0), 1.0),
        _Var("ls_pdecay", ("gamma", 10.0, 1.0), 10.0),
        _Var("period", ("normal", 1.0, 0.05), 1.0, positive=False),
        _Var("ls_psmooth", ("gamma", 4.0, 3.0), 4.0 / 3.0),
        _Var("eta_trend", ("exponential", 1.5), 1.5),
        _Var("ls_trend", ("gamma", 4.0, 1.0), 4.0),
        _Var("eta_covariates", ("halfnormal", 2.0), 2.0),
        _Var("ls_covariates", ("gamma", 2.0, 3.0), 0.5, size=ndim - 1),
        _Var("eta_res", ("exponential", 0.2), 0.2),
        _Var("ls_res", ("gamma", 2.0, 10.0), 0.2, size=ndim),
    ]


def pymc_to_natural(v: Dict[str, torch.Tensor]) -> torch.Tensor:
    """PyMC variables -> the engine's natural theta, in loadest_pymc_spec order."""
    return torch.cat([
        v["eta_per"] ** 2, 4.0 * v["ls_psmooth"] ** 2, v["period"], v["ls_pdecay"],
        v["eta_trend"] ** 2, v["ls_trend"],
        v["eta_covariates"] ** 2, v["ls_covariates"],
        v["eta_res"] ** 2, v["ls_res"]])


class MarginalPyMCB200:
    """src/discontinuum/engines/pymc.py:17-169 with find_MAP / gp.predict on the CUDA engine."""

    device_index = 0
    max_predict_chunk = 2048

    def __init__(self, model_config=None):
        self.model_config = model_config
        self.dm = None
        self.is_fitted = False
        self._engine: Optional[capi.Engine] = None
        self._factorized_at = None
        self.mp: Dict[str, np.ndarray] = {}

    # hooks a model subclass provides
    def build_model(self, X, y):
        raise NotImplementedError("This method must be implemented in a subclass")

    # -- MAP objective on the unconstrained vector
    def _split(self, u: torch.Tensor) -> Dict[str, torch.Tensor]:
        out, k = {}, 0
        for var in self.vars:
            seg = u[k:k + var.size]
            out[var.name] = torch.exp(seg) if var.positive else seg
            k += var.size
        return out

    def _neg_logp(self, u_np: np.ndarray):
        u = torch.tensor(u_np, dtype=torch.float64, requires_grad=True)
        v = self._split(u)
        nat = pymc_to_natural(v)
        th = nat.detach().numpy().astype(np.float64)
        for jit in JITTERS:
            val, grad, info = self._engine.nlml_grad(th, PYMC_JITTER + jit)
            if info == 0 and math.isfinite(val):
                break
        else:
            raise NotPSDError(f"covariance not positive definite at the MAP iterate (info={info})")
        lp = sum(var.logp(v[var.name]) for var in self.vars)
        obj = val + ((nat - nat.detach()) * torch.from_numpy(grad.copy())).sum() - lp
        obj.backward()
        return float(obj.detach()), u.grad.numpy().astype(np.float64)

    def fit(self, covariates, target, method: str = "BFGS", maxiter: Optional[int] = None):
        from scipy.optimize import minimize

        self.is_fitted = True
        self.dm.fit(target=target, covariates=covariates)
        self.X, self.y = self.dm.X, self.dm.y
        self.spec, self.vars = self.build_model(self.X, self.y)
        n = self.X.shape[0]
        if self._engine is None or self._engine.max_n < n:
            if self._engine is not None:
                self._engine.close()
            self._engine = capi.Engine(max_n=n, max_m=self.max_predict_chunk, device=self.device_index)
        self._engine.set_train(self.spec.to_c(), self.X, self.y, np.full(n, PYMC_SIGMA ** 2))
        u0 = np.concatenate([np.log(var.init) if var.positive else var.init for var in self.vars])
        opts = {"maxiter": maxiter} if maxiter is not None else {}
        res = minimize(self._neg_logp, u0, jac=True, method=method, options=opts)
        self.map_result = res
        with torch.no_grad():
            v = self._split(torch.tensor(res.x, dtype=torch.float64))
        self.mp = {k: t.numpy().copy() for k, t in v.items()}
        self._factorized_at = None

    def _theta(self) -> np.ndarray:
        return pymc_to_natural({k: torch.tensor(v, dtype=torch.float64) for k, v in self.mp.items()}).numpy().astype(np.float64)

    def _ensure_factorized(self):
        th = self._theta()
        if self._factorized_at is None or not np.array_equal(self._factorized_at, th):
            for jit in JITTERS:
                _, info = self._engine.factorize(th, PYMC_JITTER + jit)
                if info == 0:
                    break
            else:
                raise NotPSDError("covariance not positive definite at the MAP point")
            self._factorized_at = th

    def _model_space_predict(self, Xnew, pred_noise=False):
        self._ensure_factorized()
        mu, var = self._engine.predict(np.ascontiguousarray(Xnew, dtype=np.float64))
        if pred_noise:
            var = var + PYMC_SIGMA ** 2
        return mu, np.maximum(var, MIN_VARIANCE)

    @is_fitted
    def predict(self, covariates, diag=True, pred_noise=False):
        mu, var = self._model_space_predict(self.dm.Xnew(covariates), pred_noise)
        target = _assign_coords(self.dm.y_t(mu), covariates)
        se = _assign_coords(self.dm.se_t(var), covariates)
        return target, se

    @is_fitted
    def predict_grid(self, covariate: str, coord: Optional[str] = None, t_step: int = 12):
        if coord is None:
            coord = next(iter(self.dm.covariate_pipelines))
        coord_dim, covariate_dim = self.dm.get_dim(coord), self.dm.get_dim(covariate)
        x_max, x_min = self.dm.X.max(axis=0), self.dm.X.min(axis=0)
        n_cov = 18
        n_coord = int(np.round((x_max - x_min)[coord_dim] * t_step))
        x_coord = np.linspace(x_min[coord_dim], x_max[coord_dim], n_coord)
        x_cov = np.linspace(x_min[covariate_dim], x_max[covariate_dim], n_cov)
        X_grid = np.zeros((n_coord * n_cov, self.dm.X.shape[1]))
        X_grid[:, coord_dim] = np.repeat(x_coord, n_cov)     # pm.math.cartesian order: first axis slowest
        X_grid[:, covariate_dim] = np.tile(x_cov, n_coord)
        mu, _ = self._model_space_predict(X_grid, pred_noise=True)
        target = np.asarray(self.dm.y_t(mu)).reshape(n_coord, n_cov)
        index = self.dm.covariate_pipelines[coord].inverse_transform(x_coord)
        covs = self.dm.covariate_pipelines[covariate].inverse_transform(x_cov)
        return target, index, covs

    @is_fitted
    def sample(self, covariates, n=1000, diag=False, pred_noise=False, method="cholesky", tol=1e-6, seed=None):
        """Joint draws N(mu, cov) (engines/pymc.py:121-169, numpy multivariate_normal with method='cholesky')."""
        Xnew = np.ascontiguousarray(self.dm.Xnew(covariates), dtype=np.float64)
        self._ensure_factorized()
        Z = np.random.default_rng(seed).standard_normal((n, Xnew.shape[0]))
        for jit in (tol * 1e-2, tol, tol * 1e2):
            draws, info = self._engine.sample(Xnew, Z, jitter=jit + (PYMC_SIGMA ** 2 if pred_noise else 0.0))
            if info == 0:
                break
        else:
            raise NotPSDError("posterior covariance not positive definite")
        sim = np.asarray(self.dm.y_t(draws.reshape(-1))).reshape(n, -1)
        return sim


class LoadestGPMarginalPyMCB200(DataMixin, MarginalPyMCB200):
    """src/loadest_gp/models/pymc.py:12-90."""

    def __init__(self, model_config: Optional[ModelConfig] = None):
        if model_config is None:
            model_config = ModelConfig()
        super().__init__(model_config=model_config)
        self.build_datamanager(model_config)

    def build_datamanager(self, model_config: Optional[ModelConfig] = None):
        self._build_datamanager({"time": TimePipeline, "flow": LogStandardPipeline}, model_config)

    def build_model(self, X, y):
        return loadest_pymc_spec(X.shape[1]), loadest_pymc_vars(X.shape[1])
