"""MarginalB200 -- drop-in engine base class with the surface of the reference's MarginalGPyTorch
(src/discontinuum/engines/gpytorch.py:36-626): fit / predict / predict_grid / sample / save / load.

Everything O(n^2) or O(n^3) -- covariance tiles, Cholesky, inverse, gradient contraction, cross-covariance,
posterior variance and sampling -- runs in libdgp.so on the GPU (float64).  The host keeps what the reference
keeps on the host anyway: the O(P) optimiser loop (torch.optim Adam/AdamW on P ~ 10-20 raw parameters,
ReduceLROnPlateau, clipping, NaN guards, early stopping: gpytorch.py:266-444), constraint transforms and priors.
"""
from __future__ import annotations

import dataclasses
import functools
import math
import os
import warnings
from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np
import torch

from . import capi, checkpoint
from .data import DataManager, LogErrorPipeline, LogStandardPipeline, StandardErrorPipeline, StandardPipeline
from .spec import CovSpec, GPModule

MIN_VARIANCE = 1e-10          # gpytorch settings.min_variance for float64
JITTERS = (0.0, 1e-8, 1e-7, 1e-6)  # psd_safe_cholesky retry ladder in float64
# torch's single-kernel Adam: the same update, a third of the host time of the per-tensor implementation on P ~ 10-20
# one-element parameters (0.2 against 0.7 ms per step), which is a large part of an iteration at n ~ 1000
_FUSED = {"fused": True}


def _reference_model_config_shim():
    """A dataclass that unpickles as `discontinuum.engines.base.ModelConfig` (same fields), for weights_only loading of
    checkpoints written by the reference."""
    shim = dataclasses.make_dataclass("ModelConfig", [("transform", str, dataclasses.field(default="log"))])
    shim.__module__, shim.__qualname__ = "discontinuum.engines.base", "ModelConfig"
    return shim


def _push_raw(params, raw_np) -> None:
    """numpy raw parameter vector -> the module's one-element tensors."""
    with torch.no_grad():
        for p, v in zip(params, raw_np):
            p.fill_(float(v))


def _import_adam_state(optimizer_obj, params):
    """(exp_avg, exp_avg_sq, step) of a torch Adam / AdamW over one-element parameters as numpy vectors and an int
    (zeros when the optimiser has not stepped yet, e.g. a fresh fit; the loaded values when resuming)."""
    m = np.zeros(len(params))
    v = np.zeros(len(params))
    step = 0
    for k, p in enumerate(params):
        st = optimizer_obj.state.get(p, None)
        if st:
            m[k] = float(st["exp_avg"].reshape(-1)[0])
            v[k] = float(st["exp_avg_sq"].reshape(-1)[0])
            step = max(step, int(round(float(st["step"]))))
    return m, v, step


def _export_adam_state(optimizer_obj, params, m, v, step) -> None:
    """The inverse of _import_adam_state: what optimizer.state_dict() / a later resume must see."""
    if step <= 0:
        return
    for k, p in enumerate(params):
        optimizer_obj.state[p] = {"step": torch.tensor(float(step), dtype=torch.float32),
                                  "exp_avg": torch.full_like(p.detach(), float(m[k])),
                                  "exp_avg_sq": torch.full_like(p.detach(), float(v[k]))}


@dataclass
class ModelConfig:
    """src/discontinuum/engines/base.py:22-26."""
    transform: str = "log"


def is_fitted(func):
    """src/discontinuum/engines/base.py:111-120."""

    @functools.wraps(func)
    def inner(self, *args, **kwargs):
        if not self.is_fitted:
            raise RuntimeError("The model hasn't been fitted yet, call .fit().")
        return func(self, *args, **kwargs)

    return inner


class NotPSDError(RuntimeError):
    pass


class _NLML(torch.autograd.Function):
    """NLML(natural theta) evaluated by the CUDA engine; backward feeds its analytic gradient to autograd."""

    @staticmethod
    def forward(ctx, nat: torch.Tensor, owner: "MarginalB200"):
        th = nat.detach().cpu().numpy().astype(np.float64)
        for jit in JITTERS:
            val, grad, info = owner._engine.nlml_grad(th, jit)
            if info == 0 and math.isfinite(val):
                break
        else:
            raise NotPSDError(f"Matrix not positive definite after adding jitter up to {JITTERS[-1]:g} (info={info})")
        owner._last_jitter = jit
        owner._last_theta = th.copy()  # the factorisation of this theta stays resident (U, alpha): penalty adjoints reuse it
        ctx.grad = torch.from_numpy(grad.copy())
        return torch.tensor(val, dtype=torch.float64)

    @staticmethod
    def backward(ctx, g):
        return g * ctx.grad, None


class DataMixin:
    """src/discontinuum/engines/base.py:85-108."""

    def _build_datamanager(self, covariate_pipelines: dict, model_config: Optional[ModelConfig] = None):
        if model_config is None:
            model_config = ModelConfig()
        if model_config.transform == "log":
            tp, ep = LogStandardPipeline, LogErrorPipeline
        elif model_config.transform == "standard":
            tp, ep = StandardPipeline, StandardErrorPipeline
        else:
            raise ValueError("Model config transform must be 'log' or 'standard'.")
        self.dm = DataManager(target_pipeline=tp, error_pipeline=ep, covariate_pipelines=covariate_pipelines)


class MarginalB200:
    device_index = 0
    max_predict_chunk = 2048

    def __init__(self, model_config=None):
        if model_config is None:
            model_config = {}
        self.model_config = model_config
        self.dm = None
        self.is_fitted = False
        self._resume_info = None
        self._last_optimizer = None
        self._last_scheduler = None
        self._current_iteration = 0
        self._engine: Optional[capi.Engine] = None
        self._factorized_at = None
        self._last_jitter = 0.0
        self._last_theta = None
        self.fixed_noise = None
        self.history = []

    # ------------------------------------------------------------------ hooks for model subclasses
    def build_model(self, X, y, y_unc=None) -> GPModule:
        """Return a GPModule (covariance spec + raw parameters) and set self.fixed_noise[n]."""
        raise NotImplementedError("This method must be implemented in a subclass")

    def project_parameters(self, X_all: np.ndarray):
        """In-place parameter projections a model applies on every forward (rating-gp); default none."""

    # ------------------------------------------------------------------ engine plumbing
    def _bind_engine(self):
        n = self.X.shape[0]
        if self._engine is None or self._engine.max_n < n or self._engine.max_m < self.max_predict_chunk:
            if self._engine is not None:
                self._engine.close()
            self._engine = capi.Engine(max_n=n, max_m=self.max_predict_chunk, device=self.device_index)
        self._engine.set_train(self.model.spec.to_c(), self.X, self.y, self.fixed_noise)
        self._factorized_at = None

    def _objective(self, penalty_callback=None, penalty_weight=0.0):
        """-[log N(y | m, Ky) + sum log prior] / n (+ penalty), gpytorch.py:353,371-373."""
        self.project_parameters(self.X)
        nat = self.model.natural()
        nll = (_NLML.apply(nat, self) - self.model.log_prior(nat)) / self.X.shape[0]
        penalty_val = None
        if penalty_callback is not None and penalty_weight > 0.0:
            try:
                penalty_val = penalty_callback()
                if not torch.is_tensor(penalty_val):
                    penalty_val = None
            except capi.DgpError:
                raise  # bad argument / CUDA error: not a numerical failure of this iteration
            except Exception as exc:  # noqa: BLE001
                if not getattr(self, "_penalty_warned", False):
                    warnings.warn(f"penalty callback failed ({exc!r}); training continues without the penalty term", stacklevel=2)
                    self._penalty_warned = True
                penalty_val = None
        objective = nll if penalty_val is None else nll + float(penalty_weight) * penalty_val
        return objective, penalty_val

    def _nlml_grad_ladder(self, th: np.ndarray):
        """NLML and its gradient at natural theta, with psd_safe_cholesky's jitter ladder."""
        for jit in JITTERS:
            val, grad, info = self._engine.nlml_grad(th, jit)
            if info == 0 and math.isfinite(val):
                break
        else:
            raise NotPSDError(f"Matrix not positive definite after adding jitter up to {JITTERS[-1]:g} (info={info})")
        self._last_jitter = jit
        self._last_theta = th.copy()
        return val, grad

    def _objective_closed_form(self):
        """The objective of `_objective` (no penalty) and its gradient w.r.t. the raw parameters without autograd:
        the GPU returns dNLML/dnatural, the constraint and prior chain rule is O(P) closed-form numpy
        (GPModule.host_chain).  Same numbers as the autograd path (tests/test_host.py); ~10x less host time per
        iteration, which is most of an iteration at n ~ 1000."""
        self.project_parameters(self.X)
        nat, dnat, lp, dlp = self.model.host_chain()
        val, g = self._nlml_grad_ladder(np.ascontiguousarray(nat))
        n = self.X.shape[0]
        return (val - lp) / n, (g - dlp) * dnat / n

    # ------------------------------------------------------------------ checkpointing (gpytorch.py:47-160)
    def save(self, f, optimizer_obj=None, scheduler=None, extra=None) -> None:
        if optimizer_obj is None:
            optimizer_obj = getattr(self, "_last_optimizer", None)
        if scheduler is None:
            scheduler = getattr(self, "_last_scheduler", None)
        if not hasattr(self, "model"):
            raise RuntimeError("No model to save. Call fit() first.")
        # model_state_dict / likelihood_state_dict carry the reference's gpytorch key names and shapes (checkpoint.py)
        sd, lik = checkpoint.to_reference_state(self.model)
        cfg = getattr(self, "model_config", None)
        ckpt = {
            "model_class": f"{self.__class__.__module__}.{self.__class__.__name__}",
            "model_state_dict": sd,
            "likelihood_state_dict": lik,
            # the reference's layout (its parameter order, vector parameters as one tensor), so that either engine resumes it
            "optimizer_state_dict": (checkpoint.optimizer_state_to_reference(self.model, optimizer_obj.state_dict())
                                     if optimizer_obj is not None else None),
            "optimizer_layout": "reference",
            "optimizer_name": _get_optimizer_name(optimizer_obj) if optimizer_obj is not None else None,
            "optimizer_lr": optimizer_obj.param_groups[0].get("lr") if optimizer_obj is not None else None,
            "scheduler_state_dict": scheduler.state_dict() if scheduler is not None else None,
            "scheduler_name": scheduler.__class__.__name__ if scheduler is not None else None,
            "current_iteration": getattr(self, "_current_iteration", 0),
            "model_config": dataclasses.asdict(cfg) if dataclasses.is_dataclass(cfg) else cfg,  # plain dict: loads under weights_only
            "extra": extra or {},
        }
        torch.save(ckpt, f)

    @classmethod
    def load(cls, f, covariates, target, target_unc=None):
        # tensors, numbers, strings and dicts only, no pickled code -- plus the one object the reference's save() pickles, its
        # ModelConfig dataclass (discontinuum/engines/base.py:22-26), resolved to a field-for-field stand-in by name
        with torch.serialization.safe_globals([_reference_model_config_shim()]):
            ckpt = torch.load(f, map_location="cpu", weights_only=True)
        cfg = ckpt.get("model_config")
        if dataclasses.is_dataclass(cfg) and not isinstance(cfg, type):
            cfg = dataclasses.asdict(cfg)
        model = cls(ModelConfig(**cfg)) if isinstance(cfg, dict) and cfg else cls()
        model.dm.fit(target=target, covariates=covariates, target_unc=target_unc)
        model.X, model.y = model.dm.X, model.dm.y
        if target_unc is None:
            model.model = model.build_model(model.X, model.y)
        else:
            model.y_unc = model.dm.y_unc
            model.model = model.build_model(model.X, model.y, model.y_unc)
        # reference-named state dicts (from MarginalB200.save or from the reference's MarginalGPyTorch.save) or round-1 native ones
        checkpoint.load_state(model.model, ckpt["model_state_dict"], ckpt.get("likelihood_state_dict"))
        model._resume_info = {k: ckpt.get(k) for k in ("optimizer_state_dict", "optimizer_name", "optimizer_lr",
                                                       "scheduler_state_dict", "scheduler_name")}
        model._resume_info["current_iteration"] = ckpt.get("current_iteration", 0)
        # checkpoints of the reference (no marker, model_class in its packages) and of this engine since round 2 hold the
        # optimiser state in the reference's layout; round-1 checkpoints of this engine in theta order
        native_class = str(ckpt.get("model_class", "")).startswith("discontinuum_b200")
        model._resume_info["optimizer_layout"] = ckpt.get("optimizer_layout", "native" if native_class else "reference")
        model._current_iteration = ckpt.get("current_iteration", 0)
        model._bind_engine()
        model.is_fitted = True
        return model

    # ------------------------------------------------------------------ fit (gpytorch.py:162-458)
    def fit(self, covariates, target, target_unc=None, iterations: int = 100, optimizer: Optional[str] = None,
            learning_rate: Optional[float] = None, early_stopping: bool = False, patience: int = 60, scheduler: bool = True,
            resume: bool = False, penalty_callback: Optional[Callable[[], torch.Tensor]] = None, penalty_weight: float = 0.0,
            progress: bool = False):
        has_existing_model = getattr(self, "model", None) is not None and self.is_fitted
        resuming_from_checkpoint = self._resume_info is not None and has_existing_model
        resuming_from_interruption = resume and has_existing_model and not resuming_from_checkpoint
        if not resuming_from_interruption:
            self.dm.fit(target=target, covariates=covariates, target_unc=target_unc)
        self.X, self.y = self.dm.X, self.dm.y
        can_restore = resuming_from_checkpoint or resuming_from_interruption
        if not can_restore:
            if target_unc is None:
                self.model = self.build_model(self.X, self.y)
            else:
                self.y_unc = self.dm.y_unc
                self.model = self.build_model(self.X, self.y, self.y_unc)
        self._bind_engine()

        resume_info = self._resume_info or {}
        if resuming_from_interruption and self._last_optimizer is not None:
            resume_info = {
                "optimizer_name": _get_optimizer_name(self._last_optimizer),
                "optimizer_lr": self._last_optimizer.param_groups[0]["lr"] if self._last_optimizer.param_groups else None,
                "optimizer_state_dict": self._last_optimizer.state_dict(),
                "scheduler_state_dict": self._last_scheduler.state_dict() if self._last_scheduler else None,
            }
        opt_name_saved, lr_saved = resume_info.get("optimizer_name"), resume_info.get("optimizer_lr")
        opt_choice = optimizer if optimizer is not None else (opt_name_saved or "adam")
        lr_choice = learning_rate if learning_rate is not None else (lr_saved or 0.05)
        params = self.model.raw_list()
        if opt_choice == "adamw" or (opt_name_saved and opt_name_saved.lower() == "adamw"):
            optimizer_obj = torch.optim.AdamW(params, lr=lr_choice, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, **_FUSED)
        elif opt_choice == "adam" or (opt_name_saved and opt_name_saved.lower() == "adam"):
            optimizer_obj = torch.optim.Adam(params, lr=lr_choice, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4, **_FUSED)
        else:
            raise ValueError(f"Unsupported optimizer: {opt_choice!r}. Supported optimizers are 'adam' and 'adamw'.")
        if can_restore and resume_info.get("optimizer_state_dict") is not None:
            try:
                osd = resume_info["optimizer_state_dict"]
                if resume_info.get("optimizer_layout") == "reference":
                    osd = checkpoint.optimizer_state_from_reference(self.model, osd)
                optimizer_obj.load_state_dict(osd)
            except Exception:  # noqa: BLE001, S110
                pass
        scheduler_obj = None
        if scheduler:
            scheduler_obj = torch.optim.lr_scheduler.ReduceLROnPlateau(
                optimizer_obj, mode="min", factor=0.7, patience=max(20, patience // 2), threshold=1e-4,
                threshold_mode="rel", min_lr=1e-6, cooldown=10)
            if can_restore and resume_info.get("scheduler_state_dict") is not None:
                try:
                    scheduler_obj.load_state_dict(resume_info["scheduler_state_dict"])
                except Exception:  # noqa: BLE001, S110
                    pass

        start_iteration = self._current_iteration if resume else 0
        remaining = iterations - start_iteration
        if remaining <= 0:
            print(f"Model already trained for {start_iteration} iterations (>= target {iterations}). No further training needed.")
            return
        best_obj, patience_counter, min_improvement, nan_loss_counter = float("inf"), 0, 1e-6, 0
        self.history = []
        i = 0
        fast = penalty_callback is None or not penalty_weight > 0.0
        # Numpy optimiser path (default whenever there is no penalty callback): raw parameters, Adam moments and the chain rule
        # live in numpy vectors during the loop -- no tensor, autograd graph or torch optimiser call per iteration (those cost
        # ~0.4 ms, as much as a whole NLML+gradient evaluation at n ~ 1000).  torch.optim.Adam / AdamW arithmetic; the torch
        # optimiser object still owns the learning rate (ReduceLROnPlateau drives it) and receives the moments back at the
        # end, so save / resume see the same state as on the torch path (DGP_HOST_OPT=torch selects that one).
        host_opt = fast and os.environ.get("DGP_HOST_OPT", "numpy") != "torch"
        if host_opt:
            has_proj = type(self).project_parameters is not MarginalB200.project_parameters
            raw_np = np.array([float(p.detach()) for p in params], dtype=np.float64)
            m_np, v_np, steps_np = _import_adam_state(optimizer_obj, params)
            group = optimizer_obj.param_groups[0]
            b1, b2 = group["betas"]
            eps_a, wd = float(group["eps"]), float(group["weight_decay"])
            decoupled = isinstance(optimizer_obj, torch.optim.AdamW)
            n_pts = self.X.shape[0]
        try:
            for i in range(remaining):
                self._current_iteration = start_iteration + i
                if not host_opt:
                    optimizer_obj.zero_grad(set_to_none=True)
                try:
                    if host_opt:
                        if has_proj:
                            _push_raw(params, raw_np)
                            self.project_parameters(self.X)
                            raw_np = np.array([float(p.detach()) for p in params], dtype=np.float64)
                        nat, dnat, lp, dlp = self.model.host_chain_raw(raw_np)
                        val, g = self._nlml_grad_ladder(np.ascontiguousarray(nat))
                        obj_value, graw = (val - lp) / n_pts, (g - dlp) * dnat / n_pts
                    elif fast:
                        obj_value, graw = self._objective_closed_form()
                    else:
                        objective, penalty_val = self._objective(penalty_callback, penalty_weight)
                        obj_value = float(objective.item())
                except capi.DgpError:
                    raise  # rc < 0 from libdgp (bad argument / CUDA error, possibly sticky): never a "NaN iteration"
                except Exception:
                    nan_loss_counter += 1
                    if nan_loss_counter > 10:
                        raise
                    continue
                if math.isnan(obj_value) or math.isinf(obj_value):
                    nan_loss_counter += 1
                    if nan_loss_counter > 10:
                        raise RuntimeError(f"Encountered more than 10 consecutive NaN/Inf objectives at iteration {i + 1}")
                    continue
                nan_loss_counter = 0
                if fast:
                    # clip_grad_norm_(max_norm=1.0) and the NaN guard of gpytorch.py:387-415 on the gradient vector
                    total = float(np.sqrt(np.sum(graw * graw)))
                    coef = 1.0 / (total + 1e-6)
                    if not coef >= 1.0:  # (a NaN norm scales everything to NaN, as torch's clamp does)
                        graw = graw * coef
                    if np.isnan(graw).any():
                        graw = np.nan_to_num(graw, nan=0.0, posinf=0.0, neginf=0.0)
                    if not host_opt:
                        for p, gv in zip(params, graw):
                            p.grad = torch.tensor([gv], dtype=torch.float64)
                else:
                    objective.backward()
                    torch.nn.utils.clip_grad_norm_(params, max_norm=1.0)
                    if any(p.grad is not None and torch.isnan(p.grad).any() for p in params):
                        for p in params:
                            if p.grad is not None:
                                p.grad = torch.nan_to_num(p.grad, nan=0.0, posinf=0.0, neginf=0.0)
                if host_opt:
                    lr_now = float(group["lr"])
                    steps_np += 1
                    if decoupled:
                        raw_np = raw_np * (1.0 - lr_now * wd)
                        gd = graw
                    else:
                        gd = graw + wd * raw_np
                    m_np = m_np + (gd - m_np) * (1.0 - b1)
                    v_np = v_np * b2 + (1.0 - b2) * gd * gd
                    bc1, bc2 = 1.0 - b1 ** steps_np, 1.0 - b2 ** steps_np
                    raw_np = raw_np - (lr_now / bc1) * (m_np / (np.sqrt(v_np) / math.sqrt(bc2) + eps_a))
                else:
                    optimizer_obj.step()
                obj_item = obj_value
                self.history.append(obj_item)
                if scheduler_obj is not None:
                    scheduler_obj.step(obj_item)
                if obj_item < best_obj - min_improvement:
                    best_obj, patience_counter = obj_item, 0
                else:
                    patience_counter += 1
                if progress:
                    print(f"iter {start_iteration + i + 1}: obj={obj_item:.6f} lr={optimizer_obj.param_groups[0]['lr']:.1e}")
                if early_stopping and patience_counter >= patience:
                    print(f"\nEarly stopping triggered after {i + 1} iterations")
                    print(f"Best objective: {best_obj:.6f}")
                    break
        except KeyboardInterrupt:
            print(f"\nTraining interrupted at iteration {i + 1}")
            print(f"Best objective: {best_obj:.6f}")
        finally:
            if host_opt:
                _push_raw(params, raw_np)
                _export_adam_state(optimizer_obj, params, m_np, v_np, steps_np)
            self.is_fitted = True
        self._last_optimizer, self._last_scheduler = optimizer_obj, scheduler_obj
        self._factorized_at = None
        return

    # ------------------------------------------------------------------ model-space prediction (gpytorch.py:599-626)
    def _theta(self) -> np.ndarray:
        with torch.no_grad():
            return self.model.natural().numpy().astype(np.float64)

    def _ensure_factorized(self, Xnew: np.ndarray):
        self.project_parameters(np.concatenate([self.X, Xnew], axis=0))
        th = self._theta()
        if self._factorized_at is None or not np.array_equal(self._factorized_at, th):
            for jit in JITTERS:
                val, info = self._engine.factorize(th, jit)
                if info == 0:
                    break
            else:
                raise NotPSDError(f"factorisation failed (info={info})")
            self._factorized_at = th
        return th

    def _model_space_predict(self, Xnew: np.ndarray):
        """mu, var of `likelihood(model(x))` in eval mode: latent variance + learned noise (+ fixed noise only when
        m == n, the shape coincidence of SURVEY A.5), clamped at MIN_VARIANCE."""
        Xnew = np.ascontiguousarray(Xnew, dtype=np.float64)
        th = self._ensure_factorized(Xnew)
        mu, var = self._engine.predict(Xnew, want_var=True)
        spec: CovSpec = self.model.spec
        if spec.noise_theta >= 0:
            var = var + th[spec.noise_theta]
        if Xnew.shape[0] == self.X.shape[0]:
            var = var + self.fixed_noise
        return mu, np.maximum(var, MIN_VARIANCE)

    @is_fitted
    def predict(self, covariates, diag=True, pred_noise=False):
        """(target, se) in original units; diag / pred_noise are accepted and ignored like the reference (gpytorch.py:460-501)."""
        mu, var = self._model_space_predict(self.dm.Xnew(covariates))
        target = _assign_coords(self.dm.y_t(mu), covariates)
        # (this repo's data manager wraps the result itself; the reference's has no se_t: engines/gpytorch.py:497)
        se_t = getattr(self.dm, "se_t", None)
        se = _assign_coords(se_t(var) if se_t is not None else self.dm.error_pipeline.inverse_transform(var), covariates)
        return target, se

    @is_fitted
    def predict_grid(self, covariate: str, coord: Optional[str] = None, t_step: int = 12):
        """18-column grid over (coord, covariate) in model space (gpytorch.py:503-549)."""
        from .data import _cov_dict, _xr

        names = list(_cov_dict(self.dm.data.covariates))
        if coord is None:
            coord = names[0]
        coord_dim, covariate_dim = self.dm.get_dim(coord), self.dm.get_dim(covariate)
        x_max, x_min = self.dm.X.max(axis=0), self.dm.X.min(axis=0)
        n_cov = 18
        n_coord = int(np.round((x_max - x_min)[coord_dim] * t_step))
        x_coord = np.linspace(x_min[coord_dim], x_max[coord_dim], n_coord)
        x_cov = np.linspace(x_min[covariate_dim], x_max[covariate_dim], n_cov)
        X_grid = np.stack(np.meshgrid(x_coord, x_cov, indexing="ij"), axis=-1).reshape(-1, 2)
        mu, _ = self._model_space_predict(X_grid)
        target = np.asarray(self.dm.y_t(mu)).reshape(n_coord, n_cov)
        index = self.dm.covariate_pipelines[coord].inverse_transform(x_coord)
        covs = self.dm.covariate_pipelines[covariate].inverse_transform(x_cov)
        if _xr is not None and isinstance(self.dm.data.target, _xr.DataArray):
            return _xr.DataArray(target, coords=[index, covs], dims=[coord, covariate], attrs=self.dm.data.target.attrs)
        return target, index, covs

    @is_fitted
    def sample(self, covariates, n: int = 1000, seed: Optional[int] = None):
        """n joint draws of the LATENT posterior at the covariates, in original units, shape [n, m] (gpytorch.py:551-593).
        Exact Cholesky root of the posterior covariance (the reference switches to a rank-100 Lanczos root above 800 points)."""
        Xnew = np.ascontiguousarray(self.dm.Xnew(covariates), dtype=np.float64)
        self._ensure_factorized(Xnew)
        gen = torch.Generator().manual_seed(seed) if seed is not None else None
        Z = torch.randn(n, Xnew.shape[0], dtype=torch.float64, generator=gen).numpy()
        sim = None
        for jit in JITTERS:  # psd_safe_cholesky ladder on the posterior covariance
            sim, info = self._engine.sample(Xnew, Z, jit)
            if info == 0 and np.all(np.isfinite(sim)):
                break
            sim = None
        if sim is None:
            raise NotPSDError("posterior covariance not positive definite")
        data = np.asarray(self.dm.y_t(sim.reshape(-1))).reshape(n, -1)
        from .data import _cov_dict, _xr

        if _xr is not None and isinstance(covariates, _xr.Dataset):
            return _xr.DataArray(data, coords=dict(covariates.coords, draw=np.arange(n)),
                                 dims=["draw"] + list(covariates.coords), attrs=getattr(self.dm.data.target, "attrs", {}))
        return data


def _assign_coords(arr, covariates):
    from .data import _xr

    if _xr is not None and isinstance(arr, _xr.DataArray) and isinstance(covariates, _xr.Dataset):
        return arr.assign_coords(covariates.coords)
    return arr


def _get_optimizer_name(optimizer_obj):
    if isinstance(optimizer_obj, torch.optim.AdamW):
        return "adamw"
    if isinstance(optimizer_obj, torch.optim.Adam):
        return "adam"
    return optimizer_obj.__class__.__name__
