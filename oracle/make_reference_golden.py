"""TEST INFRASTRUCTURE.  Generates tests/golden/ref_models.json by running the reference's OWN model-building code
(unmodified, imported from /root/reference/src) on the gpytorch stand-in of oracle/gpytorch_standin (see its README for what
that does and does not pin).  Run here, in the build container (the GPU box has no /root/reference):

    python oracle/make_reference_golden.py

Per model (loadest-gp: loadest_gp/models/gpytorch.py:48-128; rating-gp: rating_gp/models/gpytorch.py:28-41,64-79,205-372 with
rating_gp/models/kernels.py:242-382 and discontinuum/engines/gpytorch.py:31-33) at two parameter sets (as constructed / moved by
seeded offsets): inputs, the reference-named state dicts (the checkpoint keys of MarginalGPyTorch.save), the prior covariance
matrix covar_module(x), the mean vector, the training objective -ExactMarginalLogLikelihood(likelihood, model)(model(x), y)
(discontinuum/engines/gpytorch.py:318,353) and the latent posterior mean / variance of model(x*) in eval mode.
"""
import json
import os
import sys
import types
from unittest.mock import MagicMock

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("DISCONTINUUM_REFERENCE", "/root/reference/src")
sys.path.insert(0, os.path.join(HERE, "gpytorch_standin"))




def _stub(name):
    """An importable empty module whose every attribute is a MagicMock (the reference's mixins import xarray / matplotlib /
    dataretrieval at module level and never call them here); with a real __spec__, which torch's lazy imports look up."""
    import importlib.machinery
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, loader=None, is_package=True)
    m.__path__ = []
    m.__getattr__ = lambda attr: MagicMock(name=f"{name}.{attr}")
    return m


sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests"))
import fake_xarray as _fx  # noqa: E402  (the slice of xarray's API the reference's data manager and pipelines touch)

_xr = _stub("xarray")
_xr.DataArray, _xr.Dataset = _fx.DataArray, _fx.Dataset
sys.modules["xarray"] = _xr
for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.dates", "matplotlib.ticker", "matplotlib.colors",
             "matplotlib.cm", "dataretrieval", "dataretrieval.nwis"]:
    if name not in sys.modules:
        sys.modules[name] = _stub(name)
sys.path.insert(0, REF)

import numpy as np  # noqa: E402
import torch  # noqa: E402

torch.set_default_dtype(torch.float64)   # the reference builds float32 tensors; its formulas are evaluated in double here

import gpytorch  # noqa: E402  (the stand-in)

assert "gpytorch_standin" in gpytorch.__file__
import loadest_gp.models.gpytorch as ref_loadest  # noqa: E402
import loadest_gp.models.pymc as ref_loadest_pymc  # noqa: E402  (on the pymc stand-in)
import rating_gp.models.gpytorch as ref_rating  # noqa: E402


def tolist(t):
    return t.detach().cpu().numpy().tolist()


def state(module):
    return {k: tolist(v) for k, v in module.state_dict().items() if torch.is_tensor(v)}


def evaluate(owner, model, x, y, xs):
    model.train(); owner.likelihood.train()
    mll = gpytorch.mlls.ExactMarginalLogLikelihood(owner.likelihood, model)   # engines/gpytorch.py:318
    out = model(x)
    objective = -mll(out, y)                                                # engines/gpytorch.py:353
    rec = {"K": tolist(model.covar_module(x)), "mean": tolist(out.mean), "objective": float(objective),
           "state_dict": state(model), "likelihood_state_dict": state(owner.likelihood),
           "parameter_names": [n for n, _ in model.named_parameters()]}
    model.eval(); owner.likelihood.eval()
    with torch.no_grad():
        post = model(xs)                                                     # engines/gpytorch.py:621 (latent f*)
    rec["post_mean"], rec["post_var"] = tolist(post.mean), tolist(post.variance)
    return rec


def perturb(model, seed, skip=()):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in model.named_parameters():
            if any(s in name for s in skip):
                continue
            p.add_(0.35 * torch.randn(p.shape, generator=g))


class _FakeDM:
    """What MarginalGPyTorch.fit reads from its data manager (engines/gpytorch.py:218-233): model-space arrays."""
    def __init__(self, X, y, y_unc=None):
        self.X, self.y, self.y_unc = X, y, y_unc

    def fit(self, **kw):
        pass


def run_reference_fit(cls, X, y, y_unc, iterations, seed, **fit_kw):
    """The reference's own MarginalGPyTorch.fit (engines/gpytorch.py:162-458; rating-gp enters through its override,
    rating_gp/models/gpytorch.py:81-125) on model-space arrays that are exactly representable in float32 (the loop casts its
    inputs to float32).  Returns the state dicts right after build_model, the objective passed to the scheduler at every
    iteration, the learning rate at the end and the final state dicts."""
    obj = object.__new__(cls)          # (the constructor only builds the xarray data manager)
    obj._resume_info, obj._last_optimizer, obj._last_scheduler, obj._current_iteration, obj.is_fitted = None, None, None, 0, False
    obj.dm = _FakeDM(X, y, y_unc)
    rec = {}
    build = cls.build_model

    def build_and_record(*a):
        m = build(obj, *a)
        rec["initial_state_dict"], rec["initial_likelihood_state_dict"] = state(m), state(obj.likelihood)
        torch.manual_seed(seed + 1000)   # what the loop draws from here on (penalty grids) starts from a known generator state
        return m

    obj.build_model = build_and_record
    history = []
    orig_step = torch.optim.lr_scheduler.ReduceLROnPlateau.step

    def step(self, metrics, *a, **k):
        history.append(float(metrics))
        return orig_step(self, metrics, *a, **k)

    torch.optim.lr_scheduler.ReduceLROnPlateau.step = step
    try:
        torch.manual_seed(seed)
        obj.fit(covariates=None, target=None, target_unc=(True if y_unc is not None else None), iterations=iterations, **fit_kw)
    finally:
        torch.optim.lr_scheduler.ReduceLROnPlateau.step = orig_step
    rec.update({"history": history, "final_lr": obj._last_optimizer.param_groups[0]["lr"],
                "final_state_dict": state(obj.model), "final_likelihood_state_dict": state(obj.likelihood),
                "X": X.tolist(), "y": y.tolist(), "noise": tolist(obj.likelihood.noise), "iterations": iterations})
    return rec


def run_end_to_end(cls, covariates, target, target_unc, new_covariates, iterations, seed, grid_covariate):
    """The reference model class as a user drives it: constructor (data manager with the package's pipelines,
    discontinuum/data_manager.py + pipeline.py, on the xarray stand-in), fit(covariates, target[, target_unc]) and
    predict(new covariates) in data space.  The engine's explicit float32 casts (engines/gpytorch.py:221-222,235,488-491)
    are turned into float64 for the run, so that the vectors say what the reference's formulas give, not what float32 leaves."""
    history = []
    orig_step = torch.optim.lr_scheduler.ReduceLROnPlateau.step

    def step(self, metrics, *a, **k):
        history.append(float(metrics))
        return orig_step(self, metrics, *a, **k)

    torch.optim.lr_scheduler.ReduceLROnPlateau.step = step
    f32 = torch.float32
    torch.float32 = torch.float64
    try:
        torch.manual_seed(seed)
        m = cls()
        init = {}
        build = m.build_model

        def build_and_record(*a):
            mod = build(*a)
            init["sd"], init["lik"] = state(mod), state(m.likelihood)
            return mod

        m.build_model = build_and_record
        m.fit(covariates=covariates, target=target, target_unc=target_unc, iterations=iterations)
        pred, se = m.predict(new_covariates)
        grid = m.predict_grid(grid_covariate)          # engines/gpytorch.py:500-549
    finally:
        torch.float32 = f32
        torch.optim.lr_scheduler.ReduceLROnPlateau.step = orig_step
    rec = {"X_model": m.dm.X.tolist(), "y_model": m.dm.y.tolist(), "Xnew_model": m.dm.Xnew(new_covariates).tolist(),
           "initial_state_dict": init["sd"], "initial_likelihood_state_dict": init["lik"], "history": history,
           "iterations": iterations, "predict_target": np.asarray(pred.values).tolist(), "predict_se": np.asarray(se.values).tolist(),
           "predict_attrs": dict(pred.attrs), "predict_dims": list(pred.dims),
           "grid_covariate": grid_covariate, "grid_values": np.asarray(grid.values).tolist(), "grid_dims": list(grid.dims),
           "grid_index_ns": np.asarray(grid.coords[grid.dims[0]].values).astype("datetime64[ns]").astype(np.int64).tolist(),
           "grid_covariate_values": np.asarray(grid.coords[grid.dims[1]].values, dtype=np.float64).tolist()}
    if target_unc is not None:
        rec["y_unc_model"] = m.dm.y_unc.tolist()
    return rec


def f32_exact(a, bits=10):
    """round to multiples of 2^-bits: exactly representable in float32, so the loop's float32 cast changes nothing."""
    return np.round(np.asarray(a, dtype=np.float64) * 2 ** bits) / 2 ** bits


def main():
    rng = np.random.default_rng(20261018)
    cases = []
    # ---------------------------------------------------------------- loadest-gp
    n, m = 14, 5
    t = np.sort(rng.uniform(-1.5, 1.5, n))
    X = np.stack([t, rng.standard_normal(n)], axis=1)
    y = 0.4 * np.sin(2 * np.pi * t) + 0.3 * X[:, 1] + 0.1 * rng.standard_normal(n)
    Xs = np.stack([np.linspace(-1.4, 1.6, m), rng.standard_normal(m)], axis=1)
    x_t, y_t, xs_t = torch.tensor(X), torch.tensor(y), torch.tensor(Xs)
    owner = types.SimpleNamespace()
    torch.manual_seed(0)
    model = ref_loadest.LoadestGPMarginalGPyTorch.build_model(owner, x_t, y_t)   # the reference's own build_model body
    for tag, seed in (("initial", None), ("moved", 11)):
        if seed is not None:
            perturb(model, seed)
        rec = evaluate(owner, model, x_t, y_t, xs_t)
        rec.update({"model": "loadest", "case": tag, "X": X.tolist(), "y": y.tolist(), "Xs": Xs.tolist(),
                    "noise": tolist(owner.likelihood.noise)})
        cases.append(rec)
    # ... and with three covariate columns (ARD length scales are vectors: (1, 2) and (1, 3) tensors in the state dict)
    n = 12
    t = np.sort(rng.uniform(-1.5, 1.5, n))
    X = np.concatenate([t[:, None], rng.standard_normal((n, 2))], axis=1)
    y = 0.4 * np.sin(2 * np.pi * t) + 0.3 * X[:, 1] - 0.2 * X[:, 2] + 0.1 * rng.standard_normal(n)
    Xs = np.concatenate([np.linspace(-1.4, 1.6, m)[:, None], rng.standard_normal((m, 2))], axis=1)
    x_t, y_t, xs_t = torch.tensor(X), torch.tensor(y), torch.tensor(Xs)
    owner = types.SimpleNamespace()
    model = ref_loadest.LoadestGPMarginalGPyTorch.build_model(owner, x_t, y_t)
    perturb(model, 13)
    rec = evaluate(owner, model, x_t, y_t, xs_t)
    rec.update({"model": "loadest", "case": "moved, 3 input columns", "X": X.tolist(), "y": y.tolist(), "Xs": Xs.tolist(),
                "noise": tolist(owner.likelihood.noise)})
    cases.append(rec)
    # ---------------------------------------------------------------- rating-gp
    n = 16
    t = np.sort(rng.uniform(-1.5, 1.5, n))
    stage = 1.0 + rng.uniform(0.02, 1.0, n)                      # the reference scales stage to [1, 2]
    X = np.stack([t, stage], axis=1)
    y = 0.8 + 1.6 * np.log(stage - 0.4) + 0.05 * rng.standard_normal(n)
    y_unc = rng.choice(np.array([0.02, 0.05, 0.08]), n) ** 2
    Xs = np.stack([np.linspace(-1.4, 1.6, m), 1.0 + rng.uniform(0.05, 0.95, m)], axis=1)
    x_t, y_t, xs_t, u_t = torch.tensor(X), torch.tensor(y), torch.tensor(Xs), torch.tensor(y_unc)
    owner = types.SimpleNamespace()
    torch.manual_seed(3)
    model = ref_rating.RatingGPMarginalGPyTorch.build_model(owner, x_t, y_t, u_t)
    with torch.no_grad():   # keep the power-law parameters inside their clamps so that forward() does not move them
        model.powerlaw.b.fill_(1.7); model.powerlaw.c.fill_(0.35)
    for tag, seed in (("initial", None), ("moved", 12)):
        if seed is not None:
            perturb(model, seed, skip=("powerlaw.b", "powerlaw.c"))
            with torch.no_grad():
                model.powerlaw.b.fill_(2.1); model.powerlaw.c.fill_(0.6)
        rec = evaluate(owner, model, x_t, y_t, xs_t)
        rec.update({"model": "rating", "case": tag, "X": X.tolist(), "y": y.tolist(), "Xs": Xs.tolist(),
                    "noise": tolist(owner.likelihood.noise), "second_noise": tolist(owner.likelihood.second_noise),
                    "gate_a": float(model.covar_module.kernels[0].kernels[0].a)})
        cases.append(rec)
    # ... and without a target uncertainty: the default noise of build_model (rating_gp/models/gpytorch.py:69-70)
    owner = types.SimpleNamespace()
    torch.manual_seed(8)
    model = ref_rating.RatingGPMarginalGPyTorch.build_model(owner, x_t, y_t)
    with torch.no_grad():
        model.powerlaw.b.fill_(1.9); model.powerlaw.c.fill_(0.2)
    rec = evaluate(owner, model, x_t, y_t, xs_t)
    rec.update({"model": "rating", "case": "initial, default noise", "X": X.tolist(), "y": y.tolist(), "Xs": Xs.tolist(),
                "noise": tolist(owner.likelihood.noise), "second_noise": tolist(owner.likelihood.second_noise),
                "gate_a": float(model.covar_module.kernels[0].kernels[0].a)})
    cases.append(rec)
    # ---------------------------------------------------------------- the reference's own optimiser loop
    fits = []
    n = 24
    t = f32_exact(np.sort(rng.uniform(-1.5, 1.5, n)))
    X = np.stack([t, f32_exact(rng.standard_normal(n))], axis=1)
    y = f32_exact(0.4 * np.sin(2 * np.pi * t) + 0.3 * X[:, 1] + 0.1 * rng.standard_normal(n))
    for opt in ("adam", "adamw"):
        rec = run_reference_fit(ref_loadest.LoadestGPMarginalGPyTorch, X, y, None, 30, seed=1, optimizer=opt)
        rec.update({"model": "loadest", "optimizer": opt})
        fits.append(rec)
    stage = f32_exact(1.0 + rng.uniform(0.02, 1.0, n))
    X = np.stack([t, stage], axis=1)
    y = f32_exact(0.8 + 1.6 * np.log(stage - 0.4) + 0.05 * rng.standard_normal(n))
    y_unc = rng.choice(np.array([2.0 ** -10, 2.0 ** -9, 2.0 ** -8]), n)
    rec = run_reference_fit(ref_rating.RatingGPMarginalGPyTorch, X, y, y_unc, 30, seed=5)
    rec.update({"model": "rating", "optimizer": "adam", "y_unc": y_unc.tolist()})
    fits.append(rec)
    # ... and with the monotonic-rating penalty of rating_gp/models/gpytorch.py:126-187 (random grids, finite differences)
    rec = run_reference_fit(ref_rating.RatingGPMarginalGPyTorch, X, y, y_unc, 12, seed=5, monotonic_penalty_weight=0.5, grid_size=16)
    rec.update({"model": "rating", "optimizer": "adam", "y_unc": y_unc.tolist(), "penalty_weight": 0.5, "grid_size": 16,
                "loop_seed": 1005})
    fits.append(rec)
    # ---------------------------------------------------------------- data space, end to end (data manager + pipelines too)
    e2e = []
    n, m = 40, 9
    days = np.sort(rng.uniform(0, 3650, n))
    time = np.datetime64("2000-01-01") + (days * 86400e9).astype("timedelta64[ns]")
    flow = np.exp(1.0 + 0.8 * np.sin(2 * np.pi * days / 365.25) + 0.5 * rng.standard_normal(n))
    conc = np.exp(0.3 * np.log(flow) + 0.2 * np.cos(2 * np.pi * days / 365.25) + 0.2 * rng.standard_normal(n))
    new_days = np.linspace(30, 3600, m)
    new_time = np.datetime64("2000-01-01") + (new_days * 86400e9).astype("timedelta64[ns]")
    new_flow = np.exp(1.0 + 0.8 * np.sin(2 * np.pi * new_days / 365.25))
    cov = _fx.Dataset({"flow": ("time", flow)}, coords={"time": time})
    tgt = _fx.DataArray(conc, coords={"time": time}, dims=("time",), attrs={"units": "mg/L"}, name="conc")
    new = _fx.Dataset({"flow": ("time", new_flow)}, coords={"time": new_time})
    rec = run_end_to_end(ref_loadest.LoadestGPMarginalGPyTorch, cov, tgt, None, new, 12, seed=2, grid_covariate="flow")
    rec.update({"model": "loadest", "time_ns": time.astype("datetime64[ns]").astype(np.int64).tolist(), "flow": flow.tolist(),
                "target": conc.tolist(), "new_time_ns": new_time.astype("datetime64[ns]").astype(np.int64).tolist(),
                "new_flow": new_flow.tolist()})
    e2e.append(rec)
    stage = rng.lognormal(1.0, 0.5, n)
    q = 3.0 * (stage - 0.5 * stage.min()) ** 1.6 * np.exp(0.03 * rng.standard_normal(n))
    gse = rng.choice(np.array([1.02, 1.05, 1.08]), n)
    new_stage = np.exp(np.linspace(np.log(stage.min() * 1.05), np.log(stage.max() * 0.95), m))
    cov = _fx.Dataset({"stage": ("time", stage)}, coords={"time": time})
    tgt = _fx.DataArray(q, coords={"time": time}, dims=("time",), attrs={"units": "cfs"}, name="discharge")
    unc = _fx.DataArray(gse, coords={"time": time}, dims=("time",), name="gse")
    new = _fx.Dataset({"stage": ("time", new_stage)}, coords={"time": new_time})
    rec = run_end_to_end(ref_rating.RatingGPMarginalGPyTorch, cov, tgt, unc, new, 12, seed=4, grid_covariate="stage")
    rec.update({"model": "rating", "time_ns": time.astype("datetime64[ns]").astype(np.int64).tolist(), "stage": stage.tolist(),
                "target": q.tolist(), "target_unc": gse.tolist(),
                "new_time_ns": new_time.astype("datetime64[ns]").astype(np.int64).tolist(), "new_stage": new_stage.tolist()})
    e2e.append(rec)
    # ---------------------------------------------------------------- a checkpoint written by the reference's own save()
    ckpt_path = os.path.join(os.path.dirname(HERE), "tests", "golden", "ref_checkpoint_loadest.pt")
    cov = _fx.Dataset({"flow": ("time", flow)}, coords={"time": time})
    tgt = _fx.DataArray(conc, coords={"time": time}, dims=("time",), attrs={"units": "mg/L"}, name="conc")
    history = []
    orig_step = torch.optim.lr_scheduler.ReduceLROnPlateau.step

    def step(self, metrics, *a, **k):
        history.append(float(metrics))
        return orig_step(self, metrics, *a, **k)

    torch.optim.lr_scheduler.ReduceLROnPlateau.step = step
    f32 = torch.float32
    torch.float32 = torch.float64
    try:
        torch.manual_seed(6)
        m1 = ref_loadest.LoadestGPMarginalGPyTorch()
        m1.fit(covariates=cov, target=tgt, iterations=8)
        m1.save(ckpt_path)                                        # engines/gpytorch.py:107-160
        first = list(history)
        del history[:]
        # (the reference's load() calls torch.load with the defaults, which on torch >= 2.6 refuse the pickled ModelConfig
        # dataclass the reference's save() puts into the file: allow-list it, as the error message asks)
        import discontinuum.engines.base as ref_base
        torch.serialization.add_safe_globals([ref_base.ModelConfig])
        m2 = ref_loadest.LoadestGPMarginalGPyTorch.load(ckpt_path, cov, tgt)   # engines/gpytorch.py:47-105
        m2.fit(covariates=cov, target=tgt, iterations=18, resume=True)
    finally:
        torch.float32 = f32
        torch.optim.lr_scheduler.ReduceLROnPlateau.step = orig_step
    ckpt_rec = {"file": "ref_checkpoint_loadest.pt", "first_history": first, "resumed_history": list(history),
                "time_ns": time.astype("datetime64[ns]").astype(np.int64).tolist(), "flow": flow.tolist(), "target": conc.tolist(),
                "total_iterations": 18}
    print("checkpoint: first", len(first), "iterations, resumed", len(history), "more:", history[0], "->", history[-1])
    # ... and for rating-gp, whose optimiser sees the parameters in a different order than this engine's theta (likelihood first)
    ckpt_path_r = os.path.join(os.path.dirname(HERE), "tests", "golden", "ref_checkpoint_rating.pt")
    cov_r = _fx.Dataset({"stage": ("time", stage)}, coords={"time": time})
    tgt_r = _fx.DataArray(q, coords={"time": time}, dims=("time",), attrs={"units": "cfs"}, name="discharge")
    unc_r = _fx.DataArray(gse, coords={"time": time}, dims=("time",), name="gse")
    del history[:]
    torch.optim.lr_scheduler.ReduceLROnPlateau.step = step
    torch.float32 = torch.float64
    try:
        torch.manual_seed(9)
        r1 = ref_rating.RatingGPMarginalGPyTorch()
        r1.fit(covariates=cov_r, target=tgt_r, target_unc=unc_r, iterations=8)
        r1.save(ckpt_path_r)
        first_r = list(history)
        del history[:]
        r2 = ref_rating.RatingGPMarginalGPyTorch.load(ckpt_path_r, cov_r, tgt_r, unc_r)
        r2.fit(covariates=cov_r, target=tgt_r, target_unc=unc_r, iterations=18, resume=True)
    finally:
        torch.float32 = f32
        torch.optim.lr_scheduler.ReduceLROnPlateau.step = orig_step
    ckpt_rating = {"file": "ref_checkpoint_rating.pt", "first_history": first_r, "resumed_history": list(history),
                   "time_ns": time.astype("datetime64[ns]").astype(np.int64).tolist(), "stage": stage.tolist(), "target": q.tolist(),
                   "target_unc": gse.tolist(), "total_iterations": 18}
    print("rating checkpoint: first", len(first_r), "iterations, resumed", len(history), "more:", history[0], "->", history[-1])
    # ---------------------------------------------------------------- the PyMC model (loadest_gp/models/pymc.py:30-88)
    pymc_cases = []
    for nd in (2, 3):
        n = 13
        Xp = np.concatenate([np.sort(rng.uniform(-1.5, 1.5, n))[:, None], rng.standard_normal((n, nd - 1))], axis=1)
        yp = rng.standard_normal(n)
        owner = types.SimpleNamespace()
        pm_model = ref_loadest_pymc.LoadestGPMarginalPyMC.build_model(owner, Xp, yp)
        table = [{"name": k, "prior": v.kind, "params": v.params, "size": 1 if v.shape is None else int(np.prod(v.shape)),
                  "initval": None if v.initval is None else np.atleast_1d(np.asarray(v.initval, dtype=np.float64)).tolist()}
                 for k, v in pm_model.vars.items()]
        for tag in ("a", "b"):
            vals = {t["name"]: (0.3 + 1.2 * rng.uniform(size=t["size"])).tolist() for t in table}
            nl = float(pm_model.neg_logp(vals))
            pymc_cases.append({"ndim": nd, "case": tag, "X": Xp.tolist(), "y": yp.tolist(), "variables": table, "values": vals,
                               "neg_logp": nl, "K": tolist(owner.gp.cov_func(torch.tensor(Xp)))})
    print("pymc", [(c["ndim"], c["case"], c["neg_logp"]) for c in pymc_cases])
    out = os.path.join(os.path.dirname(HERE), "tests", "golden", "ref_models.json")
    with open(out, "w") as f:
        json.dump({"generator": "oracle/make_reference_golden.py", "checkpoint": ckpt_rec, "checkpoint_rating": ckpt_rating,
                   "pymc": pymc_cases, "reference": "thodson-usgs/discontinuum (src/ as found under "
                   "/root/reference), model, engine, data-manager and pipeline code unmodified; third-party layers = "
                   "oracle/gpytorch_standin and tests/fake_xarray.py",
                   "cases": cases, "fits": fits, "end_to_end": e2e}, f)
    for r in e2e:
        print("end to end", r["model"], "objective", r["history"][0], "->", r["history"][-1], "target[:3]", r["predict_target"][:3],
              "se[:3]", r["predict_se"][:3])
    for r in fits:
        print("fit", r["model"], r["optimizer"], "objective", r["history"][0], "->", r["history"][-1], "final lr", r["final_lr"])
    print("wrote", out, "cases", [(c["model"], c["case"], c["objective"]) for c in cases])
    for c in cases:
        print(c["model"], c["case"], len(c["parameter_names"]), "parameters:", c["parameter_names"])


if __name__ == "__main__":
    main()
