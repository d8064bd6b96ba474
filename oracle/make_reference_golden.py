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


for name in ["xarray", "matplotlib", "matplotlib.pyplot", "matplotlib.dates", "matplotlib.ticker", "matplotlib.colors",
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
    out = os.path.join(os.path.dirname(HERE), "tests", "golden", "ref_models.json")
    with open(out, "w") as f:
        json.dump({"generator": "oracle/make_reference_golden.py", "reference": "thodson-usgs/discontinuum (src/ as found under "
                   "/root/reference), model and engine code unmodified, third-party layer = oracle/gpytorch_standin",
                   "cases": cases, "fits": fits}, f)
    for r in fits:
        print("fit", r["model"], r["optimizer"], "objective", r["history"][0], "->", r["history"][-1], "final lr", r["final_lr"])
    print("wrote", out, "cases", [(c["model"], c["case"], c["objective"]) for c in cases])
    for c in cases:
        print(c["model"], c["case"], len(c["parameter_names"]), "parameters:", c["parameter_names"])


if __name__ == "__main__":
    main()
