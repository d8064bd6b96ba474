"""Vectors produced by the REFERENCE's own model-building code (tests/golden/ref_models.json, made by
oracle/make_reference_golden.py in the build container, where /root/reference is readable): the unmodified `ExactGPModel`
classes, custom kernels, `PowerLawTransform` and `NoOpMean` of loadest-gp / rating-gp, executed on the gpytorch stand-in of
oracle/gpytorch_standin.  They pin what the reference owns -- which kernels act on which dimensions with which priors,
constraints and initial values, the gate and its inversion, the log warp, the power-law mean, the noise model, and the module
tree whose parameter paths are the checkpoint keys -- for the oracle, for the checkpoint key mapping and (GPU) for the engine.
"""
import json
import os

import numpy as np
import pytest
import torch

import helpers as H
from helpers import orc
from discontinuum_b200 import checkpoint, models, spec

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_models.json")
RTOL = 1e-10


def _cases():
    with open(GOLD) as f:
        return json.load(f)["cases"]


def _module_from_reference_state(case):
    """GPModule of this repo filled from the reference-named state dicts (checkpoint.load_state = MarginalB200.load's path)."""
    X = np.array(case["X"])
    if case["model"] == "loadest":
        mod = spec.GPModule(models.loadest_spec(X.shape[1]))
    else:
        b_lo, b_hi = models.stage_quantile_bounds(X[:, 1])
        mod = spec.GPModule(models.rating_spec(b_lo, b_hi))
    sd = {k: torch.tensor(v, dtype=torch.float64) for k, v in case["state_dict"].items()}
    lik = {k: torch.tensor(v, dtype=torch.float64) for k, v in case["likelihood_state_dict"].items()}
    checkpoint.load_state(mod, sd, lik)
    return mod


def _oracle_pieces(case, mod):
    X, y, noise = (torch.tensor(np.array(case[k], dtype=np.float64)) for k in ("X", "y", "noise"))
    theta = mod.natural().detach().numpy()
    if case["model"] == "loadest":
        nat = H.loadest_nat_from_theta(theta, X.shape[1])
        return X, y, noise, theta, nat, orc.loadest_cov, orc.loadest_mean, None, orc.loadest_log_prior(nat)
    nat = H.rating_nat_from_theta(theta)
    return X, y, noise, theta, nat, orc.rating_cov, orc.rating_mean, nat["noise"], orc.rating_log_prior(nat)


@pytest.mark.parametrize("idx", range(6))
def test_oracle_matches_reference_model_code(idx):
    case = _cases()[idx]
    mod = _module_from_reference_state(case)
    # every learnable parameter of the reference's module tree is consumed by the key mapping, none is left over
    table, aliases = checkpoint._table(mod)
    known = {k for k, _ in table.values()} | {a for v in aliases.values() for a in v}
    assert set(case["parameter_names"]) <= known, set(case["parameter_names"]) - known
    X, y, noise, theta, nat, cov, mean, extra, lp = _oracle_pieces(case, mod)
    K_ref, m_ref = np.array(case["K"]), np.array(case["mean"])
    K = cov(X, X, nat).numpy()
    assert np.max(np.abs(K - K_ref)) <= RTOL * np.max(np.abs(K_ref)), case["case"]
    assert np.max(np.abs(mean(X, nat).numpy() - m_ref)) <= RTOL * max(1.0, np.max(np.abs(m_ref)))
    n = X.shape[0]
    val = orc.nlml(cov, mean, nat, X, y, noise, extra_noise=extra)
    obj = float((val - lp) / n)
    assert abs(obj - case["objective"]) <= RTOL * abs(case["objective"]), (obj, case["objective"])
    # the host-side objective of the engine class (constraints + priors of GPModule) says the same
    assert abs(float(((val - mod.log_prior(mod.natural())) / n).detach()) - case["objective"]) <= RTOL * abs(case["objective"])
    # latent posterior at new points (engines/gpytorch.py:621: model(x*) in eval mode)
    Xs = torch.tensor(np.array(case["Xs"]))
    mu, _, var_lat = orc.predict(cov, mean, nat, X, y, noise, Xs, extra_noise=extra, min_variance=0.0)
    assert np.max(np.abs(mu.numpy() - np.array(case["post_mean"]))) <= 1e-8 * max(1.0, np.max(np.abs(case["post_mean"])))
    assert np.max(np.abs(var_lat.numpy() - np.array(case["post_var"]))) <= 1e-8 * max(1e-3, np.max(np.abs(case["post_var"])))
    if case["model"] == "rating":
        # constants the reference fixes in code: gate sharpness, switch-point interval from the stage quantiles, noise floor
        assert case["gate_a"] == 20.0
        sd = case["state_dict"]
        b_lo, b_hi = models.stage_quantile_bounds(np.array(case["X"])[:, 1])
        assert abs(sd["covar_module.kernels.0.kernels.0.raw_b_constraint.lower_bound"] - b_lo) <= 1e-15
        assert abs(sd["covar_module.kernels.0.kernels.0.raw_b_constraint.upper_bound"] - b_hi) <= 1e-15
        assert abs(float(nat["noise"]) - case["second_noise"][0]) <= 1e-14
        if "default noise" in case["case"]:
            assert np.allclose(np.array(case["noise"]), models.RATING_DEFAULT_NOISE, rtol=0, atol=1e-18)


def test_reference_state_round_trips_through_the_checkpoint_mapping():
    """to_reference_state(load_state(reference state)) reproduces every learnable tensor of the reference's state dicts,
    names and shapes included (what MarginalB200.save writes is what MarginalGPyTorch.load_state_dict reads)."""
    for case in _cases():
        mod = _module_from_reference_state(case)
        sd, lik = checkpoint.to_reference_state(mod)
        for name in case["parameter_names"]:
            ref = np.array(case["state_dict"][name])
            assert name in sd, name
            got = sd[name].numpy()
            assert got.shape == ref.shape, (name, got.shape, ref.shape)
            assert np.max(np.abs(got - ref)) <= 1e-12 * max(1.0, np.max(np.abs(ref))), name
        for name, v in case["likelihood_state_dict"].items():
            if name.endswith("raw_noise"):
                assert np.allclose(lik[name].numpy(), np.array(v), rtol=0, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("idx", range(6))
def test_engine_matches_reference_model_code(cuda_device, idx):
    from discontinuum_b200 import capi

    case = _cases()[idx]
    mod = _module_from_reference_state(case)
    X, y, noise = (np.array(case[k], dtype=np.float64) for k in ("X", "y", "noise"))
    theta = mod.natural().detach().numpy()
    eng = capi.Engine(max_n=X.shape[0], max_m=128)
    eng.set_train(mod.spec.to_c(), X, y, noise)
    K_ref = np.array(case["K"])
    assert np.max(np.abs(eng.covmat(theta) - K_ref)) <= 1e-12 * np.max(np.abs(K_ref))   # covar_module(x), no noise
    val, info = eng.nlml(theta)
    n = X.shape[0]
    obj = (val - float(mod.log_prior(mod.natural()).detach())) / n
    assert info == 0 and abs(obj - case["objective"]) <= 1e-9 * abs(case["objective"])
    eng.factorize(theta)
    mu, var = eng.predict(np.array(case["Xs"]))
    assert np.max(np.abs(mu - np.array(case["post_mean"]))) <= 1e-8 * max(1.0, np.max(np.abs(case["post_mean"])))
    assert np.max(np.abs(var - np.array(case["post_var"]))) <= 1e-8 * max(1e-3, np.max(np.abs(case["post_var"])))
    eng.close()


def _fits():
    with open(GOLD) as f:
        return json.load(f)["fits"]


def _raw_dict_from_state(fit, sd_key, lik_key):
    """Reference-named state dicts -> (GPModule, the oracle's raw parameter dict) through the checkpoint key mapping."""
    X = np.array(fit["X"])
    if fit["model"] == "loadest":
        mod = spec.GPModule(models.loadest_spec(X.shape[1]))
    else:
        mod = spec.GPModule(models.rating_spec(*models.stage_quantile_bounds(X[:, 1])))
    checkpoint.load_state(mod, {k: torch.tensor(v, dtype=torch.float64) for k, v in fit[sd_key].items()},
                          {k: torch.tensor(v, dtype=torch.float64) for k, v in fit[lik_key].items()})
    rawvec = np.array([float(p.detach()) for p in mod.raw_list()])
    conv = (lambda v: H.loadest_nat_from_theta(v, X.shape[1])) if fit["model"] == "loadest" else H.rating_nat_from_theta
    return mod, {k: v.clone() for k, v in conv(rawvec).items()}   # (same layout as the natural vector, raw values inside)


@pytest.mark.parametrize("idx", range(4))
def test_oracle_loop_matches_the_reference_fit_loop(idx):
    """The reference's own `MarginalGPyTorch.fit` (engines/gpytorch.py:162-458: Adam / AdamW settings, global-norm clipping,
    ReduceLROnPlateau, the rating model's in-forward clamps) run on the stand-in for 30 iterations, against the oracle's
    restatement of that loop (`fit_adam`, which the GPU tests hold `MarginalB200.fit` to): the objective at every iteration
    and every parameter at the end."""
    fit = _fits()[idx]
    X, y, noise = (torch.tensor(np.array(fit[k], dtype=np.float64)) for k in ("X", "y", "noise"))
    mod, raw = _raw_dict_from_state(fit, "initial_state_dict", "initial_likelihood_state_dict")
    kw = {}
    if fit["model"] == "rating":
        b_lo, b_hi = models.stage_quantile_bounds(np.array(fit["X"])[:, 1])
        kw = dict(b_lo=b_lo, b_hi=b_hi, h_min=float(np.array(fit["X"])[:, 1].min()))
    if "penalty_weight" in fit:   # the monotonic-rating penalty (rating_gp/models/gpytorch.py:126-187) on the same random grids
        kw.update(penalty_weight=fit["penalty_weight"], grid_size=fit["grid_size"])
        torch.manual_seed(fit["loop_seed"])
    raw_end, hist = orc.fit_adam(fit["model"], raw, X, y, noise, iterations=fit["iterations"], optimizer=fit["optimizer"], **kw)
    ref = np.array(fit["history"])
    assert len(hist) == len(ref) == fit["iterations"]
    assert np.max(np.abs(np.array(hist) - ref) / np.abs(ref)) <= 1e-8, np.max(np.abs(np.array(hist) - ref) / np.abs(ref))
    _, raw_ref = _raw_dict_from_state(fit, "final_state_dict", "final_likelihood_state_dict")
    for k, v in raw_ref.items():
        got = raw_end[k].detach().numpy()
        assert np.max(np.abs(got - v.numpy())) <= 1e-7 * max(1.0, np.max(np.abs(v.numpy()))), k


class _ModelSpaceDM:
    """What MarginalB200.fit (like MarginalGPyTorch.fit) reads from its data manager, already in model space."""
    def __init__(self, X, y, y_unc=None):
        self.X, self.y, self.y_unc = X, y, y_unc

    def fit(self, **kw):
        pass


@pytest.mark.gpu
@pytest.mark.parametrize("idx", range(4))
def test_engine_fit_matches_the_reference_fit_loop(cuda_device, idx):
    """MarginalB200.fit on the GPU, started from the reference's initial parameters, against the objective trajectory of the
    reference's own fit loop (tests/golden/ref_models.json["fits"]): <= 1e-6 relative at every one of the 30 iterations."""
    fit = _fits()[idx]
    X, y = np.array(fit["X"]), np.array(fit["y"])
    y_unc = np.array(fit["y_unc"]) if fit["model"] == "rating" else None
    base = models.LoadestGP if fit["model"] == "loadest" else models.RatingGP

    class _FromReferenceState(base):
        def build_model(self, *a):
            mod = super().build_model(*a)
            checkpoint.load_state(mod, {k: torch.tensor(v, dtype=torch.float64) for k, v in fit["initial_state_dict"].items()},
                                  {k: torch.tensor(v, dtype=torch.float64) for k, v in fit["initial_likelihood_state_dict"].items()})
            if "loop_seed" in fit:
                torch.manual_seed(fit["loop_seed"])   # the penalty grids are drawn from torch's generator, as in the reference
            return mod

    m = _FromReferenceState()
    m.dm = _ModelSpaceDM(X, y, y_unc)
    extra = dict(monotonic_penalty_weight=fit["penalty_weight"], grid_size=fit["grid_size"]) if "penalty_weight" in fit else {}
    m.fit(None, None, target_unc=(True if y_unc is not None else None), iterations=fit["iterations"], optimizer=fit["optimizer"],
          **extra)
    ref = np.array(fit["history"])
    got = np.array(m.history)
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref) / np.abs(ref)) <= 1e-6, np.max(np.abs(got - ref) / np.abs(ref))


def _e2e():
    with open(GOLD) as f:
        return json.load(f)["end_to_end"]


def _raw_inputs(rec):
    time = np.array(rec["time_ns"], dtype="int64").astype("datetime64[ns]")
    new_time = np.array(rec["new_time_ns"], dtype="int64").astype("datetime64[ns]")
    key = "flow" if rec["model"] == "loadest" else "stage"
    cov = {"time": time, key: np.array(rec[key])}
    new = {"time": new_time, key: np.array(rec["new_" + key])}
    unc = np.array(rec["target_unc"]) if "target_unc" in rec else None
    return cov, np.array(rec["target"]), unc, new


@pytest.mark.parametrize("idx", range(2))
def test_data_manager_matches_the_reference_pipelines(idx):
    """Model-space design matrix, target, target uncertainty and new-point design matrix of the reference's own DataManager
    and pipelines (discontinuum/data_manager.py, pipeline.py, the packages' data mixins; run on the xarray stand-in) against
    this repo's array-level data manager on the same raw inputs."""
    rec = _e2e()[idx]
    cov, target, unc, new = _raw_inputs(rec)
    m = models.LoadestGP() if rec["model"] == "loadest" else models.RatingGP()
    m.dm.fit(target=target, covariates=cov, target_unc=unc)
    assert np.max(np.abs(m.dm.X - np.array(rec["X_model"]))) <= 1e-11
    assert np.max(np.abs(m.dm.y - np.array(rec["y_model"]))) <= 1e-12
    assert np.max(np.abs(m.dm.Xnew(new) - np.array(rec["Xnew_model"]))) <= 1e-11
    if unc is not None:
        assert np.max(np.abs(m.dm.y_unc - np.array(rec["y_unc_model"]))) <= 1e-13


@pytest.mark.gpu
@pytest.mark.parametrize("idx", range(2))
def test_model_classes_match_the_reference_end_to_end(cuda_device, idx):
    """The user-level flow -- Model().fit(covariates, target[, target_unc], iterations=12) then predict(new covariates), raw
    data in, data-space prediction and standard error out -- against the reference's own classes run end to end
    (constructor, data manager, fit loop, predict; engines/gpytorch.py:162-458,460-499), from the same initial parameters."""
    rec = _e2e()[idx]
    cov, target, unc, new = _raw_inputs(rec)
    base = models.LoadestGP if rec["model"] == "loadest" else models.RatingGP

    class _FromReferenceState(base):
        def build_model(self, *a):
            mod = super().build_model(*a)
            checkpoint.load_state(mod, {k: torch.tensor(v, dtype=torch.float64) for k, v in rec["initial_state_dict"].items()},
                                  {k: torch.tensor(v, dtype=torch.float64) for k, v in rec["initial_likelihood_state_dict"].items()})
            return mod

    m = _FromReferenceState()
    m.fit(cov, target, target_unc=unc, iterations=rec["iterations"])
    ref = np.array(rec["history"])
    assert np.max(np.abs(np.array(m.history) - ref) / np.abs(ref)) <= 1e-6
    pred, se = m.predict(new)
    assert np.max(np.abs(np.asarray(pred) - np.array(rec["predict_target"])) / np.abs(rec["predict_target"])) <= 1e-6
    assert np.max(np.abs(np.asarray(se) - np.array(rec["predict_se"])) / np.abs(rec["predict_se"])) <= 1e-6
    # predict_grid (engines/gpytorch.py:500-549): 18 covariate columns x round(12 * time range) rows, data-space coordinates
    grid, index, covs = m.predict_grid(rec["grid_covariate"])
    ref_grid = np.array(rec["grid_values"])
    assert grid.shape == ref_grid.shape
    assert np.max(np.abs(grid - ref_grid) / np.abs(ref_grid)) <= 1e-6
    assert np.max(np.abs(np.asarray(covs, dtype=np.float64) - np.array(rec["grid_covariate_values"]))) <= 1e-9 * np.max(rec["grid_covariate_values"])
    ref_index = np.array(rec["grid_index_ns"], dtype="int64")
    got_index = np.asarray(index).astype("datetime64[ns]").astype("int64")
    assert np.max(np.abs(got_index - ref_index)) <= 2_000_000_000   # the reference rounds the grid times to the second


def test_reference_parameter_order_and_optimizer_state_layout():
    """The order in which the reference's optimiser sees the parameters (`model.parameters()` of its own module tree) and the
    conversion of an Adam state between that layout and this engine's scalar parameters, both directions."""
    for case in _cases():
        mod = _module_from_reference_state(case)
        order = checkpoint.reference_parameter_order(mod)
        assert [key for _, key, _, _ in order] == case["parameter_names"]
        for _, key, shape, idx in order:
            assert tuple(np.array(case["state_dict"][key]).shape) == shape and int(np.prod(shape, dtype=int)) == len(idx)
        params = mod.raw_list()
        opt = torch.optim.Adam(params, lr=0.05)
        for k, p in enumerate(params):
            p.grad = torch.tensor([0.1 * (k + 1)], dtype=torch.float64)
        opt.step()
        native = opt.state_dict()
        ref = checkpoint.optimizer_state_to_reference(mod, native)
        assert ref["param_groups"][0]["params"] == list(range(len(order)))
        for j, (_, key, shape, idx) in enumerate(order):
            assert tuple(ref["state"][j]["exp_avg"].shape) == shape
            assert [float(v) for v in ref["state"][j]["exp_avg"].reshape(-1)] == [float(native["state"][i]["exp_avg"]) for i in idx]
        back = checkpoint.optimizer_state_from_reference(mod, ref)
        for i in range(len(params)):
            for name in ("exp_avg", "exp_avg_sq"):
                assert float(back["state"][i][name]) == float(native["state"][i][name])
        opt2 = torch.optim.Adam(params, lr=0.05)
        opt2.load_state_dict(back)   # loads into an optimiser over the engine's parameters


def test_reference_written_checkpoint_loads_without_pickled_code():
    """tests/golden/ref_checkpoint_loadest.pt was written by the reference's own `save()` (engines/gpytorch.py:107-160, run on
    the stand-in): weights_only loading works, the state dicts map onto this engine's parameters, the optimiser state converts."""
    from discontinuum_b200 import engine

    path = os.path.join(os.path.dirname(GOLD), "ref_checkpoint_loadest.pt")
    with torch.serialization.safe_globals([engine._reference_model_config_shim()]):
        ckpt = torch.load(path, map_location="cpu", weights_only=True)
    assert ckpt["model_class"] == "loadest_gp.models.gpytorch.LoadestGPMarginalGPyTorch" and ckpt["optimizer_name"] == "adam"
    mod = spec.GPModule(models.loadest_spec(2))
    checkpoint.load_state(mod, ckpt["model_state_dict"], ckpt["likelihood_state_dict"])
    osd = checkpoint.optimizer_state_from_reference(mod, ckpt["optimizer_state_dict"])
    assert len(osd["state"]) == 10 and all(float(v["step"]) == 8.0 for v in osd["state"].values())
    torch.optim.Adam(mod.raw_list(), lr=0.05).load_state_dict(osd)


@pytest.mark.gpu
def test_resume_from_a_reference_written_checkpoint(cuda_device):
    """LoadestGP.load(<checkpoint written by the reference's save()>) + fit(resume=True): the same objective trajectory as the
    reference's own load() + fit(resume=True) -- parameters, Adam moments, step count and scheduler state all carried over."""
    with open(GOLD) as f:
        rec = json.load(f)["checkpoint"]
    time = np.array(rec["time_ns"], dtype="int64").astype("datetime64[ns]")
    cov, target = {"time": time, "flow": np.array(rec["flow"])}, np.array(rec["target"])
    m = models.LoadestGP.load(os.path.join(os.path.dirname(GOLD), rec["file"]), cov, target)
    m.fit(cov, target, iterations=rec["total_iterations"], resume=True)
    ref = np.array(rec["resumed_history"])
    got = np.array(m.history)
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref) / np.abs(ref)) <= 1e-6, np.max(np.abs(got - ref) / np.abs(ref))


def test_reference_written_rating_checkpoint_loads_without_pickled_code():
    """The same for rating-gp (tests/golden/ref_checkpoint_rating.pt, written by RatingGPMarginalGPyTorch.save()): the
    reference's optimiser holds the likelihood's noise first and the power-law / sigmoid parameters in its own module order
    (checkpoint._RATING_ORDER), so this is the layout where a wrong order would load silently into the wrong moments."""
    from discontinuum_b200 import engine

    path = os.path.join(os.path.dirname(GOLD), "ref_checkpoint_rating.pt")
    with torch.serialization.safe_globals([engine._reference_model_config_shim()]):
        ckpt = torch.load(path, map_location="cpu", weights_only=True)
    assert ckpt["model_class"].endswith("RatingGPMarginalGPyTorch")
    mod = spec.GPModule(models.rating_spec(1.1, 1.9))   # (bounds: any; only the parameter layout is checked here)
    checkpoint.load_state(mod, ckpt["model_state_dict"], ckpt["likelihood_state_dict"])
    order = checkpoint.reference_parameter_order(mod)
    ref_state = ckpt["optimizer_state_dict"]["state"]
    assert len(ref_state) == len(order)
    for j, (_, key, shape, idx) in enumerate(order):
        assert tuple(ref_state[j]["exp_avg"].shape) == shape, key
    osd = checkpoint.optimizer_state_from_reference(mod, ckpt["optimizer_state_dict"])
    assert len(osd["state"]) == len(mod.raw_list()) and all(float(v["step"]) == 8.0 for v in osd["state"].values())
    torch.optim.Adam(mod.raw_list(), lr=0.05).load_state_dict(osd)


@pytest.mark.gpu
def test_resume_from_a_reference_written_rating_checkpoint(cuda_device):
    """RatingGP.load(<checkpoint written by the reference's save()>) + fit(resume=True) follows the reference's own resumed
    trajectory (fixed per-observation noise, monotonicity penalty, likelihood-first parameter order)."""
    with open(GOLD) as f:
        rec = json.load(f)["checkpoint_rating"]
    time = np.array(rec["time_ns"], dtype="int64").astype("datetime64[ns]")
    cov, target, unc = {"time": time, "stage": np.array(rec["stage"])}, np.array(rec["target"]), np.array(rec["target_unc"])
    m = models.RatingGP.load(os.path.join(os.path.dirname(GOLD), rec["file"]), cov, target, unc)
    m.fit(cov, target, target_unc=unc, iterations=rec["total_iterations"], resume=True)
    ref = np.array(rec["resumed_history"])
    got = np.array(m.history)
    assert got.shape == ref.shape
    assert np.max(np.abs(got - ref) / np.abs(ref)) <= 1e-6, np.max(np.abs(got - ref) / np.abs(ref))


def _pymc_cases():
    with open(GOLD) as f:
        return json.load(f)["pymc"]


@pytest.mark.parametrize("idx", range(4))
def test_pymc_variant_matches_the_reference_pymc_model(idx):
    """The reference's PyMC model definition (loadest_gp/models/pymc.py:30-88: variables, priors, shapes, initial values,
    eta**2-scaled covariance terms on their dimensions, WhiteNoise(0.1)), run on the pymc stand-in, against the variable table
    of discontinuum_b200/pymc_variant.py and the oracle's restatement of the covariance and of the MAP objective."""
    from discontinuum_b200 import pymc_variant as pv

    c = _pymc_cases()[idx]
    vars_ = pv.loadest_pymc_vars(c["ndim"])
    assert [v.name for v in vars_] == [t["name"] for t in c["variables"]]
    for v, t in zip(vars_, c["variables"]):
        assert v.prior[0] == t["prior"] and v.size == t["size"], t["name"]
        want = {"halfnormal": [t["params"].get("sigma")], "normal": [t["params"].get("mu"), t["params"].get("sigma")],
                "gamma": [t["params"].get("alpha"), t["params"].get("beta")], "exponential": [t["params"].get("scale")]}[t["prior"]]
        assert [float(x) for x in v.prior[1:]] == [float(x) for x in want], t["name"]
        if t["initval"] is not None:     # explicit initval in the reference (eta_per = 1, ls_covariates = 0.5)
            assert np.allclose(v.init, np.broadcast_to(np.array(t["initval"]), v.init.shape)), t["name"]
    X, y = torch.tensor(np.array(c["X"])), torch.tensor(np.array(c["y"]))
    v = {k: torch.tensor(np.array(val, dtype=np.float64)) for k, val in c["values"].items()}
    K_ref = np.array(c["K"])
    assert np.max(np.abs(orc.pymc_loadest_cov(X, X, v).numpy() - K_ref)) <= 1e-12 * np.max(np.abs(K_ref))
    assert abs(float(orc.pymc_loadest_neg_logp(v, X, y)) - c["neg_logp"]) <= 1e-11 * abs(c["neg_logp"])
    # the engine's covariance spec at the reparameterised theta describes the same matrix (checked on the GPU elsewhere);
    # here: the log density of every variable of the table equals the stand-in's prior term
    Ky = torch.tensor(K_ref) + (0.1 ** 2 + 1e-6) * torch.eye(len(c["y"]), dtype=torch.float64)
    prior_sum = float(sum(x.logp(v[x.name]) for x in vars_))
    assert abs(-prior_sum + float(orc.nlml_from_K(Ky, y)[0]) - c["neg_logp"]) <= 1e-10 * abs(c["neg_logp"])


@pytest.mark.gpu
@pytest.mark.parametrize("idx", range(4))
def test_engine_covariance_matches_the_reference_pymc_model(cuda_device, idx):
    from discontinuum_b200 import capi, pymc_variant as pv

    c = _pymc_cases()[idx]
    X, y = np.array(c["X"]), np.array(c["y"])
    v = {k: torch.tensor(np.array(val, dtype=np.float64)) for k, val in c["values"].items()}
    eng = capi.Engine(max_n=X.shape[0], max_m=128)
    eng.set_train(pv.loadest_pymc_spec(c["ndim"]).to_c(), X, y, np.full(X.shape[0], 0.1 ** 2 + 1e-6))
    K = eng.covmat(pv.pymc_to_natural(v).numpy())
    assert np.max(np.abs(K - np.array(c["K"]))) <= 1e-12 * np.max(np.abs(c["K"]))
    eng.close()
