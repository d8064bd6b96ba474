"""Checkpoint key mapping between the B200 engine's parameter table and the gpytorch module tree the reference
checkpoints (`MarginalGPyTorch.save`, src/discontinuum/engines/gpytorch.py:147-160: `model.state_dict()` and
`likelihood.state_dict()` of the modules built at src/loadest_gp/models/gpytorch.py:48-128 and
src/rating_gp/models/gpytorch.py:64-79,205-372).

`MarginalB200.save` writes `model_state_dict` / `likelihood_state_dict` under the REFERENCE's key names and tensor
shapes, so a reference-side `load_state_dict(..., strict=False)` picks the raw parameters up (the constraint / prior
buffers gpytorch also stores are constants of the model definition and are not written); `MarginalB200.load` reads either
format -- reference-named (from either engine; unknown keys such as constraint bounds are ignored) or the round-1 native
`raw.<name>` layout.  Raw values mean the same thing on both sides: GPyTorch's Positive / GreaterThan / Interval
constraints are the softplus / softplus + lb / scaled-sigmoid transforms of spec.transform.

Module paths follow gpytorch's flattening `+` / `*` (AdditiveKernel / ProductKernel with a flat `kernels` list); the
nested form of older releases ((A + B) + C -> kernels.0.kernels.0, kernels.0.kernels.1, kernels.1) is accepted on load.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from .spec import GPModule

# Param name (vector parameters without their trailing ".<d>") -> (gpytorch key, trailing shape; -1 = number of dims)
_LOADEST = {
    "mean.constant": ("mean_module.raw_constant", ()),
    "seasonal.outputscale": ("covar_module.kernels.0.raw_outputscale", ()),
    "seasonal.periodic.lengthscale": ("covar_module.kernels.0.base_kernel.kernels.0.raw_lengthscale", (1, 1)),
    "seasonal.periodic.period_length": ("covar_module.kernels.0.base_kernel.kernels.0.raw_period_length", (1, 1)),
    "seasonal.matern52.lengthscale": ("covar_module.kernels.0.base_kernel.kernels.1.raw_lengthscale", (1, 1)),
    "covariates.outputscale": ("covar_module.kernels.1.raw_outputscale", ()),
    "covariates.rbf.lengthscale": ("covar_module.kernels.1.base_kernel.raw_lengthscale", (1, -1)),
    "residual.outputscale": ("covar_module.kernels.2.raw_outputscale", ()),
    "residual.matern32.lengthscale": ("covar_module.kernels.2.base_kernel.raw_lengthscale", (1, -1)),
}
_LOWER = "covar_module.kernels.0.kernels.1.base_kernel"   # LogWarp(shiftA + shiftB) under sigmoid_lower * ...
_UPPER = "covar_module.kernels.1.kernels.1.base_kernel"   # LogWarp(bend) under sigmoid_upper * ...
_PLAIN = "covar_module.kernels.2.base_kernel"             # LogWarp(base + periodic)
_RATING = {
    "powerlaw.a": ("powerlaw.a", (1,)),
    "powerlaw.b": ("powerlaw.b", (1,)),
    "powerlaw.c": ("powerlaw.c", (1,)),
    "likelihood.second_noise": ("likelihood.second_noise_covar.raw_noise", (1,)),
    "sigmoid.b": ("covar_module.kernels.0.kernels.0.raw_b", (1, 1)),
    "shiftA.outputscale": (_LOWER + ".kernels.0.raw_outputscale", ()),
    "shiftA.stage_matern52.lengthscale": (_LOWER + ".kernels.0.base_kernel.kernels.0.raw_lengthscale", (1, 1)),
    "shiftA.time_matern32.lengthscale": (_LOWER + ".kernels.0.base_kernel.kernels.1.raw_lengthscale", (1, 1)),
    "shiftB.outputscale": (_LOWER + ".kernels.1.raw_outputscale", ()),
    "shiftB.stage_matern52.lengthscale": (_LOWER + ".kernels.1.base_kernel.kernels.0.raw_lengthscale", (1, 1)),
    "shiftB.time_matern32.lengthscale": (_LOWER + ".kernels.1.base_kernel.kernels.1.raw_lengthscale", (1, 1)),
    "bend.outputscale": (_UPPER + ".raw_outputscale", ()),
    "bend.stage_matern52.lengthscale": (_UPPER + ".base_kernel.kernels.0.raw_lengthscale", (1, 1)),
    "bend.time_matern52.lengthscale": (_UPPER + ".base_kernel.kernels.1.raw_lengthscale", (1, 1)),
    "base.outputscale": (_PLAIN + ".kernels.0.raw_outputscale", ()),
    "base.stage_matern52.lengthscale": (_PLAIN + ".kernels.0.base_kernel.raw_lengthscale", (1, 1)),
    "periodic.outputscale": (_PLAIN + ".kernels.1.raw_outputscale", ()),
    "periodic.period_length": (_PLAIN + ".kernels.1.base_kernel.kernels.0.raw_period_length", (1, 1)),
    "periodic.lengthscale": (_PLAIN + ".kernels.1.base_kernel.kernels.0.raw_lengthscale", (1, 1)),
    "periodic.time_matern52.lengthscale": (_PLAIN + ".kernels.1.base_kernel.kernels.1.raw_lengthscale", (1, 1)),
}
# the inverted gate holds the same switch-point parameter twice more (its own registration and the shared kernel)
_RATING_ALIASES = {"covar_module.kernels.0.kernels.0.raw_b": ["covar_module.kernels.1.kernels.0.raw_b",
                                                             "covar_module.kernels.1.kernels.0.sigmoid_kernel.raw_b"]}


def _table(module: GPModule) -> Tuple[Dict[str, Tuple[str, tuple]], Dict[str, List[str]]]:
    names = {p.name for p in module.spec.params}
    if "powerlaw.a" in names:
        return _RATING, _RATING_ALIASES
    if "seasonal.outputscale" in names:
        return _LOADEST, {}
    raise ValueError("no gpytorch key mapping for this model's parameter table")


def _groups(module: GPModule) -> Dict[str, List[int]]:
    """base name -> indices of its Params in theta order ('x.lengthscale.0', 'x.lengthscale.1' -> 'x.lengthscale')."""
    out: Dict[str, List[int]] = {}
    for i, p in enumerate(module.spec.params):
        head, _, tail = p.name.rpartition(".")
        base = head if tail.isdigit() else p.name
        out.setdefault(base, []).append(i)
    return out


def _nested_alternatives(key: str) -> List[str]:
    """covar_module.kernels.K.<rest> of a flat three-term sum under the nested (A + B) + C layout."""
    pre = "covar_module.kernels."
    if not key.startswith(pre):
        return []
    k, _, rest = key[len(pre):].partition(".")
    nested = {"0": "0.kernels.0", "1": "0.kernels.1", "2": "1"}.get(k)
    return [pre + nested + "." + rest] if nested else []


def to_reference_state(module: GPModule) -> Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor]]:
    """(model_state_dict, likelihood_state_dict) with the reference's key names and shapes (raw parameters only)."""
    table, aliases = _table(module)
    raws = module.raw_list()
    model_sd: Dict[str, torch.Tensor] = {}
    for base, idx in _groups(module).items():
        key, shape = table[base]
        val = torch.cat([raws[i].detach().reshape(1) for i in idx]).to(torch.float64)
        shp = tuple(len(idx) if s == -1 else s for s in shape)
        model_sd[key] = val.reshape(shp).clone()
        for other in aliases.get(key, []):
            model_sd[other] = model_sd[key].clone()
    lik = {k[len("likelihood."):]: v for k, v in model_sd.items() if k.startswith("likelihood.")}
    return model_sd, lik


def is_native_state(sd: Dict[str, torch.Tensor]) -> bool:
    return any(k.startswith("raw.") for k in sd)


def load_state(module: GPModule, model_sd: Dict[str, torch.Tensor], likelihood_sd: Optional[Dict[str, torch.Tensor]] = None):
    """Fill the module's raw parameters from a checkpoint's state dicts: reference-named (written by either engine) or the
    native `raw.<name>` layout.  Raises KeyError naming the first parameter the checkpoint does not hold."""
    if is_native_state(model_sd):
        module.load_state_dict(model_sd)
        return
    table, _ = _table(module)
    merged = dict(model_sd)
    for k, v in (likelihood_sd or {}).items():
        merged.setdefault("likelihood." + k, v)
    groups = _groups(module)

    def resolve(nested: bool):
        out = {}
        for base, idx in groups.items():
            key, _ = table[base]
            cand = (_nested_alternatives(key) or [key]) if nested else [key]
            found = next((k for k in cand if k in merged), None)
            if found is None:
                raise KeyError(f"checkpoint holds no '{cand[0]}' (parameter '{base}'): not a state dict of this model "
                               f"(keys: {sorted(merged)[:6]} ...)")
            val = torch.as_tensor(merged[found]).detach().to(torch.float64).reshape(-1)
            if val.numel() != len(idx):
                raise ValueError(f"'{found}' has {val.numel()} elements, the model expects {len(idx)}")
            out[base] = val
        return out

    try:
        vals = resolve(False)          # flat AdditiveKernel (current gpytorch)
    except (KeyError, ValueError) as flat_error:
        try:
            vals = resolve(True)       # (A + B) + C nesting of older releases
        except (KeyError, ValueError):
            raise flat_error from None
    raws = module.raw_list()
    with torch.no_grad():
        for base, idx in groups.items():
            for j, i in enumerate(idx):
                raws[i].fill_(float(vals[base][j]))


# ---------------------------------------------------------------- optimiser state
# torch optimisers index their state by position in `model.parameters()`.  The reference's order is the registration order
# of its module tree -- gpytorch.models.ExactGP registers the likelihood first, then the model's own sub-modules in the
# order its __init__ creates them -- and a vector parameter (an ARD length scale) is ONE tensor there, while this engine keeps
# one scalar parameter per element, in theta order.  tests/test_reference_golden.py holds the two lists below to the
# reference's own `named_parameters()` (golden vectors written by its unmodified model code).
_LOADEST_ORDER = ["mean.constant", "seasonal.outputscale", "seasonal.periodic.lengthscale", "seasonal.periodic.period_length",
                  "seasonal.matern52.lengthscale", "covariates.outputscale", "covariates.rbf.lengthscale", "residual.outputscale",
                  "residual.matern32.lengthscale"]
_RATING_ORDER = ["likelihood.second_noise", "powerlaw.a", "powerlaw.b", "powerlaw.c", "sigmoid.b",
                 "shiftA.outputscale", "shiftA.stage_matern52.lengthscale", "shiftA.time_matern32.lengthscale",
                 "shiftB.outputscale", "shiftB.stage_matern52.lengthscale", "shiftB.time_matern32.lengthscale",
                 "bend.outputscale", "bend.stage_matern52.lengthscale", "bend.time_matern52.lengthscale",
                 "base.outputscale", "base.stage_matern52.lengthscale",
                 "periodic.outputscale", "periodic.lengthscale", "periodic.period_length", "periodic.time_matern52.lengthscale"]


def reference_parameter_order(module: GPModule) -> List[Tuple[str, str, tuple, List[int]]]:
    """[(base name, reference key, reference shape, indices of this engine's scalar parameters)] in the order of the
    reference's `model.parameters()`."""
    table, _ = _table(module)
    groups = _groups(module)
    order = _RATING_ORDER if table is _RATING else _LOADEST_ORDER
    out = []
    for base in order:
        key, shape = table[base]
        idx = groups[base]
        out.append((base, key, tuple(len(idx) if s == -1 else s for s in shape), idx))
    return out


def optimizer_state_to_reference(module: GPModule, sd: dict) -> dict:
    """State dict of a torch Adam / AdamW over this engine's scalar parameters (theta order) -> the layout the reference's
    optimiser over `model.parameters()` loads: its parameter order, one entry per reference tensor, reference shapes."""
    state = {}
    order = reference_parameter_order(module)
    for j, (_base, _key, shape, idx) in enumerate(order):
        parts = [sd["state"].get(i) for i in idx]
        if any(p is None for p in parts):
            continue
        entry = {"step": parts[0]["step"].clone() if torch.is_tensor(parts[0]["step"]) else parts[0]["step"]}
        for name in ("exp_avg", "exp_avg_sq", "max_exp_avg_sq"):
            if name in parts[0]:
                entry[name] = torch.cat([p[name].reshape(1) for p in parts]).reshape(shape).clone()
        state[j] = entry
    groups = [dict(g, params=list(range(len(order)))) for g in sd["param_groups"]]
    return {"state": state, "param_groups": groups}


def optimizer_state_from_reference(module: GPModule, sd: dict) -> dict:
    """The inverse: a reference-layout optimiser state dict (written by MarginalGPyTorch.save or by MarginalB200.save) ->
    one entry per scalar parameter of this engine, theta order."""
    order = reference_parameter_order(module)
    if len(sd["param_groups"]) != 1 or len(sd["param_groups"][0]["params"]) != len(order):
        raise ValueError(f"optimiser state holds {[len(g['params']) for g in sd['param_groups']]} parameters, the reference's "
                         f"module tree of this model has {len(order)}")
    ref_ids = sd["param_groups"][0]["params"]
    nparam = sum(len(idx) for _, _, _, idx in order)
    state = {}
    for j, (_base, key, _shape, idx) in enumerate(order):
        entry = sd["state"].get(ref_ids[j])
        if entry is None:
            continue
        for k, i in enumerate(idx):
            e = {"step": entry["step"].clone().to(torch.float32) if torch.is_tensor(entry["step"]) else torch.tensor(float(entry["step"]))}
            for name in ("exp_avg", "exp_avg_sq", "max_exp_avg_sq"):
                if name in entry:
                    flat = entry[name].reshape(-1)
                    if flat.numel() != len(idx):
                        raise ValueError(f"optimiser state of '{key}' has {flat.numel()} elements, the model expects {len(idx)}")
                    e[name] = flat[k].reshape(1).to(torch.float64).clone()
            state[i] = e
    groups = [dict(g, params=list(range(nparam))) for g in sd["param_groups"]]
    return {"state": state, "param_groups": groups}
