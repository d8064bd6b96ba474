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
