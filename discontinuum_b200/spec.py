"""Host-side covariance specification: the tagged tree the CUDA tile generator consumes.

A model is a sum of terms; a term is  outputscale x optional sigmoid gate x product of stationary
factors (RBF / Matern-3/2 / Matern-5/2 / periodic) on selected columns of a per-point feature
table (raw, log-warped or gate columns).  This is exactly the family the reference's models span
(src/loadest_gp/models/gpytorch.py:61-128, src/rating_gp/models/gpytorch.py:205-372,
src/rating_gp/models/kernels.py:242-382).  Hyper-parameters carry GPyTorch's constraint /
prior / initial-value semantics; the O(P) host math (constraint transforms,
priors, their chain rule) is done with torch on the CPU, the O(n^3) math never is.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import capi

LOG2PI = math.log(2.0 * math.pi)


@dataclass
class Param:
    name: str
    constraint: Tuple = ("positive",)      # ("positive",) | ("greater_than", lb) | ("interval", lo, hi) | ("none",)
    prior: Optional[Tuple] = None          # ("halfnormal", s) | ("normal", mu, s) | ("gamma", conc, rate)
    init_raw: float = 0.0
    group: str = "model"                   # "model" or "likelihood" (state-dict split, engines/gpytorch.py:149-150)


def transform(p: Param, raw: torch.Tensor) -> torch.Tensor:
    c = p.constraint
    if c[0] == "positive":
        return torch.nn.functional.softplus(raw)
    if c[0] == "greater_than":
        return torch.nn.functional.softplus(raw) + c[1]
    if c[0] == "interval":
        return c[1] + (c[2] - c[1]) * torch.sigmoid(raw)
    return raw


def inverse_transform(p: Param, value: float) -> float:
    c = p.constraint
    v = float(value)
    if c[0] == "positive":
        return v + math.log(-math.expm1(-v))
    if c[0] == "greater_than":
        v -= c[1]
        return v + math.log(-math.expm1(-v))
    if c[0] == "interval":
        u = (v - c[1]) / (c[2] - c[1])
        return math.log(u) - math.log1p(-u)
    return v


def log_prior(p: Param, x: torch.Tensor) -> torch.Tensor:
    pr = p.prior
    if pr is None:
        return x.new_zeros(())
    if pr[0] == "halfnormal":
        s = pr[1]
        return math.log(2.0) - 0.5 * LOG2PI - math.log(s) - x * x / (2.0 * s * s)
    if pr[0] == "normal":
        mu, s = pr[1], pr[2]
        return -0.5 * LOG2PI - math.log(s) - (x - mu) ** 2 / (2.0 * s * s)
    if pr[0] == "gamma":
        a, b = pr[1], pr[2]
        return a * math.log(b) + (a - 1.0) * torch.log(x) - b * x - math.lgamma(a)
    raise ValueError(pr)


@dataclass
class Factor:
    kind: int
    cols: Sequence[int]
    ls: Sequence[int]            # theta indices, one per col
    period: int = -1


@dataclass
class Term:
    scale: int
    factors: List[Factor]
    gate: int = capi.GATE_NONE
    gate_col: int = 0


@dataclass
class CovSpec:
    """Builder of a dgp_spec plus the parameter table that goes with it."""

    ndim: int
    params: List[Param] = field(default_factory=list)
    cols: List[Tuple[int, int, int, float]] = field(default_factory=list)  # (kind, src, theta, aux)
    terms: List[Term] = field(default_factory=list)
    mean_kind: int = capi.MEAN_ZERO
    mean_col: int = 0
    mean_theta: Tuple[int, ...] = ()
    noise_theta: int = -1

    def param(self, name, constraint=("positive",), prior=None, init_raw=0.0, group="model") -> int:
        self.params.append(Param(name, tuple(constraint), prior, float(init_raw), group))
        return len(self.params) - 1

    def col_copy(self, src: int) -> int:
        self.cols.append((capi.COL_COPY, src, -1, 0.0))
        return len(self.cols) - 1

    def col_log(self, src: int, eps: float) -> int:
        self.cols.append((capi.COL_LOG, src, -1, float(eps)))
        return len(self.cols) - 1

    def col_gate(self, src: int, sharpness: float, theta: int) -> int:
        self.cols.append((capi.COL_GATE, src, theta, float(sharpness)))
        return len(self.cols) - 1

    def term(self, scale: int, factors: List[Factor], gate: int = capi.GATE_NONE, gate_col: int = 0):
        self.terms.append(Term(scale, factors, gate, gate_col))

    @property
    def ntheta(self) -> int:
        return len(self.params)

    def index(self, name: str) -> int:
        for i, p in enumerate(self.params):
            if p.name == name:
                return i
        raise KeyError(name)

    def to_c(self) -> capi.DgpSpec:
        if len(self.terms) > capi.MAX_TERMS or len(self.cols) > capi.MAX_COLS or self.ntheta > capi.MAX_THETA:
            raise ValueError("covariance spec exceeds libdgp limits")
        s = capi.DgpSpec()
        s.abi = capi.ABI_VERSION
        s.ndim, s.ncols, s.nterms, s.ntheta = self.ndim, len(self.cols), len(self.terms), self.ntheta
        s.noise_theta = self.noise_theta
        s.mean_kind, s.mean_col = self.mean_kind, self.mean_col
        for k in range(4):
            s.mean_theta[k] = self.mean_theta[k] if k < len(self.mean_theta) else -1
        for i, (kind, src, th, aux) in enumerate(self.cols):
            s.col[i].kind, s.col[i].src, s.col[i].theta, s.col[i].aux = kind, src, th, aux
        for i, t in enumerate(self.terms):
            ct = s.term[i]
            ct.scale, ct.gate, ct.gate_col, ct.nfactors = t.scale, t.gate, t.gate_col, len(t.factors)
            if len(t.factors) > capi.MAX_FACTORS:
                raise ValueError("too many factors in a term")
            for j, f in enumerate(t.factors):
                cf = ct.factor[j]
                cf.kind, cf.ndims, cf.period = f.kind, len(f.cols), f.period
                if len(f.cols) > capi.MAX_FDIMS or len(f.cols) != len(f.ls):
                    raise ValueError("bad factor dims")
                for d in range(capi.MAX_FDIMS):
                    cf.col[d] = f.cols[d] if d < len(f.cols) else 0
                    cf.ls[d] = f.ls[d] if d < len(f.cols) else -1
        return s


class GPModule(torch.nn.Module):
    """Raw hyper-parameters of one model as float64 torch Parameters on the CPU (state_dict-able like
    the gpytorch module the reference checkpoints at discontinuum/engines/gpytorch.py:147-160)."""

    def __init__(self, spec: CovSpec):
        super().__init__()
        self.spec = spec
        self.raw = torch.nn.ParameterDict(
            {p.name.replace(".", "__"): torch.nn.Parameter(torch.tensor([p.init_raw], dtype=torch.float64)) for p in spec.params})
        self.training_mode = True

    def raw_list(self) -> List[torch.nn.Parameter]:
        return [self.raw[p.name.replace(".", "__")] for p in self.spec.params]

    def natural(self) -> torch.Tensor:
        """Differentiable vector of natural parameter values, in theta order."""
        return torch.cat([transform(p, r) for p, r in zip(self.spec.params, self.raw_list())])

    def log_prior(self, nat: torch.Tensor) -> torch.Tensor:
        total = nat.new_zeros(())
        for i, p in enumerate(self.spec.params):
            if p.prior is not None:
                total = total + log_prior(p, nat[i])
        return total

    # ---- closed-form host path: the same O(P) maths without building an autograd graph
    def _tables(self):
        """Per-parameter constant tables for host_chain (built once)."""
        t = getattr(self, "_host_tables", None)
        if t is None:
            P = len(self.spec.params)
            kind = np.zeros(P, dtype=np.int64)       # 0 none, 1 positive / greater_than (softplus + lb), 2 interval
            lb, lo, hi = np.zeros(P), np.zeros(P), np.ones(P)
            pk = np.zeros(P, dtype=np.int64)         # 0 none, 1 halfnormal, 2 normal, 3 gamma
            pa, pb, pc = np.zeros(P), np.ones(P), np.zeros(P)  # prior constants
            for i, p in enumerate(self.spec.params):
                c = p.constraint
                if c[0] == "positive":
                    kind[i] = 1
                elif c[0] == "greater_than":
                    kind[i], lb[i] = 1, c[1]
                elif c[0] == "interval":
                    kind[i], lo[i], hi[i] = 2, c[1], c[2]
                pr = p.prior
                if pr is None:
                    continue
                if pr[0] == "halfnormal":
                    pk[i], pb[i], pc[i] = 1, pr[1], math.log(2.0) - 0.5 * LOG2PI - math.log(pr[1])
                elif pr[0] == "normal":
                    pk[i], pa[i], pb[i], pc[i] = 2, pr[1], pr[2], -0.5 * LOG2PI - math.log(pr[2])
                elif pr[0] == "gamma":
                    pk[i], pa[i], pb[i], pc[i] = 3, pr[1], pr[2], pr[1] * math.log(pr[2]) - math.lgamma(pr[1])
                else:
                    raise ValueError(pr)
            t = self._host_tables = (kind, lb, lo, hi, pk, pa, pb, pc)
        return t

    def host_chain(self):
        """(natural[P], d natural / d raw [P], sum of log priors, d(sum log prior) / d natural [P]) in numpy float64:
        the constraint transforms and priors of `natural()` / `log_prior()` with their
        derivatives in closed form, for the optimiser loop's fast path (no autograd graph per iteration)."""
        return self.host_chain_raw(np.array([float(r.detach()) for r in self.raw_list()], dtype=np.float64))

    def host_chain_raw(self, raw: np.ndarray):
        """host_chain at the raw parameter vector `raw` (numpy, [P]) instead of the module's tensors: the optimiser loop's
        numpy path keeps the raw parameters in one array and touches no tensor per iteration."""
        kind, lb, lo, hi, pk, pa, pb, pc = self._tables()
        sig = 1.0 / (1.0 + np.exp(-raw))
        sp = np.where(raw > 20.0, raw, np.log1p(np.exp(np.minimum(raw, 20.0))))  # torch softplus, threshold 20
        nat = np.where(kind == 1, sp + lb, np.where(kind == 2, lo + (hi - lo) * sig, raw))
        dnat = np.where(kind == 1, np.where(raw > 20.0, 1.0, sig), np.where(kind == 2, (hi - lo) * sig * (1.0 - sig), 1.0))
        with np.errstate(divide="ignore", invalid="ignore"):
            lp = np.where(pk == 1, pc - nat * nat / (2.0 * pb * pb),
                          np.where(pk == 2, pc - (nat - pa) ** 2 / (2.0 * pb * pb),
                                   np.where(pk == 3, pc + (pa - 1.0) * np.log(nat) - pb * nat, 0.0)))
            dlp = np.where(pk == 1, -nat / (pb * pb),
                           np.where(pk == 2, -(nat - pa) / (pb * pb),
                                    np.where(pk == 3, (pa - 1.0) / nat - pb, 0.0)))
        return nat, dnat, float(lp.sum()), dlp

    def set_natural(self, name: str, value: float):
        i = self.spec.index(name)
        with torch.no_grad():
            self.raw_list()[i].fill_(inverse_transform(self.spec.params[i], value))

    def natural_dict(self) -> Dict[str, float]:
        with torch.no_grad():
            nat = self.natural()
        return {p.name: float(nat[i]) for i, p in enumerate(self.spec.params)}
