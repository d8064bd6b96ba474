"""Stand-in for the slice of PyMC that the reference's PyMC model touches (loadest_gp/models/pymc.py:30-88): a model context
that records named random variables with their prior, arithmetic on them (eta**2 * cov), the covariance functions and
gp.Marginal with additive composition and marginal_likelihood.  TEST INFRASTRUCTURE (see ../README.md); formulas as documented
by PyMC, evaluated with torch float64 at values supplied per variable.  `Model.neg_logp(values)` is what
pm.find_MAP minimises: -(log N(y | 0, K + noise) + sum of the priors' log densities), without Jacobian terms."""
import math

import torch

from . import gp  # noqa: F401

_STACK = []


class Model:
    def __init__(self):
        self.vars = {}
        self.observed = None

    def __enter__(self):
        _STACK.append(self)
        return self

    def __exit__(self, *a):
        _STACK.pop()

    def set_values(self, values):
        for name, var in self.vars.items():
            var.value = torch.as_tensor(values[name], dtype=torch.float64).reshape(-1)

    def neg_logp(self, values):
        self.set_values(values)
        lp = sum(var.logp() for var in self.vars.values())
        return -(lp + self.observed())


class _Expr:
    """A deferred scalar / vector expression of random variables."""
    def __init__(self, fn):
        self.fn = fn

    def eval(self):
        return self.fn()

    def __pow__(self, p):
        return _Expr(lambda: self.eval() ** p)

    def __mul__(self, other):
        if isinstance(other, gp.cov.Covariance):
            return gp.cov.Scaled(self, other)
        return _Expr(lambda: self.eval() * _val(other))

    __rmul__ = __mul__


def _val(x):
    return x.eval() if isinstance(x, _Expr) else torch.as_tensor(x, dtype=torch.float64)


class _RV(_Expr):
    def __init__(self, name, kind, params, shape=None, initval=None):
        super().__init__(lambda: self.value)
        self.name, self.kind, self.params, self.shape, self.initval = name, kind, params, shape, initval
        self.value = None
        _STACK[-1].vars[name] = self

    def logp(self):
        x = self.value
        k, p = self.kind, self.params
        if k == "halfnormal":
            s = p["sigma"]
            return (math.log(2.0) - 0.5 * math.log(2.0 * math.pi) - math.log(s) - x ** 2 / (2.0 * s ** 2)).sum()
        if k == "normal":
            return (-0.5 * math.log(2.0 * math.pi) - math.log(p["sigma"]) - (x - p["mu"]) ** 2 / (2.0 * p["sigma"] ** 2)).sum()
        if k == "gamma":
            a, b = p["alpha"], p["beta"]
            return (a * math.log(b) - math.lgamma(a) + (a - 1.0) * torch.log(x) - b * x).sum()
        if k == "exponential":
            return (-math.log(p["scale"]) - x / p["scale"]).sum()
        raise ValueError(k)


def HalfNormal(name, sigma=1.0, initval=None, shape=None, **kw):
    return _RV(name, "halfnormal", {"sigma": float(sigma)}, shape, initval)


def Normal(name, mu=0.0, sigma=1.0, initval=None, shape=None, **kw):
    return _RV(name, "normal", {"mu": float(mu), "sigma": float(sigma)}, shape, initval)


def Gamma(name, alpha=None, beta=None, initval=None, shape=None, **kw):
    return _RV(name, "gamma", {"alpha": float(alpha), "beta": float(beta)}, shape, initval)


def Exponential(name, lam=None, scale=None, initval=None, shape=None, **kw):
    return _RV(name, "exponential", {"scale": float(scale) if scale is not None else 1.0 / float(lam)}, shape, initval)
