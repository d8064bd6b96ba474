"""pm.gp.cov: stationary kernels over active dimensions, products, sums, scaling by a random-variable expression.
  ExpQuad   exp(-|d / ls|^2 / 2)
  Matern52  (1 + sqrt(5) r + 5 r^2 / 3) exp(-sqrt(5) r),  Matern32  (1 + sqrt(3) r) exp(-sqrt(3) r),   r = |d / ls|
  Periodic  exp(-sum sin^2(pi d / period) / (2 ls^2))
  WhiteNoise  sigma^2 I"""
import math

import torch


class Covariance:
    def __call__(self, X, Xs=None):
        raise NotImplementedError

    def __add__(self, other):
        return _Combo([self, other], "add")

    def __mul__(self, other):
        if isinstance(other, Covariance):
            return _Combo([self, other], "mul")
        return Scaled(other, self)

    __rmul__ = __mul__


class _Combo(Covariance):
    def __init__(self, parts, op):
        self.parts, self.op = parts, op

    def __call__(self, X, Xs=None):
        out = None
        for p in self.parts:
            v = p(X, Xs)
            out = v if out is None else (out + v if self.op == "add" else out * v)
        return out


class Scaled(Covariance):
    def __init__(self, factor, cov):
        self.factor, self.cov = factor, cov

    def __call__(self, X, Xs=None):
        f = self.factor.eval() if hasattr(self.factor, "eval") else torch.as_tensor(self.factor, dtype=torch.float64)
        return f * self.cov(X, Xs)


def _ev(x):
    return x.eval() if hasattr(x, "eval") else torch.as_tensor(x, dtype=torch.float64)


class _Stationary(Covariance):
    def __init__(self, input_dim, ls=None, ls_inv=None, active_dims=None):
        self.input_dim, self.ls = input_dim, ls
        self.active_dims = list(range(input_dim)) if active_dims is None else [int(a) for a in active_dims]

    def _slices(self, X, Xs):
        X = X[:, self.active_dims]
        Xs = X if Xs is None else Xs[:, self.active_dims]
        return X, Xs

    def _r2(self, X, Xs):
        X, Xs = self._slices(X, Xs)
        ls = _ev(self.ls)
        d = (X / ls).unsqueeze(1) - (Xs / ls).unsqueeze(0)
        return (d * d).sum(-1)


class ExpQuad(_Stationary):
    def __call__(self, X, Xs=None):
        return torch.exp(-0.5 * self._r2(X, Xs))


class Matern52(_Stationary):
    def __call__(self, X, Xs=None):
        r = torch.sqrt(self._r2(X, Xs).clamp_min(1e-30))
        return (1.0 + math.sqrt(5.0) * r + 5.0 / 3.0 * r * r) * torch.exp(-math.sqrt(5.0) * r)


class Matern32(_Stationary):
    def __call__(self, X, Xs=None):
        r = torch.sqrt(self._r2(X, Xs).clamp_min(1e-30))
        return (1.0 + math.sqrt(3.0) * r) * torch.exp(-math.sqrt(3.0) * r)


class Periodic(_Stationary):
    def __init__(self, input_dim, period, ls=None, ls_inv=None, active_dims=None):
        super().__init__(input_dim, ls=ls, active_dims=active_dims)
        self.period = period

    def __call__(self, X, Xs=None):
        X, Xs = self._slices(X, Xs)
        d = math.pi * (X.unsqueeze(1) - Xs.unsqueeze(0)) / _ev(self.period)
        ls = _ev(self.ls)
        return torch.exp(-0.5 * ((torch.sin(d) / ls) ** 2).sum(-1))


class WhiteNoise(Covariance):
    def __init__(self, sigma):
        self.sigma = sigma

    def __call__(self, X, Xs=None):
        if Xs is not None:
            return torch.zeros(X.shape[0], Xs.shape[0], dtype=torch.float64)
        return _ev(self.sigma) ** 2 * torch.eye(X.shape[0], dtype=torch.float64)
