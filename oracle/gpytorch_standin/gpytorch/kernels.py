"""Kernel base (active_dims, ARD length scales, `+` / `*` flattening) and the stationary kernels the reference composes.
Formulas as documented by gpytorch: inputs are divided by the length scale, then
  RBF      exp(-d^2 / 2)
  Matern   nu = 1/2: exp(-d) | 3/2: (1 + sqrt(3) d) exp(-sqrt(3) d) | 5/2: (1 + sqrt(5) d + 5 d^2 / 3) exp(-sqrt(5) d)
  Periodic exp(-2 sum_k sin^2(pi (x_k - x'_k) / p) / l)
  Scale    outputscale * base."""
import math

import torch

from .constraints import Positive
from .module import Module


def _dense(k):
    return k.to_dense() if hasattr(k, "to_dense") else k


class Kernel(Module):
    has_lengthscale = False

    def __init__(self, ard_num_dims=None, batch_shape=torch.Size([]), active_dims=None, lengthscale_prior=None,
                 lengthscale_constraint=None, eps=1e-6, **kwargs):
        super().__init__()
        self._batch_shape = batch_shape
        if active_dims is not None and not torch.is_tensor(active_dims):
            active_dims = torch.tensor([int(a) for a in active_dims], dtype=torch.long)
        self.register_buffer("active_dims", active_dims)
        self.ard_num_dims = ard_num_dims
        self.eps = eps
        if self.has_lengthscale:
            nls = 1 if ard_num_dims is None else int(ard_num_dims)
            self.register_parameter("raw_lengthscale", torch.nn.Parameter(torch.zeros(*batch_shape, 1, nls, dtype=torch.float64)))
            self.register_constraint("raw_lengthscale", lengthscale_constraint if lengthscale_constraint is not None else Positive())
            if lengthscale_prior is not None:
                self.register_prior("lengthscale_prior", lengthscale_prior, lambda m: m.lengthscale, lambda m, v: m._set_lengthscale(v))

    @property
    def batch_shape(self):
        return self._batch_shape

    @property
    def lengthscale(self):
        return self.raw_lengthscale_constraint.transform(self.raw_lengthscale) if self.has_lengthscale else None

    def _set_lengthscale(self, value):
        self.initialize(raw_lengthscale=self.raw_lengthscale_constraint.inverse_transform(torch.as_tensor(value, dtype=torch.float64)))

    def forward(self, x1, x2, diag=False, **params):
        raise NotImplementedError

    def __call__(self, x1, x2=None, diag=False, **params):
        if x1.ndim == 1:
            x1 = x1.unsqueeze(-1)
        if x2 is not None and x2.ndim == 1:
            x2 = x2.unsqueeze(-1)
        if self.active_dims is not None:
            x1 = x1.index_select(-1, self.active_dims)
            if x2 is not None:
                x2 = x2.index_select(-1, self.active_dims)
        if x2 is None:
            x2 = x1
        out = _dense(torch.nn.Module.__call__(self, x1, x2, **params))
        return torch.diagonal(out, dim1=-1, dim2=-2) if diag else out

    def __add__(self, other):
        ks = (list(self.kernels) if isinstance(self, AdditiveKernel) else [self]) + \
             (list(other.kernels) if isinstance(other, AdditiveKernel) else [other])
        return AdditiveKernel(*ks)

    def __mul__(self, other):
        ks = (list(self.kernels) if isinstance(self, ProductKernel) else [self]) + \
             (list(other.kernels) if isinstance(other, ProductKernel) else [other])
        return ProductKernel(*ks)


class AdditiveKernel(Kernel):
    def __init__(self, *kernels):
        super().__init__()
        self.kernels = torch.nn.ModuleList(kernels)

    def forward(self, x1, x2, **params):
        out = None
        for k in self.kernels:
            v = k(x1, x2, **params)
            out = v if out is None else out + v
        return out


class ProductKernel(Kernel):
    def __init__(self, *kernels):
        super().__init__()
        self.kernels = torch.nn.ModuleList(kernels)

    def forward(self, x1, x2, **params):
        out = None
        for k in self.kernels:
            v = k(x1, x2, **params)
            out = v if out is None else out * v
        return out


def _scaled_dist(x1, x2, ls):
    a, b = x1 / ls, x2 / ls
    d2 = ((a.unsqueeze(-2) - b.unsqueeze(-3)) ** 2).sum(-1)
    return d2


class RBFKernel(Kernel):
    has_lengthscale = True

    def forward(self, x1, x2, **params):
        return torch.exp(-0.5 * _scaled_dist(x1, x2, self.lengthscale))


class MaternKernel(Kernel):
    has_lengthscale = True

    def __init__(self, nu=2.5, **kwargs):
        if nu not in (0.5, 1.5, 2.5):
            raise RuntimeError("nu expected to be 0.5, 1.5, or 2.5")
        super().__init__(**kwargs)
        self.nu = nu

    def forward(self, x1, x2, **params):
        d = torch.sqrt(_scaled_dist(x1, x2, self.lengthscale).clamp_min(1e-30))   # (gpytorch clamps the same way: finite gradient at d = 0)
        e = torch.exp(-math.sqrt(2.0 * self.nu) * d)
        if self.nu == 0.5:
            return e
        if self.nu == 1.5:
            return (1.0 + math.sqrt(3.0) * d) * e
        return (1.0 + math.sqrt(5.0) * d + 5.0 / 3.0 * d * d) * e


class PeriodicKernel(Kernel):
    has_lengthscale = True

    def __init__(self, period_length_prior=None, period_length_constraint=None, **kwargs):
        super().__init__(**kwargs)
        nls = 1 if self.ard_num_dims is None else int(self.ard_num_dims)
        self.register_parameter("raw_period_length", torch.nn.Parameter(torch.zeros(*self.batch_shape, 1, nls, dtype=torch.float64)))
        self.register_constraint("raw_period_length", period_length_constraint if period_length_constraint is not None else Positive())
        if period_length_prior is not None:
            self.register_prior("period_length_prior", period_length_prior, lambda m: m.period_length,
                                lambda m, v: m._set_period_length(v))

    @property
    def period_length(self):
        return self.raw_period_length_constraint.transform(self.raw_period_length)

    def _set_period_length(self, value):
        self.initialize(raw_period_length=self.raw_period_length_constraint.inverse_transform(torch.as_tensor(value, dtype=torch.float64)))

    def forward(self, x1, x2, **params):
        diff = (x1 / self.period_length).unsqueeze(-2) - (x2 / self.period_length).unsqueeze(-3)
        s = torch.sin(math.pi * diff) ** 2
        return torch.exp(-2.0 * (s / self.lengthscale.unsqueeze(-2)).sum(-1))


class ScaleKernel(Kernel):
    def __init__(self, base_kernel, outputscale_prior=None, outputscale_constraint=None, **kwargs):
        super().__init__(**kwargs)
        self.base_kernel = base_kernel
        self.register_parameter("raw_outputscale", torch.nn.Parameter(torch.zeros(tuple(self.batch_shape), dtype=torch.float64)))
        self.register_constraint("raw_outputscale", outputscale_constraint if outputscale_constraint is not None else Positive())
        if outputscale_prior is not None:
            self.register_prior("outputscale_prior", outputscale_prior, lambda m: m.outputscale, lambda m, v: m._set_outputscale(v))

    @property
    def outputscale(self):
        return self.raw_outputscale_constraint.transform(self.raw_outputscale)

    def _set_outputscale(self, value):
        self.initialize(raw_outputscale=self.raw_outputscale_constraint.inverse_transform(torch.as_tensor(value, dtype=torch.float64)))

    def forward(self, x1, x2, **params):
        return self.outputscale * self.base_kernel(x1, x2, **params)
