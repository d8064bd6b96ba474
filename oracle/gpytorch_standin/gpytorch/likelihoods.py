"""FixedNoiseGaussianLikelihood: fixed per-point noise (+ one learned homoskedastic `second_noise`, GreaterThan(1e-4))."""
import torch

from .constraints import GreaterThan
from .distributions import MultivariateNormal
from .module import Module


class FixedGaussianNoise(Module):
    def __init__(self, noise):
        super().__init__()
        self.noise = torch.as_tensor(noise, dtype=torch.float64).reshape(-1)


class HomoskedasticNoise(Module):
    def __init__(self, noise_prior=None, noise_constraint=None, batch_shape=torch.Size([])):
        super().__init__()
        self.register_parameter("raw_noise", torch.nn.Parameter(torch.zeros(*batch_shape, 1, dtype=torch.float64)))
        self.register_constraint("raw_noise", noise_constraint if noise_constraint is not None else GreaterThan(1e-4))
        if noise_prior is not None:
            self.register_prior("noise_prior", noise_prior, lambda m: m.noise, lambda m, v: m._set_noise(v))

    @property
    def noise(self):
        return self.raw_noise_constraint.transform(self.raw_noise)

    def _set_noise(self, value):
        self.initialize(raw_noise=self.raw_noise_constraint.inverse_transform(torch.as_tensor(value, dtype=torch.float64)))


class FixedNoiseGaussianLikelihood(Module):
    def __init__(self, noise, learn_additional_noise=False, batch_shape=torch.Size([]), **kwargs):
        super().__init__()
        self.noise_covar = FixedGaussianNoise(noise)
        self.second_noise_covar = None
        if learn_additional_noise:
            self.second_noise_covar = HomoskedasticNoise(noise_prior=kwargs.get("noise_prior"),
                                                         noise_constraint=kwargs.get("noise_constraint"), batch_shape=batch_shape)

    @property
    def noise(self):
        return self.noise_covar.noise

    @property
    def second_noise(self):
        return self.second_noise_covar.noise if self.second_noise_covar is not None else torch.zeros(1, dtype=torch.float64)

    def forward(self, function_dist, *params, **kwargs):
        n = function_dist.mean.shape[-1]
        fixed = self.noise_covar.noise
        if fixed.shape[0] != n:
            # new points: the fixed noise is only defined at the training inputs, only the learned noise applies (callers in
            # the reference that reach this, the monotonic-penalty callback, read .mean only)
            diag = self.second_noise.expand(n)
        else:
            diag = fixed + self.second_noise
        return MultivariateNormal(function_dist.mean, function_dist.covariance_matrix + torch.diag(diag))
