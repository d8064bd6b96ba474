"""log densities of the three priors the reference models register (torch.distributions parameterisation)."""
import math

import torch


class Prior(torch.nn.Module):
    def log_prob(self, x):
        raise NotImplementedError


class NormalPrior(Prior):
    def __init__(self, loc, scale, **kw):
        super().__init__()
        self.loc, self.scale = float(loc), float(scale)

    def log_prob(self, x):
        return -0.5 * math.log(2.0 * math.pi) - math.log(self.scale) - (x - self.loc) ** 2 / (2.0 * self.scale ** 2)


class HalfNormalPrior(Prior):
    def __init__(self, scale, **kw):
        super().__init__()
        self.scale = float(scale)

    def log_prob(self, x):
        return math.log(2.0) - 0.5 * math.log(2.0 * math.pi) - math.log(self.scale) - x ** 2 / (2.0 * self.scale ** 2)


class GammaPrior(Prior):
    def __init__(self, concentration, rate, **kw):
        super().__init__()
        self.concentration, self.rate = float(concentration), float(rate)

    def log_prob(self, x):
        a, b = self.concentration, self.rate
        return a * math.log(b) - math.lgamma(a) + (a - 1.0) * torch.log(x) - b * x
