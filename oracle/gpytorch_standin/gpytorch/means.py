"""Mean base class and ConstantMean (one learnable `raw_constant`, no constraint)."""
import torch

from .module import Module


class Mean(Module):
    def forward(self, x):
        raise NotImplementedError

    def __call__(self, x):
        if x.ndim == 1:
            x = x.unsqueeze(-1)
        return torch.nn.Module.__call__(self, x)


class ConstantMean(Mean):
    def __init__(self, constant_prior=None, constant_constraint=None, batch_shape=torch.Size([]), **kwargs):
        super().__init__()
        self.register_parameter("raw_constant", torch.nn.Parameter(torch.zeros(tuple(batch_shape), dtype=torch.float64)))

    @property
    def constant(self):
        return self.raw_constant

    def forward(self, x):
        return self.constant.unsqueeze(-1).expand(x.shape[:-1])
