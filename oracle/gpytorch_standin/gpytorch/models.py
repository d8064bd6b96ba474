"""ExactGP: train mode returns the prior at the training inputs, eval mode the exact (latent) posterior, dense Cholesky."""
import torch

from .distributions import MultivariateNormal
from .module import Module


class ExactGP(Module):
    def __init__(self, train_inputs, train_targets, likelihood):
        super().__init__()
        if torch.is_tensor(train_inputs):
            train_inputs = (train_inputs,)
        self.train_inputs = tuple(t.unsqueeze(-1) if t.ndim == 1 else t for t in train_inputs)
        self.train_targets = train_targets
        self.likelihood = likelihood

    def forward(self, x):
        raise NotImplementedError

    def set_train_data(self, inputs=None, targets=None, strict=True):
        if inputs is not None:
            if torch.is_tensor(inputs):
                inputs = (inputs,)
            self.train_inputs = tuple(t.unsqueeze(-1) if t.ndim == 1 else t for t in inputs)
        if targets is not None:
            self.train_targets = targets

    def __call__(self, *args, **kwargs):
        x = args[0]
        if x.ndim == 1:
            x = x.unsqueeze(-1)
        if self.training:
            return torch.nn.Module.__call__(self, x, **kwargs)
        xt = self.train_inputs[0]
        full = torch.nn.Module.__call__(self, torch.cat([xt, x], dim=-2), **kwargs)
        n = xt.shape[-2]
        mean, cov = full.mean, full.covariance_matrix
        Kxx = self.likelihood(MultivariateNormal(mean[:n], cov[:n, :n])).covariance_matrix
        L = torch.linalg.cholesky(Kxx)
        Ks = cov[:n, n:]
        A = torch.linalg.solve_triangular(L, Ks, upper=False)
        z = torch.linalg.solve_triangular(L, (self.train_targets - mean[:n]).unsqueeze(-1), upper=False)
        return MultivariateNormal(mean[n:] + (A.transpose(-1, -2) @ z).squeeze(-1), cov[n:, n:] - A.transpose(-1, -2) @ A)
