"""pm.gp.Marginal: additive composition and marginal_likelihood (zero mean)."""
import math

import torch

from . import cov  # noqa: F401


class Marginal:
    def __init__(self, mean_func=None, cov_func=None):
        self.cov_func = cov_func

    def __add__(self, other):
        return Marginal(cov_func=self.cov_func + other.cov_func)

    def marginal_likelihood(self, name, X, y, sigma=None, noise=None, jitter=1e-6, **kw):
        import pymc
        noise = sigma if sigma is not None else noise
        X = torch.as_tensor(X, dtype=torch.float64)
        y = torch.as_tensor(y, dtype=torch.float64)
        model = pymc._STACK[-1]

        def logp():
            K = self.cov_func(X)
            Kn = noise(X) if isinstance(noise, cov.Covariance) else float(noise) ** 2 * torch.eye(X.shape[0], dtype=torch.float64)
            Ky = K + Kn + jitter * torch.eye(X.shape[0], dtype=torch.float64)   # pm.gp.Marginal adds its own jitter (1e-6)
            L = torch.linalg.cholesky(Ky)
            z = torch.linalg.solve_triangular(L, y.unsqueeze(-1), upper=False).squeeze(-1)
            return -0.5 * (z * z).sum() - torch.log(torch.diagonal(L)).sum() - 0.5 * y.shape[0] * math.log(2.0 * math.pi)

        model.observed = logp
        return None
