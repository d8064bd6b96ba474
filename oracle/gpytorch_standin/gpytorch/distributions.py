"""MultivariateNormal with a dense covariance."""
import math

import torch


class MultivariateNormal:
    def __init__(self, mean, covariance_matrix):
        self.mean = mean
        self.covariance_matrix = covariance_matrix.to_dense() if hasattr(covariance_matrix, "to_dense") else covariance_matrix

    @property
    def loc(self):
        return self.mean

    @property
    def variance(self):
        return torch.diagonal(self.covariance_matrix, dim1=-1, dim2=-2)

    @property
    def stddev(self):
        return self.variance.clamp_min(0.0).sqrt()

    def log_prob(self, value):
        r = value - self.mean
        L = torch.linalg.cholesky(self.covariance_matrix)
        z = torch.linalg.solve_triangular(L, r.unsqueeze(-1), upper=False).squeeze(-1)
        n = r.shape[-1]
        return -0.5 * (z * z).sum(-1) - torch.log(torch.diagonal(L, dim1=-1, dim2=-2)).sum(-1) - 0.5 * n * math.log(2.0 * math.pi)
