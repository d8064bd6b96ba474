"""Interval (scaled sigmoid), GreaterThan / Positive (softplus + lower bound): gpytorch's default transforms."""
import torch
import torch.nn.functional as F


class Interval(torch.nn.Module):
    def __init__(self, lower_bound, upper_bound, transform=None, inv_transform=None, initial_value=None):
        super().__init__()
        self.register_buffer("lower_bound", torch.as_tensor(float(lower_bound), dtype=torch.float64))
        self.register_buffer("upper_bound", torch.as_tensor(float(upper_bound), dtype=torch.float64))
        self.initial_value = initial_value

    def transform(self, raw):
        return self.lower_bound + (self.upper_bound - self.lower_bound) * torch.sigmoid(raw)

    def inverse_transform(self, value):
        u = (value - self.lower_bound) / (self.upper_bound - self.lower_bound)
        return torch.log(u) - torch.log1p(-u)


class GreaterThan(Interval):
    def __init__(self, lower_bound, transform=None, inv_transform=None, initial_value=None):
        super().__init__(lower_bound, float("inf"), initial_value=initial_value)

    def transform(self, raw):
        return F.softplus(raw) + self.lower_bound

    def inverse_transform(self, value):
        v = value - self.lower_bound
        return v + torch.log(-torch.expm1(-v))


class Positive(GreaterThan):
    def __init__(self, transform=None, inv_transform=None, initial_value=None):
        super().__init__(0.0, initial_value=initial_value)
