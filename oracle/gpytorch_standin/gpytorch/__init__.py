"""Stand-in for gpytorch (see ../README.md).  Dense float64 torch; gpytorch's attribute and parameter names."""
from .module import Module  # noqa: F401
from . import constraints, distributions, kernels, likelihoods, means, mlls, models, priors, settings  # noqa: F401
