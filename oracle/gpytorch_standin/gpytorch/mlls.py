"""ExactMarginalLogLikelihood: [log N(y | mean, K + noise) + sum of the registered priors' log densities] / n."""
from .module import Module


class ExactMarginalLogLikelihood(Module):
    def __init__(self, likelihood, model):
        super().__init__()
        self.likelihood, self.model = likelihood, model

    def forward(self, function_dist, target, *params, **kwargs):
        res = self.likelihood(function_dist).log_prob(target)
        for _name, module, prior, closure, _ in self.named_priors():
            res = res + prior.log_prob(closure(module)).sum()
        return res / target.shape[-1]
