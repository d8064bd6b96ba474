"""gpytorch.Module: parameters with registered constraints (`<name>_constraint` sub-modules) and priors."""
import torch


class Module(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self._priors = {}   # name -> (prior, closure(module) -> value, setting_closure)

    def register_parameter(self, name, parameter):
        super().register_parameter(name, parameter)

    def register_constraint(self, param_name, constraint):
        self.add_module(param_name + "_constraint", constraint)

    def register_prior(self, name, prior, param_or_closure, setting_closure=None):
        if isinstance(param_or_closure, str):
            pname = param_or_closure
            closure = lambda m, _p=pname: getattr(m, _p)  # noqa: E731
        else:
            closure = param_or_closure
        self.add_module(name, prior)
        self._priors[name] = (prior, closure, setting_closure)

    def initialize(self, **kwargs):
        for name, val in kwargs.items():
            if name in self._parameters:
                p = self._parameters[name]
                with torch.no_grad():
                    p.copy_(torch.as_tensor(val, dtype=p.dtype).expand_as(p))
            else:
                setattr(self, name, val)
        return self

    def named_priors(self, memo=None, prefix=""):
        """(name, module, prior, closure, setting_closure) of every prior registered in this module tree, each once."""
        if memo is None:
            memo = set()
        yield from _named_priors(self, memo, prefix)

    def __call__(self, *args, **kwargs):
        return super().__call__(*args, **kwargs)


def _named_priors(mod, memo, prefix):
    for name, (prior, closure, setter) in getattr(mod, "_priors", {}).items():
        if id(prior) not in memo:
            memo.add(id(prior))
            yield prefix + ("." if prefix else "") + name, mod, prior, closure, setter
    for mname, child in mod.named_children():   # (containers such as ModuleList are plain torch modules: walk through them)
        yield from _named_priors(child, memo, prefix + ("." if prefix else "") + mname)
