"""Context managers the reference enters around prediction (no-ops here: everything is exact and dense)."""
import contextlib


def _noop(*a, **k):
    return contextlib.nullcontext()


fast_pred_var = fast_computations = max_cholesky_size = cholesky_jitter = _noop
