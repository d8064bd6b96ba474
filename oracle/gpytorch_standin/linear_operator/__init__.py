"""Stand-in for linear_operator (see ../README.md): only `operators.MatmulLinearOperator`."""
from . import operators  # noqa: F401
