"""discontinuum_b200 -- B200-native exact-GP engine for discontinuum's marginal-likelihood path.

The CUDA engine (libdgp.so, C ABI in include/dgp.h) does all O(n^2)/O(n^3) work; this package is
the thin Python host that mirrors the reference's Marginal* engine surface.  No CPU fallback.
"""
__version__ = "0.1.0"
