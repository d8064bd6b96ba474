"""Synthetic river-like inputs in MODEL SPACE: what DataManager.X / .y / .y_unc would hand
the engine (src/discontinuum/data_manager.py:82-95) for a loadest site or a rating gauge."""
from __future__ import annotations

import numpy as np


def loadest_site(n: int, seed: int = 1000):
    """X = [centred decimal year in (-15, 15), standardised log-flow], y = standardised log-concentration."""
    rng = np.random.default_rng(seed)
    t = np.sort(rng.uniform(-15.0, 15.0, n))
    # AR(1) log-flow, rho = 0.97 per 0.01 yr, irregular spacing
    e = rng.standard_normal(n)
    q = np.empty(n)
    q[0] = e[0]
    dt = np.diff(t, prepend=t[0])
    rho = 0.97 ** (dt / 0.01)
    for i in range(1, n):
        q[i] = rho[i] * q[i - 1] + np.sqrt(max(1.0 - rho[i] ** 2, 0.0)) * e[i]
    q = 0.8 * np.sin(2 * np.pi * t) + q
    q = (q - q.mean()) / q.std()
    f = 0.6 * q + 0.4 * np.sin(2 * np.pi * t + 0.5) + 0.03 * t + 0.3 * np.sin(2 * np.pi * t) * np.tanh(q)
    y = f + 0.3 * rng.standard_normal(n)
    y = (y - y.mean()) / y.std()
    X = np.ascontiguousarray(np.stack([t, q], axis=1))
    return X, y, np.full(n, 0.1 ** 2)


def rating_gauge(n: int = 2000, seed: int = 7):
    """X = [centred decimal year, stage unit-scaled to [1, 2]], y = standardised log-discharge,
    noise = model-space variance from measurement-quality GSEs (src/rating_gp/providers/usgs.py:21-27)."""
    rng = np.random.default_rng(seed)
    t = np.sort(rng.uniform(-16.5, 16.5, n))
    s = rng.lognormal(0.0, 0.6, n)
    h = 1.0 + (s - s.min()) / (s.max() - s.min())
    c = 0.6 + 0.05 * np.sin(2 * np.pi * t / 20.0)
    hb = np.quantile(h, 0.6)
    logq = 1.0 + 1.6 * np.log(h - c) + 0.4 * np.maximum(h - hb, 0.0)
    gse = rng.choice(np.array([1.01, 1.025, 1.04, 1.06]), n)
    logq = logq + np.log(gse) * rng.standard_normal(n)
    sd = logq.std()
    y = (logq - logq.mean()) / sd
    noise = (np.log(gse) / sd) ** 2
    X = np.ascontiguousarray(np.stack([t, h], axis=1))
    return X, y, noise


def daily_grid(X: np.ndarray, m: int, seed: int = 0):
    """m prediction points: regular in time over the record, covariate interpolated from the training record."""
    t = np.linspace(X[:, 0].min(), X[:, 0].max(), m)
    cols = [t]
    for d in range(1, X.shape[1]):
        cols.append(np.interp(t, X[:, 0], X[:, d]))
    return np.ascontiguousarray(np.stack(cols, axis=1))
