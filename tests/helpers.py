"""Glue between the oracle's parameter dicts and the engine's theta vectors (test infrastructure)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import gp_oracle as orc  # noqa: E402

DT = torch.float64


def loadest_theta_from_nat(nat):
    return np.concatenate([np.atleast_1d(np.asarray(nat[k], dtype=np.float64)).ravel()
                           for k in ("mean_c", "s1", "lam", "period", "l1", "s2", "l2", "s3", "l3")])


def loadest_nat_from_theta(theta, d=2):
    th = torch.as_tensor(theta, dtype=DT)
    return {"mean_c": th[0:1], "s1": th[1:2], "lam": th[2:3], "period": th[3:4], "l1": th[4:5], "s2": th[5:6],
            "l2": th[6:6 + d - 1], "s3": th[5 + d:6 + d], "l3": th[6 + d:6 + 2 * d]}


RATING_KEYS = ("pl_a", "pl_b", "pl_c", "noise", "gate_b", "shiftA_s", "shiftA_lh", "shiftA_lt", "shiftB_s", "shiftB_lh",
               "shiftB_lt", "bend_s", "bend_lh", "bend_lt", "base_s", "base_l", "per_s", "per_period", "per_lam", "per_l")


def rating_theta_from_nat(nat):
    return np.array([float(np.asarray(nat[k]).reshape(-1)[0]) for k in RATING_KEYS])


def rating_nat_from_theta(theta):
    th = torch.as_tensor(theta, dtype=DT)
    return {k: th[i:i + 1] for i, k in enumerate(RATING_KEYS)}


def loadest_theta0(d=2):
    return loadest_theta_from_nat({k: v.numpy() for k, v in orc.loadest_natural(orc.loadest_init_raw(d)).items()})


def loadest_theta1():
    """fitted-like set of SURVEY 8d."""
    return np.array([0.05, 0.7, 1.0, 1.0, 2.0, 1.3, 0.5, 0.2, 0.3, 0.4])


def rating_theta0(b_lo, b_hi):
    raw = orc.rating_init_raw(b_lo, b_hi)
    return rating_theta_from_nat(orc.rating_natural(raw, b_lo, b_hi))


def rating_theta1(b_lo, b_hi):
    nat = dict(pl_a=0.1, pl_b=1.6, pl_c=0.55, noise=0.002, gate_b=0.5 * (b_lo + b_hi) + 0.02, shiftA_s=0.4, shiftA_lh=1.2,
               shiftA_lt=2.5, shiftB_s=0.1, shiftB_lh=2.0, shiftB_lt=0.15, bend_s=0.3, bend_lh=1.0, bend_lt=3.0,
               base_s=0.9, base_l=0.8, per_s=0.05, per_period=1.0, per_lam=0.9, per_l=4.0)
    return rating_theta_from_nat(nat)


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))


def philox_normals(seed: int, count: int) -> np.ndarray:
    """numpy restatement of csrc/dgp_panel.cuh::philox_normal: Philox4x32-10 on counter (e >> 1, 0), key = seed,
    two 53-bit uniforms, Box-Muller (cos for even e, sin for odd e)."""
    e = np.arange(count, dtype=np.uint64)
    ctr = e >> np.uint64(1)
    c0, c1 = (ctr & np.uint64(0xFFFFFFFF)), (ctr >> np.uint64(32))
    c2, c3 = np.zeros_like(c0), np.zeros_like(c0)
    k0, k1 = np.uint64(seed & 0xFFFFFFFF), np.uint64((seed >> 32) & 0xFFFFFFFF)
    M0, M1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
        k0, k1 = (k0 + np.uint64(0x9E3779B9)) & MASK, (k1 + np.uint64(0xBB67AE85)) & MASK
    a, b = (c1 << np.uint64(32)) | c0, (c3 << np.uint64(32)) | c2
    u1 = ((a >> np.uint64(11)).astype(np.float64) + 0.5) / 9007199254740992.0
    u2 = ((b >> np.uint64(11)).astype(np.float64) + 0.5) / 9007199254740992.0
    rad = np.sqrt(-2.0 * np.log(u1))
    ang = 2.0 * np.pi * u2
    return rad * np.where((e & np.uint64(1)) == 1, np.sin(ang), np.cos(ang))
