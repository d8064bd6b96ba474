"""Multi-site batch: the B200 counterpart of the reference's `fexec.map(map_retrieval, sites)`
(examples/nwqn-loadest-example/nwqn-loadest-example.py:38-128,156-159).

Sites are independent GPs: nothing is exchanged while fitting.  One process per GPU (torchrun); sites are
assigned to ranks by longest-processing-time-first on the cost iterations * n^3; inside a rank several sites are
in flight at once, each on its own libdgp handle/stream (dgp_nlml_grad_launch / _wait), so that the latency-bound
panel steps of one site overlap the trailing updates of the others.  The only collective is the final gather.
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import capi
from .engine import JITTERS, MIN_VARIANCE
from .models import LOADEST_FIXED_NOISE, loadest_spec
from .spec import GPModule


def assign_sites(costs: Sequence[float], world: int) -> List[List[int]]:
    """Greedy LPT: sites sorted by decreasing cost, each to the least-loaded rank.  Deterministic on every rank."""
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    load = [0.0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += float(costs[i])
    return out


def site_cost(n: int, iterations: int, m_predict: int = 0) -> float:
    """flop model of one site: iterations * n^3 (NLML+grad) + factorise + n^2 per predicted point."""
    return float(iterations) * float(n) ** 3 + float(n) ** 3 + float(m_predict) * float(n) ** 2


def gather_results(local: Dict[int, dict], dist=None) -> Optional[Dict[int, dict]]:
    """All ranks -> rank 0.  `dist` is torch.distributed (initialised) or None for a single process."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return dict(local)
    bucket = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(local, bucket, dst=0)
    if dist.get_rank() != 0:
        return None
    merged: Dict[int, dict] = {}
    for part in bucket:
        merged.update(part)
    return merged


@dataclass
class _SiteState:
    idx: int
    X: np.ndarray
    y: np.ndarray
    noise: np.ndarray
    module: GPModule
    engine: capi.Engine
    opt: torch.optim.Optimizer
    sched: Optional[torch.optim.lr_scheduler.ReduceLROnPlateau]
    chain: Optional[tuple] = None  # GPModule.host_chain() of the evaluation in flight
    history: List[float] = field(default_factory=list)
    failed: Optional[str] = None


def _finish_step(s: _SiteState):
    """Host side of one optimiser step (discontinuum/engines/gpytorch.py:353-420) once the GPU results are back: the
    constraint / prior chain rule in closed form (GPModule.host_chain, no autograd graph), clipping, Adam, scheduler."""
    val, grad, info = s.engine.nlml_grad_wait()
    nat, dnat, lp, dlp = s.chain
    if info != 0 or not np.isfinite(val):
        for jit in JITTERS[1:]:
            val, grad, info = s.engine.nlml_grad(nat, jit)
            if info == 0 and np.isfinite(val):
                break
        else:
            s.failed = f"not positive definite (info={info})"
            return
    n = s.X.shape[0]
    obj = (val - lp) / n
    graw = (grad - dlp) * dnat / n
    coef = 1.0 / (float(np.sqrt(np.sum(graw * graw))) + 1e-6)   # clip_grad_norm_(max_norm=1.0)
    if not coef >= 1.0:
        graw = graw * coef
    if np.isnan(graw).any():
        graw = np.nan_to_num(graw, nan=0.0, posinf=0.0, neginf=0.0)
    for p, gv in zip(s.module.raw_list(), graw):
        p.grad = torch.tensor([gv], dtype=torch.float64)
    s.opt.step()
    s.history.append(float(obj))
    if s.sched is not None:
        s.sched.step(s.history[-1])


_COMPLETION_ORDER = True  # serve sites as their evaluations finish (False: fixed round, blocking on each in turn)


def _open_site(idx, tup, device, lr, scheduler, patience, pool: Optional[List[capi.Engine]] = None,
               free_parts: Optional[List[int]] = None) -> _SiteState:
    X, y = np.ascontiguousarray(tup[0], dtype=np.float64), np.ascontiguousarray(tup[1], dtype=np.float64)
    noise = np.ascontiguousarray(tup[2], dtype=np.float64) if len(tup) > 2 else np.full(y.shape[0], LOADEST_FIXED_NOISE)
    module = GPModule(loadest_spec(X.shape[1]))
    eng = None
    if pool:  # reuse the workspace of a finished (larger) site: no cudaMalloc / cudaFree (a device-wide sync) mid-batch
        k = next((i for i, e in enumerate(pool) if e.max_n >= X.shape[0]), None)
        if k is not None:
            eng = pool.pop(k)
    if eng is None:
        if free_parts is not None:  # one engine per SM partition: take a free partition, or rebuild a pooled engine's
            if not free_parts:
                old = pool.pop(0)
                free_parts.append(old.partition)
                old.close()
            part = free_parts.pop(0)
            eng = capi.Engine(max_n=X.shape[0], max_m=2048, device=device, partition=part)
            eng.partition = part
        else:
            eng = capi.Engine(max_n=X.shape[0], max_m=2048, device=device)
    eng.set_train(module.spec.to_c(), X, y, noise)
    opt = torch.optim.Adam(module.raw_list(), lr=lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4, fused=True)
    sch = None
    if scheduler:
        sch = torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode="min", factor=0.7, patience=max(20, patience // 2),
                                                         threshold=1e-4, threshold_mode="rel", min_lr=1e-6, cooldown=10)
    return _SiteState(idx, X, y, noise, module, eng, opt, sch)


def _launch_step(s: _SiteState):
    s.opt.zero_grad(set_to_none=True)
    s.chain = s.module.host_chain()
    s.engine.nlml_grad_launch(np.ascontiguousarray(s.chain[0]))


def _close_site(s: _SiteState, predict, pool: Optional[List[capi.Engine]] = None) -> dict:
    with torch.no_grad():
        theta = s.module.natural().numpy().astype(np.float64)
    res = {"theta": theta, "history": s.history, "objective": s.history[-1] if s.history else None, "failed": s.failed,
           "n": int(s.X.shape[0])}
    if predict is not None and s.idx in predict and s.failed is None:
        for jit in JITTERS:
            _, info = s.engine.factorize(theta, jit)
            if info == 0:
                break
        mu, var = s.engine.predict(predict[s.idx])
        res["mu"], res["var"] = mu, np.maximum(var, MIN_VARIANCE)
    if pool is not None:
        pool.append(s.engine)
    else:
        s.engine.close()
    return res


def fit_sites_local(sites: Dict[int, tuple], iterations: int = 100, device: int = 0, concurrency: int = 4,
                    predict: Optional[Dict[int, np.ndarray]] = None, lr: float = 0.05, scheduler: bool = True,
                    patience: int = 60, partitions: Optional[int] = None) -> Dict[int, dict]:
    """Fit the loadest-gp model on every site of this rank.  sites: {index: (X, y[, noise])} in model space.
    Returns {index: {"theta", "objective", "history", "mu", "var"}}.

    `concurrency` sites are in flight at any time, each a pipeline of its own: as soon as a site's evaluation is
    back the host does its optimiser step and enqueues its next evaluation, while the GPU works on the others; a
    finished site is predicted, closed and replaced by the next largest one (no group barrier).  Sites are served in
    completion order (`dgp_nlml_grad_ready`), not in a fixed round: their costs differ by up to (n_max / n_min)^3, and
    a round would make every site advance at the pace of the largest one in flight.

    partitions: split the GPU's SMs into that many disjoint partitions (`capi.partition_device`) and run one site per
    partition (concurrency = partitions).  Sites sharing all SMs slow each other down 3-4x (the short dependent kernels
    of one site's panel chain wait behind the long tiles of another's inverse); on its own slice a site of this size is
    work-bound and nobody waits."""
    results: Dict[int, dict] = {}
    queue = sorted(sites, key=lambda i: -sites[i][0].shape[0])  # largest first: later sites fit the pooled workspaces
    active: List[_SiteState] = []
    pool: List[capi.Engine] = []
    free_parts: Optional[List[int]] = None
    if partitions and partitions > 1:
        nparts, _ = capi.partition_device(device, partitions)
        concurrency = nparts
        free_parts = list(range(nparts))

    def refill():
        while queue and len(active) < max(1, concurrency):
            s = _open_site(queue[0], sites[queue.pop(0)], device, lr, scheduler, patience, pool, free_parts)
            if iterations > 0:
                _launch_step(s)
            active.append(s)

    refill()
    while active:
        served = False
        for s in list(active):
            if iterations > 0 and _COMPLETION_ORDER and not s.engine.nlml_grad_ready():
                continue
            served = True
            if iterations > 0:
                _finish_step(s)
            if s.failed is None and len(s.history) < iterations:
                _launch_step(s)
                continue
            active.remove(s)
            results[s.idx] = _close_site(s, predict, pool)
            refill()
        if not served:
            time.sleep(0)  # nothing finished yet: yield, then poll again
    for e in pool:
        e.close()
    return results


def fit_sites(sites: Dict[int, tuple], iterations: int = 100, predict: Optional[Dict[int, np.ndarray]] = None,
              concurrency: int = 4, dist=None, device: int = 0) -> Optional[Dict[int, dict]]:
    """Shard `sites` (every rank holds the same dict, or at least the same keys and sizes) over the ranks of
    `dist`, fit this rank's share, gather everything on rank 0."""
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    keys = sorted(sites)
    costs = [site_cost(sites[k][0].shape[0], iterations, predict[k].shape[0] if predict and k in predict else 0) for k in keys]
    mine = [keys[i] for i in assign_sites(costs, world)[rank]]
    local = fit_sites_local({k: sites[k] for k in mine}, iterations=iterations, device=device, concurrency=concurrency,
                            predict={k: predict[k] for k in mine} if predict else None)
    return gather_results(local, dist)


# ---------------------------------------------------------------- sharded prediction grid
def shard_rows(m: int, world: int) -> List[tuple]:
    """Contiguous row ranges [start, stop) of an m-point grid, one per rank, sizes differing by at most one."""
    base, rem = divmod(int(m), int(world))
    out, start = [], 0
    for r in range(world):
        stop = start + base + (1 if r < rem else 0)
        out.append((start, stop))
        start = stop
    return out


def predict_sharded(engine, Xs: np.ndarray, dist=None, want_var: bool = True):
    """Posterior mean / latent variance over a grid sharded across ranks.  Every rank holds the same factorisation
    (fitted redundantly or loaded from the same checkpoint: `engine.factorize` was called with the same theta) and
    predicts its contiguous slice; the only collective is the final all_gather of 16 m bytes.  `engine` is a
    capi.Engine (anything with .predict(Xs, want_var) -> (mu, var)).  Returns (mu[m], var[m]) on every rank."""
    Xs = np.ascontiguousarray(Xs, dtype=np.float64)
    m = Xs.shape[0]
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    lo, hi = shard_rows(m, world)[rank]
    if hi > lo:
        mu, var = engine.predict(Xs[lo:hi], want_var)
    else:
        mu, var = np.empty(0), (np.empty(0) if want_var else None)
    if world == 1:
        return mu, var
    parts = [None] * world
    dist.all_gather_object(parts, (np.asarray(mu), None if var is None else np.asarray(var)))
    mu_all = np.concatenate([p[0] for p in parts])
    var_all = np.concatenate([p[1] for p in parts]) if want_var else None
    return mu_all, var_all


# ---------------------------------------------------------------- joint posterior draws over a sharded grid
def panel_owner(p: int, world: int) -> int:
    """Panel-cyclic distribution of the posterior covariance's block columns."""
    return p % world


def sample_sharded(engine, Xs: np.ndarray, S: int, dist=None, seed: int = 0, jitter: float = 0.0, Z: Optional[np.ndarray] = None):
    """Exact joint posterior draws [S, m] with the m x m posterior-covariance Cholesky distributed over the ranks of
    `dist` (torch.distributed, NCCL).  Every rank holds the same training factorisation (`engine.factorize` at the same
    theta) and calls this with the same arguments; every rank returns (draws, info).

    Exchange steps (the only collectives): all_gather of V' (m x n), all_reduce of the mean, one broadcast of the
    sub-diagonal rows of each factored panel from its owner, all_reduce of the partial draws.  Look-ahead: the owner of
    panel p+1 updates that panel first, factors it on a high-priority side stream and broadcasts it while every rank
    is still applying panel p to the rest of its columns.  The engine must have been created on torch's current CUDA
    stream (`capi.Engine(stream=torch.cuda.current_stream().cuda_stream)`); with a private engine stream the kernels and
    collectives are ordered through host synchronisation instead (correct, no overlap)."""
    Xs = np.ascontiguousarray(Xs, dtype=np.float64)
    m = Xs.shape[0]
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    dims = engine.dist_dims(m, S, world)
    dev = torch.device("cuda", engine.device)
    f64 = dict(dtype=torch.float64, device=dev)
    rpr = dims["rows_per_rank"]
    VT = torch.zeros((rpr * world, dims["npad"]), **f64)
    packs = [torch.empty((dims["mpad"], dims["panel_cols"]), **f64) for _ in range(2)]
    Od = torch.empty((dims["Spad"], dims["mpad"]), **f64)
    mu = torch.empty((dims["mpad"],), **f64)
    T = torch.cuda.current_stream(dev)
    shared_stream = engine.stream != 0 and engine.stream == T.cuda_stream
    side = torch.cuda.Stream(device=dev, priority=-1) if shared_stream else None  # look-ahead stream

    def fence():  # engine on a private stream: order kernels and collectives through the host instead
        if not shared_stream:
            torch.cuda.synchronize(dev)

    pw_blocks, mb, npanels = dims["panel_cols"] // 128, dims["mpad"] // 128, dims["npanels"]

    def rows_below(p):
        return (mb - min((p + 1) * pw_blocks, mb)) * 128

    tok = engine.dist_begin(Xs, S, Z, seed, jitter, rank, world, VT, Od, mu)
    try:
        engine.dist_call("vt_rows", tok, rank * rpr, min((rank + 1) * rpr, dims["mpad"]))
        if world > 1:
            fence()
            dist.all_gather_into_tensor(VT, VT[rank * rpr:(rank + 1) * rpr].clone())
            dist.all_reduce(mu)
            fence()
        engine.dist_call("sigma", tok)
        # p = -1 bootstraps panel 0; iteration p leaves panel p+1 factored and in place on every rank
        for p in range(-1, npanels - 1):
            nxt = p + 1
            owner = panel_owner(nxt, world)
            buf = packs[nxt % 2]
            rows = rows_below(nxt)
            work = None
            if owner == rank:
                if p >= 0:
                    engine.dist_call("trail", tok, p, nxt, nxt)          # panel p -> panel p+1 first
                if side is not None:
                    side.wait_event(T.record_event())
                    engine.dist_call("panel_factor", tok, nxt, buf.data_ptr(), side.cuda_stream)
                else:
                    engine.dist_call("panel_factor", tok, nxt, buf.data_ptr(), None)
            if world > 1 and rows > 0:
                if side is not None:
                    if owner != rank:
                        side.wait_event(T.record_event())              # the receive buffer is free once T got here
                    with torch.cuda.stream(side):
                        work = dist.broadcast(buf[:rows], src=owner, async_op=True)
                else:
                    fence()
                    dist.broadcast(buf[:rows], src=owner)
                    fence()
            if p >= 0:
                engine.dist_call("trail", tok, p, nxt + 1, npanels - 1)  # ... then the rest, overlapping the above
            if work is not None:
                work.wait()                                                # T waits for the broadcast
            elif side is not None and owner == rank:
                T.wait_stream(side)
            if owner != rank and rows > 0:
                engine.dist_call("panel_unpack", tok, nxt, buf.data_ptr())
        engine.dist_call("draws_partial", tok)
        if world > 1:
            fence()
            dist.all_reduce(Od)
            fence()
        out = np.empty((S, m))
        info = engine.dist_call("finish", tok, out.ctypes.data)
        if world > 1:
            t = torch.tensor([info if info > 0 else 2 ** 31 - 1], dtype=torch.int64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)  # first failing pivot over all ranks
            info = 0 if int(t) == 2 ** 31 - 1 else int(t)
        return out, info
    finally:
        engine.dist_call("end", tok)
