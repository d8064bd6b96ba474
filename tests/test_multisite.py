"""Site sharding: CPU tests of the assignment and of the final gather with world_size = 2 on gloo, plus a GPU
test that the concurrent multi-site driver reproduces single-site fits."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from discontinuum_b200 import multisite


def test_assign_sites_lpt_balanced_and_complete():
    rng = np.random.default_rng(42)
    ns = (2000 + 6000 * rng.uniform(size=128)).astype(int)
    costs = [multisite.site_cost(int(n), 100, 10958) for n in ns]
    for world in (1, 2, 4, 8):
        parts = multisite.assign_sites(costs, world)
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(128))
        loads = [sum(costs[i] for i in p) for p in parts]
        assert max(loads) / (sum(loads) / world) < 1.05
    assert multisite.assign_sites(costs, 4) == multisite.assign_sites(list(costs), 4)  # deterministic


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    costs = [float(c) for c in (9, 7, 5, 4, 3, 1)]
    mine = multisite.assign_sites(costs, world)[rank]
    local = {i: {"theta": np.full(3, float(i)), "rank": rank} for i in mine}
    merged = multisite.gather_results(local, dist)
    if rank == 0:
        assert sorted(merged) == list(range(6))
        assert all(np.array_equal(merged[i]["theta"], np.full(3, float(i))) for i in merged)
        assert {merged[i]["rank"] for i in merged} == {0, 1}
        open(os.path.join(out_dir, "ok"), "w").write("ok")
    else:
        assert merged is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_results_world2_gloo(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()


class _FakeEngine:
    """Stands in for capi.Engine on the CPU: 'predicts' a known function of the inputs."""

    def predict(self, Xs, want_var=True):
        return Xs[:, 0] * 2.0 + Xs[:, 1], (Xs[:, 0] ** 2 if want_var else None)


def _pred_worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    Xs = np.stack([np.arange(11.0), np.arange(11.0)[::-1]], axis=1)
    mu, var = multisite.predict_sharded(_FakeEngine(), Xs, dist)
    assert np.array_equal(mu, Xs[:, 0] * 2.0 + Xs[:, 1]) and np.array_equal(var, Xs[:, 0] ** 2)
    mu2, var2 = multisite.predict_sharded(_FakeEngine(), Xs[:1], dist, want_var=False)  # fewer points than ranks
    assert mu2.shape == (1,) and var2 is None
    open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_shard_rows_and_sharded_prediction_world2_gloo(tmp_path):
    assert multisite.shard_rows(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert multisite.shard_rows(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
    for m, w in ((100000, 8), (10958, 3), (7, 7)):
        rows = multisite.shard_rows(m, w)
        assert rows[0][0] == 0 and rows[-1][1] == m and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
    mp.spawn(_pred_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


@pytest.mark.gpu
def test_concurrent_sites_match_single_site_fits(cuda_device):
    from discontinuum_b200 import synthetic
    import helpers as H
    from helpers import orc

    sites = {i: synthetic.loadest_site(n, 1000 + i)[:2] for i, n in enumerate((300, 450, 700))}
    grid = {i: synthetic.daily_grid(sites[i][0], 200) for i in sites}
    res = multisite.fit_sites(sites, iterations=8, predict=grid, concurrency=3)
    assert sorted(res) == [0, 1, 2]
    for i, (X, y) in sites.items():
        raw = orc.loadest_init_raw()
        _, hist = orc.fit_adam("loadest", raw, torch.tensor(X), torch.tensor(y), orc.loadest_noise(X.shape[0]), iterations=8)
        assert res[i]["failed"] is None
        assert np.max(np.abs(np.array(res[i]["history"]) - np.array(hist)) / np.abs(hist)) <= 1e-6
        assert res[i]["mu"].shape == (200,) and np.all(res[i]["var"] > 0)


@pytest.mark.gpu
def test_sites_on_sm_partitions_match_shared_gpu(cuda_device):
    from discontinuum_b200 import synthetic

    sites = {i: synthetic.loadest_site(n, 1100 + i)[:2] for i, n in enumerate((500, 260, 380, 640))}
    shared = multisite.fit_sites_local(sites, iterations=6, concurrency=2)
    split = multisite.fit_sites_local(sites, iterations=6, partitions=2)
    for i in sites:
        assert split[i]["failed"] is None
        assert np.max(np.abs(np.array(split[i]["history"]) - np.array(shared[i]["history"]))) <= 1e-9
        assert np.max(np.abs(split[i]["theta"] - shared[i]["theta"])) <= 1e-9


@pytest.mark.gpu
def test_sample_sharded_single_rank_matches_engine_sample(cuda_device):
    """world = 1: the panel-by-panel distributed schedule reproduces dgp_sample (same Philox normals)."""
    import helpers as H
    from discontinuum_b200 import capi, models, synthetic

    n, m, S = 400, 1300, 12
    X, y, noise = synthetic.loadest_site(n, 51)
    Xs = synthetic.daily_grid(X, m) + np.array([0.0009, 0.0])
    eng = capi.Engine(max_n=n, max_m=512)
    eng.set_train(models.loadest_spec(2).to_c(), X, y, noise)
    eng.factorize(H.loadest_theta1())
    want, info = eng.sample_ex(Xs, S, Z=None, seed=99, jitter=1e-7)
    got, info2 = multisite.sample_sharded(eng, Xs, S, dist=None, seed=99, jitter=1e-7)
    assert info == 0 and info2 == 0
    assert np.max(np.abs(got - want)) <= 1e-9 * np.max(np.abs(want))
    eng.close()


def test_panel_owner_cyclic():
    assert [multisite.panel_owner(p, 4) for p in range(6)] == [0, 1, 2, 3, 0, 1]
