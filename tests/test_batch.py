"""Batched, variable-size multi-site evaluation (dgp_batch_*): every site of a batch must get exactly the numbers the
per-handle path (dgp_nlml_grad) gives it -- same tiles, same operand order, bit for bit -- for ragged mixes of sizes,
both models, per-site jitter and failing sites; and the per-handle path is what the oracle parity tests pin."""
import time

import numpy as np
import pytest

import helpers as H
from discontinuum_b200 import capi, models, synthetic

pytestmark = pytest.mark.gpu


def _single(spec, site, theta, jitter=0.0):
    X, y, noise = site
    eng = capi.Engine(max_n=X.shape[0], max_m=128)
    eng.set_train(spec.to_c(), X, y, noise)
    val, grad, info = eng.nlml_grad(theta, jitter)
    alpha = eng.alpha() if info == 0 else None
    eng.close()
    return val, grad, info, alpha


def _loadest_batch(ns, seed0=2000):
    sites = [synthetic.loadest_site(n, seed0 + k) for k, n in enumerate(ns)]
    thetas = np.stack([H.loadest_theta1() * (1.0 + 0.01 * k) for k in range(len(ns))])
    return sites, thetas


def _assert_identical(batch, sites, spec, thetas, jit=None):
    val, grad, info = batch.nlml_grad(thetas, jit)
    for k, site in enumerate(sites):
        v, g, i, a = _single(spec, site, thetas[k], 0.0 if jit is None else float(jit[k]))
        assert info[k] == i
        if i != 0:
            continue
        assert val[k] == v, (k, site[0].shape[0], val[k], v)
        assert np.array_equal(grad[k], g), (k, grad[k] - g)
        assert np.array_equal(batch.alpha(k), a)
    return val, grad, info


@pytest.mark.parametrize("ns", [(300,), (300, 450, 700, 129, 1100), (128, 128, 128), (1100, 31, 640, 513, 512, 90, 257)])
def test_batch_bit_identical_to_per_handle_loadest(cuda_device, ns):
    spec = models.loadest_spec(2)
    sites, thetas = _loadest_batch(ns)
    b = capi.BatchEngine(max_sites=len(ns), max_n=max(ns))
    b.set_train(spec.to_c(), sites)
    _assert_identical(b, sites, spec, thetas)
    # evaluate again (slabs reused), at other hyper-parameters
    _assert_identical(b, sites, spec, thetas * 1.03)
    b.close()


def test_batch_many_panels_and_regroup(cuda_device):
    """Sizes of BASELINE config 4 (several panels of 4 block columns, different end-alignment offsets), then the same
    handle re-used for a smaller group (smaller leading dimension)."""
    spec = models.loadest_spec(2)
    ns = (2565, 4100, 3333, 2048)
    sites, thetas = _loadest_batch(ns, 3000)
    b = capi.BatchEngine(max_sites=6, max_n=4100)
    b.set_train(spec.to_c(), sites)
    l0 = b.launches
    _assert_identical(b, sites, spec, thetas)
    per_eval = b.launches - l0
    ns2 = (900, 1500, 200, 1234, 77)
    sites2, thetas2 = _loadest_batch(ns2, 3100)
    b.set_train(spec.to_c(), sites2)
    _assert_identical(b, sites2, spec, thetas2)
    b.close()
    assert per_eval < 250  # one launch sequence for the four sites (a single n = 4100 site alone takes ~170)


def test_batch_rating_gauges(cuda_device):
    ns = (200, 1000, 600)
    sites = [synthetic.rating_gauge(n, 7 + k) for k, n in enumerate(ns)]
    # one spec for the batch: the interval bounds of the gate switch point only constrain the raw value on the host
    spec = models.rating_spec(1.0, 2.0)
    thetas = []
    for X, _, _ in sites:
        b_lo, b_hi = models.stage_quantile_bounds(X[:, 1])
        thetas.append(H.rating_theta1(b_lo, b_hi))
    thetas = np.stack(thetas)
    b = capi.BatchEngine(max_sites=3, max_n=1000)
    b.set_train(spec.to_c(), sites)
    _assert_identical(b, sites, spec, thetas)
    b.close()


def test_batch_per_site_jitter_and_failing_site(cuda_device):
    spec = models.loadest_spec(2)
    ns = (300, 500, 260)
    sites, thetas = _loadest_batch(ns, 4000)
    X, y, noise = sites[1]
    noise = noise.copy()
    noise[200:] = -3.0  # indefinite from row 200 on
    sites[1] = (X, y, noise)
    b = capi.BatchEngine(max_sites=3, max_n=500)
    b.set_train(spec.to_c(), sites)
    val, grad, info = _assert_identical(b, sites, spec, thetas)
    assert info[0] == 0 and info[2] == 0 and 200 < info[1] <= 500
    jit = np.array([1e-6, 3.5, 0.0])
    val2, grad2, info2 = _assert_identical(b, sites, spec, thetas, jit)
    assert np.all(info2 == 0) and np.all(np.isfinite(val2))
    assert val2[2] == val[2] and val2[0] != val[0]
    # asynchronous form
    b.nlml_grad_launch(thetas, jit)
    t0 = time.time()
    while not b.nlml_grad_ready():
        assert time.time() - t0 < 30.0
    val3, grad3, info3 = b.nlml_grad_wait()
    assert np.array_equal(val3, val2) and np.array_equal(grad3, grad2) and np.array_equal(info3, info2)
    with pytest.raises(capi.DgpError):
        b.nlml_grad_ready()
    b.close()


def test_batch_bad_arguments(cuda_device):
    with pytest.raises(capi.DgpError):
        capi.BatchEngine(max_sites=capi.BATCH_MAX_SITES + 1, max_n=256)
    b = capi.BatchEngine(max_sites=2, max_n=256)
    spec = models.loadest_spec(2)
    sites, thetas = _loadest_batch((100, 200, 150))
    with pytest.raises(capi.DgpError, match="max_sites"):
        b.set_train(spec.to_c(), sites)
    with pytest.raises(capi.DgpError, match="max_n"):
        b.set_train(spec.to_c(), [synthetic.loadest_site(300, 1)])
    b.set_train(spec.to_c(), sites[:2])
    with pytest.raises(ValueError):
        b.nlml_grad(thetas)  # three rows for two sites
    b.close()


def test_batch_small_sites_amortised_time(cuda_device):
    """16 concurrent n = 1000 evaluations in one launch sequence (BASELINE config 1 size): amortised device time per
    evaluation, printed for the record; the batch must beat 16 evaluations one after the other."""
    spec = models.loadest_spec(2)
    ns = (1000,) * 16
    sites, thetas = _loadest_batch(ns, 5000)
    b = capi.BatchEngine(max_sites=16, max_n=1000)
    b.set_train(spec.to_c(), sites)
    for _ in range(3):
        b.nlml_grad(thetas)
    t0 = time.perf_counter()
    reps = 10
    for _ in range(reps):
        b.nlml_grad(thetas)
    t_batch = (time.perf_counter() - t0) / reps
    b.close()
    eng = capi.Engine(max_n=1000, max_m=128)
    eng.set_train(spec.to_c(), *sites[0])
    for _ in range(3):
        eng.nlml_grad(thetas[0])
    t0 = time.perf_counter()
    for _ in range(reps):
        eng.nlml_grad(thetas[0])
    t_one = (time.perf_counter() - t0) / reps
    eng.close()
    print(f"n=1000: batch of 16 {t_batch * 1e3:.3f} ms = {t_batch / 16 * 1e3:.3f} ms per evaluation; single {t_one * 1e3:.3f} ms")
    assert t_batch < 16 * t_one
