"""bench.py -- NLML + hyper-parameter-gradient evaluations per second at n = 16 384 (BASELINE.json metric,
configs[2]: "loadest-gp large single site n=16k, FP64 Cholesky + hyperparameter gradient, 1xB200").

A step = one evaluation of {NLML, dNLML/dtheta} of the loadest-gp model on one synthetic river-like site.
  value   : evaluations/s with the training set already resident in HBM (set_train done before timing),
            timed with CUDA events on the stream the engine launches on.
  e2e     : the same through the C ABI with HOST buffers every step: dgp_set_train(host X, y, noise)
            + dgp_nlml_grad(host theta -> host nlml, grad).
  N > 1   : one process per GPU, every rank evaluates its own site (sites are independent: no data-path
            collective, "weak" scaling); value = total evaluations / max-over-ranks time.
  --impl reference : the CPU restatement of the reference's path (oracle, torch float64, all host threads)
            on a bounded sample of the same workload.
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

N_TRAIN = 16384
METRIC = "nlml_grad_evals_per_sec_n16k"
UNIT = "evals/s"
CPU_SAMPLE_N = 4096


def fp64_peak():
    """FP64 roofline denominator.  MEASURED_PEAKS.json (driver-written) has no FP64 entry, so the figure is
    this repo's own measurement with the same method (torch.matmul float64 8192^3, tools/fp64_peak.py)."""
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            mp = json.load(f)
        if "fp64_tflops" in mp:
            return float(mp["fp64_tflops"]), "MEASURED_PEAKS.json fp64_tflops"
    except Exception:  # noqa: BLE001
        pass
    with open(os.path.join(ROOT, "profiles", "fp64_peak_r01.json")) as f:
        return float(json.load(f)["fp64_tflops"]), "profiles/fp64_peak_r01.json (cuBLAS DGEMM 8192^3 on this pool's B200; MEASURED_PEAKS.json has no fp64 entry)"


class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, nm in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_eval_seconds(n: int, reps: int, warm: int):
    """Oracle (CPU restatement) NLML+grad at size n: dense covariance build + Cholesky + inverse + closed-form
    gradient contraction, torch float64 with every host thread."""
    import torch

    import helpers as H
    from discontinuum_b200 import synthetic
    from oracle import gp_oracle as orc

    X, y, noise = synthetic.loadest_site(n, 1000)
    Xt, yt, nt = torch.tensor(X), torch.tensor(y), torch.tensor(noise)
    nat = H.loadest_nat_from_theta(H.loadest_theta1())
    times = []
    for r in range(warm + reps):
        t0 = time.perf_counter()
        orc.nlml_grad_closed_form(orc.loadest_cov, orc.loadest_mean, nat, Xt, yt, nt)
        dt = time.perf_counter() - t0
        if r >= warm:
            times.append(dt)
    return times, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    torch.set_num_threads(os.cpu_count() or 1)
    ns = CPU_SAMPLE_N
    times, threads = cpu_eval_seconds(ns, args.steps, args.warmup)
    per = sum(times) / len(times)
    scale = (N_TRAIN / ns) ** 3
    value = 1.0 / (per * scale)
    sample = (f"oracle NLML+grad at n={ns} ({per:.2f} s/eval measured, {threads} torch threads), extrapolated to n={N_TRAIN} "
              f"by (n/{ns})^3 = {scale:.0f}x")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": per * scale * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"loadest-gp single site n={N_TRAIN}, NLML + gradient (10 hyper-parameters)", "cpu_sample_n": ns},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--n", type=int, default=N_TRAIN, help="development only; the judged workload is n=16384")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (development)")
    ap.add_argument("--no-extra", action="store_true", help="skip the predictive-points/s and sites/s legs (development)")
    ap.add_argument("--sites-per-gpu", type=int, default=16, help="sites of the NWQN-style batch fitted per GPU in the sites/s leg")
    ap.add_argument("--site-iterations", type=int, default=100)
    ap.add_argument("--predict-m", type=int, default=32768, help="grid points of the predictive-points/s leg")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch

    import helpers as H
    from discontinuum_b200 import capi, models, synthetic

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    # torchrun exports OMP_NUM_THREADS=1; the per-site host steps (torch CPU ops) were measured faster with a few threads
    torch.set_num_threads(max(1, min(8, (os.cpu_count() or 8) // max(world, 1))))
    dist = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep NCCL's banner off stdout: rank 0 prints ONE JSON line
        import torch.distributed as dist_mod

        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    n = args.n
    warm = max(args.warmup, 3)
    X, y, noise = synthetic.loadest_site(n, 1000 + rank)
    spec = models.loadest_spec(2)
    work_stream = torch.cuda.Stream(device=local)  # the engine launches on this stream; the events below are recorded on it
    torch.cuda.set_stream(work_stream)
    eng = capi.Engine(max_n=n, max_m=2048, device=local, stream=work_stream.cuda_stream)
    eng.set_train(spec.to_c(), X, y, noise)
    base = H.loadest_theta1()
    thetas = [base * (1.0 + 1e-3 * k) for k in range(warm + args.steps)]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM
    for k in range(warm):
        eng.nlml_grad(thetas[k])
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    l0 = eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    infos = []
    for k in range(args.steps):
        val, grad, info = eng.nlml_grad(thetas[warm + k])
        infos.append(info)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launches - l0
    clocks = sampler.stop()
    if any(infos) or not np.isfinite(val):
        raise RuntimeError(f"factorisation failed inside the timed region: info={infos} nlml={val}")

    # ---- phase split of one evaluation (CUDA events on the launching stream inside libdgp)
    eng.set_timing(True)
    eng.nlml_grad(base)
    phase = eng.last_timing()
    eng.set_timing(False)

    # ---- e2e: host buffers through the C ABI every step
    Xh, yh, nh = np.ascontiguousarray(X), np.ascontiguousarray(y), np.ascontiguousarray(noise)
    c_spec = spec.to_c()
    for k in range(2):
        eng.set_train(c_spec, Xh, yh, nh)
        eng.nlml_grad(thetas[k])
    barrier()
    t0 = time.perf_counter()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for k in range(args.steps):
        eng.set_train(c_spec, Xh, yh, nh)
        val2, grad2, info2 = eng.nlml_grad(thetas[warm + k])
    e3.record()
    barrier()
    ms_e2e = max(e2.elapsed_time(e3), (time.perf_counter() - t0) * 1e3)

    # ---- extra legs of the BASELINE metric: predictive points/s on this site, sites/s on an NWQN-style batch
    extra_local = [0.0, 0.0, 0.0, 0.0]
    site_info = None
    if not args.no_extra:
        m_pred = args.predict_m
        grid = synthetic.daily_grid(X, m_pred)
        _, info_f = eng.factorize(base)
        if info_f != 0:
            raise RuntimeError(f"factorize failed: info={info_f}")
        grid_d = torch.from_numpy(grid).to(f"cuda:{local}")
        eng.predict(grid_d[:4096])
        barrier()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        mu_d, var_d = eng.predict(grid_d)            # inputs and outputs resident in HBM
        p1.record()
        barrier()
        extra_local[0] = p0.elapsed_time(p1)
        t0 = time.perf_counter()
        mu_h, var_h = eng.predict(grid)              # host grid in, host mean/variance out
        torch.cuda.synchronize()
        extra_local[1] = (time.perf_counter() - t0) * 1e3
        if not (np.all(np.isfinite(mu_h)) and np.all(var_h > -1e-8) and np.allclose(mu_h, mu_d.cpu().numpy())):
            raise RuntimeError("prediction leg produced invalid output")
        eng.close()
        eng = None
        # sites/s: SURVEY 8d config 4 sizes (n_s = 2000 + 6000 u_s, seed 42), `sites_per_gpu` sites on every rank,
        # 100 Adam iterations each + mean/variance on a 10 958-point daily grid; host arrays in, host results out.
        from discontinuum_b200 import multisite

        rng = np.random.default_rng(42)
        ns_all = (2000 + 6000 * rng.uniform(size=128)).astype(int)
        S = args.sites_per_gpu * world
        keys = list(range(min(S, 128)))
        costs = [multisite.site_cost(int(ns_all[k]), args.site_iterations, 10958) for k in keys]
        mine = [keys[i] for i in multisite.assign_sites(costs, world)[rank]]  # the driver's own LPT assignment
        sites = {k: synthetic.loadest_site(int(ns_all[k]), 1000 + k) for k in mine}
        grids = {k: synthetic.daily_grid(sites[k][0], 10958) for k in mine}
        warm_site = {999: synthetic.loadest_site(512, 7)}
        multisite.fit_sites_local(warm_site, iterations=3, device=local, concurrency=1)
        barrier()
        t0 = time.perf_counter()
        res = multisite.fit_sites_local(sites, iterations=args.site_iterations, device=local, concurrency=4, predict=grids)
        summary = {k: {"theta": r["theta"], "objective": r["objective"], "failed": r["failed"]} for k, r in res.items()}
        merged = multisite.gather_results(summary, dist)  # the only collective: final gather on rank 0
        torch.cuda.synchronize()
        extra_local[2] = (time.perf_counter() - t0) * 1e3
        extra_local[3] = float(sum(float(ns_all[k]) ** 3 * args.site_iterations for k in mine))
        bad = [k for k, r in res.items() if r["failed"] is not None or not np.all(np.isfinite(r["mu"]))]
        if bad or (rank == 0 and sorted(merged) != keys):
            raise RuntimeError(f"site fits failed: {bad}")
        site_info = {"n_min": int(min(ns_all[k] for k in keys)), "n_max": int(max(ns_all[k] for k in keys))}

    t = torch.tensor([ms, ms_e2e] + extra_local[:3], dtype=torch.float64, device="cuda")
    fl = torch.tensor([extra_local[3]], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(fl, op=dist.ReduceOp.SUM)
    ms_max, ms_e2e_max = float(t[0]), float(t[1])
    value = world * args.steps / (ms_max * 1e-3)
    e2e_value = world * args.steps / (ms_e2e_max * 1e-3)
    extra = None
    if not args.no_extra:
        m_pred = args.predict_m
        pred_flop = float(m_pred) * float(n) ** 2  # variance via the triangular product, n^2 flop per point
        extra = {
            "predict": {"metric": "predictive_points_per_sec", "value": world * m_pred / (float(t[2]) * 1e-3), "unit": "points/s",
                        "e2e": world * m_pred / (float(t[3]) * 1e-3), "n_train": n, "m_grid_per_gpu": m_pred,
                        "what": "posterior mean + latent variance, grid sharded one slice per GPU",
                        "tflops": pred_flop / (float(t[2]) * 1e-3) / 1e12},
            "sites": {"metric": "sites_per_sec", "value": min(world * args.sites_per_gpu, 128) / (float(t[4]) * 1e-3), "unit": "sites/s",
                      "sites": min(world * args.sites_per_gpu, 128), "iterations": args.site_iterations, "predict_grid": 10958,
                      "fit_tflops": float(fl[0]) / (float(t[4]) * 1e-3) / 1e12, "n_range": site_info,
                      "what": "NWQN-style batch (SURVEY 8d config 4 sizes): fit + daily-grid prediction per site, host arrays in/out, "
                              "sites assigned to ranks by cost (LPT), 4 sites in flight per GPU, one final gather"}}

    if rank == 0:
        peak, peak_src = fp64_peak()
        flop = float(n) ** 3  # n^3/3 POTRF + n^3/3 triangular inverse + n^3/3 LAUUM
        dev_ms = sum(phase)
        achieved = flop / (dev_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": 80.8e9 if n == N_TRAIN else None,
                "traffic_note": "dram__bytes_read+write summed over the 471 launches of one evaluation (ncu, profiles/dram_n16384_r01d_summary.txt)",
                "kernel": "dgp::k_gemm (FP64 DMMA tile engine): every launch of one evaluation",
                "algorithmic_flop_per_eval": flop, "peak_source": "of measured: " + peak_src,
                "phases_ms": {"potrf": phase[0], "trtri": phase[1], "lauum_grad": phase[2], "rest": phase[3]},
                "phases_tflops": {"potrf": flop / 3 / phase[0] / 1e9, "trtri": flop / 3 / phase[1] / 1e9,
                                  "lauum_grad_single_launch": flop / 3 / phase[2] / 1e9}}
        cpu = None
        if not args.no_cpu:
            torch.set_num_threads(os.cpu_count() or 1)
            times, threads = cpu_eval_seconds(CPU_SAMPLE_N, 2, 1)
            per = sum(times) / len(times)
            scale = (n / CPU_SAMPLE_N) ** 3
            cpu = {"value": 1.0 / (per * scale), "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"oracle NLML+grad at n={CPU_SAMPLE_N}: {per:.2f} s/eval on {threads} threads, extrapolated to n={n} by (n/{CPU_SAMPLE_N})^3"}
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": f"loadest-gp single site n={n}, NLML + gradient (10 hyper-parameters), one site per GPU",
                           "l2": "working set 3 x n^2 x 8 B = 6.4 GB per site, far larger than the 126 MB L2 (no flush needed)",
                           "parallelism": f"independent sites x{world}"},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(n * (2 + 2) * 8 + 10 * 8),
                        "d2h_bytes_per_step": int((1 + 10) * 8)},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
                "nlml": val, "extra": extra}
        print(json.dumps(line), flush=True)
    if eng is not None:
        eng.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
