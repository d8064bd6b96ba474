"""Measure the FP64 GEMM roofline denominator (cuBLAS DGEMM through torch.matmul).

MEASURED_PEAKS.json (driver-written) has no FP64 entry; this reproduces its method
(best of 10 for the burst figure, back-to-back for 4 s for the sustained one) in float64
and writes gpurun_out/fp64_peak.json.  Run on the GPU box only.
"""
import json, sys, time
import torch

def main(n=8192):
    dev = torch.device("cuda:0")
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    c = torch.empty_like(a)
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    burst = 2.0 * n**3 / (best * 1e-3) / 1e12
    # sustained
    t0 = time.time(); k = 0
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < 4.0:
        for _ in range(5):
            torch.matmul(a, b, out=c); k += 1
        torch.cuda.synchronize()
    e1.record(); e1.synchronize()
    sustained = 2.0 * n**3 * k / (e0.elapsed_time(e1) * 1e-3) / 1e12
    # syrk-like and potrf yardsticks (stock library on the same GPU, context only)
    res = {"fp64_tflops": burst, "fp64_tflops_sustained": sustained, "n": n,
           "how": "torch.matmul float64 %d^3 (2*N^3): best of 10 (burst) and back to back for 4 s (sustained)" % n,
           "gpu_name": torch.cuda.get_device_name(0)}
    for m in (4096, 16384):
        x = torch.randn(m, m, dtype=torch.float64, device=dev)
        k_ = x @ x.T + m * torch.eye(m, dtype=torch.float64, device=dev)
        del x
        torch.linalg.cholesky(k_); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); l = torch.linalg.cholesky(k_); e1.record(); e1.synchronize()
        ms = e0.elapsed_time(e1)
        res["cusolver_potrf_ms_n%d" % m] = ms
        res["cusolver_potrf_tflops_n%d" % m] = m**3 / 3 / (ms * 1e-3) / 1e12
        e0.record(); ki = torch.cholesky_inverse(l); e1.record(); e1.synchronize()
        res["cusolver_potri_ms_n%d" % m] = e0.elapsed_time(e1)
        del k_, l, ki
    print(json.dumps(res))
    with open("gpurun_out/fp64_peak.json", "w") as f:
        json.dump(res, f, indent=1)

if __name__ == "__main__":
    main()
