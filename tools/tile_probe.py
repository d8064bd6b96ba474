"""Tile-engine micro-benchmark: time per wave of 128x64 tiles against K, with (mode -1) and without (mode 0) the C tile load.
The intercept of time-per-wave over K is the per-tile overhead the two co-resident CTAs fail to hide."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from discontinuum_b200 import capi
eng = capi.Engine(max_n=256, max_m=128)
dev = torch.device("cuda:0")
M, N = 16384, 8192
tiles = (M // 128) * (N // 64)
C = torch.zeros(M, N, dtype=torch.float64, device=dev)
for K in (128, 256, 512, 1024, 2048):
    A = torch.randn(M, K, dtype=torch.float64, device=dev) * 0.01
    B = torch.randn(N, K, dtype=torch.float64, device=dev) * 0.01
    for mode in (0, -1):
        ts = []
        for r in range(4):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            eng.gemm_nt(A, B, C, mode)
            ts.append(time.perf_counter() - t0)
        t = min(ts[1:])
        waves = tiles / 296.0
        print(f"K={K:5d} mode={mode:2d} {t*1e3:8.3f} ms  {2.0*M*N*K/t/1e12:6.2f} TF  per wave {t/waves*1e6:7.2f} us", flush=True)
