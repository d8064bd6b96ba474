import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from discontinuum_b200 import capi
eng = capi.Engine(max_n=256, max_m=128)
torch.manual_seed(0)
dev = torch.device("cuda:0")
for (M, N, K) in [(512, 512, 128), (2048, 2048, 128), (4096, 4096, 128), (4096, 4096, 64), (4096, 4096, 256), (8192, 4096, 128), (4096, 4096, 1024)]:
    A = torch.randn(M, K, dtype=torch.float64, device=dev)
    B = torch.randn(N, K, dtype=torch.float64, device=dev)
    ref = A @ B.T
    res = []
    for mode in (0, -1, 0, -1, 0):
        C0 = torch.randn(M, N, dtype=torch.float64, device=dev)
        Cm = C0.clone()
        torch.cuda.synchronize()
        eng.gemm_nt(A, B, Cm, mode)
        torch.cuda.synchronize()
        want = ref if mode == 0 else C0 - ref
        err = (Cm - want).abs()
        bad = (err > 1e-9).nonzero()
        res.append((mode, float(err.max()), int(bad.shape[0])))
        if bad.shape[0] and len(res) < 3:
            rows = torch.unique(bad[:, 0] // 128); cols = torch.unique(bad[:, 1] // 64)
            print("   bad tiles rows", rows.tolist()[:20], "cols", cols.tolist()[:20], "n bad tiles approx", bad.shape[0] / (128 * 64))
    print(M, N, K, res, flush=True)
