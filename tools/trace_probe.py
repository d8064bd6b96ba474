"""Per-launch timeline of the factorisation (development aid): DGP_TRACE event marks of the last evaluation."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
out = os.path.join(ROOT, "gpurun_out")
import numpy as np

sizes = [int(a) for a in sys.argv[1:]] or [4096, 16384]
for n in sizes:
    path = os.path.join(out, f"trace_{n}.csv")
    if os.path.exists(path):
        os.remove(path)
    os.environ["DGP_TRACE"] = path
    from discontinuum_b200 import capi, models, synthetic
    X, y, noise = synthetic.loadest_site(n, 1000)
    eng = capi.Engine(max_n=n, max_m=256)
    eng.set_train(models.loadest_spec(2).to_c(), X, y, noise)
    th = np.array([0.05, 0.7, 1.0, 1.0, 2.0, 1.3, 0.5, 0.2, 0.3, 0.4])
    for r in range(3):
        val, grad, info = eng.nlml_grad(th)
    eng.close()
    # keep the last evaluation only
    txt = open(path).read().split("# evaluation")[-1]
    open(path, "w").write("# evaluation" + txt)
    print(n, val, info, flush=True)
