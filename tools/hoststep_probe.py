import time, numpy as np, torch, sys, os
sys.path.insert(0,'/root/repo')
from discontinuum_b200.spec import GPModule
from discontinuum_b200.models import loadest_spec
def run():
    m=GPModule(loadest_spec(2))
    opt=torch.optim.Adam(m.raw_list(), lr=0.05, weight_decay=1e-4)
    sch=torch.optim.lr_scheduler.ReduceLROnPlateau(opt, mode="min", factor=0.7, patience=30, threshold=1e-4, threshold_mode="rel", min_lr=1e-6, cooldown=10)
    g=torch.from_numpy(np.random.randn(10))
    N=300
    t0=time.perf_counter()
    for i in range(N):
        opt.zero_grad(set_to_none=True)
        nat=m.natural()
        th=nat.detach().numpy().astype(np.float64)
        nll=1.0+((nat-nat.detach())*g).sum()
        obj=(nll-m.log_prior(nat))/5000
        obj.backward()
        params=m.raw_list()
        torch.nn.utils.clip_grad_norm_(params, max_norm=1.0)
        for p in params:
            if p.grad is not None and torch.isnan(p.grad).any():
                p.grad=torch.nan_to_num(p.grad)
        opt.step(); sch.step(float(obj.detach()))
    return (time.perf_counter()-t0)/N*1e3
print('OMP', os.environ.get('OMP_NUM_THREADS'), 'threads', torch.get_num_threads(), 'host step ms', round(run(),3), round(run(),3))
