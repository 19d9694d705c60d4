"""Tuning evidence for DeviceModel.solve's residual-replacement passes at the bench size: iterations, TRUE residual
and error against a much tighter solve, for several (pass_rtol, true_rtol) settings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from goldfish_b200.device_model import DeviceModel
n_el = int(sys.argv[1]) if len(sys.argv) > 1 else 201
pr, kw = bench.workload(n_el)
dm = DeviceModel(pr, **kw)
cp, th = bench.design_state(dm.sym)
dm.cp.copy_(torch.from_numpy(cp)); dm.set_theta(th)
dm.u.zero_(); dm.touch(); dm.assemble(residual=True, tangent=True)
b = -dm.R.clone()
dm.factor_preconditioner()
dm.true_rtol, dm.pass_rtol, dm.max_refine = 1e-14, 1e-6, 8
xr = dm.solve(b).clone()
print("reference solve: its", dm.last_krylov_its, "true", dm.last_true_relres)
for pass_rtol, true_rtol in ((1e-11, None), (1e-11, 1e-8), (1e-6, 1e-8), (1e-6, 1e-9), (1e-6, 1e-10), (1e-6, 1e-11), (1e-4, 1e-10), (1e-8, 1e-10)):
    dm.pass_rtol, dm.true_rtol, dm.max_refine = pass_rtol, true_rtol, 4
    if true_rtol is None:
        dm.krylov_rtol = pass_rtol
    torch.cuda.synchronize(); t = time.perf_counter()
    x = dm.solve(b)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    r = b.clone(); dm.spmv(dm.K, x, r, alpha=-1.0, beta=1.0)
    tr = float(torch.linalg.vector_norm(r) / torch.linalg.vector_norm(b))
    err = float(torch.linalg.vector_norm(x - xr) / torch.linalg.vector_norm(xr))
    print("pass_rtol %g true_rtol %s: its %d  %.1f ms  true relres %.2e  error vs tight %.2e" % (pass_rtol, true_rtol, dm.last_krylov_its, dt * 1e3, tr, err))
