"""Development aid: CUDA-event timing of the pieces of one PCG iteration at bench size."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from goldfish_b200 import _capi as capi
from goldfish_b200.device_model import DeviceModel, _ptr

n_el = int(sys.argv[1]) if len(sys.argv) > 1 else 201
pr, kw = bench.workload(n_el)
dm = DeviceModel(pr, coarse_nc=(int(os.environ["NC"]) if "NC" in os.environ else "auto"), schwarz_sub=int(os.environ.get("SUB", "64")), schwarz_layers=int(os.environ.get("LAYERS", "2")), **kw)
dm.u.zero_(); dm.assemble(residual=True, tangent=True)
dm.factor_preconditioner()
pc = dm._precond_struct(); st = dm._stream(); lib = dm.lib
N = dm.sym.N
r = dm.R.clone(); z = torch.zeros_like(r)
def T(fn, reps=10):
    for _ in range(3): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
A = dm._sw[3]; cm = dm._coarse[0]; Ac = cm._sw[3]
print("fine blocks", A["nblocks"], "max_nbr", A["max_nbr"], "max_mb", A["max_mb"], "| coarse nbr", Ac["max_nbr"], "mb", Ac["max_mb"], "Nc", cm.sym.N)
print("precond_apply total   %.3f ms" % T(lambda: capi.check(lib.gf_precond_apply(C.byref(pc), _ptr(r), _ptr(z), N, st), "pa")))
print("fine only (apply)     %.3f ms" % T(lambda: capi.check(lib.gf_schwarz_apply(pc.fine, _ptr(r), _ptr(z), N, st), "f")))
rc, zc = dm._coarse[3], dm._coarse[4]
print("coarse only (apply)   %.3f ms" % T(lambda: capi.check(lib.gf_schwarz_apply(pc.coarse, _ptr(rc), _ptr(zc), cm.sym.N, st), "c")))
print("Rt spmv               %.3f ms" % T(lambda: capi.check(lib.gf_spmv(C.byref(pc.Rt), _ptr(r), _ptr(rc), 1.0, 0.0, st), "rt")))
print("P spmv                %.3f ms" % T(lambda: capi.check(lib.gf_spmv(C.byref(pc.P), _ptr(zc), _ptr(z), 1.0, 1.0, st), "p")))
x = torch.randn(N, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
print("K spmv                %.3f ms" % T(lambda: dm.spmv(dm.K, x, y)))
print("axpby                 %.3f ms" % T(lambda: dm.axpby(1.0, x, 1.0, y)))
its0 = 0
import time
torch.cuda.synchronize(); t0 = time.time(); dm.solve(r.clone()); torch.cuda.synchronize(); t1 = time.time()
tf0 = time.time(); dm.factor_preconditioner(); torch.cuda.synchronize(); print("refactor %.1f ms" % ((time.time() - tf0) * 1e3))
print("full solve %.1f ms, its %d -> %.3f ms/it" % ((t1 - t0) * 1e3, dm.last_krylov_its, (t1 - t0) * 1e3 / dm.last_krylov_its))
