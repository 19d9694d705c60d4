"""Short, profiler-friendly pass over the hot-path kernels (used under ncu):
one R+K assembly, one linearize (dR/dCP, dR/dt), SpMV, one preconditioner
factorisation and 20 PCG iterations on the cylinder workload."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from goldfish_b200.device_model import DeviceModel

n_el = int(sys.argv[1]) if len(sys.argv) > 1 else 48
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
pr, kw = bench.workload(n_el)
dm = DeviceModel(pr, **kw)
for rep in range(reps):
    dm.touch()
    dm.assemble(residual=True, tangent=True, functionals=True)
    dm.assemble(shape=True, thickness=True)
    x = torch.ones(dm.sym.N, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
    for _ in range(3):
        dm.spmv(dm.K, x, y)
    rhs = dm.R.clone()
    try:
        dm.solve(rhs, refactor=True, max_it=20)
    except Exception as e:
        print("pcg stopped:", type(e).__name__)
torch.cuda.synchronize()
print("done N=%d elements=%d nnz=%d" % (dm.sym.N, dm.sym.num_elements, dm.K.nnz))
