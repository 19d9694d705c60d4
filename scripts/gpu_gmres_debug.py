import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from goldfish_b200 import problems
from goldfish_b200.device_model import DeviceModel

def true_rel(dm, b, x):
    r = b.clone(); dm.spmv(dm.K, x, r, alpha=-1.0, beta=1.0)
    return float(torch.linalg.vector_norm(r) / torch.linalg.vector_norm(b))

for name, pr in (("tbeam8", problems.tbeam(num_el=8, body_force=(0, 0, 1.0))), ("cyl16", problems.cylinder(n_el=16)), ("cyl12", problems.cylinder(n_el=12))):
    for precond in ("schwarz", "jacobi"):
        dm = DeviceModel(pr, precond=precond)
        dm.set_u(np.zeros(dm.sym.N)); dm.assemble(residual=True, tangent=True)
        b = -dm.R.clone()
        xc = dm.solve(b).clone()
        print(name, precond, "N", dm.sym.N, "coarse_nc", dm.coarse_nc, "cg its", dm.last_krylov_its, "true", true_rel(dm, b, xc))
        for restart in (120, 60, 10):
            dm._gm = None
            x = torch.zeros_like(b)
            try:
                its, rel = dm._gmres(b, x, 1e-10, 3000 if precond == "schwarz" else 600, restart=restart)
                msg = ""
            except Exception as e:
                its, rel, msg = -1, -1, str(e)[:80]
            print("   gmres restart", restart, "its", its, "rel", rel, "true", true_rel(dm, b, x), msg)
