"""Development aid: Schwarz-PCG vs Jacobi-PCG vs oracle LU, timings at growing size."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from goldfish_b200 import problems
from goldfish_b200.device_model import DeviceModel

def run(name, pr, oracle=True, **kw):
    dm = DeviceModel(pr, precond="schwarz", **kw)
    dm.u.zero_(); dm.assemble(residual=True, tangent=True)
    rhs = -dm.R.clone()
    torch.cuda.synchronize(); t0 = time.time()
    dm.factor_preconditioner(); torch.cuda.synchronize(); t1 = time.time()
    x = dm.solve(rhs); torch.cuda.synchronize(); t2 = time.time()
    A = dm._sw[3]
    msg = "%s nc=%d N=%d blocks=%d max_nbr=%d max_mb=%d band=%.2f GB | factor %.3fs solve %.3fs its=%d relres=%.2e" % (
        name, dm.coarse_nc, dm.sym.N, A["nblocks"], A["max_nbr"], A["max_mb"], A["band_len"] * 8 / 1e9, t1 - t0, t2 - t1, dm.last_krylov_its, dm.last_relres)
    if oracle:
        from oracle.model import OracleModel
        om = OracleModel(pr); K = om.stiffness(); xo = om.solve(K, -om.residual())
        msg += " | err vs LU %.2e" % (np.linalg.norm(x.cpu().numpy() - xo) / np.linalg.norm(xo))
    # true residual
    y = torch.empty_like(rhs); dm.spmv(dm.K, x, y)
    msg += " | true relres %.2e" % float(torch.linalg.vector_norm(y - rhs) / torch.linalg.vector_norm(rhs))
    print(msg, flush=True)
    return dm

if __name__ == "__main__":
    sizes = [int(a) for a in sys.argv[1:]] or [16, 32]
    run("tbeam10", problems.tbeam(num_el=10, body_force=(0, 0, 1.)))
    run("slr4", problems.scordelis_lo(num_el=4))
    run("plate", problems.plate(os.path.join(os.path.dirname(__file__), "..", "tests/golden/plate_c1_input.npz")))
    for ne in sizes:
        for sub, nc in ((64, 16), (64, 24)):
            t0 = time.time(); pr = problems.cylinder(n_el=ne)
            dm = run("cyl%d sub%d" % (ne, sub), pr, oracle=(ne <= 8), coarse_nc=nc, schwarz_sub=sub)
            torch.cuda.synchronize(); t1 = time.time(); dm.factor_preconditioner(); torch.cuda.synchronize()
            print("   refactor %.3fs" % (time.time() - t1))
            print("   setup+run wall", time.time() - t0, flush=True)
            del dm; torch.cuda.empty_cache()
