"""Ad-hoc GPU parity run (development aid; the asserted versions live in tests/)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from goldfish_b200 import problems
from goldfish_b200.device_model import DeviceModel
from oracle.model import OracleModel

def rel(a, b):
    return float(abs(a - b).max() / max(abs(b).max(), 1e-300))

def check(name, pr, opt_field=(), surf=()):
    print("=====", name, flush=True)
    dm = DeviceModel(pr, opt_field=opt_field, shopt_surf_inds=surf)
    om = OracleModel(pr)
    rng = np.random.default_rng(0)
    u = 1e-2 * rng.standard_normal(om.N); u[om.bc_global] = 0
    om.set_u(u); dm.set_u(u)
    dm.assemble(residual=True, tangent=True, functionals=True, shape=bool(opt_field), thickness=True)
    torch.cuda.synchronize()
    R = dm.R.cpu().numpy(); Ro = om.residual()
    print("R   ", rel(R, Ro))
    Kd = dm.K.to_scipy(); Ko = om.stiffness()
    print("K   ", float(abs(Kd - Ko).max() / abs(Ko).max()), "pattern", np.array_equal(Kd.indices, Ko.indices))
    print("W,V ", dm.wv_sum.cpu().numpy(), om.energy(), om.volume())
    Td = dm.T.to_scipy(); To = om.dRdt()
    print("T   ", float(abs(Td - To).max() / abs(To).max()))
    print("dWdu", rel(dm.dWdu.cpu().numpy(), om.dWdu(apply_bcs=False)))
    print("dWdt", rel(dm.dWdt.cpu().numpy()[:om.n_th], om.dWdt()), "dVdt", rel(dm.dVdt.cpu().numpy()[:om.n_th], om.dVdt()))
    for i, f in enumerate(opt_field):
        A = dm.P[i].to_scipy()
        if dm.penP[i] is not None:
            A = A + dm.penP[i][0].to_scipy()
        Ao = om.dRdCP(f, surf[i])
        print("P%d  " % f, float(abs(A - Ao).max() / abs(Ao).max()),
              "dWdP", rel(dm.dWdP[i].cpu().numpy(), om.dWdCP(f, surf[i])),
              "dVdP", rel(dm.dVdP[i].cpu().numpy(), om.dVdCP(f, surf[i])))
    # spmv
    x = rng.standard_normal(om.N); xd = torch.from_numpy(x).cuda(); yd = torch.empty_like(xd)
    dm.spmv(dm.K, xd, yd); print("spmv", rel(yd.cpu().numpy(), Ko @ x))
    xt = rng.standard_normal(om.N); xtd = torch.from_numpy(xt).cuda(); ytd = torch.empty(om.n_th, dtype=torch.float64, device="cuda")
    dm.spmv(dm.T, xtd, ytd, transpose=True); print("spmvT", rel(ytd.cpu().numpy(), To.T @ xt))
    # newton
    t0 = time.time(); ud = dm.newton(verbose=True).cpu().numpy(); t1 = time.time()
    uo = om.solve_nonlinear()
    print("newton u relerr", np.linalg.norm(ud - uo) / np.linalg.norm(uo), "time", t1 - t0, "kits", dm.newton_krylov_its, "hist", dm.newton_history)

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    check("tbeam4", problems.tbeam(num_el=4, body_force=(0, 0, 1.), thickness_kind="iga"), [0, 1, 2], [[0, 1]] * 3)
    check("slr4", problems.scordelis_lo(num_el=4), [1], [[0, 3, 4]])
    check("plate", problems.plate(os.path.join(os.path.dirname(__file__), "..", "tests/golden/plate_c1_input.npz")))
