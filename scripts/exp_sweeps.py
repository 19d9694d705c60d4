"""Tuning tool: time the Schwarz preconditioner application and one PCG solve of the C3 tangent
system for a given sweep kernel / sub-domain shape.  usage: exp_sweeps.py n_el mode sub [layers [coarse_nc]]"""
import sys, time, json
import numpy as np
import torch

sys.path.insert(0, ".")
from goldfish_b200 import problems
from goldfish_b200.device_model import DeviceModel

n_el = int(sys.argv[1]); mode = sys.argv[2]
sub = [int(x) for x in sys.argv[3].split(",")]; sub = sub[0] if len(sub) == 1 else tuple(sub)
layers = int(sys.argv[4]) if len(sys.argv) > 4 else 2
nc = sys.argv[5] if len(sys.argv) > 5 else "auto"
nc = nc if nc == "auto" else int(nc)
t0 = time.time()
dm = DeviceModel(problems.cylinder(n_el=n_el), schwarz_sub=sub, schwarz_layers=layers, coarse_nc=nc)
dm.set_u(np.zeros(dm.sym.N))
dm.assemble(residual=True, tangent=True)
dm.factor_preconditioner()
dm.set_sweep_mode(mode)
torch.cuda.synchronize()
t_setup = time.time() - t0
A = dm._sw[3]
b = -dm.R.clone()
z = torch.empty_like(b)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for _ in range(3):
    dm.precond_apply(b, z)
ev[0].record()
for _ in range(20):
    dm.precond_apply(b, z)
ev[1].record(); torch.cuda.synchronize()
pc_ms = ev[0].elapsed_time(ev[1]) / 20
coarse_only_ms = None
if mode == "single" and dm.coarse_nc > 0:
    dm._schwarz().debug_flags = 4 | 32          # timing only: coarse sweeps without the fine ones
    dm.precond_apply(b, z)
    ev[0].record()
    for _ in range(20):
        dm.precond_apply(b, z)
    ev[1].record(); torch.cuda.synchronize()
    coarse_only_ms = ev[0].elapsed_time(ev[1]) / 20
    dm.set_sweep_mode(mode)
ev[0].record()
dm.factor_preconditioner()
ev[1].record(); torch.cuda.synchronize()
fac_ms = ev[0].elapsed_time(ev[1])
dm.solve(b)
ev[0].record()
x = dm.solve(b)
ev[1].record(); torch.cuda.synchronize()
print(json.dumps({"n_el": n_el, "N": dm.sym.N, "mode": mode, "sub": sub, "layers": layers, "coarse_nc": dm.coarse_nc, "blocks": A["nblocks"],
                  "band32_GB": A["band_len"] * 4 / 1e9, "max_n_pad": A["max_n_pad"], "max_mb": A["max_mb"],
                  "precond_ms": pc_ms, "precond_coarse_only_ms": coarse_only_ms, "factor_ms": fac_ms, "pcg_its": dm.last_krylov_its,
                  "solve_ms": ev[0].elapsed_time(ev[1]), "relres": dm.last_relres, "setup_s": t_setup,
                  "x_norm": float(x.norm())}), flush=True)
