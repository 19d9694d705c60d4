"""bench.py -- analysis+adjoint iterations/s on the synthetic non-matching
multi-patch cylinder (BASELINE.json configs[2], SURVEY.md section 8d "C3").

One *step* = one optimizer iteration of the reference stack:
  design vars -> Newton solve of R(u)=0 from u=0 (rtol 1e-3, max 30)
  -> W_int, V -> linearize (dR/du, dR/dCP_f for f=0,1,2, dR/dt)
  -> dW/du, dW/dCP, dW/dt -> adjoint solve K^T lam = dW/du
  -> total gradients dW/dp - (dR/dp)^T lam.

  python bench.py --gpus N --steps K --warmup W [--n-el NE] [--impl reference]

`value`  : steps/s with the design variables already resident in HBM.
`e2e`    : steps/s through the reference-facing facade (NonMatchingOpt +
           operations) with HOST numpy arrays in and out (H2D of the design
           variables and D2H of state, objective and gradients every step).
`roofline`: CSR SpMV (the kernel the Krylov solves spend their time in),
           algorithmic bytes 12 nnz + 24 N + 8 per launch over its CUDA-event time.
`cpu_baseline` / `--impl reference`: the restated reference CPU path (oracle/c: C++/OpenMP assembly incl.
           penalty coupling + multifrontal LU), every number a FULL iteration on the mesh it names.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-el", type=int, default=int(os.environ.get("GF_BENCH_NEL", "201")),
                    help="elements per patch side; 201 = BASELINE configs[2] (8 patches, ~1.03 M DOF)")
    ap.add_argument("--topology", default="4x2", help="patches around x along the cylinder (4x2 = BASELINE configs[2]), or "
                    "'wingbox' = BASELINE configs[3] (40 patches, --dofs sets the size)")
    ap.add_argument("--dofs", type=float, default=1.0e7, help="--topology wingbox: target number of displacement dofs")
    ap.add_argument("--cpu-n-el", type=int, default=64, help="mesh of the bounded cpu_baseline sample in our arm's line")
    ap.add_argument("--no-trend", action="store_true", help="--impl reference: skip the three smaller meshes")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the (untimed) parity checks at the benched size")
    ap.add_argument("--parity-sharded", action="store_true", help="N > 1: run the full parity checks (FD, tight Newton, sharded vs single) too")
    return ap.parse_args()


def topo(args):
    if args.topology.lower() == "wingbox":
        return ("wingbox", float(args.dofs))
    a, b = args.topology.lower().split("x")
    return int(a), int(b)


def workload_name(n_el, n_circ=4, n_axial=2):
    if n_circ == "wingbox":
        return ("wingbox_40p_%.3gM: synthetic wing box (BASELINE configs[3]): 2 skins x 10 segments + 3 spars + 17 ribs = 40 non-matching "
                "bicubic patches, 163 intersections (T- and X-junctions interior to the patches), shape fields 0,1,2 + per-patch thickness"
                % (n_axial / 1e6))
    if (n_circ, n_axial) == (4, 2):
        return ("cylinder_4x2_ne%d: synthetic 8-patch non-matching bicubic NURBS cylinder (BASELINE configs[2]), 12 intersections, "
                "shape fields 0,1,2 + per-patch thickness" % n_el)
    n_itf = n_circ * n_axial + n_circ * (n_axial - 1)
    return ("cylinder_%dx%d_ne%d: synthetic %d-patch non-matching bicubic NURBS cylinder (BASELINE configs[3]-like, coupling-heavy), "
            "%d intersections, shape fields 0,1,2 + per-patch thickness" % (n_circ, n_axial, n_el, n_circ * n_axial, n_itf))


def workload(n_el, n_circ=4, n_axial=2):
    """BASELINE configs[2] (default 4 x 2 patches); --topology 8x5 gives the 40-patch, 72-intersection stand-in for
    configs[3] (n_el 286 there is ~10 M DOF)."""
    from goldfish_b200 import problems
    if n_circ == "wingbox":
        pr = problems.wingbox(target_dofs=n_axial)
        return pr, dict(opt_field=[0, 1, 2], shopt_surf_inds=[list(range(len(pr["patches"])))] * 3)
    pr = problems.cylinder(n_el=n_el, n_circ=n_circ, n_axial=n_axial, R=1.0, L=2.0 * n_axial, E=68e9, nu=0.35, h_th=1e-2,
                           pressure_like_load=(0.0, 0.0, -1.0e3), quad_deg_const=3, thickness_kind="const")
    kw = dict(opt_field=[0, 1, 2], shopt_surf_inds=[list(range(n_circ * n_axial))] * 3)
    return pr, kw


def problems_dofs(pr):
    from goldfish_b200 import problems
    return problems.num_dofs(pr)


def design_state(S):
    """CP perturbation U(-1e-3,1e-3) h_e (seed 0), thickness t(1+0.1U) (seed 1)."""
    rng = np.random.default_rng(0)
    cp = S.cp0.copy()
    for P in S.patches:
        he = 1.0 / max(P.neu, P.nev)
        cp[P.cp_off:P.cp_off + P.ncp, :3] += rng.uniform(-1e-3, 1e-3, (P.ncp, 3)) * he
    th = S.theta0 * (1 + 0.1 * np.random.default_rng(1).uniform(-1, 1, S.n_th))
    return cp, th


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.samples, self.stop, self.index = [], False, index
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.samples.append([x.strip() for x in o.strip().split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def __enter__(self):
        self.t.start(); return self

    def __exit__(self, *a):
        self.stop = True; self.t.join(timeout=6)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if len(s) >= 6 and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) >= 6 and s[1].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            if len(s) >= 6:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


class Watchdog:
    """Guard for the optional, untimed sections that follow the measurement (parity checks, sharded-vs-single run,
    extra configurations): if a section does not return within its allowance -- e.g. a collective some rank never
    enters -- the bench line measured so far is printed with a note and every rank leaves with exit code 0, instead
    of the whole run being killed at the driver's limit with nothing printed."""

    def __init__(self, rank):
        self.rank, self.line, self.timer = rank, None, None

    def _fire(self, label, seconds):
        try:
            if self.rank == 0 and self.line is not None:
                self.line.setdefault("notes", []).append("section '%s' did not finish within %d s; line printed by the watchdog" % (label, seconds))
                sys.stdout.write(json.dumps(self.line) + "\n"); sys.stdout.flush()
        finally:
            os._exit(0)

    def start(self, seconds, label):
        self.cancel()
        self.timer = threading.Timer(seconds, self._fire, args=(label, seconds))
        self.timer.daemon = True
        self.timer.start()

    def cancel(self):
        if self.timer is not None:
            self.timer.cancel(); self.timer = None


# ----------------------------------------------------------------------------- our arm
SPMV_NODE_TRAFFIC_C3 = 1481301648   # dram read 1469338000 + write 11963648 B of one k_spmv_node launch at n_el = 201 (profiles/r2_ncu_full_k_spmv_node.csv)
SWEEP_TRAFFIC_C3 = 7155089760      # dram read 7143036000 + write 12053760 B of one k_sw_solve1 launch at n_el = 201 (profiles/r2_ncu_full_k_sw_solve1.csv)


class Step:
    """One analysis+adjoint iteration on the device model (inputs resident in HBM)."""

    def __init__(self, dm):
        import torch
        self.dm, self.torch = dm, torch
        S = dm.sym
        self.bc_idx = torch.from_numpy(S.bc_list.astype(np.int64)).to(dm.device)
        self.lam = torch.zeros(S.N, dtype=torch.float64, device=dm.device)
        self.rhs = torch.zeros(S.N, dtype=torch.float64, device=dm.device)
        self.gP = [torch.zeros(n, dtype=torch.float64, device=dm.device) for n in S.P_ncols]
        self.gT = torch.zeros(S.n_th, dtype=torch.float64, device=dm.device)
        self.info = {}
        self.newton_rtol = 1e-3          # the reference's default (operations/disp_imop.py:38)

    def __call__(self, timers=None):
        import ctypes as C
        from goldfish_b200 import _capi as capi
        dm, S = self.dm, self.dm.sym
        torch = self.torch

        def mark(name):
            if timers is not None:
                e = torch.cuda.Event(enable_timing=True); e.record(); timers.append((name, e))
        mark("start")
        dm.newton(max_it=30, rtol=self.newton_rtol, accept_stagnation=self.newton_rtol < 1e-6)
        mark("newton (assemble R,K + factor + PCG per iteration)")
        kits = list(dm.newton_krylov_its)
        trel = list(dm.newton_true_relres)
        dm.ensure(tangent=True, functionals=True, shape=True, thickness=True)   # K of the last Newton iterate is reused
        mark("linearize (K, W, V, dR/dCP x3, dR/dt, dW/d*)")
        self.rhs.copy_(dm.dWdu)
        capi.check(dm.lib.gf_mask_vec(C.byref(dm.model), C.c_void_p(self.rhs.data_ptr()), dm._stream()), "mask")
        dm.solve(self.rhs, self.lam)
        mark("adjoint solve")
        kits.append(dm.last_krylov_its); trel.append(dm.last_true_relres)
        for i in range(len(S.opt_field)):
            self.gP[i].copy_(dm.dWdP[i][:S.P_ncols[i]])
            dm.spmv_global(dm.P[i], self.lam, self.gP[i], alpha=-1.0, beta=1.0, transpose=True)
            if dm.penP[i] is not None:
                dm.spmv_global(dm.penP[i][0], self.lam, self.gP[i], alpha=-1.0, beta=1.0, transpose=True)
        self.gT.copy_(dm.dWdt[:S.n_th])
        dm.spmv_global(dm.T, self.lam, self.gT, alpha=-1.0, beta=1.0, transpose=True)
        mark("gradient products (dR/dp)^T lam")
        self.info = {"newton_its": len(dm.newton_history) - 1, "krylov_its": kits, "true_relres": trel,
                     "gmres_fallback_used": bool(dm.fallback_used)}


def e2e_step(nm, ops, cp_host, th_host):
    """Same iteration through the reference-facing facade with host arrays."""
    disp, wint, vol = ops
    for f in nm.opt_field:
        nm.update_CPIGA(cp_host[f], f)
    nm.update_h_th(th_host)
    u = disp.solve_nonlinear(max_it=30, rtol=1e-3)
    nm.update_uIGA(u)
    W, V = wint.Wint(), vol.volume()
    disp.linearize()
    dWdu = wint.dWintduIGA()
    lam = np.zeros_like(u)
    disp.solve_linear_rev(dWdu, lam)
    d_in = [np.zeros(nm.vec_scalar_iga_dof) for _ in nm.opt_field] + [np.zeros(nm.h_th_dof)]
    disp.apply_linear_rev(d_in, None, lam)
    grads = [wint.dWintdCPIGA(f) - d_in[i] for i, f in enumerate(nm.opt_field)] + [wint.dWintdh_th() - d_in[-1]]
    return W, V, grads, u


def build_facade(pr, kw, symbolic=None, device_model=None):
    from goldfish_b200.nonmatching_opt import NonMatchingOpt, SplinePatch, Thickness, ShellLoad
    from goldfish_b200.operations import DispImOpeartion, IntEnergyExOperation, VolumeExOperation
    splines = [SplinePatch(P["knots"], P["p"], P["cp"], P["quad_deg"], P["bc_dofs"]) for P in pr["patches"]]
    nm = NonMatchingOpt(splines, pr["E"], [Thickness(P["thickness"]["kind"], P["thickness"]["values"]) for P in pr["patches"]], pr["nu"])
    nm.set_shopt_surf_inds(kw["opt_field"], kw["shopt_surf_inds"])
    nm.set_thickness_opt(var_thickness=False)
    nm.create_mortar_meshes([len(it["xi"][0]) - 1 for it in pr["interfaces"]])
    nm.mortar_meshes_setup([it["patches"] for it in pr["interfaces"]], [it["xi"] for it in pr["interfaces"]],
                           pr["penalty_coefficient"], 1)
    nm.set_residuals([ShellLoad(body_force=P["body_force"]) for P in pr["patches"]])
    nm._symbolic = symbolic          # same topology as the device-resident arm: reuse its symbolic phase
    nm._device_model = device_model  # ... and (large runs) its whole device model
    return nm, (DispImOpeartion(nm), IntEnergyExOperation(nm), VolumeExOperation(nm))


NEWTON_TIGHT = 1e-8     # |R| / |R0| of the parity runs (the solves hold a true relative residual of 1e-10)


def parity_checks(dm, step, torch, world):
    """Correctness evidence AT THE BENCHED SIZE (not timed): true residuals of the state and adjoint solves,
    symmetry of the assembled tangent, and central finite differences of W_int (each side a full Newton solve)
    against the adjoint total gradient -- the reference's own check (check_totals in
    demos_csdl_alpha/thickness_opt/plate_const_th_opt_wint.py:220-223) -- for one patch thickness and one
    shape direction.  Newton is converged to NEWTON_TIGHT here so that FD and adjoint differentiate the same state."""
    import ctypes as C
    from goldfish_b200 import _capi as capi
    S = dm.sym
    out = {}
    step.newton_rtol = NEWTON_TIGHT
    step()
    out["newton_history_tight"] = [float(h) for h in dm.newton_history]
    out["newton_stopped_at_fp64_floor"] = bool(dm.newton_stagnated)
    tr = step.info["true_relres"]
    # the first two Newton solves are the ones the timed (rtol 1e-3) iteration does; later ones act on a residual that
    # already sits at its FP64 floor
    out["true_relres_state"] = max(tr[:min(2, len(tr) - 1)]) if len(tr) > 1 else None
    out["true_relres_state_all_tight_solves"] = tr[:-1]
    out["true_relres_adjoint"] = tr[-1]
    out["recurrence_rtol"] = dm.krylov_rtol
    W0 = float(dm.wv_sum[0].item())
    gT = step.gT.clone(); gP = [g.clone() for g in step.gP]
    # symmetry of K without forming K^T: |x.(K y) - y.(K x)| / (|x| |K y|) for two random vectors
    g = torch.Generator(device="cpu"); g.manual_seed(3)
    x = torch.randn(S.N, dtype=torch.float64, generator=g).to(dm.device)
    y = torch.randn(S.N, dtype=torch.float64, generator=g).to(dm.device)
    Kx = torch.zeros_like(x); Ky = torch.zeros_like(x)
    dm.spmv_global(dm.K, x, Kx); dm.spmv_global(dm.K, y, Ky)
    out["K_asymmetry"] = abs(dm.dot(x, Ky) - dm.dot(y, Kx)) / (dm.dot(x, x) ** 0.5 * dm.dot(Ky, Ky) ** 0.5)
    del x, y, Kx, Ky

    def W_at(theta=None, cp=None):
        th0, cp0 = dm.theta.clone(), dm.cp.clone()
        if theta is not None:
            dm.theta.copy_(theta)
        if cp is not None:
            dm.cp.copy_(cp)
        dm.touch()
        dm.newton(max_it=30, rtol=NEWTON_TIGHT, accept_stagnation=True)
        W = float(dm.wv_sum[0].item())
        dm.theta.copy_(th0); dm.cp.copy_(cp0); dm.touch()
        return W
    # (1) thickness of one patch (const thickness: one design variable per patch)
    ip = min(3, S.n_th - 1)
    h = 1e-3 * float(dm.theta[ip].item())
    tp, tm = dm.theta.clone(), dm.theta.clone()
    tp[ip] += h; tm[ip] -= h
    fd_t = (W_at(theta=tp) - W_at(theta=tm)) / (2 * h)
    out["fd_dWdt"] = {"dof": int(ip), "adjoint": float(gT[ip].item()), "central_fd": fd_t,
                      "relerr": abs(fd_t - float(gT[ip].item())) / abs(float(gT[ip].item()))}
    # (2) a smooth shape direction on patch 0, field 2 (z): bump sin(pi a/n_u) sin(pi b/n_v)
    if S.opt_field:
        f = S.opt_field[-1]; fi = S.opt_field.index(f)
        P = S.patches[S.shopt_surf_inds[fi][0]]
        I = np.tile(np.arange(P.n_u), P.n_v); J = np.repeat(np.arange(P.n_v), P.n_u)
        d = np.sin(np.pi * I / (P.n_u - 1)) * np.sin(np.pi * J / (P.n_v - 1))
        dvec = np.zeros(S.P_ncols[fi]); dvec[P.pcol_off[f]:P.pcol_off[f] + P.ncp] = d
        dd = torch.from_numpy(dvec).to(dm.device)
        adj = float((gP[fi] * dd).sum().item())
        eps = 1e-4                         # metres; the cylinder has R = 1, t = 1e-2
        dcp = torch.zeros_like(dm.cp).view(-1, 4)      # design variables = homogeneous coordinates cpFuncs[f] (update_CPIGA)
        dcp[P.cp_off:P.cp_off + P.ncp, f] = torch.from_numpy(d).to(dm.device)
        fd_p = (W_at(cp=(dm.cp.view(-1, 4) + eps * dcp).reshape(dm.cp.shape)) - W_at(cp=(dm.cp.view(-1, 4) - eps * dcp).reshape(dm.cp.shape))) / (2 * eps)
        out["fd_dWdCP"] = {"field": int(f), "patch": int(P.index), "direction": "sin x sin bump", "adjoint": adj, "central_fd": fd_p,
                           "relerr": abs(fd_p - adj) / abs(adj)}
    out["fd_grad_relerr"] = max(v["relerr"] for k, v in out.items() if k.startswith("fd_d"))
    out["W_int"] = W0
    step.newton_rtol = 1e-3
    return out


def ranks_vs_single(torch, dist, world, rank, n_el=20):
    """Sharded run against an unsharded run of the same (small) problem on every rank's own GPU:
    max relative difference of u and of the total gradients (printed in the bench line at N > 1)."""
    from goldfish_b200.device_model import DeviceModel
    pr, kw = workload(n_el)
    res = {}
    outs = []
    for distributed in (True, False):
        dm = DeviceModel(pr, distributed=distributed, **kw)
        cp, th = design_state(dm.sym)
        dm.cp.copy_(torch.from_numpy(cp)); dm.set_theta(th)
        st = Step(dm); st.newton_rtol = NEWTON_TIGHT
        st()
        outs.append((dm.u.clone(), [g.clone() for g in st.gP] + [st.gT.clone()], st.info["krylov_its"]))
    (ua, ga, ka), (ub, gb, kb) = outs
    rel = lambda a, b: float(torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b))
    res = {"n_el": n_el, "dofs": int(ua.numel()), "u": rel(ua, ub), "gradients": max(rel(a, b) for a, b in zip(ga, gb)),
           "krylov_its_sharded": ka, "krylov_its_single": kb}
    t = torch.tensor([res["u"], res["gradients"]], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res["u"], res["gradients"] = float(t[0]), float(t[1])
    return res


def small_configs(torch):
    """BASELINE configs[0] (plate thickness opt, C1) and configs[1] (T-beam shape opt, C2): the
    same analysis+adjoint step at the reference's own (tiny, latency-bound) sizes."""
    from goldfish_b200 import problems
    from goldfish_b200.device_model import DeviceModel
    out = {}
    cases = {"C1_plate_thickness_N1449": (problems.plate(os.path.join(ROOT, "tests", "golden", "plate_c1_input.npz")), {}),
             "C2_tbeam_shape_N648": (problems.tbeam(num_el=10), dict(opt_field=[0], shopt_surf_inds=[[0, 1]]))}
    for name, (pr, kw) in cases.items():
        st = Step(DeviceModel(pr, **kw))
        for _ in range(2):
            st()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            st()
        b.record(); torch.cuda.synchronize()
        out[name] = {"ms_per_step": a.elapsed_time(b) / 5, "newton_its": st.info["newton_its"], "krylov_its": st.info["krylov_its"]}
    return out


def time_kernel(torch, fn, reps, flush):
    """Average CUDA-event time of `fn` (ms) with an L2 flush between launches."""
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for _ in range(3):
        fn()
    for a, b in ev:
        flush.add_(1.0)
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return float(np.mean([a.elapsed_time(b) for a, b in ev]))


def cpu_reference_iteration(pr, kw, timings=None):
    """The restated reference CPU path on the host cores (oracle/cpu_port.py): compiled C++/OpenMP shell
    quadrature + penalty coupling + CSR scatter (oracle/c/kl_cpu.cpp), multifrontal sparse LU on a nested
    dissection ordering standing in for MUMPS (oracle/c/mf_lu.cpp, OpenMP + the image's OpenBLAS) for every
    Newton step and -- transposed and re-factorised, as the reference does (utils/opt_utils.py:199-204) -- for
    the adjoint.  Same step as ours: Newton from u=0 to 1e-3, W/V, K, dR/dCP x3, dR/dt, adjoint, total gradients.
    Set-up (symbolic phase, nested dissection) is outside the timed iteration, as it is for the GPU arm."""
    from oracle.cpu_port import CpuModel
    cm = CpuModel(pr, **kw)
    cp, th = design_state(cm.S)
    cm.cp[:] = cp; cm.theta[:] = th
    lu = cm.direct_solver()
    dt, _ = cm.iteration(timings=timings)
    info = {"dofs": int(cm.S.N), "newton_its": int(cm.newton_its), "lu_flops": lu.flops, "lu_bytes": int(lu.lu_bytes),
            "fronts": int(lu.nfronts), "max_front": int(lu.max_front), "elements": int(cm.S.num_elements),
            "nnz_K": int(cm.S.K_indptr[-1]), "nq": int(cm.S.nq), "W_int": float(cm.W)}
    return dt, info


def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count()


def run_reference(args, rank):
    """--impl reference: the restated reference CPU path, timed IN FULL on the configuration it names
    (config.workload), with all host threads.  One iteration at C3 is minutes of CPU work, so the arm runs as
    many full iterations as fit a ~4 minute budget (at least one, at most --steps) and reports that count in
    `steps`; nothing is extrapolated.  A trend over three smaller meshes of the same topology (each run in full)
    is printed beside it."""
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm uses every core the process may run on
    os.environ.setdefault("GFO_THREADS", str(cpu_threads()))
    budget = float(os.environ.get("GF_REF_BUDGET_S", "240"))
    pr, kw = workload(args.n_el, *topo(args))
    times, phases, info = [], [], None
    t_all = time.perf_counter()
    while len(times) < max(1, args.steps):
        tm = {}
        dt, info = cpu_reference_iteration(pr, kw, tm)
        times.append(dt); phases.append(tm)
        if time.perf_counter() - t_all + dt > budget:
            break
    t = float(np.mean(times))
    trend = []
    if not args.no_trend:
        for ne in (24, 48, 96):
            if ne >= args.n_el or args.topology.lower() == "wingbox":
                continue
            prs, kws = workload(ne, *topo(args))
            dts, inf = cpu_reference_iteration(prs, kws)
            trend.append({"n_el": ne, "dofs": inf["dofs"], "s_per_iter": dts, "newton_its": inf["newton_its"]})
        trend.append({"n_el": args.n_el, "dofs": info["dofs"], "s_per_iter": t, "newton_its": info["newton_its"]})
    expo = None
    if len(trend) >= 3:
        x = np.log([r["dofs"] for r in trend]); y = np.log([r["s_per_iter"] for r in trend])
        expo = float(np.polyfit(x, y, 1)[0])
    val = 1.0 / t
    ncores = cpu_threads()
    sample = ("%d full analysis+adjoint iteration(s) of the restated CPU path on %s (N=%d): C++/OpenMP assembly (shells + "
              "penalty) and multifrontal LU (nested dissection, %d fronts, %.2e flops and %.1f GB per factorisation, fresh LU per "
              "Newton step + re-factorised transpose for the adjoint) on %d threads; nothing extrapolated"
              % (len(times), workload_name(args.n_el, *topo(args)).split(":")[0], info["dofs"], info["fronts"], info["lu_flops"],
                 info["lu_bytes"] / 1e9, ncores))
    line = {"impl": "reference", "metric": "analysis+adjoint iters/s", "value": val, "unit": "iters/s", "n_gpus": args.gpus,
            "steps": len(times), "warmup": 0, "steps_requested": args.steps, "warmup_requested": args.warmup,
            "ms_per_step": 1e3 * t, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.n_el, *topo(args)), "dofs": info["dofs"], "elements": info["elements"],
                       "nnz_K": info["nnz_K"], "quad_pts_per_element": info["nq"]},
            "cpu_baseline": {"value": val, "unit": "iters/s", "cores": ncores, "kind": "port", "sample": sample},
            "cpu_phase_s": {k: round(float(np.mean([p[k] for p in phases])), 3) for k in phases[0]},
            "newton_its": info["newton_its"], "W_int": info["W_int"],
            "cpu_trend": trend, "cpu_trend_exponent_in_dofs": expo,
            "e2e": {"value": val, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline_sample(args, torch, DeviceModel):
    """`cpu_baseline` of our arm's line (N = 1)."""
    # bounded CPU sample: ONE FULL iteration of the restated reference CPU path on a smaller mesh of the
    # same topology (value at THAT size, nothing scaled), with this GPU path timed on the same mesh
    # beside it; the full-size CPU number is `bench.py --impl reference`.
    if args.topology.lower() == "wingbox":
        ne, sample_topo = 0, ("wingbox", min(1.2e5, float(args.dofs)))
    else:
        ne, sample_topo = min(args.cpu_n_el, args.n_el), topo(args)
    prs, kws = workload(ne, *sample_topo)
    tm = {}
    dt, inf = cpu_reference_iteration(prs, kws, tm)
    dms = DeviceModel(prs, **kws)
    cps, ths = design_state(dms.sym)
    dms.cp.copy_(torch.from_numpy(cps)); dms.set_theta(ths)
    sts = Step(dms)
    for _ in range(2):
        sts()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        sts()
    b.record(); torch.cuda.synchronize()
    gpu_s = a.elapsed_time(b) / 3e3
    return {"value": 1.0 / dt, "unit": "iters/s", "cores": cpu_threads(), "kind": "port",
                            "sample": "one FULL analysis+adjoint iteration of the restated CPU path (C++/OpenMP assembly incl. penalty, "
                                      "multifrontal LU per Newton step + re-factorised adjoint, %d threads) on %s "
                                      "(N=%d, %.1f s, %d Newton its); NOT scaled to the headline size -- the same mesh on this GPU "
                                      "path takes %.4f s (same_config_gpu_value)"
                                      % (cpu_threads(), workload_name(ne, *sample_topo).split(":")[0], inf["dofs"], dt, inf["newton_its"], gpu_s),
                            "sample_config": {"workload": workload_name(ne, *sample_topo), "dofs": inf["dofs"]},
                            "cpu_phase_s": {k: round(v, 3) for k, v in tm.items()},
                            "same_config_gpu_value": 1.0 / gpu_s, "same_config_gpu_krylov_its": sts.info["krylov_its"],
                            "W_int_cpu": inf["W_int"], "W_int_gpu": float(dms.wv_sum[0].item())}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank)
    import torch
    import torch.distributed as dist
    import __graft_entry__ as g
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        g.build()                       # no-op when the in-tree .so is up to date
    if world > 1:
        dist.barrier()                  # nobody loads the library before rank 0 has (re)built it
    from goldfish_b200 import _capi as capi
    from goldfish_b200.device_model import DeviceModel
    lib = capi.load()
    t_setup = time.perf_counter()
    setup = {}

    def log(msg):
        setup[msg] = round(time.perf_counter() - t_setup, 1)
        if rank == 0:
            print("[bench %7.1fs] %s" % (time.perf_counter() - t_setup, msg), file=sys.stderr, flush=True)
    pr, kw = workload(args.n_el, *topo(args))
    log("problem built")
    dm = DeviceModel(pr, lean=problems_dofs(pr) > 4e6, **kw)
    log("symbolic phase + upload")
    S = dm.sym
    cp, th = design_state(S)
    dm.cp.copy_(torch.from_numpy(cp)); dm.set_theta(th)
    step = Step(dm)
    flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float64, device="cuda")  # 512 MB > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step()
        if i == 0:
            torch.cuda.synchronize(); log("first step (incl. Schwarz + coarse set-up)")
    barrier()
    log("warm-up done")
    l0 = lib.gf_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as cs:
        step_timers = [[] for _ in range(args.steps)]     # four event records per step: phase split of the timed steps
        e0.record()
        for k in range(args.steps):
            step(step_timers[k])
        e1.record()
        barrier()
    launches = lib.gf_launch_count() - l0
    ms = e0.elapsed_time(e1)
    log("timed steps done")
    if world > 1:
        t = torch.tensor([ms], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    value = args.steps / (ms * 1e-3)              # one patch-sharded problem over all ranks (strong scaling)

    # ---- e2e through the facade, host arrays ----
    nm, ops = build_facade(pr, kw, symbolic=S, device_model=dm)
    cp_host = {f: np.concatenate([cp[P.cp_off:P.cp_off + P.ncp, f] for P in S.patches]) for f in kw["opt_field"]}
    for _ in range(max(1, args.warmup - 1)):
        e2e_step(nm, ops, cp_host, th)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = e2e_step(nm, ops, cp_host, th)
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); e2e_s = float(t.item())
    h2d = 8 * (sum(v.size for v in cp_host.values()) + th.size + S.N)            # design vars + update_uIGA
    d2h = 8 * (S.N * 3 + sum(S.P_ncols) * 2 + S.n_th * 3 + 2)                    # u, dWdu, lam-products, gradients, W, V

    # ---- roofline of the dominant kernel: CSR SpMV ----
    x = torch.randn(S.N, dtype=torch.float64, device="cuda"); y = torch.empty_like(x)
    dm.assemble(tangent=True)
    spmv_ms = time_kernel(torch, lambda: dm.spmv_node(x, y), 20, flush)
    spmv_row_ms = time_kernel(torch, lambda: dm.spmv(dm.K, x, y), 20, flush)
    spmv_bytes = 12 * dm.K.nnz + 24 * S.N + 8
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = spmv_bytes / (spmv_ms * 1e-3) / 1e9
    # ---- roofline of the kernel with the largest share of the step: the fine Schwarz sweeps (k_sw_solve1) ----
    # algorithmic bytes per launch = every solve-form panel tile (FP32) once forward and once backward, the
    # diagonal blocks D_j (FP64) once, the block-local vector in (index + gathered value) and out
    import ctypes as C
    from goldfish_b200 import _capi as capi
    sw = dm._schwarz(); A = dm._sw[3]
    sweep_bytes = 2 * int(A["mbj"].sum()) * 64 * 64 * 4 + int(A["nbr"].sum()) * 64 * 64 * 8 + int(A["n_y"]) * (4 + 8 + 8)
    rsw = dm.R.clone()
    try:
        sweep_ms = time_kernel(torch, lambda: capi.check(dm.lib.gf_schwarz_sweeps(C.byref(sw), C.c_void_p(rsw.data_ptr()), dm._stream()),
                                                         "gf_schwarz_sweeps"), 20, flush)
    except capi.GoldfishError:          # blocks too large / too wide for the one-CTA-per-block kernel: group kernel in use
        sweep_ms = float("nan")
    sweep_gbs = sweep_bytes / (sweep_ms * 1e-3) / 1e9
    # assembly kernels (FP64-pipe bound; reported beside the roofline object)
    asm_ms = time_kernel(torch, lambda: dm.assemble(tangent=True, residual=True), 5, flush)
    nq = S.nq
    asm_flops = S.num_elements * nq * 29376.0
    torch.cuda.synchronize()
    phases = {}
    for timers in step_timers:
        for i in range(1, len(timers)):
            phases[timers[i][0]] = phases.get(timers[i][0], 0.0) + timers[i - 1][1].elapsed_time(timers[i][1]) / len(step_timers)
    phases = {k: round(v, 2) for k, v in phases.items()}
    fac_ms = time_kernel(torch, lambda: dm.factor_preconditioner(), 2, flush)
    log("kernel timings done")
    # invariants of the last timed step: comparable across the N = 1, 2, 4, 8 lines of one scaling run
    invariants = {"W_int": float(dm.wv_sum[0].item()), "V": float(dm.wv_sum[1].item()),
                  "u_norm": dm.dot(dm.u, dm.u) ** 0.5, "lam_norm": dm.dot(step.lam, step.lam) ** 0.5,
                  "grad_thickness_norm": float(torch.linalg.vector_norm(step.gT).item()),
                  "grad_shape_norms": [float(torch.linalg.vector_norm(g).item()) for g in step.gP],
                  "true_relres": step.info.get("true_relres")}
    kernels = {"phase_ms": phases, "precond_factor_ms": fac_ms, "spmv_ms": spmv_ms, "spmv_gbs": achieved, "sweeps_ms": sweep_ms, "sweeps_gbs": sweep_gbs,
               "assemble_RK_ms": asm_ms, "assemble_RK_material_tflops": asm_flops / (asm_ms * 1e-3) / 1e12,
               "assemble_RK_gbs_algorithmic": (8 * dm.K.nnz + 8 * (4 * S.n_scalar + S.N + S.n_th)) / (asm_ms * 1e-3) / 1e9,
               "newton_its": step.info.get("newton_its"), "krylov_its": step.info.get("krylov_its")}

    if rank == 0:
        line = {"metric": "analysis+adjoint iters/s", "value": value, "unit": "iters/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(args.n_el, *topo(args)),
                           "dofs": int(S.N), "elements": int(S.num_elements), "nnz_K": int(dm.K.nnz), "quad_pts_per_element": int(nq),
                           "cache": "512 MB flush buffer written between timed kernel launches; step working set > L2 at n_el >= 64",
                           "parallelism": "1 GPU" if world == 1 else "patch-sharded over %d GPUs (rows + Schwarz blocks owned, vectors replicated, NCCL all-reduce)" % world},
                "e2e": {"value": args.steps / e2e_s, "unit": "iters/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "kernel": "k_sw_solve1", "achieved": sweep_gbs, "peak": peak, "unit": "GB/s",
                             "frac": sweep_gbs / peak,
                             # dram__bytes_read.sum + dram__bytes_write.sum of ONE k_sw_solve1 launch on this workload
                             # (ncu --set full, profiles/r1_ncu_sweeps_c3.txt)
                             "traffic": SWEEP_TRAFFIC_C3 if (args.n_el == 201 and world == 1) else None,
                             "algorithmic_bytes": int(sweep_bytes), "launch_ms": sweep_ms,
                             "launches_per_step": int(sum(step.info.get("krylov_its") or []) + len(step.info.get("krylov_its") or [])),
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst copy)" if peaks else "fallback 6650 GB/s"},
                "roofline_spmv": {"bound": "hbm", "kernel": "k_spmv_node (tangent product of the Krylov loop)", "achieved": achieved,
                                  "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                                  "traffic": SPMV_NODE_TRAFFIC_C3 if (args.n_el == 201 and world == 1) else None,
                                  "algorithmic_bytes": int(spmv_bytes),
                                  # the node-wise kernel reads one column index and one x entry per THREE non-zeros:
                                  # bytes it has to move = (8 + 4/3) nnz + 24 N; its bandwidth on that count
                                  "bytes_moved_model": int((8 + 4.0 / 3.0) * dm.K.nnz + 24 * S.N),
                                  "achieved_on_bytes_moved": ((8 + 4.0 / 3.0) * dm.K.nnz + 24 * S.N) / (spmv_ms * 1e-3) / 1e9,
                                  "launch_ms": spmv_ms,
                                  "rowwise_k_spmv_ms": spmv_row_ms, "rowwise_k_spmv_gbs": spmv_bytes / (spmv_row_ms * 1e-3) / 1e9},
                "parity": None, "invariants": invariants, "setup_s": setup,
                "kernels": kernels, "clocks": cs.summary()}
    else:
        line = None
    # ---------------- optional, untimed sections; each under the watchdog ----------------
    wd = Watchdog(rank); wd.line = line
    parity = None
    if not args.no_parity:
        if world == 1 or args.parity_sharded:
            wd.start(300, "parity checks at the benched size")
            try:
                parity = parity_checks(dm, step, torch, world)
            except Exception as e:          # reported in the line, never hidden
                parity = {"error": "%s: %s" % (type(e).__name__, e), "newton_history": [float(h) for h in getattr(dm, "newton_history", [])]}
                step.newton_rtol = 1e-3
            wd.cancel()
        else:
            # N > 1: the finite-difference / tight-Newton checks run on the N = 1 line (same workload); this line carries
            # the invariants (W_int, |u|, |lambda|, gradient norms: equal to the N = 1 line's to solver tolerance) and
            # the sharded-vs-single comparison below (`--parity-sharded` runs the full checks sharded as well)
            parity = {"mode": "N > 1: invariants + sharded-vs-single run; FD / tight-Newton checks are on the N = 1 line"}
        if line is not None:
            line["parity"] = parity
        if world > 1:
            wd.start(240, "sharded vs single-GPU run of a small model")
            try:
                rvs = ranks_vs_single(torch, dist, world, rank)
            except Exception as e:
                rvs = {"error": "%s: %s" % (type(e).__name__, e)}
            wd.cancel()
            parity["ranks_vs_single"] = rvs
    log("parity done")
    if rank == 0:
        if world == 1:
            wd.start(120, "other configurations")
            try:
                line["other_configs"] = small_configs(torch)
            except Exception as e:
                line["other_configs"] = {"error": "%s: %s" % (type(e).__name__, e)}
            wd.cancel()
        if not args.no_cpu_baseline and world == 1:     # reported at N = 1 only (bench contract)
            wd.start(300, "cpu_baseline sample")
            try:
                line["cpu_baseline"] = cpu_baseline_sample(args, torch, DeviceModel)
            except Exception as e:
                line["cpu_baseline"] = {"error": "%s: %s" % (type(e).__name__, e)}
            wd.cancel()
        print(json.dumps(line), flush=True)
        wd.line = None                      # printed: the watchdog must not print it again
    if world > 1:
        wd.start(60, "process-group shutdown")
        dist.destroy_process_group()
        wd.cancel()


if __name__ == "__main__":
    main()
