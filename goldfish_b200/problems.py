"""Problem definitions (plain-data dictionaries) for fixtures and benchmarks.

A *problem* is what the reference scripts build with igakit/tIGAr/PENGoLINS
before the hot path starts: spline patches (knots, homogeneous control points,
zero-DoFs), material, thickness description, loads and the non-matching
intersections with their mortar parametric coordinates.

  tbeam()          /root/reference/GOLDFISH/tests/test_tbeam.py:38-117 (BASELINE C2)
  scordelis_lo()   /root/reference/GOLDFISH/tests/test_slr.py:6-142
  plate()          /root/reference/demos_csdl_alpha/thickness_opt/plate_const_th_opt_wint.py:12-150
                   (BASELINE C1; geometry + intersection tables from tests/golden/plate_c1_input.npz)
  cylinder()       synthetic 8-patch non-matching cylinder (BASELINE C3)
  wingbox()        synthetic 40-patch wing box with T- and X-junction intersections (BASELINE C4)
  twisted_beam()   MacNeal-Harder twisted beam on two non-matching patches (known answer, tests only)
  cantilever_shear()  geometrically nonlinear cantilever under end shear (Sze et al. 2004; known answer, tests only)
  hemisphere()     pinched hemisphere with an 18-degree hole on four non-matching NURBS patches (known answer, tests only)

Layout conventions: scalar CP index a = i + j*n_u; patch-local vector dof
= field*n_cp + a; global dofs = patches concatenated in list order
(/root/reference/GOLDFISH/nonmatching_opt.py:50-65).
"""
import numpy as np
from . import bsplines as bsp


def _side_dofs(n_u, n_v, direction, side, n_layers=1):
    """tIGAr ``getSideDofs(direction, side, nLayers)`` on the scalar CP grid."""
    I, J = np.meshgrid(np.arange(n_u), np.arange(n_v), indexing="ij")
    idx = I if direction == 0 else J
    n = n_u if direction == 0 else n_v
    sel = idx < n_layers if side == 0 else idx >= n - n_layers
    return np.sort((I + J * n_u)[sel])


def _patch_from_surface(srf, quad_deg, thickness, bc=(), body_force=(0.0, 0.0, 0.0)):
    n_u, n_v = srf.control.shape[0], srf.control.shape[1]
    ncp = n_u * n_v
    bc_dofs = []
    for field, direction, side, n_layers in bc:
        bc_dofs.append(field * ncp + _side_dofs(n_u, n_v, direction, side, n_layers))
    bc_dofs = np.unique(np.concatenate(bc_dofs)) if bc_dofs else np.zeros(0, dtype=np.int64)
    return dict(p=tuple(srf.degree), knots=(srf.knots[0].copy(), srf.knots[1].copy()),
                cp=srf.flat_control(), bc_dofs=bc_dofs.astype(np.int64), quad_deg=int(quad_deg),
                thickness=dict(thickness), body_force=tuple(body_force))


def mortar_coords(ends, n_cells):
    """PENGoLINS ``move_mortar_mesh`` with two end points: n_cells+1 equally
    spaced mortar vertices in the parametric space of one side."""
    ends = np.asarray(ends, dtype=np.float64)
    if ends.shape[0] != 2:
        return ends.copy()
    s = np.linspace(0.0, 1.0, n_cells + 1)[:, None]
    return ends[0][None, :] * (1.0 - s) + ends[1][None, :] * s


def _ruled_quad(pts, n_el0, n_el1, p):
    """igakit: ruled(line(pts0,pts1), line(pts2,pts3)); elevate to p; refine
    uniformly (/root/reference/GOLDFISH/tests/test_tbeam.py:5-16)."""
    srf = bsp.ruled(bsp.line(pts[0], pts[1]), bsp.line(pts[2], pts[3]))
    srf.elevate(0, p - srf.degree[0]); srf.elevate(1, p - srf.degree[1])
    srf.refine(0, np.linspace(0, 1, n_el0 + 1)[1:-1])
    srf.refine(1, np.linspace(0, 1, n_el1 + 1)[1:-1])
    return srf


def tbeam(num_el=10, p=3, E=1.0e7, nu=0.0, h_th=0.1, penalty_coefficient=1.0e3,
          body_force=(0.0, 0.0, 0.0), tip_load=-10.0, thickness_kind="const",
          L=20.0, w=2.0, h=2.0, quad_deg_const=3):
    """Two-patch T-beam (flange + web), one intersection."""
    pts0 = [[-w / 2., 0., 0.], [w / 2., 0., 0.], [-w / 2., L, 0.], [w / 2., L, 0.]]
    pts1 = [[0., 0., 0.], [0., 0., -h], [0., L, 0.], [0., L, -h]]
    n0, n1 = num_el, num_el + 1
    srf0 = _ruled_quad(pts0, n0 // 2, n0, p)
    srf1 = _ruled_quad(pts1, n1 // 2, n1, p)
    bc = [(f, 1, 0, 1) for f in range(3)]  # pin side 0 of direction 1, all fields
    th = dict(kind=thickness_kind, values=h_th)
    patches = [_patch_from_surface(s, quad_deg_const * p, th, bc, body_force) for s in (srf0, srf1)]
    n_m = 2 * n1
    interfaces = [dict(patches=(0, 1),
                       xi=(mortar_coords([[0.5, 0.0], [0.5, 1.0]], n_m),
                           mortar_coords([[0.0, 0.0], [0.0, 1.0]], n_m)))]
    # PointSource(spline0.V.sub(2), Point(1.,1.), -tip_load): adds -tip_load to R
    point_loads = [dict(patch=0, field=2, xi=(1.0, 1.0), value=-tip_load)] if tip_load else []
    return dict(name="tbeam", patches=patches, E=E, nu=nu, interfaces=interfaces,
                penalty_coefficient=penalty_coefficient, point_loads=point_loads, edge_loads=[])


def _roof_patch(num_el, p, R, angle_lim, z_lim):
    a = (np.radians(angle_lim[0]), np.radians(angle_lim[1]))
    C = bsp.circle_arc([0, 0, z_lim[0]], R, a)
    T = bsp.circle_arc([0, 0, z_lim[1]], R, a)
    S = bsp.ruled(C, T)
    S.elevate(0, p - S.degree[0]); S.elevate(1, p - S.degree[1])
    new = np.linspace(0, 1, num_el + 1)[1:-1]
    S.refine(0, new); S.refine(1, new)
    return S


def scordelis_lo(num_el=6, p=3, penalty_coefficient=1.0e3, quad_deg_const=2):
    """Nine-patch non-matching Scordelis-Lo roof; QoI_ref = 0.3006 is the
    vertical displacement at the mid-point of the free edge
    (/root/reference/GOLDFISH/tests/test_slr.py:40-50)."""
    L, R = 50.0, 25.0
    E, nu, h_th = 4.32e8, 0.0, 0.25
    f = (0.0, -90.0, 0.0)
    angles = [50, 80, 100, 130]
    angle_lims = [angles[0:2], angles[1:3], angles[2:4]] * 3
    z = [0, L / 4, 3 * L / 4, L]
    z_lims = [z[0:2]] * 3 + [z[1:3]] * 3 + [z[2:4]] * 3
    ne = num_el
    nels = [ne, ne - 2, ne - 1, ne + 2, ne + 1, ne + 3, ne - 1, ne, ne - 2]
    bcs = [[1, 0]] * 3 + [[0, 0]] * 3 + [[0, 1]] * 3
    patches = []
    for i in range(9):
        S = _roof_patch(nels[i], p, R, angle_lims[i], z_lims[i])
        bc = []
        for field in (0, 1):
            for side in (0, 1):
                if bcs[i][side] == 1:
                    bc.append((field, 1, side, 1))
        P = _patch_from_surface(S, quad_deg_const * p, dict(kind="const", values=h_th), bc, f)
        if i == 0:  # fix_z_node: pin z displacement of control point 0
            ncp = S.control.shape[0] * S.control.shape[1]
            P["bc_dofs"] = np.unique(np.concatenate([P["bc_dofs"], [2 * ncp + 0]])).astype(np.int64)
        patches.append(P)
    mapping = [[0, 1], [1, 2], [3, 4], [4, 5], [6, 7], [7, 8],
               [0, 3], [3, 6], [1, 4], [4, 7], [2, 5], [5, 8]]
    h_locs = [[[0., 1.], [1., 1.]], [[0., 0.], [1., 0.]]]
    v_locs = [[[1., 0.], [1., 1.]], [[0., 0.], [0., 1.]]]
    interfaces = []
    for j, (a, b) in enumerate(mapping):
        n_m = 3 * (nels[a] + nels[b])
        locs = v_locs if j < 6 else h_locs
        interfaces.append(dict(patches=(a, b), xi=(mortar_coords(locs[0], n_m), mortar_coords(locs[1], n_m))))
    return dict(name="scordelis_lo", patches=patches, E=E, nu=nu, interfaces=interfaces,
                penalty_coefficient=penalty_coefficient, point_loads=[], edge_loads=[],
                qoi=dict(patch=3, xi=(0.0, 0.5), field=1, ref=0.3006))


def plate(npz_path, E=68e9, nu=0.35, h_th=1.0e-2, penalty_coefficient=1.0e3, load=-100.0,
          quad_deg_const=4, thickness_kind="linear"):
    """Six-strip non-matching plate of the CSDL thickness-optimisation demo."""
    d = np.load(npz_path)
    patches = []
    ns = int(d["num_patches"])
    for s in range(ns):
        pu, pv = [int(x) for x in d[f"p{s}_deg"]]
        ku, kv = d[f"p{s}_ku"], d[f"p{s}_kv"]
        n_u, n_v = len(ku) - pu - 1, len(kv) - pv - 1
        ncp = n_u * n_v
        bc = np.zeros(0, dtype=np.int64)
        if s == 0:  # clampedBC(side=0, direction=0): field 0 one layer, fields 1,2 two layers
            bc = np.unique(np.concatenate(
                [f * ncp + _side_dofs(n_u, n_v, 0, 0, 1 if f == 0 else 2) for f in range(3)]))
        patches.append(dict(p=(pu, pv), knots=(ku, kv), cp=d[f"p{s}_cp"], bc_dofs=bc,
                            quad_deg=quad_deg_const * pu,
                            thickness=dict(kind=thickness_kind, values=h_th),
                            body_force=(0.0, 0.0, 0.0)))
    interfaces = []
    for i, (a, b) in enumerate(d["mapping_list"]):
        interfaces.append(dict(patches=(int(a), int(b)), xi=(d[f"int{i}_xi0"], d[f"int{i}_xi1"])))
    edge_loads = [dict(patch=ns - 1, direction=0, side=1, traction=(0.0, 0.0, load))]
    return dict(name="plate", patches=patches, E=E, nu=nu, interfaces=interfaces,
                penalty_coefficient=penalty_coefficient, point_loads=[], edge_loads=edge_loads)


def cylinder(n_el=32, p=3, n_circ=4, n_axial=2, R=1.0, L=4.0, E=68e9, nu=0.35, h_th=1.0e-2,
             penalty_coefficient=1.0e3, pressure_like_load=(0.0, 0.0, -1.0e3), quad_deg_const=3,
             thickness_kind="const", arc_deg=360.0, jitter=True):
    """Synthetic non-matching multi-patch cylinder (BASELINE C3 topology,
    SURVEY.md section 8d): n_circ arcs x n_axial axial segments, bicubic, patch s
    has n_el + delta_s elements per side (delta_s = s mod 8) so that meshes do
    not match along interfaces; one end (z = 0) clamped with two CP layers;
    dead load per unit area.  arc_deg < 360 gives an open panel."""
    patches, nels = [], []
    da = arc_deg / n_circ
    closed = abs(arc_deg - 360.0) < 1e-12
    for ia in range(n_axial):
        for ic in range(n_circ):
            s = ic + ia * n_circ
            ne = n_el + ((s % 8) if jitter else 0)
            nels.append(ne)
            S = _roof_patch(ne, p, R, [ic * da, (ic + 1) * da], [ia * L / n_axial, (ia + 1) * L / n_axial])
            bc = []
            if ia == 0:
                bc = [(f, 1, 0, 2) for f in range(3)]
            patches.append(_patch_from_surface(S, quad_deg_const * p, dict(kind=thickness_kind, values=h_th),
                                               bc, pressure_like_load))
    v_locs = [[[1., 0.], [1., 1.]], [[0., 0.], [0., 1.]]]
    h_locs = [[[0., 1.], [1., 1.]], [[0., 0.], [1., 0.]]]
    interfaces = []
    for ia in range(n_axial):
        for ic in range(n_circ):
            a = ic + ia * n_circ
            if ic + 1 < n_circ or (closed and n_circ > 1):
                b = (ic + 1) % n_circ + ia * n_circ
                n_m = 2 * max(nels[a], nels[b])
                interfaces.append(dict(patches=(a, b), xi=(mortar_coords(v_locs[0], n_m), mortar_coords(v_locs[1], n_m))))
            if ia + 1 < n_axial:
                b = ic + (ia + 1) * n_circ
                n_m = 2 * max(nels[a], nels[b])
                interfaces.append(dict(patches=(a, b), xi=(mortar_coords(h_locs[0], n_m), mortar_coords(h_locs[1], n_m))))
    return dict(name=f"cylinder_{n_circ}x{n_axial}_ne{n_el}", patches=patches, E=E, nu=nu,
                interfaces=interfaces, penalty_coefficient=penalty_coefficient, point_loads=[],
                edge_loads=[])


def wingbox(h=0.05, n_seg=10, n_spar=3, n_rib=17, L=10.0, C=2.0, H=0.4, p=3, E=68e9, nu=0.35, h_th=1.0e-2,
            penalty_coefficient=1.0e3, pressure=-1.0e3, quad_deg_const=3, thickness_kind="const", target_dofs=None):
    """Synthetic wing-box-like shell (BASELINE configs[3], SURVEY.md section 8d "C4"): a rectangular box beam of
    span L (x), chord C (y) and height H (z) made of
      2 skins x n_seg spanwise segments  (segments meet edge to edge),
      n_spar full-span spars             (top / bottom edges lie INSIDE every skin segment: T-junctions),
      n_rib ribs between the outer spars (edges inside a skin segment and on the outer spars: T-junctions;
                                          the inner spars pass through the rib interior: X-junctions),
    i.e. 2 n_seg + n_spar + n_rib non-matching bicubic patches (40 by default) and
    2 (n_seg - 1) + 2 n_spar n_seg + 2 n_rib + n_rib n_spar intersections (163), most of them interior to at least
    one patch -- the situation of the reference's wing models (demos_csdl_alpha/ex_caddee/wing_int_data.npz: 93 of
    124 interface sides in a patch interior).  Every patch has its own element size (h scaled by 1 + 0.04 (s mod 6))
    so no two meshes match.  Root (x = 0) clamped with two CP layers, dead pressure on the upper skin.
    target_dofs: choose h so that the model has about that many displacement dofs."""
    if target_dofs is not None:
        dims = [(L / n_seg, C)] * (2 * n_seg) + [(L, H)] * n_spar + [(0.8 * C if n_spar > 1 else C, H)] * n_rib

        def count(hh):
            return sum(3 * (max(4, int(np.ceil(a / (hh * (1 + 0.04 * (s % 6)))))) + p) * (max(4, int(np.ceil(b / (hh * (1 + 0.04 * (s % 6)))))) + p)
                       for s, (a, b) in enumerate(dims))
        h = float(np.sqrt(3.0 * sum(a * b for a, b in dims) / target_dofs))
        for _ in range(6):
            h *= float(np.sqrt(count(h) / target_dofs))
    xs = np.linspace(0.0, L, n_seg + 1)
    y_sp = np.linspace(0.1 * C, 0.9 * C, n_spar) if n_spar > 1 else np.array([0.5 * C])
    ya, yb = float(y_sp[0]), float(y_sp[-1])
    # rib stations: evenly spread, nudged away from the segment boundaries
    x_rib = (np.arange(n_rib) + 0.37) * L / n_rib
    for r in range(n_rib):
        d = np.abs(xs - x_rib[r])
        if d.min() < 0.08 * L / n_seg:
            x_rib[r] += 0.13 * L / n_seg
    th = dict(kind=thickness_kind, values=h_th)
    patches, meta = [], []          # meta: (kind, index data, n_el0, n_el1)

    def add(pts, Lu, Lv, bc, load):
        s = len(patches)
        hs = h * (1.0 + 0.04 * (s % 6))
        n0 = max(4, int(np.ceil(Lu / hs))); n1 = max(4, int(np.ceil(Lv / hs)))
        srf = _ruled_quad(pts, n0, n1, p)
        patches.append(_patch_from_surface(srf, quad_deg_const * p, th, bc, load))
        meta.append((n0, n1))
        return s
    clamp = [(f, 0, 0, 2) for f in range(3)]
    skin = {}
    for side, z in (("lo", 0.0), ("up", H)):
        for k in range(n_seg):
            pts = [[xs[k], 0, z], [xs[k + 1], 0, z], [xs[k], C, z], [xs[k + 1], C, z]]
            skin[side, k] = add(pts, xs[k + 1] - xs[k], C, clamp if k == 0 else [], (0.0, 0.0, pressure) if side == "up" else (0.0, 0.0, 0.0))
    spar = [add([[0, y, 0], [L, y, 0], [0, y, H], [L, y, H]], L, H, clamp, (0.0, 0.0, 0.0)) for y in y_sp]
    rib = [add([[x, ya, 0], [x, yb, 0], [x, ya, H], [x, yb, H]], yb - ya, H, [], (0.0, 0.0, 0.0)) for x in x_rib]
    interfaces = []

    def itf(a, endsA, b, endsB, nA, nB):
        n_m = 2 * max(nA, nB, 2)
        interfaces.append(dict(patches=(a, b), xi=(mortar_coords(endsA, n_m), mortar_coords(endsB, n_m))))
    for side in ("lo", "up"):
        v_edge = 0.0 if side == "lo" else 1.0
        for k in range(n_seg - 1):                                   # skin segment | skin segment
            a, b = skin[side, k], skin[side, k + 1]
            itf(a, [[1., 0.], [1., 1.]], b, [[0., 0.], [0., 1.]], meta[a][1], meta[b][1])
        for j, y in enumerate(y_sp):                                 # spar edge on the inside of every skin segment
            for k in range(n_seg):
                a, b = skin[side, k], spar[j]
                itf(a, [[0., y / C], [1., y / C]], b, [[xs[k] / L, v_edge], [xs[k + 1] / L, v_edge]],
                    meta[a][0], int(np.ceil(meta[b][0] / n_seg)))
        for r, x in enumerate(x_rib):                                # rib edge on the inside of one skin segment
            k = int(np.searchsorted(xs, x) - 1)
            a, b = skin[side, k], rib[r]
            u = (x - xs[k]) / (xs[k + 1] - xs[k])
            itf(a, [[u, ya / C], [u, yb / C]], b, [[0., v_edge], [1., v_edge]], int(np.ceil(meta[a][1] * (yb - ya) / C)), meta[b][0])
    for r, x in enumerate(x_rib):                                    # rib | spar (vertical lines)
        for j, y in enumerate(y_sp):
            a, b = spar[j], rib[r]
            ur = (y - ya) / (yb - ya) if yb > ya else 0.5
            itf(a, [[x / L, 0.], [x / L, 1.]], b, [[ur, 0.], [ur, 1.]], meta[a][1], meta[b][1])
    return dict(name=f"wingbox_{len(patches)}p_h{h:.4g}", patches=patches, E=E, nu=nu, interfaces=interfaces,
                penalty_coefficient=penalty_coefficient, point_loads=[], edge_loads=[])


def _twisted_strip(s0, s1, L, w, ne_u, ne_v, p=3):
    """Bicubic patch of the 90-degree twisted strip X = (r cos phi, r sin phi, s), phi = pi s / (2 L), s in [s0, s1]:
    exact in r (linear, degree elevated), cubic B-spline interpolation at the Greville points in s."""
    kv = np.concatenate([np.zeros(p), np.linspace(0, 1, ne_v + 1), np.ones(p)])
    n = len(kv) - p - 1
    g = np.array([kv[j + 1:j + p + 1].mean() for j in range(n)])
    span, B = bsp.basis_window(kv, p, g, 0)
    Cm = np.zeros((n, n))
    for i in range(n):
        Cm[i, span[i] - p:span[i] + 1] = B[i, 0]
    s = s0 + (s1 - s0) * g
    phi = 0.5 * np.pi * s / L
    coef = np.linalg.solve(Cm, np.stack([np.cos(phi), np.sin(phi), s], 1))
    ctrl = np.zeros((2, n, 4))
    for i, r in enumerate((-0.5 * w, 0.5 * w)):
        ctrl[i, :, 0] = r * coef[:, 0]; ctrl[i, :, 1] = r * coef[:, 1]; ctrl[i, :, 2] = coef[:, 2]; ctrl[i, :, 3] = 1.0
    srf = bsp.NURBSSurface([[0., 0., 1., 1.], kv], [1, p], ctrl)
    srf.elevate(0, p - 1)
    srf.refine(0, np.linspace(0, 1, ne_u + 1)[1:-1])
    return srf


def twisted_beam(ne=16, tip_force=(0.0, 1.0, 0.0), L=12.0, w=1.1, t=0.32, E=29.0e6, nu=0.22, penalty_coefficient=1.0e3):
    """MacNeal-Harder twisted beam (90-degree twist, root clamped, unit tip force) as TWO non-matching bicubic patches
    coupled by the penalty method: a known answer on doubly curved geometry with nu != 0.  Reference tip displacements
    in the load direction: 5.424e-3 for the in-plane load (tip width direction, global y) and 1.754e-3 for the
    out-of-plane load (global x); a Kirchhoff-Love shell (no transverse shear) gives 0.995 of both."""
    th = dict(kind="const", values=t)
    cuts = [0.0, 0.45 * L, L]
    nes = [(3, ne), (4, ne + 3)]
    patches = []
    for k in range(2):
        srf = _twisted_strip(cuts[k], cuts[k + 1], L, w, nes[k][0], nes[k][1])
        bc = [(f, 1, 0, 2) for f in range(3)] if k == 0 else []
        patches.append(_patch_from_surface(srf, 9, th, bc, (0.0, 0.0, 0.0)))
    n_m = 2 * max(nes[0][0], nes[1][0]) + 4
    itf = [dict(patches=(0, 1), xi=(mortar_coords([[0., 1.], [1., 1.]], n_m), mortar_coords([[0., 0.], [1., 0.]], n_m)))]
    loads = [dict(patch=1, field=f, xi=(0.5, 1.0), value=-tip_force[f]) for f in range(3) if tip_force[f] != 0.0]
    return dict(name="twisted_beam", patches=patches, E=E, nu=nu, interfaces=itf, penalty_coefficient=penalty_coefficient,
                point_loads=loads, edge_loads=[])


def _arc3(P0, P2, center):
    """Rational quadratic arc from P0 to P2 on the circle around `center` (sweep < 180 deg): homogeneous control points."""
    P0, P2, center = (np.asarray(v, dtype=np.float64) for v in (P0, P2, center))
    a, b = P0 - center, P2 - center
    R = np.linalg.norm(a)
    w = np.cos(0.5 * np.arccos(np.clip(a @ b / (R * R), -1.0, 1.0)))
    m = a + b
    m = m / np.linalg.norm(m) * (R / w)
    c = np.zeros((3, 4))
    for r, (P, wt) in enumerate(((a, 1.0), (m, w), (b, 1.0))):
        c[r, :3] = (center + P) * wt; c[r, 3] = wt
    return c


def _sphere_patch(R, th0, th1, ph0, ph1, ne_u, ne_v, p=3):
    """Exact NURBS patch of a sphere: u = azimuth in [ph0, ph1], v = polar angle in [th0, th1] (biquadratic rational
    surface of revolution, elevated to degree p, refined)."""
    mer = _arc3([R * np.sin(th0), 0, R * np.cos(th0)], [R * np.sin(th1), 0, R * np.cos(th1)], [0, 0, 0])
    ctrl = np.zeros((3, 3, 4))
    for j in range(3):
        wj = mer[j, 3]; xj = mer[j, 0] / wj; zj = mer[j, 2] / wj
        az = _arc3([xj * np.cos(ph0), xj * np.sin(ph0), zj], [xj * np.cos(ph1), xj * np.sin(ph1), zj], [0, 0, zj])
        for i in range(3):
            wi = az[i, 3]
            ctrl[i, j, :3] = az[i, :3] / wi * (wi * wj); ctrl[i, j, 3] = wi * wj
    k = [0., 0., 0., 1., 1., 1.]
    srf = bsp.NURBSSurface([k, k], [2, 2], ctrl)
    srf.elevate(0, p - 2); srf.elevate(1, p - 2)
    srf.refine(0, np.linspace(0, 1, ne_u + 1)[1:-1]); srf.refine(1, np.linspace(0, 1, ne_v + 1)[1:-1])
    return srf


def hemisphere(ne=16, R=10.0, t=0.04, E=6.825e7, nu=0.3, F=2.0, hole_deg=18.0, penalty_coefficient=1.0e3):
    """Pinched hemisphere with an 18-degree hole (shell obstacle course, MacNeal-Harder): four NON-MATCHING exact NURBS
    patches (90 degrees of azimuth each, ne + k elements per side) closed into a ring by penalty coupling, alternating
    radial forces +-F on the equator at 0, 90, 180, 270 degrees (patch corners), six statically determinate supports.
    Reference radial displacement under the loads: 0.0940.  Inextensional bending of a doubly curved rational surface."""
    th = dict(kind="const", values=t)
    patches, nes = [], []
    for k in range(4):
        nes.append(ne + k)
        srf = _sphere_patch(R, np.radians(hole_deg), np.radians(90.0), k * np.pi / 2, (k + 1) * np.pi / 2, ne + k, ne + k)
        patches.append(_patch_from_surface(srf, 9, th, [], (0.0, 0.0, 0.0)))
    itf = []
    for k in range(4):
        a, b = k, (k + 1) % 4
        n_m = 2 * max(nes[a], nes[b])
        itf.append(dict(patches=(a, b), xi=(mortar_coords([[1., 0.], [1., 1.]], n_m), mortar_coords([[0., 0.], [0., 1.]], n_m))))
    loads = []
    for k in range(4):                                   # corner (u = 0, v = 1) of patch k = equator at azimuth k * 90 deg
        er = np.array([np.cos(k * np.pi / 2), np.sin(k * np.pi / 2)])
        sgn = 1.0 if k % 2 == 0 else -1.0
        loads += [dict(patch=k, field=f, xi=(0.0, 1.0), value=-sgn * F * er[f]) for f in range(2) if abs(er[f]) > 1e-12]
    P0, P2 = patches[0], patches[2]
    n_u, n_v = len(P0["knots"][0]) - 4, len(P0["knots"][1]) - 4
    ncp, a0, a1 = n_u * n_v, (n_v - 1) * n_u, (n_v - 1) * n_u + n_u - 1
    P0["bc_dofs"] = np.array([2 * ncp + a0, 2 * ncp + a1, ncp + a0, a1], dtype=np.int64)      # u_z at 0 and 90 deg, tangential there
    n_u2, n_v2 = len(P2["knots"][0]) - 4, len(P2["knots"][1]) - 4
    ncp2, b0 = n_u2 * n_v2, (n_v2 - 1) * n_u2
    P2["bc_dofs"] = np.array([2 * ncp2 + b0, ncp2 + b0], dtype=np.int64)                      # u_z and tangential u_y at 180 deg
    return dict(name="hemisphere", patches=patches, E=E, nu=nu, interfaces=itf, penalty_coefficient=penalty_coefficient,
                point_loads=loads, edge_loads=[])


def cantilever_shear(P=4.0, ne=8, L=10.0, b=1.0, t=0.1, E=1.2e6, nu=0.0, penalty_coefficient=1.0e3):
    """Cantilever strip under an end shear force (Sze, Liu & Lo 2004, the standard geometrically NONLINEAR shell
    benchmark: L = 10, b = 1, t = 0.1, E = 1.2e6, nu = 0, P_max = 4) as two non-matching patches along the length;
    total force P as a dead traction on the tip edge.  Tip deflections (-u_x, u_z): P = 1: (0.563, 3.015),
    P = 2: (1.603, 4.933), P = 4: (3.286, 6.698)."""
    th = dict(kind="const", values=t)
    cuts = [0.0, 0.4 * L, L]
    nes = [(2, ne), (3, ne + 3)]
    patches = []
    for k in range(2):
        pts = [[cuts[k], 0, 0], [cuts[k], b, 0], [cuts[k + 1], 0, 0], [cuts[k + 1], b, 0]]     # u: width (y), v: length (x)
        srf = _ruled_quad(pts, nes[k][0], nes[k][1], 3)
        bc = [(f, 1, 0, 2) for f in range(3)] if k == 0 else []
        patches.append(_patch_from_surface(srf, 9, th, bc, (0.0, 0.0, 0.0)))
    itf = [dict(patches=(0, 1), xi=(mortar_coords([[0., 1.], [1., 1.]], 10), mortar_coords([[0., 0.], [1., 0.]], 10)))]
    edge = [dict(patch=1, direction=1, side=1, traction=(0.0, 0.0, P / b))]
    return dict(name="cantilever_shear", patches=patches, E=E, nu=nu, interfaces=itf, penalty_coefficient=penalty_coefficient,
                point_loads=[], edge_loads=edge)


def num_dofs(problem):
    return sum(3 * (len(P["knots"][0]) - P["p"][0] - 1) * (len(P["knots"][1]) - P["p"][1] - 1)
               for P in problem["patches"])
