"""Shared problem set-ups for the tests (BASELINE configs at oracle-friendly sizes)."""
import os
import numpy as np
from goldfish_b200 import problems

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def tbeam_small():
    """C2 topology (2 patches, 1 intersection), 4 elements, body force + tip load, spline thickness."""
    return (problems.tbeam(num_el=4, body_force=(0.0, 0.0, 1.0), thickness_kind="iga"),
            dict(opt_field=[0, 1, 2], shopt_surf_inds=[[0, 1]] * 3))


def tbeam_c2():
    """BASELINE C2: test_tbeam.py fixture (num_el=10, N=648), opt_field=[0]."""
    return problems.tbeam(num_el=10), dict(opt_field=[0], shopt_surf_inds=[[0, 1]])


def slr_small():
    """Scordelis-Lo, 9 NURBS patches / 12 intersections, constant thickness, one shape field on 3 patches."""
    return problems.scordelis_lo(num_el=4), dict(opt_field=[1], shopt_surf_inds=[[0, 3, 4]])


def wingbox_small():
    """C4 topology in miniature: 2 x 2 skin segments, 2 spars, 1 rib = 7 non-matching patches, 14 intersections of
    which 12 lie in the interior of a patch (T-junctions; the rib/spar lines are edge-to-interior)."""
    pr = problems.wingbox(h=0.45, n_seg=2, n_spar=2, n_rib=1, L=2.0, C=1.0, H=0.4)
    # shape variables: one field on a skin segment, a spar and the rib (keeps the numpy oracle's jets affordable)
    return pr, dict(opt_field=[1], shopt_surf_inds=[[1, 4, 6]])


def plate_c1():
    """BASELINE C1: six-strip plate, V_linear thickness, edge traction, quad_deg 12."""
    return problems.plate(os.path.join(GOLDEN, "plate_c1_input.npz")), dict()


def random_state(N, bc, seed=0, scale=1e-2):
    u = scale * np.random.default_rng(seed).standard_normal(N)
    u[bc] = 0.0
    return u
