"""CPU: host set-up of the patch-sharded path -- what every rank builds for itself must add up to the unsharded set-up.

  * Symbolic(own_patches=...) builds the coupling gather lists only for destinations in own patches: over the ranks they
    partition the unsharded lists (same items, same positions).
  * SchwarzSetup(own_patches=...) builds exactly the blocks the unsharded set-up builds for those patches.
  * the sub-domain shape and coarse-level rules give the documented choices.
"""
import numpy as np

import cases
from goldfish_b200 import problems, coarse
from goldfish_b200.partition import lpt_partition, shard_symbolic
from goldfish_b200.schwarz import SchwarzSetup
from goldfish_b200.symbolic import Symbolic


def _owner(pr, world):
    nel = [(len(np.unique(P["knots"][0])) - 1) * (len(np.unique(P["knots"][1])) - 1) for P in pr["patches"]]
    return lpt_partition(nel, world)


def test_own_patch_symbolic_partitions_the_coupling_lists():
    pr, _ = cases.wingbox_small()
    kw = dict(opt_field=[0, 2], shopt_surf_inds=[list(range(len(pr["patches"])))] * 2)
    world = 3
    owner = _owner(pr, world)
    full = Symbolic(pr, **kw)
    tot = dict(nR=0, nK=0, itemsK=0, itemsR=0, destP=0, itemsP=0)
    seenK = []
    for rank in range(world):
        S = Symbolic(pr, own_patches=(owner == rank), **kw)
        # patterns do not depend on the filter
        assert np.array_equal(S.K_indptr, full.K_indptr) and np.array_equal(S.K_indices, full.K_indices)
        sh = shard_symbolic(S, owner, rank)
        ref = shard_symbolic(full, owner, rank)                      # filtering the unsharded lists must give the same
        for k in ("R_ptr", "R_item", "R_row", "K_ptr", "K_item", "K_pos"):
            assert np.array_equal(sh["pen"][k], ref["pen"][k]), k
        assert [len(rd["item_eval"]) for pp in sh["penP"] for rd in pp["rounds"]] is not None
        nP = sum(rd["n_dest"] for rd in sh["penP"][0]["rounds"])
        nP_ref = sum(rd["n_dest"] for rd in ref["penP"][0]["rounds"])
        assert nP == nP_ref
        tot["nR"] += sh["pen"]["nR"]; tot["nK"] += sh["pen"]["nK"]
        tot["itemsK"] += len(sh["pen"]["K_item"]); tot["itemsR"] += len(sh["pen"]["R_item"])
        tot["destP"] += nP
        seenK.append(sh["pen"]["K_pos"])
    assert tot["nR"] == full.pen["nR"] and tot["nK"] == full.pen["nK"]
    assert tot["itemsK"] == len(full.pen["K_item"]) and tot["itemsR"] == len(full.pen["R_item"])
    assert tot["destP"] == sum(rd["n_dest"] for rd in full.penP[0]["rounds"])
    allpos = np.concatenate(seenK)
    assert np.array_equal(np.sort(allpos[allpos >= 0]), np.sort(full.pen["K_pos"][full.pen["K_pos"] >= 0]))


def test_penalty_rounds_have_disjoint_destinations():
    pr, _ = cases.wingbox_small()
    S = Symbolic(pr, opt_field=[1], shopt_surf_inds=[list(range(len(pr["patches"])))])
    pp = S.penP[0]
    assert len(pp["rounds"]) >= 2                      # T-junction intersections share destinations
    seen_all = []
    for rd in pp["rounds"]:
        pos = rd["pos"][rd["pos"] >= 0]
        assert len(np.unique(pos)) == len(pos)         # inside a round every CSR slot has one owner thread
        seen_all.append(pos)
    assert len(np.unique(np.concatenate(seen_all))) <= pp["nnz"]


def test_own_patch_schwarz_blocks_equal_the_unsharded_ones():
    pr, _ = cases.wingbox_small()
    S = Symbolic(pr)
    world = 3
    owner = _owner(pr, world)
    full = SchwarzSetup(S, layers=2, sub=(8, 16))
    for rank in range(world):
        mine = SchwarzSetup(S, layers=2, sub=(8, 16), own_patches=(owner == rank))
        ref = [b for b in full.blocks if owner[b["patch"]] == rank]
        assert len(mine.blocks) == len(ref) > 0
        for a, b in zip(mine.blocks, ref):
            assert np.array_equal(a["nodes"], b["nodes"]) and np.array_equal(a["mbj"], b["mbj"]) and np.array_equal(a["glob"], b["glob"])


def test_subdomain_and_coarse_rules():
    S = Symbolic(problems.cylinder(n_el=201))
    picks = [SchwarzSetup.choose_subdomains(S.patches, w) for w in (1, 2, 4, 8)]
    assert picks == [(24, 96), (24, 48), (24, 24), (12, 24)]       # bandwidth bound on one GPU, chain bound on many
    pr = problems.cylinder(n_el=201)
    r1, r8 = coarse.coarsening_ratio(pr, 24000), coarse.coarsening_ratio(pr, 48000)
    assert 7.0 <= r8 <= r1 < 8.0
    cpr, P = coarse.build(pr, ratio=r1)
    assert P.shape[0] == S.N and 20000 < P.shape[1] <= 24000
    # long thin patches keep their aspect ratio: the spars of the wing box get few coarse elements across their height
    wb = problems.wingbox(target_dofs=2.0e5)
    r = coarse.coarsening_ratio(wb, 24000)
    cwb, Pw = coarse.build(wb, ratio=r)
    spar = cwb["patches"][20]
    neu, nev = len(np.unique(spar["knots"][0])) - 1, len(np.unique(spar["knots"][1])) - 1
    assert neu > 3 * nev and nev >= 8 and Pw.shape[1] <= 24000


def test_coarse_design_restriction_and_refresh_trigger():
    """refresh_coarse(): the coarse control net of the CURRENT design is the least-squares restriction of the fine net --
    the initial design gives back the coarse net `build` made, a design that lies in the coarse space is reproduced
    exactly, thickness becomes the patch mean; the automatic trigger fires only for a changed design and a clear growth."""
    pr = problems.cylinder(n_el=24)
    S = Symbolic(pr)
    cpr, P = coarse.build(pr, ratio=7.0)
    fine = [(Pp.n_u, Pp.n_v, Pp.cp_off, Pp.th_off, Pp.nth) for Pp in S.patches]
    cpc, thc = coarse.restrict_design(cpr, S.cp0, S.theta0, fine)
    assert np.abs(cpc - np.concatenate([p["cp"] for p in cpr["patches"]])).max() < 1e-10
    assert np.allclose(thc, [p["thickness"]["values"] for p in cpr["patches"]])
    # a perturbation taken from the coarse space is recovered exactly: restrict(prolong(dXc)) = dXc
    Sc = Symbolic(cpr)
    rng = np.random.default_rng(0)
    dXc = rng.standard_normal((Sc.n_scalar, 4)) * 1e-3
    dX = np.zeros_like(S.cp0)
    for Pf, Pc in zip(S.patches, Sc.patches):
        blk = P[Pf.dof_off:Pf.dof_off + Pf.ncp, Pc.dof_off:Pc.dof_off + Pc.ncp]
        dX[Pf.cp_off:Pf.cp_off + Pf.ncp] = blk @ dXc[Pc.cp_off:Pc.cp_off + Pc.ncp]
    cpc2, _ = coarse.restrict_design(cpr, S.cp0 + dX, 1.1 * S.theta0, fine)
    assert np.abs(cpc2 - (cpc + dXc)).max() < 1e-10
    assert coarse.refresh_due(120, 60, True) and not coarse.refresh_due(120, 60, False)
    assert not coarse.refresh_due(100, 60, True) and not coarse.refresh_due(500, None, True)
