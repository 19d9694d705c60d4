"""Patch -> GPU assignment for the patch-sharded multi-GPU path.

Shell terms are strictly per patch (the tangent is block diagonal before
coupling, /root/reference/GOLDFISH/nonmatching_opt.py:813-823), so whole
patches are the unit of distribution: greedy longest-processing-time packing
on the element count.  Deterministic.
"""
import numpy as np


def lpt_partition(weights, nparts):
    """owner[i] in [0, nparts): heaviest item first onto the lightest part
    (ties: lower part index, lower item index)."""
    w = np.asarray(weights, dtype=np.float64)
    order = np.lexsort((np.arange(len(w)), -w))
    load = np.zeros(nparts)
    owner = np.zeros(len(w), dtype=np.int32)
    for i in order:
        p = int(np.argmin(load))
        owner[i] = p
        load[p] += w[i]
    return owner


def filter_ragged(ptr, items, keep):
    """Keep the segments `keep` of a CSR-like (ptr, items) list."""
    ptr = np.asarray(ptr, dtype=np.int64)
    keep = np.asarray(keep, dtype=bool)
    lens = np.diff(ptr)[keep]
    new_ptr = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=new_ptr[1:])
    if len(lens) == 0:
        return new_ptr, items[:0]
    starts = ptr[:-1][keep]
    idx = np.repeat(starts - new_ptr[:-1], lens) + np.arange(new_ptr[-1])
    return new_ptr, items[idx]


def shard_symbolic(S, owner, rank):
    """What one rank keeps of the (replicated) symbolic phase:
    elements of its own patches (per colour), the coupling gather destinations
    whose ROW belongs to an own patch, and its contiguous row ranges."""
    own = np.asarray(owner) == rank
    out = {}
    keep = own[S.elem_patch[S.color_elem]]
    color_of = np.repeat(np.arange(S.num_colors), np.diff(S.color_ptr))
    out["color_elem"] = S.color_elem[keep]
    out["color_ptr"] = np.searchsorted(color_of[keep], np.arange(S.num_colors + 1)).astype(np.int32)
    rng = []
    for P in S.patches:
        if not own[P.index]:
            continue
        if rng and rng[-1][1] == P.dof_off:
            rng[-1][1] = P.dof_off + 3 * P.ncp
        else:
            rng.append([P.dof_off, P.dof_off + 3 * P.ncp])
    out["own_ranges"] = np.asarray(rng, dtype=np.int64).reshape(-1, 2)
    pen = dict(S.pen)
    if pen["n_eval"] > 0:
        kR = own[pen["R_dest_patch"]]
        pen["R_ptr"], pen["R_item"] = filter_ragged(pen["R_ptr"], pen["R_item"], kR)
        pen["R_row"] = np.ascontiguousarray(pen["R_row"][kR]); pen["nR"] = int(kR.sum())
        kK = own[pen["K_dest_patch"]]
        pen["K_ptr"], pen["K_item"] = filter_ragged(pen["K_ptr"], pen["K_item"], kK)
        pen["K_pos"] = np.ascontiguousarray(pen["K_pos"][kK]); pen["nK"] = int(kK.sum())
    out["pen"] = pen
    penP = []
    done = {}                    # the fields of one patch list share their gather lists: filter them once
    for pp in getattr(S, "penP", []):
        pp = dict(pp)
        rid = id(pp.get("rounds"))
        if rid in done:
            pp["rounds"], pp["n_dest_own"] = done[rid]
            penP.append(pp)
            continue
        rounds = []
        for rd in pp.get("rounds", []):
            rd = dict(rd)
            kP = own[rd["dest_patch"]]
            ptr2, ev2 = filter_ragged(rd["ptr"], rd["item_eval"], kP)
            _, code2 = filter_ragged(rd["ptr"], rd["item_code"], kP)
            rd.update(ptr=ptr2, item_eval=ev2, item_code=code2, pos=np.ascontiguousarray(rd["pos"][kP]), n_dest=int(kP.sum()))
            rounds.append(rd)
        pp["rounds"] = rounds
        pp["n_dest_own"] = int(sum(rd["n_dest"] for rd in rounds))
        done[rid] = (rounds, pp["n_dest_own"])
        penP.append(pp)
    out["penP"] = penP
    return out
