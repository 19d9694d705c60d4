"""ORACLE / CPU BASELINE (test + measurement infrastructure, never the product path).

Fill-reducing ordering and symbolic phase of the CPU sparse direct solver that stands in for MUMPS
(`solve_nonmatching_mat(..., solver='direct')`, /root/reference/GOLDFISH/utils/opt_utils.py:176,204:
PETSc PCLU + MUMPS = multifrontal LU on a nested-dissection ordering).  MUMPS/METIS are not in this image, so
the ordering is restated here: recursive coordinate bisection of the control-point graph with vertex
separators, which is what a graph partitioner finds on these tensor-product shells.

  nested_dissection(G, X, leaf)  -> list of fronts in post-order: (own nodes, parent index)
  symbolic(fronts, G)            -> per-front update (border) node lists
"""
import numpy as np


def _rows(G, nodes):
    """(neighbour list, owner index into `nodes`) of the CSR rows `nodes`."""
    ip = G.indptr
    cnt = (ip[nodes + 1] - ip[nodes]).astype(np.int64)
    tot = int(cnt.sum())
    starts = np.cumsum(cnt) - cnt
    idx = np.arange(tot, dtype=np.int64) - np.repeat(starts, cnt) + np.repeat(ip[nodes].astype(np.int64), cnt)
    return G.indices[idx], np.repeat(np.arange(len(nodes)), cnt)


def nested_dissection(G, X, leaf=64):
    """G: scipy CSR adjacency of the scalar control-point graph (symmetric pattern, any values);
    X: (n, 3) coordinates.  Returns (fronts, parent): fronts[t] = node array owned by tree node t
    (leaf interior or separator), listed in post-order (children before parents); parent[t] = index or -1."""
    n = G.shape[0]
    label = np.zeros(n, dtype=np.int64)          # scratch: 0 = outside the current set, 1 = A, 2 = B
    fronts, parent = [], []

    def rec(nodes):
        if len(nodes) <= leaf:
            fronts.append(nodes); parent.append(-1)
            return len(fronts) - 1
        Xs = X[nodes]
        ax = int(np.argmax(Xs.max(0) - Xs.min(0)))
        order = np.argsort(Xs[:, ax], kind="stable")
        half = len(nodes) // 2
        A, B = nodes[order[:half]], nodes[order[half:]]
        label[A] = 1; label[B] = 2
        nb, own = _rows(G, A)
        touches = np.zeros(len(A), dtype=bool)
        touches[own[label[nb] == 2]] = True
        label[A] = 0; label[B] = 0
        S, A2 = A[touches], A[~touches]
        if len(A2) == 0 or len(S) == 0:          # disconnected halves or degenerate split
            if len(S) == 0:
                ia = rec(A); ib = rec(B)
                fronts.append(np.zeros(0, dtype=nodes.dtype)); parent.append(-1)
                me = len(fronts) - 1
                parent[ia] = me; parent[ib] = me
                return me
            fronts.append(nodes); parent.append(-1)
            return len(fronts) - 1
        ia = rec(A2); ib = rec(B)
        fronts.append(S); parent.append(-1)
        me = len(fronts) - 1
        parent[ia] = me; parent[ib] = me
        return me

    import sys
    sys.setrecursionlimit(max(10000, sys.getrecursionlimit()))
    # connected components are handled by the degenerate-split branch
    rec(np.arange(n, dtype=np.int64))
    return fronts, np.asarray(parent, dtype=np.int64)


def symbolic(fronts, parent, G):
    """Border (update) node set of every front: the later-eliminated nodes its Schur complement touches.
    Returns (pos, border): pos[node] = elimination position, border[t] = sorted positions (> own range)."""
    n = G.shape[0]
    perm = np.concatenate(fronts)
    assert len(perm) == n and len(np.unique(perm)) == n
    pos = np.empty(n, dtype=np.int64); pos[perm] = np.arange(n)
    first = np.cumsum([0] + [len(f) for f in fronts])
    nt = len(fronts)
    border = [None] * nt
    children = [[] for _ in range(nt)]
    for t, p in enumerate(parent):
        if p >= 0:
            children[p].append(t)
    for t in range(nt):
        own_hi = first[t + 1]
        parts = []
        if len(fronts[t]):
            nb, _ = _rows(G, fronts[t])
            pn = pos[nb]
            parts.append(pn[pn >= own_hi])
        for c in children[t]:
            bc = border[c]
            parts.append(bc[bc >= own_hi])
        border[t] = np.unique(np.concatenate(parts)) if parts else np.zeros(0, dtype=np.int64)
    return pos, first, border, children
