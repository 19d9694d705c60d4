"""CPU, world_size = 2 over gloo: the patch-sharding logic of the multi-GPU path
(goldfish_b200/partition.py).  Each rank shards the replicated symbolic phase;
together the shards must cover every element / coupling destination exactly once."""
import os
import sys
import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import cases
    from goldfish_b200.symbolic import Symbolic
    from goldfish_b200.partition import lpt_partition, shard_symbolic
    pr, kw = cases.slr_small()
    S = Symbolic(pr, **kw)
    owner = lpt_partition([P.nel for P in S.patches], world)
    sh = shard_symbolic(S, owner, rank)
    mine = dict(elems=np.sort(sh["color_elem"]).tolist(),
                nR=sh["pen"]["nR"], nK=sh["pen"]["nK"], items_K=int(len(sh["pen"]["K_item"])),
                rows=sh["own_ranges"].tolist(), owner=owner.tolist(),
                nP=[pp.get("n_dest_own", 0) for pp in sh["penP"]])
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    # colour lists stay conflict free and sorted by colour
    cp = sh["color_ptr"]
    assert cp[0] == 0 and cp[-1] == len(sh["color_elem"]) and np.all(np.diff(cp) >= 0)
    if rank == 0:
        allel = sorted(sum((g["elems"] for g in gathered), []))
        ok = (allel == list(range(S.num_elements))
              and sum(g["nR"] for g in gathered) == S.pen["nR"]
              and sum(g["nK"] for g in gathered) == S.pen["nK"]
              and sum(g["items_K"] for g in gathered) == len(S.pen["K_item"])
              and all(g["owner"] == gathered[0]["owner"] for g in gathered)
              and [sum(g["nP"][i] for g in gathered) for i in range(len(S.penP))] == [sum(rd["n_dest"] for rd in pp["rounds"]) for pp in S.penP])
        covered = np.zeros(S.N, dtype=int)
        for g in gathered:
            for b0, b1 in g["rows"]:
                covered[b0:b1] += 1
        q.put(bool(ok and np.all(covered == 1)))
    dist.barrier()
    dist.destroy_process_group()


def test_sharding_covers_everything_once():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


def test_lpt_is_balanced_and_deterministic():
    from goldfish_b200.partition import lpt_partition, filter_ragged
    w = [40000 + 400 * i for i in range(8)]
    for n in (1, 2, 4, 8):
        o = lpt_partition(w, n)
        loads = np.bincount(o, weights=w, minlength=n)
        assert loads.max() <= 1.1 * loads.mean() and np.array_equal(o, lpt_partition(w, n))
    ptr, items = filter_ragged(np.array([0, 2, 2, 5]), np.arange(5), np.array([True, False, True]))
    assert ptr.tolist() == [0, 2, 5] and items.tolist() == [0, 1, 2, 3, 4]
    ptr, items = filter_ragged(np.array([0, 2, 2, 5]), np.arange(5), np.array([False, True, False]))
    assert ptr.tolist() == [0, 0] and len(items) == 0


def _pcg_worker(rank, world, port, q):
    """The exchange steps of the sharded Krylov solve, emulated with scipy over gloo:
    each rank multiplies its OWN rows and solves its OWN Schwarz blocks, both results are summed with an
    all-reduce, every rank then holds identical full vectors and computes the dot products redundantly."""
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    import scipy.sparse.linalg as spla
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from goldfish_b200 import problems, _capi as capi
    from goldfish_b200.partition import lpt_partition, shard_symbolic
    from goldfish_b200.schwarz import SchwarzSetup
    from oracle.cpu_port import CpuModel
    pr = problems.cylinder(n_el=6)
    cm = CpuModel(pr)
    S = cm.S
    cm.set_u(np.zeros(S.N))
    cm.shell(capi.GF_OUT_R | capi.GF_OUT_K)
    K = cm.K_matrix().tocsr(); b = -cm.residual()
    owner = lpt_partition([P.nel for P in S.patches], world)
    sh = shard_symbolic(S, owner, rank)
    own = np.zeros(S.N, dtype=bool)
    for b0, b1 in sh["own_ranges"]:
        own[b0:b1] = True
    K_own = K[np.nonzero(own)[0]]
    SW = SchwarzSetup(S, layers=2, sub=8)
    n_all = len(SW.blocks)
    SW.keep_blocks([owner[bl["patch"]] == rank for bl in SW.blocks])
    lus = []
    for bl in SW.blocks:
        g = bl["glob"][bl["glob"] >= 0]
        lus.append((g, spla.splu(K[g][:, g].tocsc())))

    def allreduce(v):
        t = torch.from_numpy(v); dist.all_reduce(t); return t.numpy()

    def matvec(p):
        y = np.zeros(S.N); y[own] = K_own @ p
        return allreduce(y)

    def precond(r):
        z = np.zeros(S.N)
        for g, lu in lus:
            z[g] += lu.solve(r[g])
        return allreduce(z)

    x = np.zeros(S.N); r = b.copy(); z = precond(r); p = z.copy(); rz = r @ z; bn = np.linalg.norm(b); its = 0
    for its in range(1, 401):
        Ap = matvec(p); a = rz / (p @ Ap); x += a * p; r -= a * Ap
        if np.linalg.norm(r) < 1e-10 * bn:
            break
        z = precond(r); rz2 = r @ z; p = z + (rz2 / rz) * p; rz = rz2
    counts = [None] * world
    dist.all_gather_object(counts, (its, len(SW.blocks), float(np.linalg.norm(x))))
    if rank == 0:
        xe = spla.splu(K.tocsc()).solve(b)
        ok = (np.linalg.norm(x - xe) < 1e-6 * np.linalg.norm(xe)                 # the sharded solve is the LU solution
              and len({c[0] for c in counts}) == 1 and len({c[2] for c in counts}) == 1   # identical on every rank
              and sum(c[1] for c in counts) == n_all and its < 200)               # every block solved exactly once
        q.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_krylov_exchange_over_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_pcg_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True


def test_forty_patch_topology_shards_over_eight_ranks():
    """BASELINE configs[3] stand-in (bench.py --topology 8x5: 40 patches, 72 intersections): the symbolic phase,
    the Schwarz blocks and the coarse level set up, and the 8-rank sharding covers every element, coupling
    destination, row and block exactly once (several patches per rank, ragged loads)."""
    import bench
    from goldfish_b200.symbolic import Symbolic
    from goldfish_b200.partition import lpt_partition, shard_symbolic
    from goldfish_b200.schwarz import SchwarzSetup
    from goldfish_b200 import coarse
    pr, kw = bench.workload(5, 8, 5)
    S = Symbolic(pr, **kw)
    assert len(S.patches) == 40 and len(pr["interfaces"]) == 72
    owner = lpt_partition([P.nel for P in S.patches], 8)
    loads = np.bincount(owner, weights=[P.nel for P in S.patches], minlength=8)
    assert loads.min() > 0 and loads.max() <= 1.25 * loads.mean()
    elems, covered, nR, nK = [], np.zeros(S.N, dtype=int), 0, 0
    for rank in range(8):
        sh = shard_symbolic(S, owner, rank)
        elems.append(np.asarray(sh["color_elem"]))
        nR += sh["pen"]["nR"]; nK += sh["pen"]["nK"]
        for b0, b1 in sh["own_ranges"]:
            covered[b0:b1] += 1
    assert np.array_equal(np.sort(np.concatenate(elems)), np.arange(S.num_elements))
    assert np.all(covered == 1) and nR == S.pen["nR"] and nK == S.pen["nK"]
    SW = SchwarzSetup(S)
    per_rank = [sum(owner[b["patch"]] == r for b in SW.blocks) for r in range(8)]
    assert sum(per_rank) == len(SW.blocks) == 40 and min(per_rank) >= 1
    cpr, P = coarse.build(pr, nc=3)
    assert P.shape == (S.N, Symbolic(cpr).N)
