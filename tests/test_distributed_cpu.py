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
                nP=[pp.get("n_dest", 0) for pp in sh["penP"]])
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
              and [sum(g["nP"][i] for g in gathered) for i in range(len(S.penP))] == [pp["n_dest"] for pp in S.penP])
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
