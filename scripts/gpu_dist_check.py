"""torchrun --nproc-per-node N scripts/gpu_dist_check.py : the patch-sharded path against the
single-GPU path on the same problem (rank 0 also runs the unsharded model)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
from goldfish_b200.device_model import DeviceModel

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
n_el = int(sys.argv[1]) if len(sys.argv) > 1 else 24
pr, kw = bench.workload(n_el)
dm = DeviceModel(pr, **kw)                      # sharded
step = bench.Step(dm)
step(); torch.cuda.synchronize()
t0 = time.time(); step(); torch.cuda.synchronize(); t1 = time.time()
if rank == 0:
    ref = DeviceModel(pr, distributed=False, **kw)
    rstep = bench.Step(ref); rstep(); torch.cuda.synchronize()
    t2 = time.time(); rstep(); torch.cuda.synchronize(); t3 = time.time()
    rel = lambda a, b: float(torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b))
    print("world", world, "N", dm.sym.N, "own patches", dm.own_patches, "its", step.info, rstep.info)
    print("u", rel(dm.u, ref.u), "lam", rel(step.lam, rstep.lam), "gT", rel(step.gT, rstep.gT),
          "gP", [rel(a, b) for a, b in zip(step.gP, rstep.gP)], "W,V", dm.wv_sum.tolist(), ref.wv_sum.tolist())
    print("step time sharded %.3f s, single %.3f s" % (t1 - t0, t3 - t2))
    import json
    print("RESULT " + json.dumps({"u": rel(dm.u, ref.u), "lam": rel(step.lam, rstep.lam), "gT": rel(step.gT, rstep.gT),
                                  "gP": max(rel(a, b) for a, b in zip(step.gP, rstep.gP)),
                                  "its_sharded": step.info["krylov_its"], "its_single": rstep.info["krylov_its"]}))
dist.barrier()
dist.destroy_process_group()
