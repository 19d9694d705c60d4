import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
code = r"""
import sys, os, torch, numpy as np
sys.path.insert(0, os.getcwd())
import bench
from goldfish_b200.device_model import DeviceModel
pr, kw = bench.workload(int(os.environ.get('NE','128')))
dm = DeviceModel(pr, precond='jacobi', **kw); dm.assemble(tangent=True)
x = torch.randn(dm.sym.N, dtype=torch.float64, device='cuda'); y = torch.empty_like(x)
flush = torch.zeros(64*1024*1024, dtype=torch.float64, device='cuda')
ms = bench.time_kernel(torch, lambda: dm.spmv(dm.K, x, y), 20, flush)
b = 12*dm.K.nnz + 24*dm.sym.N + 8
print('variant', os.environ.get('GF_SPMV_VARIANT'), 'grid', os.environ.get('GF_SPMV_GRID'), 'ms %.4f' % ms, 'GB/s %.0f' % (b/ms/1e6))
"""
for v in ("0", "1", "2"):
    for g in ("888", "1024", "1776", "2368", "4096"):
        env = dict(os.environ, GF_SPMV_VARIANT=v, GF_SPMV_GRID=g)
        print(subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True).stdout.strip(), flush=True)
