"""Round-2 candidate: node-wise tangent product (gf_spmv_node) against the default row-wise gf_spmv on the bench
workload -- bitwise comparison of y and CUDA-event timing of both.  usage: gpu_spmv_node_check.py [n_el]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from goldfish_b200.device_model import DeviceModel

n_el = int(sys.argv[1]) if len(sys.argv) > 1 else 201
pr, kw = bench.workload(n_el)
dm = DeviceModel(pr, precond="jacobi", **kw)
dm.assemble(tangent=True)
x = torch.randn(dm.sym.N, dtype=torch.float64, device="cuda")
y0 = torch.empty_like(x); y1 = torch.empty_like(x)
dm.spmv(dm.K, x, y0); dm.spmv_node(x, y1)
torch.cuda.synchronize()
print("bitwise equal:", bool(torch.equal(y0, y1)), " max |diff|:", float((y0 - y1).abs().max()))
flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float64, device="cuda")
b = 12 * dm.K.nnz + 24 * dm.sym.N + 8
for name, fn in (("gf_spmv", lambda: dm.spmv(dm.K, x, y0)), ("gf_spmv_node", lambda: dm.spmv_node(x, y1))):
    ms = bench.time_kernel(torch, fn, 20, flush)
    print("%-13s %.4f ms  %.0f GB/s on the 12 B/nnz algorithmic count" % (name, ms, b / ms / 1e6))
