"""FP64 peak of this B200 (DFMA chains and DMMA m8n8k4) -> gpurun_out/fp64_peak.json (copied to profiles/)."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from goldfish_b200 import _capi as capi
lib = capi.load()
sms = torch.cuda.get_device_properties(0).multi_processor_count
out = torch.zeros(sms * 8 * 256, dtype=torch.float64, device="cuda")
res = {"gpu": torch.cuda.get_device_name(0), "sms": sms}
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for mode, name in ((0, "dfma"), (1, "dmma_m8n8k4")):
    best = 0.0
    for ctas in (4, 8):
        for rep in range(6):
            fl = C.c_double(0)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            capi.check(lib.gf_peak_fp64(mode, sms * ctas, 20000, C.c_void_p(out.data_ptr()), C.byref(fl), st), "peak")
            b.record(); torch.cuda.synchronize()
            best = max(best, fl.value / (a.elapsed_time(b) * 1e-3) / 1e12)
    res[name + "_tflops"] = best
print(json.dumps(res))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/fp64_peak.json", "w"))
