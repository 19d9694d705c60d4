#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_c3_1gpu.log 2>&1; echo "rc=$?" >> gpurun_out/bench_c3_1gpu.log
tail -c 3000 gpurun_out/bench_c3_1gpu.log
timeout 1200 python bench.py --topology wingbox --dofs 1e6 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_wingbox_1M_1gpu.log 2>&1; echo "rc=$?" >> gpurun_out/bench_wingbox_1M_1gpu.log
tail -c 3000 gpurun_out/bench_wingbox_1M_1gpu.log
