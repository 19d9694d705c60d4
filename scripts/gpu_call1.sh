#!/bin/bash
# round-2 first GPU call: environment facts, GPU tests, SpMV node-wise check, FP64 peak, bench with parity
mkdir -p gpurun_out
{ nproc; free -g; nvidia-smi -L; nvidia-smi --query-gpu=memory.total --format=csv; } > gpurun_out/box.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python scripts/gpu_spmv_node_check.py > gpurun_out/spmv_node.log 2>&1
timeout 300 python scripts/gpu_fp64_peak.py > gpurun_out/fp64_peak.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_c3.log
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/spmv_node.log gpurun_out/fp64_peak.log; tail -c 3000 gpurun_out/bench_c3.log
