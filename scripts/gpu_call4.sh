#!/bin/bash
# 2-GPU call: GPU tests (incl. the 2-rank parity test), sharded-vs-single check, sub-domain experiments, 1- and 2-GPU bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/gpu_dist_check.py 48 > gpurun_out/dist2_check.log 2>&1; echo "dist rc=$?" >> gpurun_out/dist2_check.log
tail -6 gpurun_out/dist2_check.log
for sub in 24,96 24,48 24,24 12,24; do timeout 300 python scripts/exp_sweeps.py 201 single $sub 2 28 2>&1 | tail -1 >> gpurun_out/r2_sweep_tuning_subdomains.jsonl; done
cat gpurun_out/r2_sweep_tuning_subdomains.jsonl
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_c3_1gpu.log 2>&1; echo "bench rc=$?" >> gpurun_out/bench_c3_1gpu.log
tail -c 5000 gpurun_out/bench_c3_1gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_c3_2gpu.log 2>&1; echo "bench2 rc=$?" >> gpurun_out/bench_c3_2gpu.log
tail -c 4000 gpurun_out/bench_c3_2gpu.log
