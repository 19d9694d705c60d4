#!/bin/bash
mkdir -p gpurun_out
{ nproc; free -g; nvidia-smi -L; } > gpurun_out/box2.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --topology wingbox --dofs 1e6 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_wingbox_1M_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/bench_wingbox_1M_2gpu.log
tail -c 4000 gpurun_out/bench_wingbox_1M_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_c3_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/bench_c3_2gpu.log
tail -c 3000 gpurun_out/bench_c3_2gpu.log
