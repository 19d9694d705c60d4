#!/bin/bash
# 8 x B200: BASELINE configs[2] (C3, strong scaling point) and configs[3] (C4: 40-patch wing box, 10 M dofs)
mkdir -p gpurun_out
{ nproc; free -g; nvidia-smi -L; nvidia-smi topo -m; } > gpurun_out/box8.txt 2>&1
export NCCL_DEBUG=WARN
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 2 --no-cpu-baseline > gpurun_out/bench_c3_8gpu.log 2>&1; echo "rc=$?" >> gpurun_out/bench_c3_8gpu.log
tail -c 3500 gpurun_out/bench_c3_8gpu.log
( while true; do free -g | sed -n 2p; nvidia-smi --query-gpu=memory.used --format=csv,noheader | tr '\n' ' '; echo; sleep 20; done ) > gpurun_out/c4_mem_trace.txt 2>&1 &
MON=$!
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --topology wingbox --dofs 1e7 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_c4_8gpu.log 2>&1; echo "rc=$?" >> gpurun_out/bench_c4_8gpu.log
kill $MON
tail -c 6000 gpurun_out/bench_c4_8gpu.log
tail -5 gpurun_out/c4_mem_trace.txt
