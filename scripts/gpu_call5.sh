#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --topology wingbox --dofs 1e6 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_wingbox_1M_1gpu.log 2>&1; echo "rc=$?" >> gpurun_out/bench_wingbox_1M_1gpu.log
tail -c 3500 gpurun_out/bench_wingbox_1M_1gpu.log
# ncu: launch list of one profiler-friendly pass at C3, then full captures of the hot kernels
timeout 600 python scripts/profile_step.py 201 1 > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches_c3.csv python scripts/profile_step.py 201 1 > gpurun_out/ncu_launches.log 2>&1
for k in k_spmv_node k_sw_solve1 k_shell_k2 k_shell_p2; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -c 2 -o gpurun_out/r2_$k -f python scripts/profile_step.py 201 1 > gpurun_out/ncu_$k.log 2>&1
  echo "ncu $k rc=$?"
done
ls -la gpurun_out/*.ncu-rep
