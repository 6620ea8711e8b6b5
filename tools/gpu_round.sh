#!/bin/bash
# One GPU visit: tests, bench, ncu launch list, ncu full capture of the dominant kernel.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -15
echo "=== bench"
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench.json
echo "=== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
echo "=== ncu full KA"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 20 -c 2 -o gpurun_out/prof_ka python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out
