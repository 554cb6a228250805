#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02c_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-op-profile --no-extras > gpurun_out/r02c_bench_under_ncu.log 2>&1
echo "launch list rc=$?"; wc -l gpurun_out/r02c_launches.csv
