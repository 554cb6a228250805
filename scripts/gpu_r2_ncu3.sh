#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
# launch list (time only) of the bench command
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02b_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-op-profile --no-extras > gpurun_out/r02b_bench_under_ncu.log 2>&1
echo "launch list rc=$?"
# full metric set for one warm eager step
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off \
    -o gpurun_out/r02b_step_full -f python scripts/profile_step.py > gpurun_out/r02b_ncu_step.log 2>&1
echo "ncu full rc=$?"; tail -n 2 gpurun_out/r02b_ncu_step.log; ls -la gpurun_out/r02b_step_full.ncu-rep
