#!/usr/bin/env bash
# Round-end profiling pass (run under gpurun on ONE GPU): launch list of the bench command + `--set full`
# captures of the dominant kernels at config-2 size.  Summarise afterwards with scripts/summarize_profiles.py.
set -u
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r01_launches_final.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-op-profile > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_umma|k_layernorm_act_bwd" --launch-skip 10 --launch-count 10 \
    -o gpurun_out/r01_step_cfg2 -f python scripts/profile_step.py > gpurun_out/ncu_step.log 2>&1
PAGES=512 REPS=2 ncu --set full --clock-control none --import-source on -k regex:k_spmm_paged_pk --launch-skip 2 --launch-count 2 \
    -o gpurun_out/r01_spmm_cfg2 -f python scripts/profile_spmm.py > gpurun_out/ncu_spmm.log 2>&1
tail -n 2 gpurun_out/ncu_step.log; tail -n 2 gpurun_out/ncu_spmm.log
