#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_umma_gemm_pair|k_umma_dw" --launch-skip 12 --launch-count 8 \
    -o gpurun_out/r02_umma_step -f python scripts/profile_step.py > gpurun_out/r2_ncu_step.log 2>&1
echo "ncu rc=$?"; tail -n 3 gpurun_out/r2_ncu_step.log; ls -la gpurun_out/r02_umma_step.ncu-rep
