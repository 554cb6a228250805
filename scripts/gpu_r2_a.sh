#!/usr/bin/env bash
# Round-2 GPU pass A: first contact of the CTA-pair projection kernel, then the full GPU suite, bench and role traces.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_gpu.csv 2>&1
timeout 240 python scripts/pair_smoke.py > gpurun_out/r2a_pair_smoke.log 2>&1
rc=$?
echo "pair_smoke rc=$rc"; tail -n 15 gpurun_out/r2a_pair_smoke.log
if [[ $rc -eq 0 ]]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1
  echo "pytest rc=$?"; tail -n 8 gpurun_out/r2a_pytest.log
  timeout 600 python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
  echo "bench rc=$?"; cut -c 1-600 gpurun_out/r2a_bench.json
  export GTE_LIB=$PWD/gnn_tableextraction_b200/libgte_b200_exp.so
  GTE_UMMA_DBG=1 GTE_UMMA_PAIR=1 timeout 180 python scripts/umma_trace.py > gpurun_out/r2a_trace_pair.log 2>&1
  GTE_UMMA_DBG=1 GTE_UMMA_PAIR=0 timeout 180 python scripts/umma_trace.py > gpurun_out/r2a_trace_single.log 2>&1
  tail -n 12 gpurun_out/r2a_trace_pair.log
else
  # diagnose with the single-CTA kernels only
  GTE_TEST_NO_PAIR=1 timeout 900 python -m pytest tests -m gpu -x -q -k "not pair" > gpurun_out/r2a_pytest_single.log 2>&1
  echo "pytest(single) rc=$?"; tail -n 8 gpurun_out/r2a_pytest_single.log
fi
