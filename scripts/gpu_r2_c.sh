#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 240 python scripts/pair_smoke.py > gpurun_out/r2c_pair_smoke.log 2>&1
rc=$?
echo "pair_smoke rc=$rc"; tail -n 14 gpurun_out/r2c_pair_smoke.log
if [[ $rc -eq 0 ]]; then
  export GTE_LIB=$PWD/gnn_tableextraction_b200/libgte_b200_exp.so
  GTE_UMMA_DBG=1 GTE_UMMA_PAIR=1 timeout 180 python scripts/umma_trace.py > gpurun_out/r2c_trace_pair.log 2>&1
  grep -E "event ms|tile 2|span" gpurun_out/r2c_trace_pair.log | cut -c 1-60,150-400
  unset GTE_LIB
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1
  echo "pytest rc=$?"; tail -n 8 gpurun_out/r2c_pytest.log
  timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err
  echo "bench rc=$?"; cut -c 1-300 gpurun_out/r2c_bench.json
fi

