#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export GTE_LIB=$PWD/gnn_tableextraction_b200/libgte_b200_exp.so
for dbg in 1 17 65 67 3 69; do
  GTE_UMMA_DBG=$dbg GTE_UMMA_PAIR=1 timeout 180 python scripts/umma_trace.py > gpurun_out/r2d_trace_pair_dbg$dbg.log 2>&1
  echo "== dbg $dbg"; grep -E "event ms|tile 2|tile period" gpurun_out/r2d_trace_pair_dbg$dbg.log | cut -c 1-30,56-140,150-400
done
timeout 120 ./scripts/micro/mma_rate > gpurun_out/r2_mma_rate2.log 2>&1; grep -E "N=224" gpurun_out/r2_mma_rate2.log
