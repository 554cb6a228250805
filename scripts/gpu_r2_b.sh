#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 120 ./scripts/micro/mma_rate > gpurun_out/r2_mma_rate.log 2>&1; echo "mma_rate rc=$?"
export GTE_LIB=$PWD/gnn_tableextraction_b200/libgte_b200_exp.so
for dbg in 1 3 5 9 17 31; do
  GTE_UMMA_DBG=$dbg GTE_UMMA_PAIR=1 timeout 180 python scripts/umma_trace.py > gpurun_out/r2b_trace_pair_dbg$dbg.log 2>&1
  echo "== dbg $dbg"; grep -E "event ms|tile 2" gpurun_out/r2b_trace_pair_dbg$dbg.log
done
