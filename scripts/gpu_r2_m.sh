#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export GTE_LIB=$PWD/gnn_tableextraction_b200/libgte_b200_exp.so
for d in 1 3 5 9 15; do
  echo "##### GTE_UMMA_DBG=$d"
  GTE_UMMA_DBG=$d timeout 120 python scripts/umma_trace.py 2>&1 | grep -v "^ cta 0 tile [234]"
done
