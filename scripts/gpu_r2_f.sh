#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 90 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "bwd_weight or class_layer" > gpurun_out/r2f_pytest_dw.log 2>&1
rc=$?; echo "pytest dw rc=$rc"; tail -n 3 gpurun_out/r2f_pytest_dw.log
[[ $rc -ne 0 ]] && exit 0
export GTE_LIB=$PWD/gnn_tableextraction_b200/libgte_b200_exp.so
timeout 60 python scripts/dw_trace.py 2>&1 | tail -4
for dbg in 1 33; do echo "== pair GTE_UMMA_DBG=$dbg"; GTE_UMMA_DBG=$dbg GTE_UMMA_PAIR=1 timeout 60 python scripts/umma_trace.py 2>&1 | grep -E "event ms|tile 2" | cut -c 1-30,150-260; done
for dbg in 1 33; do echo "== single GTE_UMMA_DBG=$dbg"; GTE_UMMA_DBG=$dbg GTE_UMMA_PAIR=0 timeout 60 python scripts/umma_trace.py 2>&1 | grep -E "event ms|tile 2" | cut -c 1-30,150-260; done
