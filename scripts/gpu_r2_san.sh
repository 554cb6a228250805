#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
export PYTHONDONTWRITEBYTECODE=1
# memcheck over the whole GPU suite
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 --launch-timeout 0 \
  python -m pytest tests -m gpu -q -x > gpurun_out/r2_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2_memcheck.log | tail -n 4
# initcheck (reads of uninitialised device memory) over the model-level tests and the combined-operand kernels
timeout 1500 compute-sanitizer --tool initcheck --error-exitcode 3 --launch-timeout 0 \
  python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -m gpu -q -x -k "comb or golden or dropout or captured or class_layer" \
  > gpurun_out/r2_initcheck.log 2>&1
echo "initcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/r2_initcheck.log | tail -n 4
grep -m 12 -A 12 "Uninitialized" gpurun_out/r2_initcheck.log | head -60
