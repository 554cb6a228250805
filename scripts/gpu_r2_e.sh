#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "bwd_weight or class_layer" > gpurun_out/r2e_pytest_dw.log 2>&1
rc=$?; echo "pytest dw rc=$rc"; tail -n 6 gpurun_out/r2e_pytest_dw.log
if [[ $rc -eq 0 ]]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1
  echo "pytest rc=$?"; tail -n 6 gpurun_out/r2e_pytest.log
fi
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2e_bench.json'))
print({k:d[k] for k in ['value','ms_per_step','launches_per_step']}, d['e2e']['value'])
for r in d['ops']: print(r['op'], r['ms'], r['share'])
print(d.get('extras'))
PY
tail -n 3 gpurun_out/r2e_bench.err
