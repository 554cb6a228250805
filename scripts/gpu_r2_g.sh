#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "umma" > gpurun_out/r2g_pytest_umma.log 2>&1
rc=$?; echo "pytest umma rc=$rc"; tail -n 3 gpurun_out/r2g_pytest_umma.log
[[ $rc -ne 0 ]] && exit 0
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 3 gpurun_out/r2g_pytest.log
timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2g_bench.json'))
print({k:d[k] for k in ['value','ms_per_step','launches_per_step']}, d['e2e']['value'])
for r in d['ops']: print(r['op'], r['ms'], r['share'])
PY
tail -n 3 gpurun_out/r2g_bench.err
