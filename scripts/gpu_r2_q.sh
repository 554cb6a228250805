#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 4 gpurun_out/r2q_pytest.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2q_bench.json'))
print({k:d[k] for k in ['value','ms_per_step','launches_per_step']}, d['e2e']['value'])
for r in d['ops']: print(r['op'], r['ms'], r['share'])
PY
