#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 4 gpurun_out/r2n_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 3
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r2n_bench.json 2> gpurun_out/r2n_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2n_bench.json'))
print({k:d[k] for k in ['value','ms_per_step','launches_per_step']}, d['e2e']['value'])
PY
