#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "dropout or bwd_weight" > gpurun_out/r2i_pytest_new.log 2>&1
rc=$?; echo "pytest new rc=$rc"; tail -n 12 gpurun_out/r2i_pytest_new.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 4 gpurun_out/r2i_pytest.log
timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2i_bench.json'))
print({k:d[k] for k in ['value','ms_per_step','launches_per_step']}, d['e2e']['value'])
for r in d['ops'][:6]: print(r['op'], r['ms'], r['share'])
PY
timeout 600 python bench.py --mode infer --steps 2 --warmup 1 > gpurun_out/r2i_infer.json 2> gpurun_out/r2i_infer.err
echo "infer rc=$?"; cut -c 1-400 gpurun_out/r2i_infer.json; tail -n 3 gpurun_out/r2i_infer.err
