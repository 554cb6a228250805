#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 4 gpurun_out/r2k_pytest.log
timeout 600 python bench.py > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2k_bench.json'))
print({k:d[k] for k in ['value','ms_per_step','launches_per_step']}, d['e2e']['value'], d['cpu_baseline'])
print(d['roofline']); print(d['conv_roofline'])
for r in d['ops']: print(r['op'], r['ms'], r['share'])
print(d['extras'])
PY
timeout 600 python bench.py --mode infer --steps 2 --warmup 1 > gpurun_out/r2k_infer.json 2> gpurun_out/r2k_infer.err
echo "infer rc=$?"; cut -c 1-230 gpurun_out/r2k_infer.json; tail -n 3 gpurun_out/r2k_infer.err
timeout 300 python bench.py --pages 4096 --steps 10 --warmup 3 --no-cpu-baseline --no-op-profile --no-extras > gpurun_out/r2k_bench_4096.json 2> gpurun_out/r2k_bench_4096.err
echo "bench 4096 rc=$?"; cut -c 1-230 gpurun_out/r2k_bench_4096.json; tail -n 3 gpurun_out/r2k_bench_4096.err
