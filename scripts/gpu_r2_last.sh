#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -n 3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
timeout 400 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2z_bench.json'))
print({k:d[k] for k in ['value','ms_per_step','launches_per_step','gpu_launches']}, d['e2e']['value'], d['cpu_baseline']['value'])
print(d['roofline']['bound'], round(d['roofline']['frac'],3), d['roofline']['traffic'], d['conv_roofline']['frac'], d['clocks'])
PY
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | cut -c 1-200
