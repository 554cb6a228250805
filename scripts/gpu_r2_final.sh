#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
timeout 600 python bench.py > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench.json'))
print({k:d[k] for k in ['value','ms_per_step','launches_per_step','gpu_launches']}, d['e2e'], d['cpu_baseline'])
print(d['roofline']); print(d['conv_roofline']); print(d['clocks'])
for r in d['ops']: print(r['op'], r['ms'], r['share'], r['GBs'], r['TFLOPs'])
print(d['extras'])
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2f_ref.json 2> gpurun_out/r2f_ref.err
echo "ref rc=$?"; cut -c 1-400 gpurun_out/r2f_ref.json
timeout 300 python bench.py --pages 4096 --steps 10 --warmup 3 --no-cpu-baseline --no-op-profile --no-extras > gpurun_out/r2f_bench_4096.json 2> gpurun_out/r2f_bench_4096.err
echo "bench 4096 rc=$?"; cut -c 1-230 gpurun_out/r2f_bench_4096.json
timeout 600 python bench.py --mode infer --steps 2 --warmup 1 > gpurun_out/r2f_infer.json 2> gpurun_out/r2f_infer.err
echo "infer rc=$?"; cut -c 1-230 gpurun_out/r2f_infer.json
