#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 120 python scripts/pair_smoke.py > gpurun_out/r2h_pair_smoke.log 2>&1
rc=$?; echo "pair_smoke rc=$rc"; tail -n 3 gpurun_out/r2h_pair_smoke.log
[[ $rc -ne 0 ]] && exit 0
GTE_LIB=$PWD/gnn_tableextraction_b200/libgte_b200_exp.so GTE_UMMA_DBG=1 GTE_UMMA_PAIR=1 timeout 60 python scripts/umma_trace.py 2>&1 | grep -E "event ms|tile 2|period" | cut -c 1-60,150-300
timeout 120 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "umma" > gpurun_out/r2h_pytest_umma.log 2>&1
rc=$?; echo "pytest umma rc=$rc"; tail -n 3 gpurun_out/r2h_pytest_umma.log
[[ $rc -ne 0 ]] && exit 0
timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2h_bench.json'))
print({k:d[k] for k in ['value','ms_per_step','launches_per_step']}, d['e2e']['value'])
for r in d['ops']: print(r['op'], r['ms'], r['share'])
PY
