#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "dp_allreduce" > gpurun_out/r2j_pytest_dp.log 2>&1
rc=$?; echo "pytest dp rc=$rc"; tail -n 5 gpurun_out/r2j_pytest_dp.log
[[ $rc -ne 0 ]] && exit 0
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dp_peer_check.py > gpurun_out/r2j_dp_peer.log 2>&1
echo "dp_peer rc=$?"; tail -n 6 gpurun_out/r2j_dp_peer.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --no-extras > gpurun_out/r2j_bench_n2.json 2> gpurun_out/r2j_bench_n2.err
echo "bench n2 rc=$?"; cut -c 1-330 gpurun_out/r2j_bench_n2.json; tail -n 2 gpurun_out/r2j_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --mode infer --steps 2 --warmup 1 > gpurun_out/r2j_infer_n2.json 2> gpurun_out/r2j_infer_n2.err
echo "infer n2 rc=$?"; cut -c 1-200 gpurun_out/r2j_infer_n2.json; tail -n 2 gpurun_out/r2j_infer_n2.err
