#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29521 scripts/dp_peer_check.py > gpurun_out/r2n_dp_peer_n$N.log 2>&1
echo "dp_peer rc=$?"; tail -n 3 gpurun_out/r2n_dp_peer_n$N.log
timeout 200 $TR --master-port 29522 bench.py --gpus $N --no-extras --no-op-profile > gpurun_out/r2n_bench_n$N.json 2> gpurun_out/r2n_bench_n$N.err
echo "bench rc=$?"; cut -c 1-260 gpurun_out/r2n_bench_n$N.json; tail -n 2 gpurun_out/r2n_bench_n$N.err
timeout 200 $TR --master-port 29523 bench.py --gpus $N --global-pages 4096 --no-extras --no-op-profile > gpurun_out/r2n_bench_strong_n$N.json 2> gpurun_out/r2n_bench_strong_n$N.err
echo "bench strong rc=$?"; cut -c 1-260 gpurun_out/r2n_bench_strong_n$N.json
timeout 200 $TR --master-port 29524 bench.py --gpus $N --mode infer --steps 2 --warmup 1 > gpurun_out/r2n_infer_n$N.json 2> gpurun_out/r2n_infer_n$N.err
echo "infer rc=$?"; cut -c 1-220 gpurun_out/r2n_infer_n$N.json; tail -n 2 gpurun_out/r2n_infer_n$N.err
