#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "=== pytest subset"; timeout 300 python -m pytest tests -m gpu -q -x -k "small_sizes or edge_values or mismatch or stream or full_size or bounce_cpp or helloworld_cpp" > gpurun_out/pytest_gpu_n2_final.log 2>&1; echo rc=$?; tail -c 800 gpurun_out/pytest_gpu_n2_final.log
echo "=== bench n2"; timeout 200 $TR --master-port 29861 bench.py --gpus 2 > gpurun_out/bench_n2_final.json 2>gpurun_out/bench_n2_final.err; echo rc=$?; tail -1 gpurun_out/bench_n2_final.json | cut -c1-700
echo "=== nvls n2 256MiB"; timeout 100 $TR --master-port 29862 bench.py --gpus 2 --algo nvls --no-e2e > gpurun_out/bench_n2_nvls_final.json 2>/dev/null; tail -1 gpurun_out/bench_n2_nvls_final.json | cut -c1-200
