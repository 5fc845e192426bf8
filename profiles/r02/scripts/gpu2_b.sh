#!/bin/bash
# 2 GPUs: the full bench after the registration-cache flush, the NCCL deadline and the host p2p fast path
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "=== bench n2"; timeout 200 $TR --nproc-per-node 2 --master-port 29761 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2_b.json 2> gpurun_out/r2_bench_n2_b.err; echo rc=$?; grep '^{"metric' gpurun_out/r2_bench_n2_b.json | tail -c 2400; tail -2 gpurun_out/r2_bench_n2_b.err
echo "=== p2p tests"; timeout 150 python -m pytest tests -m gpu -q -x --timeout 140 -k "bounce or isend" > gpurun_out/r2_pytest_gpu_n2_b.log 2>&1; echo rc=$?; tail -3 gpurun_out/r2_pytest_gpu_n2_b.log
