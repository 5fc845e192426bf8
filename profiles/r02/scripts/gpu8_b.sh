#!/bin/bash
# 8 GPUs, final: bench (all parity flags, e2e + roofline, pageable, secondary, NCCL line) and the LL two-phase sweep
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "=== bench n8"; timeout 400 $TR --nproc-per-node 8 --master-port 29851 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8_b.json 2> gpurun_out/r2_bench_n8_b.err; echo rc=$?; grep '^{"metric' gpurun_out/r2_bench_n8_b.json | tail -c 5000; tail -2 gpurun_out/r2_bench_n8_b.err
rm -f gpurun_out/r2_sweep_n8_small2.jsonl
echo "=== small"; timeout 150 $TR --nproc-per-node 8 --master-port 29852 tools/sweep.py --out gpurun_out/r2_sweep_n8_small2.jsonl --colls allreduce,latency --algos ll,nvls --sizes 1024,4096,16384,32768,65536,131072,262144 > gpurun_out/r2_s8d.log 2>&1; echo rc=$?; tail -1 gpurun_out/r2_s8d.log | cut -c1-200
