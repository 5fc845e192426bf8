#!/bin/bash
# 4 GPUs: bench (parity / e2e / secondary / NCCL line), LL two-phase sweep, big-message variants, link probe
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "=== bench n4"; timeout 300 $TR --nproc-per-node 4 --master-port 29941 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err; echo rc=$?; tail -c 2600 gpurun_out/r2_bench_n4.json; tail -2 gpurun_out/r2_bench_n4.err
rm -f gpurun_out/r2_sweep_n4_*.jsonl
echo "=== small"; timeout 150 $TR --nproc-per-node 4 --master-port 29942 tools/sweep.py --out gpurun_out/r2_sweep_n4_small.jsonl --colls allreduce --algos ll,nvls,oneshot --sizes 1024,4096,32768,65536,131072,262144,524288,1048576 > gpurun_out/r2_s4a.log 2>&1; echo rc=$?; tail -1 gpurun_out/r2_s4a.log | cut -c1-200
echo "=== big + link"; timeout 200 $TR --nproc-per-node 4 --master-port 29943 tools/sweep.py --out gpurun_out/r2_sweep_n4_big.jsonl --colls allreduce,link --algos twoshot,smem,nvls --sizes 67108864,268435456,1073741824 --max-bytes 1073741824 > gpurun_out/r2_s4b.log 2>&1; echo rc=$?; tail -1 gpurun_out/r2_s4b.log | cut -c1-200
