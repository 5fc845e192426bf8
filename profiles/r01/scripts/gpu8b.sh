#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "=== pytest subset (8 GPUs)"; timeout 240 python -m pytest tests -m gpu -q -x -k "world_of_8 or full_size" > gpurun_out/pytest_gpu_n8b.log 2>&1; echo rc=$?; tail -c 600 gpurun_out/pytest_gpu_n8b.log
echo "=== bench n8 auto"; timeout 200 $TR --nproc-per-node 8 --master-port 29641 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8b.json 2> gpurun_out/bench_n8b.err; echo rc=$?; tail -1 gpurun_out/bench_n8b.json | cut -c1-1500
rm -f gpurun_out/sweep8b_*.jsonl
echo "=== nvls/smem own_block"; timeout 200 $TR --nproc-per-node 8 --master-port 29642 tools/sweep.py --out gpurun_out/sweep8b_own.jsonl --colls allreduce --algos nvls,smem --sizes 134217728,268435456,536870912,1073741824 --param-sets "own_block_bytes=65536|own_block_bytes=1048576|own_block_bytes=16777216|own_block_bytes=1073741824" > gpurun_out/s8a.log 2>&1; echo rc=$?; tail -1 gpurun_out/s8a.log | cut -c1-200
echo "=== nvls blocks/unroll"; timeout 200 $TR --nproc-per-node 8 --master-port 29643 tools/sweep.py --out gpurun_out/sweep8b_blk.jsonl --colls allreduce --algos nvls --sizes 67108864,268435456,1073741824 --blocks 64,96,128,148 --param-sets "nvls_unroll=2|nvls_unroll=4|nvls_unroll=8" > gpurun_out/s8b.log 2>&1; echo rc=$?; tail -1 gpurun_out/s8b.log | cut -c1-200
echo "=== auto sweep n8"; timeout 300 $TR --nproc-per-node 8 --master-port 29644 tools/sweep.py --out gpurun_out/sweep8b_auto.jsonl --algos auto > gpurun_out/s8c.log 2>&1; echo rc=$?; tail -1 gpurun_out/s8c.log | cut -c1-200
echo "=== bench n4 auto"; timeout 200 $TR --nproc-per-node 4 --master-port 29645 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/bench_n4b.json 2> gpurun_out/bench_n4b.err; echo rc=$?; tail -1 gpurun_out/bench_n4b.json | cut -c1-400
