#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "=== pytest (4 GPUs)"; timeout 500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_n4.log 2>&1; echo rc=$?; tail -c 800 gpurun_out/pytest_gpu_n4.log
rm -f gpurun_out/sweep4c.jsonl
echo "=== n4 nvls configs vs smem"; timeout 200 $TR --nproc-per-node 4 --master-port 29741 tools/sweep.py --out gpurun_out/sweep4c.jsonl --colls allreduce --algos nvls,smem --sizes 4194304,16777216,67108864,268435456,1073741824 --blocks 32,64,148 --param-sets "nvls_unroll=1|nvls_unroll=2|nvls_unroll=4" > gpurun_out/s4c.log 2>&1; echo rc=$?; tail -1 gpurun_out/s4c.log | cut -c1-200
echo "=== bench n4 auto"; timeout 200 $TR --nproc-per-node 4 --master-port 29742 bench.py --gpus 4 > gpurun_out/bench_n4c.json 2> gpurun_out/bench_n4c.err; echo rc=$?; tail -1 gpurun_out/bench_n4c.json | cut -c1-300
echo "=== auto sweep n4"; timeout 300 $TR --nproc-per-node 4 --master-port 29743 tools/sweep.py --out gpurun_out/sweep4c_auto.jsonl --algos auto > gpurun_out/s4d.log 2>&1; echo rc=$?; tail -1 gpurun_out/s4d.log | cut -c1-200
