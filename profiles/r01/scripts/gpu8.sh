#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "=== pytest subset (8 GPUs)"; timeout 300 python -m pytest tests -m gpu -q -x -k "world_of_8 or full_size or edge_values or helloworld" > gpurun_out/pytest_gpu_n8.log 2>&1; echo rc=$?; tail -c 1500 gpurun_out/pytest_gpu_n8.log
rm -f gpurun_out/sweep_n8.jsonl gpurun_out/sweep_n4.jsonl
echo "=== bench n8"; timeout 200 $TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo rc=$?; tail -1 gpurun_out/bench_n8.json; tail -2 gpurun_out/bench_n8.err
echo "=== sweep n8"; timeout 400 $TR --nproc-per-node 8 --master-port 29542 tools/sweep.py > gpurun_out/sweep_n8.log 2>&1; echo rc=$?; tail -2 gpurun_out/sweep_n8.log | cut -c1-300
echo "=== sweep n4"; timeout 200 $TR --nproc-per-node 4 --master-port 29543 tools/sweep.py --colls allreduce,bcast --min-bytes 1048576 > gpurun_out/sweep_n4.log 2>&1; echo rc=$?; tail -1 gpurun_out/sweep_n4.log | cut -c1-300
echo "=== bench n8 twoshot / smem"; for a in twoshot smem; do timeout 100 $TR --nproc-per-node 8 --master-port 2955${#a} bench.py --gpus 8 --steps 20 --warmup 5 --algo $a --no-e2e > gpurun_out/bench_n8_$a.json 2>/dev/null; tail -1 gpurun_out/bench_n8_$a.json | cut -c1-200; done
echo "=== ref n8"; timeout 200 python bench.py --impl reference --gpus 8 --steps 3 --warmup 3 | tail -1 | cut -c1-400
