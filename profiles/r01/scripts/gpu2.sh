#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L; nvidia-smi topo -m | head -6
export B200MPI_DEBUG=1
echo "=== pytest"; timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_n2.log 2>&1; echo rc=$?; tail -c 1500 gpurun_out/pytest_gpu_n2.log
echo "=== sweep"; rm -f gpurun_out/sweep_n2.jsonl
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sweep.py > gpurun_out/sweep_n2.log 2>&1; echo rc=$?; tail -c 1500 gpurun_out/sweep_n2.log
echo "=== bench n2"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo rc=$?; tail -1 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
echo "=== ref n2"; timeout 200 python bench.py --impl reference --gpus 2 --steps 3 --warmup 3 | tail -1
