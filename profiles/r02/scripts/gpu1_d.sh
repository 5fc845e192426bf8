#!/bin/bash
# 1 GPU, last check of the round: the whole GPU suite on the final tree, then the N=1 contract line
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 260 python -m pytest tests -m gpu -q -x --timeout 240 > gpurun_out/r2_pytest_gpu_d.log 2>&1; echo "pytest rc=$?"; tail -c 400 gpurun_out/r2_pytest_gpu_d.log
timeout 60 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2_bench_n1_d.json 2> gpurun_out/r2_bench_n1_d.err; echo "bench rc=$?"; tail -c 700 gpurun_out/r2_bench_n1_d.json
