#!/bin/bash
# 1 GPU: the GPU parity suite (worlds share device 0) after the round-2 core changes
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r2_pytest_gpu_a.log 2>&1; echo "pytest rc=$?"
tail -c 3000 gpurun_out/r2_pytest_gpu_a.log
