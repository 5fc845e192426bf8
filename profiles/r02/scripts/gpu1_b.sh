#!/bin/bash
# 1 GPU: re-check the suite after the fence-free start barrier; the same-config reference arm for 8 ranks on this host's cores
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 500 -k "smoke or collectives_small or mismatch or ll_allreduce or (reduce_scatter and 2-) or world_of_3 or caller_stream or tags" > gpurun_out/r2_pytest_gpu_b.log 2>&1; echo "pytest rc=$?"; tail -c 600 gpurun_out/r2_pytest_gpu_b.log
echo "=== ref n8 (same config)"; timeout 400 python bench.py --impl reference --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_ref_n8.json; tail -1 gpurun_out/r2_ref_n8.json | cut -c1-1300
