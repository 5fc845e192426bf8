#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "=== pytest (critical subset)"; timeout 110 python -m pytest tests -m gpu -q -x -k "unaligned or smoke or mismatch or world_of_3" > gpurun_out/pytest_gpu_final3.log 2>&1; echo rc=$?; tail -c 1500 gpurun_out/pytest_gpu_final3.log
echo "=== bench"; timeout 60 python bench.py --no-cpu-baseline > gpurun_out/bench_n1f.json 2>gpurun_out/bench_n1f.err; echo rc=$?; tail -1 gpurun_out/bench_n1f.json | cut -c1-300
