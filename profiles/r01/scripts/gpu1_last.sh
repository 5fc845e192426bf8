#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 60 python -m pytest tests -m gpu -q -x -k "smoke or small_sizes_all_algorithms" > gpurun_out/pytest_last.log 2>&1; echo rc=$?; tail -c 600 gpurun_out/pytest_last.log
timeout 25 python bench.py --no-cpu-baseline --no-e2e 2>/dev/null | cut -c1-200
