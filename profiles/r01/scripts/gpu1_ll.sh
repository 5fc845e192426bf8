#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 70 python -m pytest tests -m gpu -q -x -k "ll_allreduce" > gpurun_out/pytest_ll.log 2>&1; echo rc=$?; tail -c 2500 gpurun_out/pytest_ll.log
