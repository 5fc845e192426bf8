#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "=== pytest"; timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_n1c.log 2>&1; echo rc=$?; tail -c 800 gpurun_out/pytest_gpu_n1c.log
echo "=== smoke"; timeout 200 python __graft_entry__.py smoke 2>&1 | tail -3
echo "=== bench"; timeout 200 python bench.py > gpurun_out/bench_n1c.json 2>gpurun_out/bench_n1c.err; tail -1 gpurun_out/bench_n1c.json | cut -c1-300
echo "=== ncu launches"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_n1.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1; echo rc=$?; tail -3 gpurun_out/launches_n1.csv | cut -c1-300
echo "=== ncu full copy kernel"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:copy_bytes -s 4 -c 2 -f -o gpurun_out/prof_copy_n1 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu2.log 2>&1; echo rc=$?; ls -la gpurun_out/*.ncu-rep
