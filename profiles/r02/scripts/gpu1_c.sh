#!/bin/bash
# 1 GPU: full GPU suite after the LL rewrite; bench N=1; ncu launch list + --set full of the shipped copy kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 700 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r2_pytest_gpu_c.log 2>&1; echo "pytest rc=$?"; tail -c 500 gpurun_out/r2_pytest_gpu_c.log
echo "=== bench n1"; timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo rc=$?; tail -c 2500 gpurun_out/r2_bench_n1.json
echo "=== ncu launch list"; timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 8 -c 24 --csv --log-file gpurun_out/r2_launches_bench_n1.csv python bench.py --gpus 1 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_ncu_b1.log 2>&1; echo rc=$?; tail -4 gpurun_out/r2_launches_bench_n1.csv | cut -c1-250
echo "=== ncu --set full copy kernel"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:copy_bytes -s 10 -c 2 -o gpurun_out/r2_prof_copy_n1 -f python bench.py --gpus 1 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/r2_ncu_b2.log 2>&1; echo rc=$?; ls -la gpurun_out/r2_prof_copy_n1.ncu-rep
