#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "=== pytest"; timeout 700 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_final1.log 2>&1; echo rc=$?; tail -c 1500 gpurun_out/pytest_gpu_final1.log
rm -f gpurun_out/sweep1_copy.jsonl
echo "=== copy variants"; timeout 200 python tools/sweep.py --out gpurun_out/sweep1_copy.jsonl --colls allreduce --algos twoshot --sizes 67108864,268435456,1073741824 --param-sets "copy_variant=0|copy_variant=1|copy_variant=2|copy_variant=3|copy_variant=4|copy_variant=5|copy_variant=6|copy_variant=7" > gpurun_out/s1c.log 2>&1; echo rc=$?; tail -2 gpurun_out/s1c.log | cut -c1-200
echo "=== bench"; timeout 200 python bench.py > gpurun_out/bench_n1d.json 2>gpurun_out/bench_n1d.err; tail -1 gpurun_out/bench_n1d.json | cut -c1-1200
