#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "=== pytest subset"; timeout 400 python -m pytest tests -m gpu -q -x -k "helloworld or small_sizes or edge_values or unaligned or full_size or tags" > gpurun_out/pytest_gpu_n2b.log 2>&1; echo rc=$?; tail -c 1800 gpurun_out/pytest_gpu_n2b.log
rm -f gpurun_out/sweep2b.jsonl
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "=== sweep smem/twoshot hi"; timeout 200 $TR --master-port 29521 tools/sweep.py --out gpurun_out/sweep2b.jsonl --colls allreduce --algos twoshot,smem --min-bytes 65536 --tag hi > gpurun_out/s1.log 2>&1; echo rc=$?; tail -2 gpurun_out/s1.log
echo "=== sweep twoshot lo"; timeout 200 $TR --master-port 29522 tools/sweep.py --out gpurun_out/sweep2b.jsonl --colls allreduce --algos twoshot --min-bytes 1048576 --params "twoshot_unroll=0" --tag lo > gpurun_out/s2.log 2>&1; echo rc=$?; tail -1 gpurun_out/s2.log
for u in 1 2 8; do
echo "=== sweep nvls unroll $u"; timeout 200 $TR --master-port 2953$u tools/sweep.py --out gpurun_out/sweep2b.jsonl --colls allreduce --algos nvls --min-bytes 16777216 --params "nvls_unroll=$u" --tag u$u > gpurun_out/s3$u.log 2>&1; echo rc=$?; tail -1 gpurun_out/s3$u.log
done
