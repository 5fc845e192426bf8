#!/bin/bash
# 2 GPUs: switch paths + LL over real NVLink, bench N=2 (parity / secondary / NCCL line), link probes, small sweeps
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L | wc -l
nvidia-smi topo -m > gpurun_out/r2_topo_n2.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "=== pytest (2 GPUs)"; timeout 500 python -m pytest tests -m gpu -q -x --timeout 400 -k "switch_paths or ll_allreduce or isend or (collectives_small and 2) or (host_slice and 2) or (reduce_scatter and 2-)" > gpurun_out/r2_pytest_gpu_n2.log 2>&1; echo rc=$?; tail -c 1500 gpurun_out/r2_pytest_gpu_n2.log
echo "=== bench n2"; timeout 300 $TR --nproc-per-node 2 --master-port 29741 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo rc=$?; tail -c 3000 gpurun_out/r2_bench_n2.json; tail -5 gpurun_out/r2_bench_n2.err
rm -f gpurun_out/r2_sweep_n2_*.jsonl
echo "=== link probes"; timeout 120 $TR --nproc-per-node 2 --master-port 29742 tools/sweep.py --out gpurun_out/r2_sweep_n2_link.jsonl --colls link --max-bytes 268435456 > gpurun_out/r2_s2a.log 2>&1; echo rc=$?; tail -3 gpurun_out/r2_s2a.log | cut -c1-250
echo "=== allreduce big"; timeout 200 $TR --nproc-per-node 2 --master-port 29743 tools/sweep.py --out gpurun_out/r2_sweep_n2_big.jsonl --colls allreduce --algos twoshot,smem,nvls,hybrid --sizes 67108864,268435456 --param-sets "hybrid_p2p_permille=500|hybrid_p2p_permille=700" > gpurun_out/r2_s2b.log 2>&1; echo rc=$?; tail -2 gpurun_out/r2_s2b.log | cut -c1-250
echo "=== allreduce small"; timeout 200 $TR --nproc-per-node 2 --master-port 29744 tools/sweep.py --out gpurun_out/r2_sweep_n2_small.jsonl --colls allreduce,latency --algos ll,oneshot,twoshot,nvls --sizes 1024,8192,32768,131072,262144 > gpurun_out/r2_s2c.log 2>&1; echo rc=$?; tail -2 gpurun_out/r2_s2c.log | cut -c1-250
echo "=== ref n2 (same config)"; timeout 300 python bench.py --impl reference --gpus 2 --steps 20 --warmup 5 | tail -1 | cut -c1-900
