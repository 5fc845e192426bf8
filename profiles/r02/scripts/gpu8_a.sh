#!/bin/bash
# 8 GPUs: switch paths parity, bench (parity / e2e roofline / secondary / NCCL line), hybrid + LL + bcast/allgather sweeps, ncu on rank 0
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi -L | wc -l
nvidia-smi topo -m > gpurun_out/r2_topo_n8.txt 2>&1
lscpu | grep -i "numa\|model name\|^CPU(s)" > gpurun_out/r2_lscpu.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "=== pytest (8 GPUs)"; timeout 420 python -m pytest tests -m gpu -q -x --timeout 400 -k "switch_paths or world_of_8" > gpurun_out/r2_pytest_gpu_n8.log 2>&1; echo rc=$?; tail -c 800 gpurun_out/r2_pytest_gpu_n8.log
echo "=== bench n8"; timeout 400 $TR --nproc-per-node 8 --master-port 29841 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo rc=$?; tail -c 4500 gpurun_out/r2_bench_n8.json; tail -3 gpurun_out/r2_bench_n8.err
echo "=== bench n8, 4 MiB pipeline chunks (e2e only)"; timeout 200 $TR --nproc-per-node 8 --master-port 29845 bench.py --gpus 8 --steps 10 --warmup 3 --no-parity --no-secondary --params "pipe_chunk_bytes=4194304" > gpurun_out/r2_bench_n8_chunk4.json 2> /dev/null; echo rc=$?; python -c "import json; d=json.loads(open('gpurun_out/r2_bench_n8_chunk4.json').read().strip().splitlines()[-1]); print(d['e2e'], d['e2e_pageable'])"
rm -f gpurun_out/r2_sweep_n8_*.jsonl
echo "=== hybrid"; timeout 200 $TR --nproc-per-node 8 --master-port 29842 tools/sweep.py --out gpurun_out/r2_sweep_n8_hybrid.jsonl --colls allreduce --algos nvls,hybrid --sizes 268435456,1073741824 --param-sets "hybrid_p2p_permille=100|hybrid_p2p_permille=150|hybrid_p2p_permille=200|hybrid_p2p_permille=250|hybrid_p2p_permille=300|hybrid_p2p_permille=400|hybrid_p2p_permille=200,hybrid_p2p_blocks=32|hybrid_p2p_permille=200,hybrid_p2p_blocks=64|hybrid_p2p_permille=200,hybrid_p2p_blocks=0,nvls_max_blocks=48" > gpurun_out/r2_s8a.log 2>&1; echo rc=$?; tail -1 gpurun_out/r2_s8a.log | cut -c1-250
echo "=== small/mid"; timeout 200 $TR --nproc-per-node 8 --master-port 29843 tools/sweep.py --out gpurun_out/r2_sweep_n8_small.jsonl --colls allreduce,latency --algos ll,oneshot,twoshot,nvls --sizes 1024,4096,16384,32768,65536,131072,262144,524288,1048576,4194304 > gpurun_out/r2_s8b.log 2>&1; echo rc=$?; tail -1 gpurun_out/r2_s8b.log | cut -c1-250
echo "=== bcast/allgather/link"; timeout 300 $TR --nproc-per-node 8 --master-port 29844 tools/sweep.py --out gpurun_out/r2_sweep_n8_bcag.jsonl --colls bcast,allgather,link --sizes 1048576,8388608,67108864,268435456,1073741824 > gpurun_out/r2_s8c.log 2>&1; echo rc=$?; tail -1 gpurun_out/r2_s8c.log | cut -c1-250
echo "=== ncu rank 0, nvls 256 MiB"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__registers_per_thread,lts__t_sectors_srcunit_tex_aperture_peer.sum,lts__t_sectors_aperture_peer.sum"
NCU_PORT_BASE=17100 tools/ncu_rank0.sh 8 gpurun_out/r2_ncu_nvls_n8.csv "$M" allreduce -- --colls allreduce --algos nvls --sizes 268435456
