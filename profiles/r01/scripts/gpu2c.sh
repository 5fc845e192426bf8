#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
echo "=== pytest subset"; timeout 400 python -m pytest tests -m gpu -q -x -k "small_sizes or edge_values or unaligned or full_size or helloworld" > gpurun_out/pytest_gpu_n2c.log 2>&1; echo rc=$?; tail -c 1200 gpurun_out/pytest_gpu_n2c.log
echo "=== examples"; make -s -C examples; timeout 120 examples/bin/gompirun 2 examples/bin/helloworld 2>&1 | sort | tail -8; timeout 200 examples/bin/gompirun 2 examples/bin/bounce 2>&1 | tail -5
echo "=== bench n2"; timeout 200 $TR --master-port 29561 bench.py --gpus 2 > gpurun_out/bench_n2c.json 2>gpurun_out/bench_n2c.err; tail -1 gpurun_out/bench_n2c.json | cut -c1-900
rm -f gpurun_out/sweep2c.jsonl
for ob in 65536 1048576 16777216; do
echo "=== sweep own_block $ob"; timeout 200 $TR --master-port 2957${#ob} tools/sweep.py --out gpurun_out/sweep2c.jsonl --colls allreduce --algos twoshot,smem,nvls --min-bytes 4194304 --params "own_block_bytes=$ob" --tag ob$ob > gpurun_out/s4.log 2>&1; echo rc=$?; tail -1 gpurun_out/s4.log | cut -c1-200
done
echo "=== ncu single-pass on rank 0 (rank 1 unprofiled)"
A="127.0.0.1:7100,127.0.0.1:7101"
for algo in smem nvls twoshot; do
for pass in dram nvl; do
  if [ $pass = dram ]; then M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum"; else M="gpu__time_duration.sum,nvlrx__bytes.sum,nvltx__bytes.sum"; fi
  B200MPI_WATCHDOG_S=20 timeout 150 python tools/sweep.py --out gpurun_out/ncu_side.jsonl --colls allreduce --algos $algo --min-bytes 268435456 --max-bytes 268435456 -mpi-addr 127.0.0.1:7101 -mpi-alladdr $A -mpi-gpu 1 > gpurun_out/ncu_r1.log 2>&1 &
  R1=$!
  B200MPI_WATCHDOG_S=20 timeout 150 ncu --metrics $M --clock-control none --cache-control none -k regex:allreduce -s 2 -c 6 --csv --log-file gpurun_out/ncu_${algo}_${pass}_n2.csv python tools/sweep.py --out gpurun_out/ncu_side.jsonl --colls allreduce --algos $algo --min-bytes 268435456 --max-bytes 268435456 -mpi-addr 127.0.0.1:7100 -mpi-alladdr $A -mpi-gpu 0 > gpurun_out/ncu_r0.log 2>&1
  echo "ncu $algo $pass rc=$?"; wait $R1; tail -4 gpurun_out/ncu_${algo}_${pass}_n2.csv | cut -c1-260
done; done
