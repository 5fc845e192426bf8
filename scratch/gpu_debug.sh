#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out/trace
export B200MPI_DEBUG=1 B200MPI_WATCHDOG_S=15 WORLD_TIMEOUT=70
i=0
for spec in "2 edge_values" "2 collectives --sizes 0,1,257 --kinds host --dtypes f32" "2 collectives --sizes 0,1,257,65537 --kinds heap --dtypes f32"; do
  i=$((i+1))
  export B200MPI_TEST_TRACE=gpurun_out/trace/t$i
  echo "=== $spec"
  timeout 100 python scratch/run_world.py $spec > gpurun_out/trace/out$i.log 2>&1
  tail -c 3000 gpurun_out/trace/out$i.log
  for f in gpurun_out/trace/t$i.rank*; do echo "--- $f"; tail -5 $f; done
done
