#!/bin/bash
# first functional pass on one GPU: worlds of 1, 2, 4 ranks sharing device 0
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
export B200MPI_DEBUG=1
for spec in "1 smoke" "2 smoke" "1 collectives" "2 collectives" "2 p2p --sizes 0,1,10,100,1000,10000,100000,1000000" "2 helloworld" "2 tags" "2 edge_values" "2 unaligned" "4 collectives --sizes 0,1,5,257,65537 --kinds heap" "4 helloworld" "4 edge_values" "3 collectives --sizes 3,257,4099 --kinds heap"; do
  echo "=== $spec"
  timeout 600 python scratch/run_world.py $spec 2>&1 | tail -40
done
