#!/bin/bash
# ncu on rank 0 of an N-rank world, the other ranks unprofiled (never wrap a multi-rank launcher in ncu).
#   tools/ncu_rank0.sh N OUT_CSV METRICS KERNEL_REGEX -- <sweep.py arguments>
# Single-pass metric lists only: the kernels rendezvous across ranks, so a replayed launch would
# run against peers that have moved on (it still terminates -- flags only grow -- but its timing
# would mean nothing).
N=$1; OUT=$2; METRICS=$3; KREGEX=$4; shift 4; [ "$1" = "--" ] && shift
cd "$(dirname "$0")/.."
BASE=${NCU_PORT_BASE:-17100}
A=""; for r in $(seq 0 $((N-1))); do A="$A${A:+,}127.0.0.1:$((BASE+r))"; done
PIDS=""
for r in $(seq 1 $((N-1))); do
  B200MPI_WATCHDOG_S=${B200MPI_WATCHDOG_S:-40} timeout 200 python tools/sweep.py --out gpurun_out/ncu_side.jsonl "$@" -mpi-addr 127.0.0.1:$((BASE+r)) -mpi-alladdr $A -mpi-gpu $r > gpurun_out/ncu_r$r.log 2>&1 &
  PIDS="$PIDS $!"
done
B200MPI_WATCHDOG_S=${B200MPI_WATCHDOG_S:-40} timeout 200 ncu --metrics "$METRICS" --clock-control none --cache-control none -k "regex:$KREGEX" -s ${NCU_SKIP:-3} -c ${NCU_COUNT:-6} --csv --log-file "$OUT" \
  python tools/sweep.py --out gpurun_out/ncu_side.jsonl "$@" -mpi-addr 127.0.0.1:$BASE -mpi-alladdr $A -mpi-gpu 0 > gpurun_out/ncu_r0.log 2>&1
RC=$?
for p in $PIDS; do wait $p; done
echo "ncu rank0 rc=$RC"; tail -n 8 "$OUT" | cut -c1-300
