#!/bin/bash
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
echo "=== pytest"; timeout 700 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_final2.log 2>&1; echo rc=$?; tail -c 2500 gpurun_out/pytest_gpu_final2.log
echo "=== bench"; timeout 200 python bench.py > gpurun_out/bench_n1e.json 2>gpurun_out/bench_n1e.err; echo rc=$?; tail -1 gpurun_out/bench_n1e.json | cut -c1-1300
echo "=== memcheck (world of 1: allreduce copy path + self send/recv)"; timeout 150 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "
import __graft_entry__ as g, numpy as np, threading
import mpi_b200 as mpi
mpi.api._reset_for_tests(mpi.Cuda(Gpu=0)); mpi.Init()
x=np.arange(100003,dtype=np.float32); s=mpi.Alloc(x.size,np.float32).copy_from_host(x); r=mpi.Alloc(x.size,np.float32)
mpi.Allreduce(s,r); assert np.array_equal(r.to_host(),x)
got={}
t=threading.Thread(target=lambda: got.setdefault('v', mpi.Receive(np.zeros(x.size,dtype=np.float32),0,3))); t.start(); mpi.Send(x,0,3); t.join(); assert np.array_equal(got['v'],x)
mpi.Finalize(); print('memcheck-body-ok')
" 2>&1 | tail -6
