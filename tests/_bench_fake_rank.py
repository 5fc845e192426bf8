"""One rank of a CPU dry run of bench.py: installs tests/fake_b200mpi.FakeLib where mpi_b200 looks for
libb200mpi.so and calls bench.main() with the arguments given.  TEST INFRASTRUCTURE (see fake_b200mpi.py)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import fake_b200mpi  # noqa: E402

fake = fake_b200mpi.install()
import bench  # noqa: E402

if os.environ.get("FAKE_NCCL") == "1":
    # a stand-in for libnccl.so.2 with the five entry points bench.py binds, so that the comparison
    # block (init with retry, timing loops, result cross-check, deadline guard) executes on the CPU
    import ctypes

    class FakeNccl:
        class _Fn:
            def __init__(self, f):
                self.f, self.argtypes, self.restype = f, None, None

            def __call__(self, *a):
                return self.f(*a)

        def __init__(self):
            self.ncclGetVersion = self._Fn(lambda v: (setattr(v._obj, "value", 22703), 0)[1])
            self.ncclGetUniqueId = self._Fn(lambda u: 0)
            self.ncclCommInitRank = self._Fn(lambda comm, n, uid, rank: (setattr(comm._obj, "value", 1), 0)[1])
            self.ncclAllReduce = self._Fn(lambda s, r, cnt, dt, op, comm, stream: fake.b200mpi_allreduce(s, r, cnt, 2, 0, 1))
            self.ncclCommDestroy = self._Fn(lambda comm: 0)

    real_cdll = ctypes.CDLL
    ctypes.CDLL = lambda name, *a, **k: FakeNccl() if "nccl" in str(name) else real_cdll(name, *a, **k)

sys.argv = ["bench.py"] + sys.argv[1:]
sys.exit(bench.main())
