"""One rank of a CPU dry run of bench.py: installs tests/fake_b200mpi.FakeLib where mpi_b200 looks for
libb200mpi.so and calls bench.main() with the arguments given.  TEST INFRASTRUCTURE (see fake_b200mpi.py)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import fake_b200mpi  # noqa: E402

fake_b200mpi.install()
import bench  # noqa: E402

sys.argv = ["bench.py"] + sys.argv[1:]
sys.exit(bench.main())
