"""mpi_b200 -- B200-native transport behind the btracey/mpi API.

`import mpi_b200 as mpi` gives the reference's package surface (mpi.go) over libb200mpi.so:
Init, Finalize, Rank, Size, Send, Receive (+ Recv, Bcast, Allreduce, Allgather, Barrier,
ReduceScatter, Reduce, Alltoall, Isend/Wait).
The shared library (mpi_b200/lib/libb200mpi.so, built by mpi_b200/csrc/Makefile) is loaded on
first use; nothing here computes or moves payload bytes in Python.
"""
from ._lib import LIB_PATH, load  # noqa: F401
from .api import (  # noqa: F401
    MAX, MIN, SUM, Alloc, Allgather, Allreduce, Alltoall, Barrier, Bcast, Cuda, DeviceSlice, Finalize, Init,
    Interface, Isend, MpiError, Rank, Raw, Receive, Recv, Reduce, ReduceScatter, Register, Send, Size,
    TagExists, Wait)
from . import flags  # noqa: F401
