"""TEST INFRASTRUCTURE: a stand-in for libb200mpi.so that lets bench.py's host logic run on a CPU box.

It implements the C ABI names bench.py and mpi_b200/api.py call, on host memory, with gloo
(torch.distributed, CPU) as the transport between rank processes and the CPU oracle's
reduction orders as the arithmetic.  It exists so that `pytest -m "not gpu"` can execute the
whole bench script (world of 1 and world of 2: parity bookkeeping, JSON contract line, deadline
plumbing, reference-arm wiring) where there is no GPU.  Nothing here is a data path of the
product: mpi_b200 never imports it; tests/_bench_fake_rank.py installs it into
`mpi_b200._lib._lib` for the duration of one test process.
"""
import ctypes
import os
import time

import numpy as np

from oracle import oracle as O

U8, I64, F32, F64 = 0, 1, 2, 3
NP = {U8: np.uint8, I64: np.int64, F32: np.float32, F64: np.float64}
ALGO_ONESHOT, ALGO_TWOSHOT, ALGO_RING, ALGO_NVLS, ALGO_SMEM, ALGO_LL, ALGO_HYBRID = 1, 2, 3, 4, 5, 6, 7


def _addr(p):
    if p is None:
        return 0
    if isinstance(p, int):
        return p
    if hasattr(p, "value"):
        return p.value or 0
    return int(p)


def _view(p, count, dtype):
    """numpy view of `count` elements at address p (0 elements -> empty array)."""
    dt = np.dtype(NP[dtype] if isinstance(dtype, int) else dtype)
    if count == 0:
        return np.empty(0, dtype=dt)
    buf = (ctypes.c_char * (count * dt.itemsize)).from_address(_addr(p))
    return np.frombuffer(buf, dtype=dt, count=count)


def _out(ref):
    """the ctypes object behind a ctypes.byref(...) argument (or the object itself)"""
    return getattr(ref, "_obj", ref)


class FakeLib:
    def __init__(self):
        self.rank, self.n = -1, 0
        self.blocks = {}      # address -> ctypes buffer kept alive
        self.algo = [0, 0, 0, 0]
        self.params = {"hybrid_p2p_permille": 0, "host_register": 0}
        self.launches = 0
        self.t0 = 0.0
        self.err = b""
        self.dist = None
        self.nvls = os.environ.get("FAKE_NVLS", "0") == "1"  # pretend there is a multicast mapping (switch-order sums)

    # ---- lifecycle ------------------------------------------------------------------------------
    def b200mpi_init(self, addr, alladdr, password, timeout_ns, gpu):
        addrs = [a for a in alladdr.decode().split(",") if a]
        if not addrs:
            self.rank, self.n = 0, 1
            return 0
        addrs.sort()
        self.rank, self.n = addrs.index(addr.decode()), len(addrs)
        import torch.distributed as dist
        dist.init_process_group("gloo", rank=self.rank, world_size=self.n)
        self.dist = dist
        return 0

    def b200mpi_finalize(self):
        if self.dist is not None:
            self.dist.barrier()
            self.dist.destroy_process_group()
        self.rank, self.n = -1, 0
        return 0

    def b200mpi_rank(self):
        return self.rank

    def b200mpi_size(self):
        return self.n

    def b200mpi_device(self):
        return 0

    def b200mpi_version(self):
        return 200

    def b200mpi_last_error(self):
        return self.err

    def b200mpi_numa_node(self):
        return 0

    # ---- memory ---------------------------------------------------------------------------------
    def b200mpi_alloc(self, nbytes, out):
        buf = ctypes.create_string_buffer(max(int(nbytes), 1) + 64)
        a = (ctypes.addressof(buf) + 63) & ~63
        self.blocks[a] = buf
        _out(out).value = a
        return 0

    def b200mpi_free(self, p):
        self.blocks.pop(_addr(p), None)
        return 0

    b200mpi_host_alloc = b200mpi_alloc
    b200mpi_host_free = b200mpi_free

    def b200mpi_memcpy(self, dst, src, nbytes, kind):
        ctypes.memmove(_addr(dst), _addr(src), int(nbytes))
        return 0

    def b200mpi_heap_info(self, total, used, nvls):
        if total is not None:
            _out(total).value = 1 << 34
        if used is not None:
            _out(used).value = sum(len(b) for b in self.blocks.values())
        if nvls is not None:
            _out(nvls).value = 1 if self.nvls else 0
        return 0

    # ---- tuning / measurement --------------------------------------------------------------------
    def b200mpi_set_algo(self, coll, algo):
        self.algo[coll] = algo
        return 0

    def b200mpi_get_algo(self, coll, count, dtype):
        forced = self.algo[coll]
        nbytes = count * np.dtype(NP[dtype]).itemsize
        if coll == 0:
            if forced in (ALGO_NVLS, ALGO_HYBRID) and not self.nvls:
                forced = 0  # no multicast mapping in this fake world
            if forced == ALGO_HYBRID and self.params.get("hybrid_p2p_permille", 0) <= 0:
                forced = ALGO_NVLS
            if forced == ALGO_LL and nbytes > (256 << 10):
                forced = 0
            if forced:
                return forced
            return ALGO_LL if nbytes <= (256 << 10) else ALGO_TWOSHOT
        if coll == 2:
            return forced if forced == ALGO_RING or (forced == ALGO_NVLS and self.nvls) else ALGO_ONESHOT
        return forced if forced in (ALGO_ONESHOT, ALGO_TWOSHOT) else ALGO_TWOSHOT

    def b200mpi_set_param(self, name, value):
        self.params[name.decode()] = int(value)
        return 0

    def b200mpi_get_param(self, name, out):
        _out(out).value = self.params.get(name.decode(), 0)
        return 0

    def b200mpi_set_max_blocks(self, blocks):
        return 0

    def b200mpi_get_stream(self, out):
        _out(out).value = 1
        return 0

    def b200mpi_set_stream(self, s):
        return 0

    def b200mpi_timer_start(self):
        self.t0 = time.perf_counter()
        return 0

    def b200mpi_timer_stop(self, ms):
        _out(ms).value = (time.perf_counter() - self.t0) * 1e3
        return 0

    def b200mpi_stream_sync(self):
        return 0

    def b200mpi_launch_count(self):
        return self.launches

    def b200mpi_pcie_probe(self, nbytes, iters, up, down, both):
        time.sleep(float(os.environ.get("FAKE_SLOW_PROBE_S", "0")))  # lets a test outlast bench.py's --deadline
        _out(up).value, _out(down).value, _out(both).value = 50.0, 50.0, 40.0
        return 0

    def b200mpi_link_probe(self, nbytes, mode, iters, ms):
        if self.dist is not None:
            self.dist.barrier()
        _out(ms).value = nbytes / 700e9 * 1e3 * (2 if mode == 4 else 1)
        return 0

    # ---- transport helpers -------------------------------------------------------------------------
    def _gather(self, arr):
        import torch
        if self.n == 1:
            return [np.array(arr, copy=True)]
        t = torch.from_numpy(np.ascontiguousarray(arr).view(np.uint8).copy())
        outs = [torch.empty_like(t) for _ in range(self.n)]
        self.dist.all_gather(outs, t)
        return [o.numpy().view(arr.dtype) for o in outs]

    def _order(self, algo, count, dtype):
        if algo == ALGO_RING:
            return O.ORDER_RING
        if algo in (ALGO_NVLS, ALGO_HYBRID):
            return O.ORDER_F64  # "the switch picks the order": any order within the tolerance
        if algo == ALGO_ONESHOT:
            nvec = -(-count // (16 // np.dtype(NP[dtype]).itemsize))
            return O.ORDER_TREE if (self.n in (2, 4, 8) and nvec <= 4096) else O.ORDER_RANK
        return O.ORDER_RANK

    # ---- collectives --------------------------------------------------------------------------------
    def b200mpi_allreduce(self, send, recv, count, dtype, op, memkind):
        self.launches += 1
        x = _view(send, count, dtype)
        ins = self._gather(x)
        algo = self.b200mpi_get_algo(0, count, dtype) if self.n > 1 else ALGO_TWOSHOT
        res = O.allreduce(ins, op=op, order=self._order(algo, count, dtype)) if count else x
        _view(recv, count, dtype)[:] = res
        return 0

    def b200mpi_allreduce_async(self, send, recv, count, dtype, op):
        return self.b200mpi_allreduce(send, recv, count, dtype, op, 1)

    def b200mpi_bcast(self, buf, count, dtype, root, memkind):
        self.launches += 1
        v = _view(buf, count, dtype)
        if self.n > 1 and count:
            v[:] = self._gather(v)[root]
        return 0

    def b200mpi_bcast_async(self, buf, count, dtype, root):
        return self.b200mpi_bcast(buf, count, dtype, root, 1)

    def b200mpi_allgather(self, send, recv, count, dtype, memkind):
        self.launches += 1
        parts = self._gather(_view(send, count, dtype))
        _view(recv, count * self.n, dtype)[:] = np.concatenate(parts) if count else []
        return 0

    def b200mpi_allgather_async(self, send, recv, count, dtype):
        return self.b200mpi_allgather(send, recv, count, dtype, 1)

    def b200mpi_reduce_scatter(self, send, recv, count, dtype, op, memkind):
        self.launches += 1
        ins = self._gather(_view(send, count * self.n, dtype))
        _view(recv, count, dtype)[:] = O.reduce_scatter(ins, self.rank, op=op)
        return 0

    def b200mpi_barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        return 0

    # ---- point to point -------------------------------------------------------------------------------
    def b200mpi_send(self, buf, count, dtype, dest, tag, memkind):
        import torch
        self.dist.send(torch.tensor([count, dtype], dtype=torch.int64), dest, tag=tag)
        if count:
            self.dist.send(torch.from_numpy(_view(buf, count, dtype).view(np.uint8).copy()), dest, tag=tag)
        return 0

    def b200mpi_recv(self, buf, capacity, count_out, dtype, src, tag, memkind):
        import torch
        hdr = torch.zeros(2, dtype=torch.int64)
        self.dist.recv(hdr, src, tag=tag)
        count = int(hdr[0])
        if count_out is not None:
            _out(count_out).value = count
        if count:
            t = torch.empty(count * np.dtype(NP[dtype]).itemsize, dtype=torch.uint8)
            self.dist.recv(t, src, tag=tag)
            _view(buf, count, dtype)[:] = t.numpy().view(NP[dtype])
        return 0


def install():
    """Put a FakeLib where mpi_b200._lib.load() looks first.  Returns it."""
    from mpi_b200 import _lib
    fake = FakeLib()
    _lib._lib = fake
    os.environ.setdefault("B200MPI_FAKE", "1")
    return fake
