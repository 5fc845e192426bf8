"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes front end of liboracle.so (collectives.c, gob.c, ref_tcp.c) plus a numpy twin of the
collective semantics.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs import this module; mpi_b200 never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")

U8, I64, F32, F64, STRING = 0, 1, 2, 3, 4
SUM, MAX, MIN = 0, 1, 2
ORDER_RANK, ORDER_TREE, ORDER_RING, ORDER_F64 = 0, 1, 2, 3
COLL_ALLREDUCE, COLL_BCAST, COLL_ALLGATHER, COLL_PINGPONG, COLL_ALLREDUCE_NAIVE = 0, 1, 2, 3, 4
NP2DT = {np.dtype(np.uint8): U8, np.dtype(np.int64): I64, np.dtype(np.float32): F32, np.dtype(np.float64): F64}
DT2NP = {v: k for k, v in NP2DT.items()}

_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = ctypes.CDLL(_SO)
        c = ctypes
        L.oracle_allreduce.argtypes = [c.c_int, c.c_int, c.c_int, c.c_int, c.c_size_t, c.POINTER(c.c_void_p), c.c_void_p]
        L.oracle_allgather.argtypes = [c.c_int, c.c_int, c.c_size_t, c.POINTER(c.c_void_p), c.c_void_p]
        L.oracle_reduce_scatter.argtypes = [c.c_int, c.c_int, c.c_int, c.c_int, c.c_int, c.c_size_t, c.POINTER(c.c_void_p), c.c_void_p]
        L.oracle_alltoall.argtypes = [c.c_int, c.c_int, c.c_int, c.c_size_t, c.POINTER(c.c_void_p), c.c_void_p]
        L.oracle_splitmix64.argtypes = [c.c_uint64, c.c_uint64]
        L.oracle_splitmix64.restype = c.c_uint64
        for f in ("oracle_fill_i64", "oracle_fill_f32", "oracle_fill_f64"):
            getattr(L, f).argtypes = [c.c_uint64, c.c_size_t, c.c_void_p]
            getattr(L, f).restype = None
        for f in ("oracle_fill_i64_at", "oracle_fill_f32_at", "oracle_fill_f64_at"):
            getattr(L, f).argtypes = [c.c_uint64, c.c_size_t, c.c_size_t, c.c_void_p]
            getattr(L, f).restype = None
        for f in ("gob_put_uint", "gob_put_int", "gob_put_float"):
            getattr(L, f).restype = c.c_size_t
        L.gob_put_uint.argtypes = [c.c_void_p, c.c_uint64]
        L.gob_put_int.argtypes = [c.c_void_p, c.c_int64]
        L.gob_put_float.argtypes = [c.c_void_p, c.c_double]
        L.gob_payload_bound.argtypes = [c.c_int, c.c_size_t]
        L.gob_payload_bound.restype = c.c_size_t
        L.gob_encode_payload.argtypes = [c.c_int, c.c_void_p, c.c_size_t, c.c_void_p]
        L.gob_encode_payload.restype = c.c_size_t
        L.gob_decode_payload.argtypes = [c.c_int, c.c_void_p, c.c_size_t, c.c_void_p, c.c_size_t, c.POINTER(c.c_size_t)]
        L.gob_encode_envelope.argtypes = [c.c_int64, c.c_void_p, c.c_size_t, c.c_void_p]
        L.gob_encode_envelope.restype = c.c_size_t
        L.gob_put_struct_typedef.argtypes = [c.c_void_p, c.c_char_p, c.c_int, c.c_int, c.POINTER(c.c_char_p), c.POINTER(c.c_int)]
        L.gob_put_struct_typedef.restype = c.c_size_t
        L.ref_bench.argtypes = [c.c_int, c.c_int, c.c_int, c.c_size_t, c.c_int, c.c_int, c.c_uint64, c.POINTER(c.c_double), c.c_void_p]
        L.ref_bench_procs.argtypes = L.ref_bench.argtypes
        _lib = L
    return _lib


# ---- inputs ----------------------------------------------------------------------------------
def fill(dtype, seed, count):
    """The synthetic buffers of SURVEY.md 8(d): splitmix64(seed, i) as int64 'indices', or mapped to
    uniform [0,1) floats (24 random bits for f32, 53 for f64)."""
    dt = np.dtype(dtype)
    out = np.empty(count, dtype=dt)
    fn = {I64: "oracle_fill_i64", F32: "oracle_fill_f32", F64: "oracle_fill_f64"}[NP2DT[dt]]
    getattr(lib(), fn)(ctypes.c_uint64(seed & (2**64 - 1)), count, out.ctypes.data)
    return out


def fill_at(dtype, seed, start, count):
    """Elements [start, start+count) of fill(dtype, seed, .): block-wise checks of large buffers."""
    dt = np.dtype(dtype)
    out = np.empty(count, dtype=dt)
    fn = {I64: "oracle_fill_i64_at", F32: "oracle_fill_f32_at", F64: "oracle_fill_f64_at"}[NP2DT[dt]]
    getattr(lib(), fn)(ctypes.c_uint64(seed & (2**64 - 1)), start, count, out.ctypes.data)
    return out


# ---- C oracle --------------------------------------------------------------------------------
def allreduce(inputs, op=SUM, order=ORDER_RANK):
    n = len(inputs)
    arrs = [np.ascontiguousarray(a) for a in inputs]
    dt = arrs[0].dtype
    out = np.empty_like(arrs[0])
    ptrs = (ctypes.c_void_p * n)(*[a.ctypes.data for a in arrs])
    rc = lib().oracle_allreduce(NP2DT[dt], op, order, n, arrs[0].size, ptrs, out.ctypes.data)
    if rc:
        raise ValueError("oracle_allreduce: unsupported dtype %s" % dt)
    return out


def allgather(inputs):
    return np.concatenate([np.ascontiguousarray(a).reshape(-1) for a in inputs])


def bcast(root_buf):
    return np.array(root_buf, copy=True)


def reduce_scatter(inputs, me, op=SUM, order=ORDER_RANK):
    """inputs[r] holds n blocks; returns what rank `me` receives (block `me` reduced over ranks)."""
    n = len(inputs)
    arrs = [np.ascontiguousarray(a) for a in inputs]
    dt = arrs[0].dtype
    count = arrs[0].size // n
    out = np.empty(count, dtype=dt)
    ptrs = (ctypes.c_void_p * n)(*[a.ctypes.data for a in arrs])
    if lib().oracle_reduce_scatter(NP2DT[dt], op, order, n, me, count, ptrs, out.ctypes.data):
        raise ValueError("oracle_reduce_scatter: bad arguments")
    return out


def alltoall(inputs, me):
    """inputs[r] holds n blocks; returns what rank `me` receives: block `me` of every rank, in rank order."""
    n = len(inputs)
    arrs = [np.ascontiguousarray(a) for a in inputs]
    dt = arrs[0].dtype
    count = arrs[0].size // n
    out = np.empty(count * n, dtype=dt)
    ptrs = (ctypes.c_void_p * n)(*[a.ctypes.data for a in arrs])
    lib().oracle_alltoall(NP2DT[dt], n, me, count, ptrs, out.ctypes.data)
    return out


def reduce_scatter_np(inputs, me, op=SUM, order=ORDER_RANK):
    n = len(inputs)
    count = np.asarray(inputs[0]).size // n
    return allreduce_np([np.asarray(a).reshape(-1)[me * count:(me + 1) * count] for a in inputs], op=op, order=order)


def alltoall_np(inputs, me):
    n = len(inputs)
    count = np.asarray(inputs[0]).size // n
    return np.concatenate([np.asarray(a).reshape(-1)[me * count:(me + 1) * count] for a in inputs])


# ---- numpy twin (independent restatement used to cross-check the C oracle) --------------------
def _apply(op, a, b):
    if op == SUM:
        if a.dtype == np.int64:
            return (a.view(np.uint64) + b.view(np.uint64)).view(np.int64)
        return a + b
    if op == MAX:
        return np.where(b > a, b, a)
    return np.where(b < a, b, a)


def allreduce_np(inputs, op=SUM, order=ORDER_RANK):
    n = len(inputs)
    xs = [np.ascontiguousarray(a).reshape(-1) for a in inputs]
    dt = xs[0].dtype
    count = xs[0].size
    with np.errstate(over="ignore", invalid="ignore"):
        if order == ORDER_TREE and n in (2, 4, 8):
            x = list(xs)
            m = 1
            while m < n:
                for r in range(0, n, 2 * m):
                    x[r] = _apply(op, x[r], x[r + m])
                m *= 2
            return x[0]
        if order == ORDER_F64 and op == SUM and dt != np.int64:
            wide = np.longdouble if dt == np.float64 else np.float64
            acc = xs[0].astype(wide)
            for r in range(1, n):
                acc = acc + xs[r].astype(wide)
            return acc.astype(dt)
        if order == ORDER_RING:
            epv = 16 // dt.itemsize
            groups = count // epv
            per = -(-groups // n) if groups else 0
            out = np.empty_like(xs[0])
            for c in range(n):
                lo = min(per * c, groups) * epv
                hi = min(per * (c + 1), groups) * epv
                acc = xs[c][lo:hi]
                for k in range(1, n):
                    acc = _apply(op, acc, xs[(c + k) % n][lo:hi])
                out[lo:hi] = acc
            t = groups * epv
            acc = xs[0][t:]
            for r in range(1, n):
                acc = _apply(op, acc, xs[r][t:])
            out[t:] = acc
            return out
        acc = xs[0]
        for r in range(1, n):
            acc = _apply(op, acc, xs[r])
        return np.array(acc, copy=True)


# ---- gob ---------------------------------------------------------------------------------------
def gob_encode(arr_or_bytes):
    if isinstance(arr_or_bytes, str):
        raw = np.frombuffer(arr_or_bytes.encode(), dtype=np.uint8)
        dt = STRING
    elif isinstance(arr_or_bytes, (bytes, bytearray)):
        raw = np.frombuffer(bytes(arr_or_bytes), dtype=np.uint8)
        dt = U8
    else:
        raw = np.ascontiguousarray(arr_or_bytes)
        dt = NP2DT[raw.dtype]
    out = np.empty(lib().gob_payload_bound(dt, raw.size), dtype=np.uint8)
    n = lib().gob_encode_payload(dt, raw.ctypes.data if raw.size else None, raw.size, out.ctypes.data)
    return out[:n].tobytes()


def gob_decode(stream, dtype, capacity):
    dt = STRING if dtype is str else NP2DT[np.dtype(dtype)]
    npdt = np.uint8 if dtype is str else np.dtype(dtype)
    out = np.empty(capacity, dtype=npdt)
    src = np.frombuffer(stream, dtype=np.uint8)
    cnt = ctypes.c_size_t(0)
    rc = lib().gob_decode_payload(dt, src.ctypes.data, src.size, out.ctypes.data, capacity, ctypes.byref(cnt))
    if rc:
        raise ValueError("gob_decode_payload rc=%d (count %d)" % (rc, cnt.value))
    out = out[: cnt.value]
    return out.tobytes().decode() if dtype is str else out


# ---- restated reference TCP path ---------------------------------------------------------------
def ref_bench(coll, dtype, n, count, iters=3, warmup=1, seed=0xB2000000, processes=False):
    """Times the restated reference path; returns (seconds_per_iter, rank 0's final buffer).
    processes=True: one OS process per rank (as gompirun starts them) instead of threads."""
    dt = np.dtype(dtype)
    total = count * n if coll == COLL_ALLGATHER else count
    out = np.empty(total, dtype=dt)
    secs = ctypes.c_double(0)
    fn = lib().ref_bench_procs if processes else lib().ref_bench
    rc = fn(coll, NP2DT[dt], n, count, iters, warmup, ctypes.c_uint64(seed), ctypes.byref(secs), out.ctypes.data)
    if rc:
        raise RuntimeError("ref_bench failed rc=%d" % rc)
    return secs.value, out
