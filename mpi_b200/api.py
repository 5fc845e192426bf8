"""Python mirror of the reference's package-level API (/root/reference/mpi.go:56-182).

Same names, same argument meaning, same blocking behaviour:

    Init() Finalize() Rank() Size() Send(data, destination, tag) Receive(data, source, tag)
    Register(impl)  Interface  Raw  TagExists                       (the reference's surface)
    Recv  Bcast(data, root)  Allreduce(send, recv, op)  Allgather(send, recv)  Barrier()   (new)
    ReduceScatter(send, recv, op)  Reduce(send, recv, op, root)  Alltoall(send, recv)      (new)
    Isend(data, destination, tag) + Wait(destination, tag)     (the Send/Wait design commented out
                                                                in the reference, mpi.go:132-152)

Where Go returns `error`, Python raises MpiError (Go's "implementations may panic", mpi.go:20-21,
maps to the same exception).  `data` is type-switched exactly like the cgo shim
(go/mpi/cuda.go): float64 / float32 / int64 / uint8 numpy arrays and bytes-likes travel as typed
buffers; `str` is encoded to bytes first (the "anything else is gob-encoded" path of the reference);
DeviceSlice is memory from the library's device heap and takes the zero-copy path.
"""
import abc
import ctypes
import sys

import numpy as np

from . import _lib as L
from . import flags as _flags

SUM, MAX, MIN = L.SUM, L.MAX, L.MIN


class MpiError(Exception):
    def __init__(self, code, message):
        super().__init__("%s (b200mpi error %d)" % (message, code))
        self.code = code
        self.message = message


class TagExists(MpiError):
    """mpi.go:174-182: the tag already has a concurrent request between the two nodes."""

    def __init__(self, code, message, tag):
        super().__init__(code, message)
        self.Tag = tag


class Raw(bytes):
    """mpi.go:75: send as raw bytes, no encoding."""


_NP2DT = {np.dtype(np.uint8): L.U8, np.dtype(np.int64): L.I64, np.dtype(np.float32): L.F32, np.dtype(np.float64): L.F64}
_DT2NP = {v: k for k, v in _NP2DT.items()}


def _check(rc, tag=None):
    if rc == 0:
        return
    msg = L.last_error()
    if rc == L.ERR_TAG_EXISTS:
        raise TagExists(rc, msg, tag)
    raise MpiError(rc, msg)


class DeviceSlice:
    """`count` elements of `dtype` in this rank's peer-mapped device heap (b200mpi_alloc)."""

    def __init__(self, count, dtype, ptr=None, owner=True):
        self.dtype = np.dtype(dtype)
        if self.dtype not in _NP2DT:
            raise TypeError("unsupported element type %s" % self.dtype)
        self.count = int(count)
        self._owner = owner
        if ptr is None:
            p = ctypes.c_void_p()
            _check(L.load().b200mpi_alloc(max(self.nbytes, 1), ctypes.byref(p)))
            ptr = p.value
        self.ptr = ptr

    @property
    def nbytes(self):
        return self.count * self.dtype.itemsize

    def __len__(self):
        return self.count

    def __getitem__(self, sl):
        if not isinstance(sl, slice):
            raise TypeError("DeviceSlice supports slicing only")
        start, stop, step = sl.indices(self.count)
        if step != 1:
            raise ValueError("DeviceSlice slices must be contiguous")
        return DeviceSlice(max(stop - start, 0), self.dtype, self.ptr + start * self.dtype.itemsize, owner=False)

    def copy_from_host(self, arr):
        a = np.ascontiguousarray(arr, dtype=self.dtype).reshape(-1)
        if a.size != self.count:
            raise ValueError("size mismatch: %d vs %d" % (a.size, self.count))
        if a.size:
            _check(L.load().b200mpi_memcpy(self.ptr, a.ctypes.data, a.nbytes, 0))
        return self

    def to_host(self):
        out = np.empty(self.count, dtype=self.dtype)
        if out.size:
            _check(L.load().b200mpi_memcpy(out.ctypes.data, self.ptr, out.nbytes, 1))
        return out

    def free(self):
        if self._owner and self.ptr:
            _check(L.load().b200mpi_free(self.ptr))
        self.ptr = None


def Alloc(count, dtype):
    return DeviceSlice(count, dtype)


def _describe(data):
    """-> (pointer, count, dtype code, memkind, keepalive, kind tag)."""
    if isinstance(data, DeviceSlice):
        return data.ptr, data.count, _NP2DT[data.dtype], L.DEVICE, data, "device"
    if isinstance(data, np.ndarray):
        if data.dtype not in _NP2DT:
            raise TypeError("mpi: unsupported element type %s (want float64, float32, int64 or uint8)" % data.dtype)
        if not data.flags["C_CONTIGUOUS"]:
            raise ValueError("mpi: array must be contiguous")
        return (data.ctypes.data if data.size else None), data.size, _NP2DT[data.dtype], L.HOST, data, "ndarray"
    if isinstance(data, str):
        b = np.frombuffer(data.encode("utf-8"), dtype=np.uint8)
        return (b.ctypes.data if b.size else None), b.size, L.U8, L.HOST, b, "str"
    if isinstance(data, (bytes, bytearray, memoryview)):
        b = np.frombuffer(data, dtype=np.uint8)
        return (b.ctypes.data if b.size else None), b.size, L.U8, L.HOST, b, "bytes"
    raise TypeError("mpi: cannot send a %s" % type(data).__name__)


class Interface(abc.ABC):
    """mpi.go:163-170."""

    @abc.abstractmethod
    def Init(self): ...

    @abc.abstractmethod
    def Finalize(self): ...

    @abc.abstractmethod
    def Rank(self): ...

    @abc.abstractmethod
    def Size(self): ...

    @abc.abstractmethod
    def Send(self, data, destination, tag): ...

    @abc.abstractmethod
    def Receive(self, data, source, tag): ...


class Cuda(Interface):
    """The B200 implementation behind the facade; plays the role of `Network`
    (network.go:25-39): public fields win over flags (network.go:69-90)."""

    def __init__(self, Addr="", Addrs=None, Timeout=0, Password="", Gpu=None):
        self.Addr = Addr
        self.Addrs = list(Addrs) if Addrs else []
        self.Timeout = Timeout  # nanoseconds
        self.Password = Password
        self.Gpu = Gpu

    # -- lifecycle ---------------------------------------------------------------------------
    def _use_flags(self):
        f = _flags.parse(sys.argv[1:])
        if not self.Password:
            self.Password = f.password
        if not self.Timeout:
            self.Timeout = f.inittimeout
        if not self.Addr:
            self.Addr = f.addr
        if not self.Addrs:
            self.Addrs = list(f.alladdr)
        if self.Gpu is None:
            self.Gpu = f.gpu

    def Init(self):
        self._use_flags()
        lib = L.load()
        _check(lib.b200mpi_init(self.Addr.encode(), ",".join(self.Addrs).encode(), self.Password.encode(),
                                int(self.Timeout), int(-1 if self.Gpu is None else self.Gpu)))

    def Finalize(self):
        _check(L.load().b200mpi_finalize())

    def Rank(self):
        return L.load().b200mpi_rank()

    def Size(self):
        return L.load().b200mpi_size()

    # -- point to point ----------------------------------------------------------------------
    def Send(self, data, destination, tag):
        ptr, count, dt, kind, keep, _ = _describe(data)
        _check(L.load().b200mpi_send(ptr, count, dt, int(destination), int(tag), kind), tag)

    def Isend(self, data, destination, tag):
        """mpi.go:132-143 (commented design): returns once `data` may be modified again, without
        waiting for the receiver; Wait(destination, tag) collects the confirmation."""
        ptr, count, dt, kind, keep, _ = _describe(data)
        _check(L.load().b200mpi_isend(ptr, count, dt, int(destination), int(tag), kind), tag)

    def Wait(self, destination, tag):
        """mpi.go:146-152 (commented design)."""
        _check(L.load().b200mpi_wait(int(destination), int(tag)), tag)

    def Receive(self, data, source, tag):
        """Fills `data` and returns the received value.  numpy arrays and bytearrays are filled
        in place when large enough (a view of the received length is returned), otherwise a new
        object of the sent length is returned -- the analogue of gob resizing the destination
        slice (network.go:597).  Pass the type `str` / `bytes` (or an instance) to receive one."""
        lib = L.load()
        n = ctypes.c_size_t(0)
        if isinstance(data, DeviceSlice):
            _check(lib.b200mpi_recv(data.ptr, data.count, ctypes.byref(n), _NP2DT[data.dtype], int(source), int(tag), L.DEVICE), tag)
            return data[: n.value]
        want_str = data is str or isinstance(data, str)
        want_bytes = data is bytes or isinstance(data, bytes)
        if want_str or want_bytes:
            buf = np.empty(256, dtype=np.uint8)
        elif isinstance(data, bytearray):
            buf = np.frombuffer(data, dtype=np.uint8) if len(data) else np.empty(0, dtype=np.uint8)
        elif isinstance(data, np.ndarray):
            if data.dtype not in _NP2DT:
                raise TypeError("mpi: unsupported element type %s" % data.dtype)
            buf = data.reshape(-1)
        else:
            raise TypeError("mpi: cannot receive into a %s" % type(data).__name__)
        dt = _NP2DT[buf.dtype]
        rc = lib.b200mpi_recv(buf.ctypes.data if buf.size else None, buf.size, ctypes.byref(n), dt, int(source), int(tag), L.HOST)
        if rc == L.ERR_TRUNCATE:  # message is still posted: retry with the right size
            buf = np.empty(n.value, dtype=buf.dtype)
            rc = lib.b200mpi_recv(buf.ctypes.data, buf.size, ctypes.byref(n), dt, int(source), int(tag), L.HOST)
            if isinstance(data, bytearray) and rc == 0:
                data[:] = buf.tobytes()
                return data
        _check(rc, tag)
        out = buf[: n.value]
        if want_str:
            return out.tobytes().decode("utf-8")
        if want_bytes:
            return out.tobytes()
        if isinstance(data, bytearray):
            del data[n.value:]
            return data
        return out

    # -- collectives (the optional upgrade interface hinted at by mpi.go:69-71) ---------------
    def Bcast(self, data, root):
        ptr, count, dt, kind, keep, tagk = _describe(data)
        if tagk in ("str", "bytes") and not isinstance(data, (bytearray, memoryview)):
            raise TypeError("mpi: Bcast needs a mutable buffer (numpy array, bytearray or DeviceSlice)")
        _check(L.load().b200mpi_bcast(ptr, count, dt, int(root), kind))
        return data

    def Allreduce(self, send, recv, op=SUM):
        sp, sc, sdt, sk, _, _ = _describe(send)
        rp, rc_, rdt, rk, _, _ = _describe(recv)
        if sc != rc_ or sdt != rdt or sk != rk:
            raise ValueError("mpi: Allreduce send and recv must have the same length, type and memory kind")
        _check(L.load().b200mpi_allreduce(sp, rp, sc, sdt, int(op), sk))
        return recv

    def Allgather(self, send, recv):
        sp, sc, sdt, sk, _, _ = _describe(send)
        rp, rc_, rdt, rk, _, _ = _describe(recv)
        if rc_ != sc * self.Size() or sdt != rdt or sk != rk:
            raise ValueError("mpi: Allgather recv must hold Size()*len(send) elements of the same type and memory kind")
        _check(L.load().b200mpi_allgather(sp, rp, sc, sdt, sk))
        return recv

    def ReduceScatter(self, send, recv, op=SUM):
        sp, sc, sdt, sk, _, _ = _describe(send)
        rp, rc_, rdt, rk, _, _ = _describe(recv)
        if sc != rc_ * self.Size() or sdt != rdt or sk != rk:
            raise ValueError("mpi: ReduceScatter send must hold Size()*len(recv) elements of the same type and memory kind")
        _check(L.load().b200mpi_reduce_scatter(sp, rp, rc_, sdt, int(op), sk))
        return recv

    def Reduce(self, send, recv, op=SUM, root=0):
        sp, sc, sdt, sk, _, _ = _describe(send)
        if recv is None:
            rp = None
        else:
            rp, rc_, rdt, rk, _, _ = _describe(recv)
            if sc != rc_ or sdt != rdt or sk != rk:
                raise ValueError("mpi: Reduce send and recv must have the same length, type and memory kind")
        _check(L.load().b200mpi_reduce(sp, rp, sc, sdt, int(op), int(root), sk))
        return recv

    def Alltoall(self, send, recv):
        sp, sc, sdt, sk, _, _ = _describe(send)
        rp, rc_, rdt, rk, _, _ = _describe(recv)
        n = self.Size()
        if sc != rc_ or sc % max(n, 1) or sdt != rdt or sk != rk:
            raise ValueError("mpi: Alltoall send and recv must both hold Size() equal blocks of the same type and memory kind")
        _check(L.load().b200mpi_alltoall(sp, rp, sc // n, sdt, sk))
        return recv

    def Barrier(self):
        _check(L.load().b200mpi_barrier())


# ---- package-level facade: mpi.go:56-159 -------------------------------------------------------
_mpier = Cuda()
_register_called = False


def Register(impl):
    """mpi.go:61-67: install an implementation; a second call panics."""
    global _mpier, _register_called
    _mpier = impl
    if _register_called:
        raise RuntimeError("register called more than once")
    _register_called = True


def _reset_for_tests(impl=None):
    global _mpier, _register_called
    _mpier = impl if impl is not None else Cuda()
    _register_called = False


def Init():
    return _mpier.Init()


def Finalize():
    return _mpier.Finalize()


def Rank():
    return _mpier.Rank()


def Size():
    return _mpier.Size()


def Send(data, destination, tag):
    return _mpier.Send(data, destination, tag)


def Receive(data, source, tag):
    return _mpier.Receive(data, source, tag)


Recv = Receive  # north_star spells it Recv; the reference spells it Receive (mpi.go:157)


def _collective(name):
    fn = getattr(_mpier, name, None)
    if fn is None:
        raise MpiError(L.ERR_UNSUPPORTED, "registered implementation has no %s" % name)
    return fn


def Bcast(data, root):
    return _collective("Bcast")(data, root)


def Allreduce(send, recv, op=SUM):
    return _collective("Allreduce")(send, recv, op)


def Allgather(send, recv):
    return _collective("Allgather")(send, recv)


def Barrier():
    return _collective("Barrier")()


def ReduceScatter(send, recv, op=SUM):
    return _collective("ReduceScatter")(send, recv, op)


def Reduce(send, recv, op=SUM, root=0):
    return _collective("Reduce")(send, recv, op, root)


def Alltoall(send, recv):
    return _collective("Alltoall")(send, recv)


def Isend(data, destination, tag):
    return _collective("Isend")(data, destination, tag)


def Wait(destination, tag):
    return _collective("Wait")(destination, tag)
