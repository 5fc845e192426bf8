"""Host-side logic that mirrors the reference: flag parsing (flags.go), rank assignment and
handshake errors (network.go:94-109, 343-351), launcher argv (gompirun.go:77-83), Register
(mpi.go:61-67)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import mpi_b200 as mpi
from mpi_b200 import _lib as L
from mpi_b200 import flags, launcher

from _launch import free_ports

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---- flags.go ----------------------------------------------------------------------------------
def test_flag_defaults():
    f = flags.parse([])
    assert (f.addr, f.alladdr, f.inittimeout, f.protocol, f.password, f.gpu) == ("", [], 0, "tcp", "", -1)


def test_flag_forms_and_append():
    f = flags.parse(["-mpi-addr", ":5001", "--mpi-alladdr=:5000,:5001", "-mpi-alladdr", ":5003", "-mpi-inittimeout=1.5s",
                     "-mpi-password", "s3", "-mpi-protocol=tcp", "-other", "x", "-mpi-gpu", "3"])
    assert f.addr == ":5001"
    assert f.alladdr == [":5000", ":5001", ":5003"]  # AddrsFlag.Set appends (flags.go:22-27)
    assert f.inittimeout == 1_500_000_000
    assert f.password == "s3" and f.gpu == 3
    assert f.rest == ["-other", "x"]
    assert flags.parse(["-mpi-alladdr", "a,,b"]).alladdr == ["a", "", "b"]  # strings.Split keeps empties


@pytest.mark.parametrize("text,ns", [("0", 0), ("300ms", 300_000_000), ("1h2m3.5s", 3723_500_000_000), ("2us", 2000),
                                     ("1.5h", 5400_000_000_000), ("-1s", -1_000_000_000), ("1m30s", 90_000_000_000), (".5s", 500_000_000)])
def test_parse_duration(text, ns):
    assert flags.parse_duration(text) == ns


@pytest.mark.parametrize("bad", ["", "1", "s", "1x", "1 s", "abc"])
def test_parse_duration_rejects(bad):
    with pytest.raises(ValueError):
        flags.parse_duration(bad)


# ---- gompirun ------------------------------------------------------------------------------------
def test_launcher_addresses_and_argv():
    addrs = launcher.addresses(4)
    assert addrs == [":6000", ":6001", ":6002", ":6003"]  # gompirun.go:45-51
    assert sorted(addrs) == addrs                          # rank == launch index
    argv = launcher.child_argv("prog", ["-x", "1"], addrs[2], addrs, gpu=2)
    assert argv == ["prog", "-x", "1", "-mpi-addr", ":6002", "-mpi-alladdr", ":6000,:6001,:6002,:6003", "-mpi-gpu", "2"]
    assert launcher.main(["x"]) == 2 and launcher.main(["zero", "prog"]) == 2 and launcher.main(["0", "prog"]) == 2


# ---- Register ------------------------------------------------------------------------------------
def test_register_twice_panics():
    class Fake(mpi.Interface):
        def Init(self): return None
        def Finalize(self): return None
        def Rank(self): return 7
        def Size(self): return 9
        def Send(self, data, destination, tag): return None
        def Receive(self, data, source, tag): return None
    try:
        mpi.api._reset_for_tests()
        mpi.Register(Fake())
        assert mpi.Rank() == 7 and mpi.Size() == 9
        with pytest.raises(mpi.MpiError):
            mpi.Allreduce(np.zeros(1), np.zeros(1))  # implementation without the collective upgrade
        with pytest.raises(RuntimeError, match="more than once"):
            mpi.Register(Fake())
    finally:
        mpi.api._reset_for_tests()


# ---- bootstrap errors (control plane only; each case in a fresh process) ---------------------------
def _init(addr, alladdr, password="", timeout_ns=2_000_000_000):
    code = (
        "import sys, mpi_b200 as mpi\n"
        "from mpi_b200 import _lib as L\n"
        "lib = L.load()\n"
        "rc = lib.b200mpi_init(%r.encode(), %r.encode(), %r.encode(), %d, -2)\n"
        "print('RC', rc, '|', L.last_error(), '| rank', lib.b200mpi_rank(), 'size', lib.b200mpi_size())\n"
        "if rc == 0: lib.b200mpi_finalize()\n") % (addr, alladdr, password, timeout_ns)
    return subprocess.Popen([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                            cwd=ROOT, env=dict(os.environ, PYTHONPATH=ROOT))


def _out(p):
    return p.communicate(timeout=60)[0]


def test_default_is_one_node_on_5000():
    out = _out(_init("", ""))
    assert "RC 0" in out and "rank 0 size 1" in out  # network.go:55-58


def test_duplicate_addresses_rejected():
    out = _out(_init(":7001", ":7001,:7001"))
    assert "RC -3" in out and "not unique" in out  # network.go:96-98


def test_local_address_must_be_listed():
    out = _out(_init(":7009", ":7001,:7002"))
    assert "RC -3" in out and "not in global list" in out  # network.go:103-105


def test_rank_is_index_in_sorted_address_list():
    p = free_ports(3)
    a = ["127.0.0.1:%d" % x for x in p]
    shuffled = ",".join([a[2], a[0], a[1]])
    procs = [_init(a[i], shuffled, timeout_ns=20_000_000_000) for i in (1, 2, 0)]
    outs = [_out(q) for q in procs]
    assert "rank 1 size 3" in outs[0] and "rank 2 size 3" in outs[1] and "rank 0 size 3" in outs[2], outs


def test_password_mismatch_is_rejected():
    p = free_ports(2)
    a = ["127.0.0.1:%d" % x for x in p]
    procs = [_init(a[0], ",".join(a), "alpha", 5_000_000_000), _init(a[1], ",".join(a), "beta", 5_000_000_000)]
    outs = [_out(q) for q in procs]
    assert all("RC 0" not in o for o in outs), outs
    assert any("RC -4" in o and "password" in o for o in outs), outs  # network.go:343-346


def test_init_times_out_when_a_peer_never_shows_up():
    p = free_ports(2)
    a = ["127.0.0.1:%d" % x for x in p]
    out = _out(_init(a[0], ",".join(a), "", 700_000_000))
    assert "RC -5" in out and "timed out" in out  # network.go:223-231


def test_more_than_eight_ranks_rejected():
    addrs = ",".join(":%d" % (7100 + i) for i in range(9))
    out = _out(_init(":7100", addrs))
    assert "RC -3" in out and "at most 8" in out


# ---- the collectives added along the reference's conventions: argument checks before any device work --------
def test_new_collectives_check_shapes_before_touching_the_library():
    c = mpi.Cuda()
    a4, a8 = np.zeros(4, dtype=np.float32), np.zeros(8, dtype=np.float32)
    with pytest.raises(ValueError):
        c.Allreduce(a4, a8)                       # lengths differ
    with pytest.raises(ValueError):
        c.Allreduce(a4, np.zeros(4, dtype=np.float64))  # element types differ
    with pytest.raises(ValueError):
        c.Alltoall(a4, a8)                        # send and recv must match
    with pytest.raises(ValueError):
        c.Reduce(a4, a8, mpi.SUM, 0)
    with pytest.raises(TypeError):
        c.Allreduce([1.0, 2.0], a4)               # not a typed buffer: never silently converted
    with pytest.raises(TypeError):
        c.Bcast("immutable", 0)                   # Bcast needs a mutable buffer
    with pytest.raises(ValueError):
        mpi.DeviceSlice.__getitem__(mpi.DeviceSlice(8, np.float32, ptr=4096, owner=False), slice(0, 8, 2))  # contiguous slices only


def test_data_calls_before_init_report_not_initialised():
    lib = L.load()
    assert lib.b200mpi_rank() == -1 and lib.b200mpi_size() == 0       # mpi.go:110-118
    x = np.zeros(4, dtype=np.float32)
    for call in (lambda: mpi.Allreduce(x, x), lambda: mpi.Send(x, 0, 1), lambda: mpi.Isend(x, 0, 1), lambda: mpi.Wait(0, 1),
                 lambda: mpi.Reduce(x, x, mpi.SUM, 0), lambda: mpi.Barrier()):
        with pytest.raises(mpi.MpiError) as e:
            call()
        assert e.value.code == L.ERR_NOT_INIT
