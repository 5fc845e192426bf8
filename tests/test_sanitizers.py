"""ThreadSanitizer and AddressSanitizer over the host-side control plane (the
reference has no race detection at all; its handshake runs two goroutines per rank)."""
import os
import subprocess
import tempfile

import pytest

from _launch import free_ports

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = [os.path.join(ROOT, "tests", "cpp", "ctrl_sanitize.cpp"), os.path.join(ROOT, "mpi_b200", "csrc", "ctrl.cpp")]


def _build(flag):
    exe = os.path.join(tempfile.mkdtemp(prefix="b200mpi-san-"), "ctrl_" + flag)
    err = ""
    for cxx in ("/usr/bin/g++", "g++", os.environ.get("CXX", "c++")):  # not every toolchain ships the sanitizer runtimes
        cmd = [cxx, "-std=c++17", "-O1", "-g", "-fsanitize=" + flag, "-fno-omit-frame-pointer", "-pthread", "-o", exe] + SRC
        try:
            r = subprocess.run(cmd, capture_output=True, text=True)
        except FileNotFoundError:
            continue
        if r.returncode == 0:
            return exe
        err = r.stderr[-300:]
    pytest.skip("cannot build with -fsanitize=%s here: %s" % (flag, err))


@pytest.mark.parametrize("flag", ["thread", "address"])
@pytest.mark.parametrize("n", [1, 3])
def test_control_plane_under_sanitizer(flag, n):
    exe = _build(flag)
    ports = free_ports(n)
    addrs = ["127.0.0.1:%d" % p for p in ports]
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1 exitcode=66", ASAN_OPTIONS="detect_leaks=1:exitcode=67")
    procs = [subprocess.Popen([exe, a, ",".join(addrs), "pw"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env) for a in addrs]
    outs = [p.communicate(timeout=120) for p in procs]
    for r, (p, (so, se)) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, "rank %d rc=%d\n%s\n%s" % (r, p.returncode, so, se[-3000:])
        assert "ok" in so and "WARNING: ThreadSanitizer" not in se and "ERROR: AddressSanitizer" not in se


@pytest.mark.parametrize("flag", ["thread", "address"])
def test_host_helpers_under_sanitizer(flag):
    """mpi_b200/csrc/hostutil.h: the copy pool that stages pageable host slices (submit / wait / the
    waiter helping / shutdown) and the NUMA helpers, race- and memory-checked on the CPU."""
    exe = os.path.join(tempfile.mkdtemp(prefix="b200mpi-san-"), "hostutil_" + flag)
    src = os.path.join(ROOT, "tests", "cpp", "hostutil_test.cpp")
    built = False
    for cxx in ("/usr/bin/g++", "g++", os.environ.get("CXX", "c++")):
        try:
            r = subprocess.run([cxx, "-std=c++17", "-O1", "-g", "-fsanitize=" + flag, "-fno-omit-frame-pointer", "-pthread", "-o", exe, src], capture_output=True, text=True)
        except FileNotFoundError:
            continue
        if r.returncode == 0:
            built = True
            break
    if not built:
        pytest.skip("cannot build with -fsanitize=%s here" % flag)
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=1 exitcode=66", ASAN_OPTIONS="detect_leaks=1:exitcode=67")
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0 and "hostutil ok" in out.stdout, out.stdout + out.stderr[-3000:]
    assert "WARNING: ThreadSanitizer" not in out.stderr and "ERROR: AddressSanitizer" not in out.stderr

