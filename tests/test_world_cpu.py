"""CPU plumbing for worlds of more than one rank: bootstrap over TCP loopback (the reference's
configs[0], examples/helloworld world_size=2, needs no GPU for this part), control-plane barrier,
data calls refusing without a device, and the torchrun-style rank mapping bench.py uses."""
import os
import subprocess
import sys

import pytest

from _launch import assert_world_ok, run_world

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("n", [1, 2, 4, 8])
def test_control_plane_world(n):
    res = run_world(n, "control_only", args=["--control-only"], timeout=120)
    assert_world_ok(res)
    assert [r["rank_reported"] for r in res] == list(range(n))
    assert all(r["size_reported"] == n for r in res)


GLOO_RANK = r'''
import os, sys
sys.path.insert(0, %(root)r)
import torch.distributed as dist
import bench
from mpi_b200 import _lib as L
rank, world, local, addr, addrs = bench.world_from_env(None)
dist.init_process_group("gloo", rank=rank, world_size=world)
lib = L.load()
rc = lib.b200mpi_init(addr.encode(), ",".join(addrs).encode(), b"", 30 * 10**9, -2)
assert rc == 0, L.last_error()
mine = [lib.b200mpi_rank(), lib.b200mpi_size()]
got = [None] * world
dist.all_gather_object(got, mine)
assert got == [[r, world] for r in range(world)], got   # library rank == torchrun RANK on every rank
assert lib.b200mpi_barrier() == 0
lib.b200mpi_finalize()
dist.destroy_process_group()
print("GLOO-OK", rank)
'''


def test_torchrun_env_maps_to_library_ranks_world_of_2_gloo():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), LOCAL_RANK=str(r), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), PYTHONPATH=ROOT)
        procs.append(subprocess.Popen([sys.executable, "-c", GLOO_RANK % {"root": ROOT}], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all("GLOO-OK" in o for o in outs), outs


def test_reference_arm_prints_the_contract_line():
    import json
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "4", "--warmup", "3",
                          "--bytes", str(1 << 20)], capture_output=True, text=True, cwd=ROOT, timeout=300)
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "GB/s" and line["n_gpus"] == 2 and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["parity_ok"] is True
    assert line["e2e"]["h2d_bytes_per_step"] == 0
    # same config as the GPU arm: bytes per rank, rank count, steps and warm-up as asked (nothing capped at this size)
    assert line["config"]["bytes_per_rank"] == 1 << 20 and line["steps"] == 4 and line["warmup"] == 3 and line["config"]["steps_capped"] is False
    assert 1 <= line["cpu_baseline"]["cores"] <= len(os.sched_getaffinity(0))


def test_fd_channel_ignores_strangers():
    """The abstract unix sockets that carry the allocation handles are visible in /proc/net/unix.
    A local process that connects to them without the secret agreed over the password-checked TCP
    mesh (ctrl.cpp) is dropped and the world still comes up."""
    import socket
    import threading
    import time
    stop = threading.Event()
    hits = []

    def stranger():
        seen = set()
        while not stop.is_set():
            try:
                with open("/proc/net/unix") as f:
                    names = [ln.split()[-1] for ln in f if "@b200mpi." in ln]
            except OSError:
                names = []
            for nm in names:
                if nm in seen:
                    continue
                seen.add(nm)
                try:
                    c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
                    c.settimeout(1.0)
                    c.connect("\0" + nm[1:])
                    c.sendall((7).to_bytes(4, "little") + b"\0" * 20)  # claims to be rank 7, wrong secret
                    hits.append(nm)
                    c.close()
                except OSError:
                    pass
            time.sleep(0.0005)

    th = threading.Thread(target=stranger, daemon=True)
    th.start()
    try:
        for _ in range(3):
            res = run_world(4, "control_only", args=["--control-only"], timeout=120)
            assert_world_ok(res)
    finally:
        stop.set()
        th.join(timeout=5)
    # the stranger usually gets a connection in; either way every world came up
    print("stranger connections:", len(hits))
