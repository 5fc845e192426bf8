"""gompirun re-targeted: launch one rank per local GPU.

    python -m mpi_b200.launcher N program [args...]

Mirrors /root/reference/mpirun/gompirun/gompirun.go:28-93: N child processes of `program`, each with
`-mpi-addr <own> -mpi-alladdr <comma list>` appended AFTER the user's arguments, stdio inherited,
wait for all.  Differences, all additive: N defaults to / is capped by the number of visible GPUs
when it is given as 0 or "auto"; every child also gets `-mpi-gpu <rank % ngpus>`; a child that exits
non-zero makes the launcher exit non-zero (the reference ignores exit codes, gompirun.go:89).
The native twin is mpirun/gompirun.cpp.
"""
import os
import subprocess
import sys

BASE_PORT = 6000  # gompirun.go:45


def gpu_count():
    """Number of GPUs without creating a CUDA context in the launcher."""
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis is not None and vis.strip() != "":
        return len([v for v in vis.split(",") if v.strip() != ""])
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return len([l for l in out.splitlines() if l.startswith("GPU ")])
    except Exception:  # noqa: BLE001
        return 0


def addresses(n, base_port=BASE_PORT):
    """":6000", ":6001", ... (gompirun.go:46-51).  Rank = index in the SORTED list
    (network.go:94-109), which equals launch order as long as all ports have the same number of
    digits -- true for n <= 8 from 6000."""
    return [":%d" % (base_port + i) for i in range(n)]


def child_argv(program, user_args, addr, addrs, gpu=None):
    """gompirun.go:77-83: user args first, then the two mpi flags."""
    argv = [program] + list(user_args) + ["-mpi-addr", addr, "-mpi-alladdr", ",".join(addrs)]
    if gpu is not None:
        argv += ["-mpi-gpu", str(gpu)]
    return argv


def launch(n, program, user_args, base_port=BASE_PORT, python=False):
    ngpu = gpu_count()
    if n <= 0:
        n = max(ngpu, 1)
    addrs = addresses(n, base_port)
    procs = []
    for i, a in enumerate(addrs):
        argv = child_argv(program, user_args, a, addrs, gpu=(i % ngpu) if ngpu else None)
        if python or program.endswith(".py"):
            argv = [sys.executable] + argv
        procs.append(subprocess.Popen(argv))  # stdin/stdout/stderr inherited (gompirun.go:86-88)
    rc = 0
    for p in procs:
        p.wait()
        if p.returncode != 0:
            rc = 1
    return rc


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) < 2:
        sys.stderr.write("less than two arguments, must have at least number of nodes and executable\n")
        return 2
    try:
        n = 0 if argv[0] == "auto" else int(argv[0])
    except ValueError:
        sys.stderr.write("error parsing nNodes: %r\n" % argv[0])
        return 2
    if argv[0] != "auto" and n < 1:
        sys.stderr.write("number of nodes must be positive\n")
        return 2
    return launch(n, argv[1], argv[2:], base_port=int(os.environ.get("GOMPIRUN_BASE_PORT", BASE_PORT)))


if __name__ == "__main__":
    sys.exit(main())
