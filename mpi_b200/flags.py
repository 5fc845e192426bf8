"""The five command-line flags of the reference (/root/reference/flags.go:10-50), parsed with the
rules of Go's `flag` package: `-name value`, `-name=value`, one or two dashes, parsing stops at the
first non-flag argument or at `--`.  Unknown flags are left for the program.

    -mpi-addr         address of the local running process
    -mpi-alladdr      comma separated addresses of all processes (may repeat; values append)
    -mpi-inittimeout  Go duration ("1.5s", "300ms", "2m"), 0 = wait forever
    -mpi-protocol     kept for compatibility; the control plane is always tcp
    -mpi-password     compared verbatim during the handshake (network.go:343-346)
    -mpi-gpu          (new, additive) CUDA ordinal for this rank; default rank % device count
"""
import re

_UNITS = {"ns": 1, "us": 1_000, "µs": 1_000, "μs": 1_000, "ms": 1_000_000,
          "s": 1_000_000_000, "m": 60_000_000_000, "h": 3_600_000_000_000}


def parse_duration(text):
    """time.ParseDuration: a signed sequence of decimal numbers each with a unit; returns ns."""
    s = text.strip()
    if s in ("0", "+0", "-0"):
        return 0
    sign = 1
    if s and s[0] in "+-":
        sign = -1 if s[0] == "-" else 1
        s = s[1:]
    if not s:
        raise ValueError("time: invalid duration %r" % text)
    total = 0.0
    pos = 0
    pat = re.compile(r"(\d+\.?\d*|\.\d+)(ns|us|µs|μs|ms|s|m|h)")
    while pos < len(s):
        m = pat.match(s, pos)
        if not m:
            raise ValueError("time: invalid duration %r" % text)
        total += float(m.group(1)) * _UNITS[m.group(2)]
        pos = m.end()
    return sign * int(round(total))


class Flags:
    def __init__(self):
        self.addr = ""            # FlagAddr
        self.alladdr = []         # FlagAllAddrs (AddrsFlag appends, flags.go:22-27)
        self.inittimeout = 0      # FlagInitTimeout, nanoseconds
        self.protocol = "tcp"     # FlagProtocol
        self.password = ""        # FlagPassword
        self.gpu = -1
        self.rest = []            # arguments that are not ours

    @property
    def alladdr_csv(self):
        return ",".join(self.alladdr)


_NAMES = {"mpi-addr", "mpi-alladdr", "mpi-inittimeout", "mpi-protocol", "mpi-password", "mpi-gpu"}


def parse(argv):
    """Parse a list of arguments (no program name).  Returns Flags; foreign args go to .rest."""
    f = Flags()
    i = 0
    while i < len(argv):
        a = argv[i]
        name = None
        if a.startswith("--") and len(a) > 2:
            name = a[2:]
        elif a.startswith("-") and len(a) > 1 and not a.startswith("--"):
            name = a[1:]
        if a == "--":
            f.rest.extend(argv[i + 1:])
            break
        value = None
        if name is not None and "=" in name:
            name, value = name.split("=", 1)
        if name not in _NAMES:
            f.rest.append(a)
            i += 1
            continue
        if value is None:
            if i + 1 >= len(argv):
                raise ValueError("flag needs an argument: -%s" % name)
            value = argv[i + 1]
            i += 1
        if name == "mpi-addr":
            f.addr = value
        elif name == "mpi-alladdr":
            f.alladdr.extend(value.split(","))
        elif name == "mpi-inittimeout":
            f.inittimeout = parse_duration(value)
        elif name == "mpi-protocol":
            f.protocol = value
        elif name == "mpi-password":
            f.password = value
        elif name == "mpi-gpu":
            f.gpu = int(value)
        i += 1
    return f
