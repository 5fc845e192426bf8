#!/usr/bin/env python
"""bench.py -- Allreduce(float32, sum) bus bandwidth, the metric BASELINE.json names.

    python bench.py --gpus N --steps K --warmup W            (N>1: one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N ...            (restated reference TCP path, host cores)

One JSON line on stdout (rank 0).  A "step" is one Allreduce of S bytes per rank (default
S = 256 MiB, the headline point) on synthetic uniform[0,1) buffers; out of place, so the input
is identical every step.  Send+recv buffers (2 x 256 MiB) exceed the 126 MB L2, so no flush is
needed between steps.

  value     N >= 2: bus bandwidth, nccl-tests convention: S/t * 2(N-1)/N, GB = 1e9 B, device
            resident buffers, t = CUDA-event time over K back-to-back steps / K, max over ranks.
            N == 1: the collective degenerates to a local copy; the bus factor is 0 there, so the
            line reports algorithm bandwidth S/t and says so in config.note.
  e2e       the same quantity through the blocking public call with HOST (pinned) buffers:
            H2D of the input and D2H of the result inside every step.
  roofline  N == 1: HBM (read S + write S per launch) against MEASURED_PEAKS.json hbm_gbs;
            N >= 2: NVLink, busbw against 900 GB/s nominal (measured peer copy ~770 GB/s) plus the
            HBM side ((3 - 1/N) * S per launch).
  cpu_baseline / --impl reference: oracle/ref_tcp.c, the restated gob-over-TCP-loopback path of
            the reference (it has no Allreduce; composed as a ring over Send/Receive), on a
            bounded sample.  This is the only use of oracle/ here besides the parity spot check.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

SEED = 0xB2000000
NVLINK_NOMINAL_GBS = 900.0   # per direction per GPU (B200_PROFILING.md)
NVLINK_MEASURED_GBS = 770.0  # peer copy measured on this pool (B200_PROFILING.md)
HBM_FALLBACK_GBS = 6650.0


def world_from_env(args):
    """torchrun env (RANK/LOCAL_RANK/WORLD_SIZE/MASTER_*) or a single rank."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world == 1:
        return rank, world, local, "", []
    mport = int(os.environ.get("MASTER_PORT", "29500"))
    base = 20000 + (mport * 7 + 13) % 30000  # 5 digits for every rank: sorted order == rank order
    addrs = ["127.0.0.1:%d" % (base + r) for r in range(world)]
    return rank, world, local, addrs[rank], addrs


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu):
        super().__init__(daemon=True)
        self.gpu = gpu
        self.rows = []
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:  # noqa: BLE001
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for k, nm in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:  # noqa: BLE001
                continue
        # the sampler also sees idle samples around the region: "under load" = top half
        sm.sort()
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": (load[len(load) // 2] if load else None), "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:  # noqa: BLE001
        return {"hbm_gbs": HBM_FALLBACK_GBS}, "fallback"


def measured_traffic(n, nbytes, kernel):
    """dram read+write bytes per launch from the committed ncu captures (profiles/traffic.json)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            for e in json.load(f)["entries"]:
                if e["n"] == n and e["bytes"] == nbytes and e["kernel"] == kernel:
                    return e["traffic_bytes"], e["source"]
    except Exception:  # noqa: BLE001
        pass
    return None, None


def reference_arm(n, count, dtype, steps, warmup):
    """Times oracle/ref_tcp.c (restated reference) with n ranks (threads) on this host."""
    from oracle import oracle as O
    t0 = time.time()
    secs, out = O.ref_bench(O.COLL_ALLREDUCE, dtype, n, count, iters=steps, warmup=warmup, seed=SEED)
    ins = [O.fill(dtype, SEED + r, count) for r in range(n)]
    want = O.allreduce(ins, order=O.ORDER_F64)
    ok = bool(np.allclose(out, want, rtol=1e-6, atol=0))
    return secs, ok, time.time() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bytes", type=int, default=256 << 20, help="message size S per rank")
    ap.add_argument("--algo", default="auto")
    ap.add_argument("--cpu-sample-bytes", type=int, default=16 << 20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    rank, world, local, addr, addrs = world_from_env(args)
    n = world
    dtype = np.float32
    count = args.bytes // 4
    S = count * 4
    bus = (2.0 * (n - 1) / n) if n > 1 else 1.0
    metric = "allreduce_f32_sum_busbw" if n > 1 else "allreduce_f32_sum_algbw"
    workload = "Allreduce float32 sum, %d MiB per rank, %d GPU(s)" % (S >> 20, n)

    # ------------------------------------------------------------------ reference arm
    if args.impl == "reference":
        if rank != 0:
            return 0
        rn = max(args.gpus, 1)
        sample = min(args.cpu_sample_bytes, S)
        rcount = sample // 4
        secs, ok, wall = reference_arm(rn, rcount, dtype, max(1, min(args.steps, 5)), 1)
        rbus = (2.0 * (rn - 1) / rn) if rn > 1 else 1.0
        val = rcount * 4 / secs * rbus / 1e9
        line = {
            "impl": "reference", "metric": "allreduce_f32_sum_busbw" if rn > 1 else "allreduce_f32_sum_algbw",
            "value": val, "unit": "GB/s", "n_gpus": rn, "steps": max(1, min(args.steps, 5)), "warmup": 1,
            "ms_per_step": secs * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "Allreduce float32 sum, %d MiB per rank, %d host rank(s) over TCP loopback" % (sample >> 20, rn),
                       "note": "restated reference path (oracle/ref_tcp.c): gob encode/decode + 2 TCP conns per pair + ack, ring allreduce composed from Send/Receive; bounded sample of the %d MiB workload" % (S >> 20)},
            "cpu_baseline": {"value": val, "unit": "GB/s", "cores": 2 * rn, "kind": "port", "sample": "%d MiB per rank, %d iterations" % (sample >> 20, max(1, min(args.steps, 5))), "parity_ok": ok},
            "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "host_cores": os.cpu_count(),
        }
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm
    os.environ.setdefault("B200MPI_HEAP_BYTES", str(4 * S + (512 << 20)))
    import mpi_b200 as mpi
    from mpi_b200 import _lib as L
    lib = L.load()
    mpi.api._reset_for_tests(mpi.Cuda(Addr=addr, Addrs=addrs, Timeout=120 * 10**9, Gpu=local))
    mpi.Init()
    algo_ids = {"auto": 0, "oneshot": 1, "twoshot": 2, "ring": 3, "nvls": 4, "smem": 5}
    lib.b200mpi_set_algo(L.COLL_ALLREDUCE, algo_ids[args.algo])
    from oracle import oracle as O  # input generator + parity spot check only

    x = O.fill(dtype, SEED + rank, count)
    send = mpi.Alloc(count, dtype).copy_from_host(x)
    recv = mpi.Alloc(count, dtype)
    algo_used = lib.b200mpi_get_algo(L.COLL_ALLREDUCE, count, L.F32) if n > 1 else 0

    def run_steps(k):
        for _ in range(k):
            rc = lib.b200mpi_allreduce_async(send.ptr, recv.ptr, count, L.F32, L.SUM)
            if rc:
                raise RuntimeError(L.last_error())

    def max_over_ranks(v):
        a = np.array([v], dtype=np.float64)
        o = np.zeros(1, dtype=np.float64)
        mpi.Allreduce(a, o, mpi.MAX)
        return float(o[0])

    run_steps(args.warmup)
    if lib.b200mpi_stream_sync():
        raise RuntimeError(L.last_error())
    # parity spot check on the warm-up result (first 1 Mi elements) against the oracle
    head = min(count, 1 << 20)
    got = recv[:head].to_host()
    want = O.allreduce([O.fill(dtype, SEED + r, head) for r in range(n)], order=O.ORDER_F64)
    parity_ok = bool(np.allclose(got, want, rtol=1e-6, atol=0))

    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    mpi.Barrier()
    l0 = lib.b200mpi_launch_count()
    ms = ctypes.c_float(0)
    lib.b200mpi_timer_start()
    run_steps(args.steps)
    if lib.b200mpi_timer_stop(ctypes.byref(ms)):
        raise RuntimeError(L.last_error())
    launches = int(lib.b200mpi_launch_count() - l0)
    mpi.Barrier()
    t_step = max_over_ranks(ms.value / 1e3 / args.steps)

    algbw = S / t_step / 1e9
    value = algbw * bus

    # ---- end to end: blocking public call, pinned HOST buffers, H2D + D2H inside every step
    e2e = None
    if not args.no_e2e:
        hs, hr = ctypes.c_void_p(), ctypes.c_void_p()
        lib.b200mpi_host_alloc(S, ctypes.byref(hs))
        lib.b200mpi_host_alloc(S, ctypes.byref(hr))
        ctypes.memmove(hs.value, x.ctypes.data, S)
        k_e2e = max(3, min(args.steps, 10))
        for _ in range(2):
            if lib.b200mpi_allreduce(hs, hr, count, L.F32, L.SUM, L.HOST):
                raise RuntimeError(L.last_error())
        mpi.Barrier()
        t0 = time.perf_counter()
        for _ in range(k_e2e):
            if lib.b200mpi_allreduce(hs, hr, count, L.F32, L.SUM, L.HOST):
                raise RuntimeError(L.last_error())
        t_e2e = max_over_ranks((time.perf_counter() - t0) / k_e2e)
        mpi.Barrier()
        e2e = {"value": S / t_e2e / 1e9 * bus, "unit": "GB/s", "h2d_bytes_per_step": S, "d2h_bytes_per_step": S,
               "ms_per_step": t_e2e * 1e3, "steps": k_e2e}
        lib.b200mpi_host_free(hs)
        lib.b200mpi_host_free(hr)

    # keep the GPU under the same load a little longer so nvidia-smi (100 ms period) sees it:
    # the timed region itself is only K x ~0.1-0.7 ms
    n_load = int(min(max(0.6 / t_step, 10), 20000))  # same count on every rank (t_step is the max over ranks)
    run_steps(n_load)
    if lib.b200mpi_stream_sync():
        raise RuntimeError(L.last_error())
    clocks = sampler.stop()
    clocks["window"] = "timed region + e2e region + 0.6 s of the same launches (sampler period 100 ms)"
    peaks, peak_kind = measured_peaks()
    hbm_peak = float(peaks.get("hbm_gbs", HBM_FALLBACK_GBS))
    if n == 1:
        roof = {"bound": "hbm", "achieved": 2 * S / t_step / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": 2 * S / t_step / 1e9 / hbm_peak, "traffic": None, "peak_source": peak_kind,
                "kernel": "copy_bytes_kernel", "algorithmic_bytes_per_launch": 2 * S}
    else:
        hbm_bytes = (3.0 - 1.0 / n) * S
        roof = {"bound": "nvlink", "achieved": value, "peak": NVLINK_NOMINAL_GBS, "unit": "GB/s",
                "frac": value / NVLINK_NOMINAL_GBS, "frac_of_measured_peer_copy": value / NVLINK_MEASURED_GBS,
                "traffic": None, "peak_source": "nominal NVLink 5 per direction per GPU; measured peer copy %.0f GB/s" % NVLINK_MEASURED_GBS,
                "kernel": "allreduce_%s_kernel" % L.ALGO_NAMES.get(algo_used, "?"),
                "nvls_link_bytes_per_launch": (1.0 + 1.0 / n) * S if L.ALGO_NAMES.get(algo_used) == "nvls" else None,
                "algorithmic_bytes_per_launch": 2.0 * (n - 1) / n * S,
                "hbm": {"achieved": hbm_bytes / t_step / 1e9, "peak": hbm_peak, "frac": hbm_bytes / t_step / 1e9 / hbm_peak, "peak_source": peak_kind}}

    traffic, tsrc = measured_traffic(n, S, roof["kernel"])
    if traffic is not None:
        roof["traffic"] = traffic
        roof["traffic_source"] = tsrc
    info = (ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_int())
    lib.b200mpi_heap_info(ctypes.byref(info[0]), ctypes.byref(info[1]), ctypes.byref(info[2]))

    cpu = None
    if rank == 0 and n == 1 and not args.no_cpu_baseline:
        sample = min(args.cpu_sample_bytes, S)
        secs, ok, wall = reference_arm(1, sample // 4, dtype, 3, 1)
        cpu = {"value": sample / secs / 1e9, "unit": "GB/s", "cores": 2, "kind": "port",
               "sample": "%d MiB, 3 iterations, world of 1 = Send/Receive to self (gob encode + decode)" % (sample >> 20), "parity_ok": ok}

    send.free()
    recv.free()
    mpi.Finalize()
    if rank != 0:
        return 0
    line = {
        "metric": metric, "value": value, "unit": "GB/s", "n_gpus": n, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": t_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "bytes_per_rank": S, "algo": L.ALGO_NAMES.get(algo_used, "copy") if n > 1 else "local copy (world of 1)",
                   "nvls": bool(info[2].value), "l2": "inputs+outputs (%d MiB) exceed L2, no flush" % (2 * S >> 20),
                   "note": ("world of 1: Allreduce degenerates to a device copy, busbw factor 2(N-1)/N is 0, value is algbw S/t" if n == 1
                            else "busbw = S/t * 2(N-1)/N (nccl-tests convention)")},
        "algbw_gbs": algbw, "aggregate_gbs": algbw * n,
        "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roof, "parity_ok": parity_ok,
    }
    if cpu is not None:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))
    return 0


if __name__ == "__main__":
    sys.exit(main())
