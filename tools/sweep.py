#!/usr/bin/env python
"""Size x algorithm sweep of every collective and of Send/Receive.

Run one process per rank, either under torchrun or under the gompirun-style launcher:
    python -m torch.distributed.run --nproc-per-node N tools/sweep.py [--out FILE]
    python -m mpi_b200.launcher N tools/sweep.py [--out FILE]
Rank 0 appends one JSON line per (collective, algorithm, size) to --out (default
gpurun_out/sweep_n<N>.jsonl).  Device-resident heap buffers; time = CUDA events on the library
stream over back-to-back launches, max over ranks; `t_call_us` = wall clock of one blocking call.
busbw conventions as nccl-tests (allreduce 2(n-1)/n, allgather (n-1)/n of the total, bcast 1).
Inputs are rank-constant (x_r[i] = r+1) so every result is verified exactly in O(1) reads.
"""
import argparse
import ctypes
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--max-bytes", type=int, default=1 << 30)
    ap.add_argument("--min-bytes", type=int, default=1 << 10)
    ap.add_argument("--colls", default="allreduce,bcast,allgather,p2p")
    ap.add_argument("--algos", default="oneshot,twoshot,ring,nvls,smem")
    ap.add_argument("--blocks", default="")
    ap.add_argument("--params", default="", help="name=value;name=value passed to b200mpi_set_param")
    ap.add_argument("--tag", default="")
    ap.add_argument("--param-sets", default="", help="a=1,b=2|a=3,b=4 : the allreduce sweep is repeated for every set")
    ap.add_argument("--sizes", default="", help="explicit byte sizes, comma separated")
    args, rest = ap.parse_known_args()
    if "RANK" in os.environ:
        rank, world, local, addr, addrs = bench.world_from_env(None)
        gpu = local
    else:
        sys.argv = [sys.argv[0]] + rest
        rank = world = None
        addr, addrs, gpu = "", [], None
    os.environ.setdefault("B200MPI_HEAP_BYTES", str(3 * args.max_bytes + (512 << 20)))
    import mpi_b200 as mpi
    from mpi_b200 import _lib as L
    lib = L.load()
    mpi.api._reset_for_tests(mpi.Cuda(Addr=addr, Addrs=addrs, Timeout=120 * 10**9, Gpu=gpu))
    mpi.Init()
    rank, n = mpi.Rank(), mpi.Size()
    out_path = args.out or os.path.join(ROOT, "gpurun_out", "sweep_n%d.jsonl" % n)
    if rank == 0:
        os.makedirs(os.path.dirname(out_path), exist_ok=True)
    info = (ctypes.c_size_t(), ctypes.c_size_t(), ctypes.c_int())
    lib.b200mpi_heap_info(ctypes.byref(info[0]), ctypes.byref(info[1]), ctypes.byref(info[2]))
    nvls = bool(info[2].value)
    params = {}
    for kv in [x for x in args.params.split(";") if x]:
        k, v = kv.split("=")
        params[k] = int(v)
        if lib.b200mpi_set_param(k.encode(), int(v)):
            raise RuntimeError(L.last_error())
    ALGOS = {"auto": 0, "oneshot": 1, "twoshot": 2, "ring": 3, "nvls": 4, "smem": 5, "ll": 6, "hybrid": 7}

    def emit(d):
        if rank == 0:
            d.update({"n": n, "nvls_available": nvls, "params": params, "tag": args.tag})
            with open(out_path, "a") as f:
                f.write(json.dumps(d) + "\n")
            print(json.dumps(d), flush=True)

    def maxr(v):
        a = np.array([v], dtype=np.float64)
        o = np.zeros(1, dtype=np.float64)
        mpi.Allreduce(a, o, mpi.MAX)
        return float(o[0])

    def timed(fn_async, iters, warm):
        for _ in range(warm):
            fn_async()
        if lib.b200mpi_stream_sync():
            raise RuntimeError(L.last_error())
        mpi.Barrier()
        ms = ctypes.c_float()
        lib.b200mpi_timer_start()
        for _ in range(iters):
            fn_async()
        if lib.b200mpi_timer_stop(ctypes.byref(ms)):
            raise RuntimeError(L.last_error())
        return maxr(ms.value * 1e-3 / iters)

    def iters_for(b):
        return (50, 10) if b <= (1 << 20) else (20, 5) if b <= (64 << 20) else (8, 3)

    sizes = []
    b = args.min_bytes
    while b <= args.max_bytes:
        sizes.append(b)
        b *= 2
    if args.sizes:
        sizes = [int(x) for x in args.sizes.split(",")]
    param_sets = [dict((kv.split("=")[0], int(kv.split("=")[1])) for kv in ps.split(",") if kv) for ps in args.param_sets.split("|") if ps] or [{}]
    colls = args.colls.split(",")
    maxc = args.max_bytes // 4
    blocks_list = [int(x) for x in args.blocks.split(",") if x] or [0]

    if "allreduce" in colls:
        send = mpi.Alloc(maxc, np.float32).copy_from_host(np.full(maxc, rank + 1, dtype=np.float32))
        recv = mpi.Alloc(maxc, np.float32)
        want = np.float32(n * (n + 1) // 2)
        for ps in param_sets:
          for k, v in ps.items():
            params[k] = v
            if lib.b200mpi_set_param(k.encode(), int(v)):
                raise RuntimeError(L.last_error())
          for algo in args.algos.split(","):
              if algo in ("nvls", "hybrid") and not nvls:
                  continue
              if n == 1 and algo != "twoshot":
                  continue
              lib.b200mpi_set_algo(L.COLL_ALLREDUCE, ALGOS[algo])
              for nb in blocks_list:
                  lib.b200mpi_set_max_blocks(nb)
                  for S in sizes:
                      if algo == "oneshot" and S > (8 << 20):
                          continue
                      if algo == "ll" and S > (256 << 10):
                          continue
                      if algo == "hybrid" and S < (1 << 20):
                          continue
                      cnt = S // 4
                      it, wm = iters_for(S)
                      t = timed(lambda: lib.b200mpi_allreduce_async(send.ptr, recv.ptr, cnt, L.F32, L.SUM), it, wm)
                      t0 = time.perf_counter()
                      for _ in range(5):
                          lib.b200mpi_allreduce(send.ptr, recv.ptr, cnt, L.F32, L.SUM, L.DEVICE)
                      tc = maxr((time.perf_counter() - t0) / 5)
                      probe = np.concatenate([recv[:min(cnt, 64)].to_host(), recv[max(cnt - 64, 0):cnt].to_host(), recv[cnt // 2:cnt // 2 + 1].to_host()])
                      ok = bool(np.all(probe == want))
                      emit({"coll": "allreduce", "dtype": "f32", "algo": algo, "bytes": S, "t_us": t * 1e6, "t_call_us": tc * 1e6,
                            "algbw_gbs": S / t / 1e9, "busbw_gbs": S / t / 1e9 * (2 * (n - 1) / n if n > 1 else 1), "ok": ok, "max_blocks": nb, "pset": dict(ps),
                            "algo_used": L.ALGO_NAMES.get(lib.b200mpi_get_algo(L.COLL_ALLREDUCE, cnt, L.F32))})
        lib.b200mpi_set_algo(L.COLL_ALLREDUCE, 0)
        lib.b200mpi_set_max_blocks(0)
        send.free()
        recv.free()

    if "link" in colls and n > 1:
        # raw NVLink rates of the copy kernel between neighbours: the roofline of the P2P collectives
        names = {0: "all ranks pull from rank+1", 1: "all ranks push to rank+1", 2: "rank 0 pulls from rank 1 (one direction busy)",
                 3: "rank 0 pushes to rank 1 (one direction busy)", 4: "all ranks pull and push at once"}
        for S in (64 << 20, 256 << 20):
            for mode in range(5):
                ms = ctypes.c_float()
                if lib.b200mpi_link_probe(S, mode, 10, ctypes.byref(ms)):
                    raise RuntimeError(L.last_error())
                t = maxr(ms.value * 1e-3)
                per_dir = (2 if mode == 4 else 1) * S / t / 1e9 if t > 0 else 0.0
                emit({"coll": "link", "algo": names[mode], "mode": mode, "bytes": S, "t_us": t * 1e6, "gbs_per_direction": per_dir, "busbw_gbs": per_dir, "ok": True})

    if "latency" in colls and n > 1:
        # blocking-call latency of small Allreduce (what a Go caller sees per call): device-resident and
        # pinned-host buffers, every algorithm incl. the experimental LL path; median of 200 calls
        hin, hout = ctypes.c_void_p(), ctypes.c_void_p()
        lib.b200mpi_host_alloc(1 << 16, ctypes.byref(hin))
        lib.b200mpi_host_alloc(1 << 16, ctypes.byref(hout))
        ctypes.memset(hin.value, 0, 1 << 16)
        dsend = mpi.Alloc(1 << 14, np.float32).copy_from_host(np.full(1 << 14, rank + 1, dtype=np.float32))
        drecv = mpi.Alloc(1 << 14, np.float32)
        for algo in ("auto", "oneshot", "twoshot", "nvls", "ll"):
            if algo == "nvls" and not nvls:
                continue
            lib.b200mpi_set_algo(L.COLL_ALLREDUCE, ALGOS[algo])
            for S in (8, 1024, 32768):
                cnt = S // 4
                for kind, sp, rp, mk in (("device", dsend.ptr, drecv.ptr, L.DEVICE), ("host", hin, hout, L.HOST)):
                    ts = []
                    for i in range(220):
                        mpi.Barrier() if i % 50 == 0 else None
                        t0 = time.perf_counter()
                        rc = lib.b200mpi_allreduce(sp, rp, cnt, L.F32, L.SUM, mk)
                        ts.append(time.perf_counter() - t0)
                        if rc:
                            raise RuntimeError(L.last_error())
                    ts = sorted(ts[20:])
                    emit({"coll": "latency", "dtype": "f32", "algo": algo, "kind": kind, "bytes": S, "t_call_us_median": maxr(ts[len(ts) // 2]) * 1e6,
                          "t_call_us_p10": maxr(ts[len(ts) // 10]) * 1e6, "t_us": maxr(ts[len(ts) // 2]) * 1e6, "busbw_gbs": 0.0, "ok": True})
        lib.b200mpi_set_algo(L.COLL_ALLREDUCE, 0)
        lib.b200mpi_host_free(hin)
        lib.b200mpi_host_free(hout)
        dsend.free()
        drecv.free()

    if "bcast" in colls and n > 1:
        buf = mpi.Alloc(maxc, np.float32)
        for algo in ("oneshot", "twoshot", "nvls", "nvls_root"):
            if algo.startswith("nvls") and not nvls:
                continue
            lib.b200mpi_set_param(b"bcast_nvls2", 0 if algo == "nvls_root" else 1)
            lib.b200mpi_set_algo(L.COLL_BCAST, ALGOS["nvls" if algo == "nvls_root" else algo])
            for S in sizes:
                if algo == "oneshot" and S > (64 << 20):
                    continue
                cnt = S // 4
                buf[:cnt].copy_from_host(np.full(cnt, 7.0 if rank == 0 else -1.0, dtype=np.float32)) if S <= (64 << 20) else None
                it, wm = iters_for(S)
                t = timed(lambda: lib.b200mpi_bcast_async(buf.ptr, cnt, L.F32, 0), it, wm)
                probe = np.concatenate([buf[:min(cnt, 64)].to_host(), buf[max(cnt - 64, 0):cnt].to_host()])
                ok = bool(np.all(probe == 7.0)) if S <= (64 << 20) else None
                emit({"coll": "bcast", "dtype": "f32", "algo": algo, "bytes": S, "t_us": t * 1e6, "algbw_gbs": S / t / 1e9, "busbw_gbs": S / t / 1e9, "ok": ok})
        lib.b200mpi_set_algo(L.COLL_BCAST, 0)
        buf.free()

    if "allgather" in colls and n > 1:
        per_max = min(args.max_bytes // n, 128 << 20)
        send = mpi.Alloc(per_max // 8, np.int64).copy_from_host(np.full(per_max // 8, rank + 1, dtype=np.int64))
        recv = mpi.Alloc(per_max // 8 * n, np.int64)
        for algo in ("oneshot", "ring", "nvls"):
            if algo == "nvls" and not nvls:
                continue
            lib.b200mpi_set_algo(L.COLL_ALLGATHER, ALGOS[algo])
            for S in [s for s in sizes if s <= per_max] + [1000000 * 8]:
                if S > per_max:
                    continue
                cnt = S // 8
                it, wm = iters_for(S * n)
                t = timed(lambda: lib.b200mpi_allgather_async(send.ptr, recv.ptr, cnt, L.I64), it, wm)
                got = recv[:cnt * n].to_host() if S * n <= (64 << 20) else None
                ok = bool(all(np.all(got[r * cnt:(r + 1) * cnt] == r + 1) for r in range(n))) if got is not None else None
                tot = S * n
                emit({"coll": "allgather", "dtype": "i64", "algo": "push" if algo == "oneshot" else algo, "bytes_per_rank": S, "bytes": tot, "t_us": t * 1e6,
                      "algbw_gbs": tot / t / 1e9, "busbw_gbs": tot / t / 1e9 * (n - 1) / n, "ok": ok})
        lib.b200mpi_set_algo(L.COLL_ALLGATHER, 0)
        send.free()
        recv.free()

    if "p2p" in colls and n > 1:
        # bounce (examples/bounce/bounce.go:85-138): even sends, odd returns; device-resident float64
        maxp = min(args.max_bytes, 256 << 20)
        msg = mpi.Alloc(maxp // 8, np.float64).copy_from_host(np.arange(maxp // 8, dtype=np.float64))
        rcv = mpi.Alloc(maxp // 8, np.float64)
        even = rank % 2 == 0
        peer = rank + 1 if even else rank - 1
        ladder = sorted(set([0, 8, 80, 800, 8000, 80000, 800000, 8000000, 1 << 20, 16 << 20, maxp]))
        for S in ladder:
            if peer >= n or S > maxp:
                continue
            cnt = S // 8
            reps = 100 if S <= (1 << 20) else 10
            for phase in range(2):  # 0 warm-up, 1 timed
                mpi.Barrier()
                t0 = time.perf_counter()
                for _ in range(reps if phase else 3):
                    if even:
                        mpi.Send(msg[:cnt], peer, 0)
                        mpi.Receive(rcv[:cnt], peer, 0)
                    else:
                        mpi.Receive(rcv[:cnt], peer, 0)
                        mpi.Send(rcv[:cnt], peer, 0)
                dt = (time.perf_counter() - t0) / reps
            dt = maxr(dt)
            ok = bool(np.array_equal(rcv[:min(cnt, 256)].to_host(), np.arange(min(cnt, 256), dtype=np.float64))) if cnt else True
            emit({"coll": "bounce", "dtype": "f64", "algo": "pull", "bytes": S, "round_trip_us": dt * 1e6, "algbw_gbs": (2 * S / dt / 1e9) if S else 0.0,
                  "busbw_gbs": (2 * S / dt / 1e9) if S else 0.0, "ok": ok, "reps": reps})
        # host-slice ping-pong at the headline 1 MiB float64 point (what an unmodified Go caller does)
        h = np.arange(131072, dtype=np.float64)
        hr = np.zeros(131072, dtype=np.float64)
        if peer < n:
            for phase in range(2):
                mpi.Barrier()
                t0 = time.perf_counter()
                for _ in range(20 if phase else 3):
                    if even:
                        mpi.Send(h, peer, 1)
                        hr = mpi.Receive(hr, peer, 1)
                    else:
                        hr = mpi.Receive(hr, peer, 1)
                        mpi.Send(hr, peer, 1)
                dt = (time.perf_counter() - t0) / 20
            dt = maxr(dt)
            emit({"coll": "bounce_host", "dtype": "f64", "algo": "staged", "bytes": 1 << 20, "round_trip_us": dt * 1e6, "ok": bool(np.array_equal(hr, h))})
        msg.free()
        rcv.free()
    mpi.Barrier()
    mpi.Finalize()


if __name__ == "__main__":
    main()
