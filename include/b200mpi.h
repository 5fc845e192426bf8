/*
 * b200mpi.h -- C ABI of libb200mpi.so: the B200-native transport that sits behind the
 * btracey/mpi Go facade (mpi.Init/Finalize/Rank/Size/Send/Receive + the new
 * Recv/Bcast/Allreduce/Allgather entry points).
 *
 * This is the drop-in boundary.  Every entry point is what the cgo shim (go/mpi/cuda.go,
 * see INTEGRATION.md) binds for one method of the reference's `mpi.Interface`
 * (/root/reference/mpi.go:163-170) or for one of the collectives the reference only stubs
 * (/root/reference/mpi.go:130, mpi.go:69-71).  Plain pointers and sizes only; no C++ or torch
 * types cross this line.
 *
 * Conventions
 *   - every call returns 0 (B200MPI_OK) or a negative b200mpi_error; the text of the last
 *     error on the calling thread is b200mpi_last_error().
 *   - all data calls block until the caller's buffers may be reused (mpi.go:47-48), except the
 *     *_async forms which only enqueue on the library stream (b200mpi_stream_sync completes them).
 *   - caller owns every buffer; HOST pointers are used only for the duration of the call
 *     (cgo pointer-passing rule); count==0 with buf==NULL is legal (bounce.go:33 sends length 0).
 *   - p2p calls (send/recv) are thread-safe provided {peer,tag} pairs are unique among
 *     in-flight calls (mpi.go:121-125); collectives are called by one thread per rank, in the
 *     same order on every rank.
 */
#ifndef B200MPI_H_
#define B200MPI_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200MPI_VERSION 200 /* 0.2.0: + isend/wait, reduce_scatter, reduce, alltoall, probes, get_param (additive) */
#define B200MPI_MAX_RANKS 8 /* "the 8 GPUs of one box" */

typedef enum {
  B200MPI_U8 = 0,  /* []byte / mpi.Raw / gob-encoded anything-else (e.g. string) */
  B200MPI_I64 = 1, /* []int64   */
  B200MPI_F32 = 2, /* []float32 */
  B200MPI_F64 = 3  /* []float64 */
} b200mpi_dtype;

typedef enum { B200MPI_SUM = 0, B200MPI_MAX = 1, B200MPI_MIN = 2 } b200mpi_op;

typedef enum {
  B200MPI_HOST = 0,  /* pageable or pinned host memory (a Go slice) */
  B200MPI_DEVICE = 1 /* device memory; zero-copy when it lies inside the b200mpi_alloc heap */
} b200mpi_memkind;

typedef enum {
  B200MPI_COLL_ALLREDUCE = 0,
  B200MPI_COLL_BCAST = 1,
  B200MPI_COLL_ALLGATHER = 2,
  B200MPI_COLL_REDUCE_SCATTER = 3
} b200mpi_coll;

typedef enum {
  B200MPI_ALGO_AUTO = 0,    /* size/topology based choice (default) */
  B200MPI_ALGO_ONESHOT = 1, /* allreduce: every rank reads all n buffers (latency path);
                               bcast: everyone pulls from root; allgather: direct push */
  B200MPI_ALGO_TWOSHOT = 2, /* allreduce: owner-reduces-slice + push to all, one fused pass;
                               bcast: pull-slice-from-root + push to all, one fused pass */
  B200MPI_ALGO_RING = 3,    /* allreduce/allgather: 2(n-1) / (n-1) neighbour steps */
  B200MPI_ALGO_NVLS = 4,    /* multimem.ld_reduce / multimem.st through the NVSwitch */
  B200MPI_ALGO_TWOSHOT_SMEM = 5, /* allreduce two-shot with cp.async.bulk shared-memory staging */
  B200MPI_ALGO_LL = 6,       /* allreduce <= 256 KiB: flag-in-data cells, no barrier (the small-message default) */
  B200MPI_ALGO_HYBRID = 7    /* allreduce: NVLS for most of the message + fused P2P two-shot for the rest, one kernel */
} b200mpi_algo;

typedef enum {
  B200MPI_OK = 0,
  B200MPI_ERR_ARG = -1,         /* bad argument (rank/dtype/op/NULL with count>0 ...) */
  B200MPI_ERR_NOT_INIT = -2,    /* call before init / after finalize */
  B200MPI_ERR_BOOTSTRAP = -3,   /* address list / listen / dial / handshake failure */
  B200MPI_ERR_PASSWORD = -4,    /* "bad password" (network.go:343-346) */
  B200MPI_ERR_TIMEOUT = -5,     /* init timeout (network.go:223-231) or device-flag watchdog */
  B200MPI_ERR_TAG_EXISTS = -6,  /* duplicate in-flight {peer,tag} (mpi.go:174-182) */
  B200MPI_ERR_CUDA = -7,        /* a CUDA runtime/driver call failed */
  B200MPI_ERR_NOMEM = -8,       /* symmetric heap exhausted */
  B200MPI_ERR_TRUNCATE = -9,    /* recv capacity smaller than the message */
  B200MPI_ERR_NO_DEVICE = -10,  /* no usable CUDA device (there is no CPU fallback) */
  B200MPI_ERR_UNSUPPORTED = -11,/* dtype/op/algo combination not implemented */
  B200MPI_ERR_PEER = -12        /* a peer reported failure / mismatched collective */
} b200mpi_error;

/* ---- lifecycle: mpi.Init / Finalize / Rank / Size  (mpi.go:96-118, network.go:41-65) -------- */

/* addr / alladdr_csv / password / timeout_ns are the values of -mpi-addr, -mpi-alladdr,
 * -mpi-password, -mpi-inittimeout (flags.go:44-50).  Empty alladdr_csv => single rank on
 * ":5000" (network.go:55-58).  rank = index of addr in the sorted address list
 * (network.go:94-109).  gpu: device ordinal, -1 = rank % deviceCount.
 * gpu == -2 starts the control plane only (no CUDA): rank/size/barrier work, data calls
 * return B200MPI_ERR_NO_DEVICE.  Used by CPU-only plumbing tests. */
int b200mpi_init(const char* addr, const char* alladdr_csv, const char* password,
                 int64_t timeout_ns, int gpu);
int b200mpi_finalize(void);
int b200mpi_rank(void); /* -1 before init (mpi.go:110-112) */
int b200mpi_size(void); /*  0 before init (mpi.go:116-118) */
int b200mpi_device(void); /* CUDA ordinal bound to this rank, -1 if none */
int b200mpi_version(void);
const char* b200mpi_last_error(void); /* thread-local, never NULL */

/* ---- memory -------------------------------------------------------------------------------- */

/* Device memory from this rank's peer-mapped heap (size: env B200MPI_HEAP_BYTES, default 2 GiB).
 * Buffers from here take the zero-copy path; any other device pointer is staged through it. */
int b200mpi_alloc(size_t bytes, void** dptr);
int b200mpi_free(void* dptr);
/* Pinned host memory on the NUMA node of this rank's GPU (optional; plain malloc'ed / Go memory
 * works too: it is staged through a pinned bounce ring by helper threads). */
int b200mpi_host_alloc(size_t bytes, void** hptr);
int b200mpi_host_free(void* hptr);
/* Synchronous copies for harnesses that have no CUDA binding of their own. kind: 0 H2D, 1 D2H, 2 D2D */
int b200mpi_memcpy(void* dst, const void* src, size_t bytes, int kind);
int b200mpi_heap_info(size_t* total, size_t* used, int* nvls_available);
int b200mpi_numa_node(void); /* NUMA node of the bound GPU (sysfs), -1 unknown */

/* ---- point to point: mpi.Send / mpi.Receive (mpi.go:126,157; network.go:518-602) ----------- */

/* Synchronous (rendezvous) send: returns once the matching recv has taken the data, like the
 * reference's ack wait (network.go:569).  dest == own rank is legal (network.go:545-548). */
int b200mpi_send(const void* buf, size_t count, int dtype, int dest, int tag, int memkind);
/* The Send/Wait pair sketched in the reference (mpi.go:132-152, commented out there): isend
 * returns once the data has left the caller's buffer ("sent on connection", here: staged in this
 * rank's device heap), without waiting for the receiver; wait blocks until `dest` confirmed the
 * message with `tag` and frees the {dest, tag} pair.  Every isend must be followed by one wait. */
int b200mpi_isend(const void* buf, size_t count, int dtype, int dest, int tag, int memkind);
int b200mpi_wait(int dest, int tag);
/* Blocks for message (src, tag).  *count_out = elements sent (gob resizes the destination,
 * network.go:597); more than `capacity` elements => B200MPI_ERR_TRUNCATE, *count_out = needed. */
int b200mpi_recv(void* buf, size_t capacity, size_t* count_out, int dtype, int src, int tag,
                 int memkind);

/* ---- collectives (new API; semantics in SURVEY.md 8(c), oracle/collectives.c) -------------- */

int b200mpi_bcast(void* buf, size_t count, int dtype, int root, int memkind);
/* send == recv means in place. */
int b200mpi_allreduce(const void* send, void* recv, size_t count, int dtype, int op, int memkind);
/* recv holds size()*count_per_rank elements, rank r's block at r*count_per_rank. */
int b200mpi_allgather(const void* send, void* recv, size_t count_per_rank, int dtype, int memkind);
/* send holds size()*count_per_rank elements; rank j receives op over r of rank r's block j
 * (rank order).  recv may be the caller's own block of send (in place).  This is the first half
 * of the two-shot Allreduce on its own. */
int b200mpi_reduce_scatter(const void* send, void* recv, size_t count_per_rank, int dtype, int op, int memkind);
/* Allreduce whose result lands on `root` only (recv may be NULL elsewhere). */
int b200mpi_reduce(const void* send, void* recv, size_t count, int dtype, int op, int root, int memkind);
/* send and recv hold size()*count_per_rank elements; block j of send becomes block rank() of rank j's recv. */
int b200mpi_alltoall(const void* send, void* recv, size_t count_per_rank, int dtype, int memkind);
int b200mpi_barrier(void);

/* Enqueue-only forms on the library stream (DEVICE memkind only). */
int b200mpi_bcast_async(void* buf, size_t count, int dtype, int root);
int b200mpi_allreduce_async(const void* send, void* recv, size_t count, int dtype, int op);
int b200mpi_allgather_async(const void* send, void* recv, size_t count_per_rank, int dtype);
int b200mpi_reduce_scatter_async(const void* send, void* recv, size_t count_per_rank, int dtype, int op);
int b200mpi_stream_sync(void); /* waits and reports device-side watchdog errors */

/* ---- tuning / measurement ------------------------------------------------------------------ */

int b200mpi_set_algo(int coll, int algo); /* force an algorithm (benches, tests) */
int b200mpi_get_algo(int coll, size_t count, int dtype); /* what AUTO resolves to (>=1) or <0 */
int b200mpi_set_max_blocks(int blocks); /* cap grid size (0 = default: SMs x occupancy) */
/* Tuning knobs for sweeps: "twoshot_unroll" (0|1), "nvls_unroll" (1|2|4|8), "nvls_min_ranks",
 * "nvls_max_blocks", "oneshot_max_bytes", "ll_max_bytes", "stage_chunk", "pipe_min_bytes",
 * "pipe_chunk_bytes", "bounce_chunk_bytes", "host_threads", "own_block_bytes", "copy_variant",
 * "hybrid_p2p_permille", "hybrid_p2p_blocks", "hybrid_min_bytes", "bcast_nvls2", "bcast_nvls_min",
 * "allgather_nvls_min", "host_register" (pin pageable host slices in place, cached; 0 drops the cache),
 * "p2p_fast" (host-slice Send/Receive through mapped pinned buffers), "watchdog_ms". */
int b200mpi_set_param(const char* name, int64_t value);
int b200mpi_get_param(const char* name, int64_t* value); /* also "sm_count", "shared_device" (read only) */
/* cudaStream_t used for collectives; set NULL to restore the library's own stream. */
int b200mpi_get_stream(void** stream);
int b200mpi_set_stream(void* stream);
/* CUDA-event stopwatch on the collective stream. */
int b200mpi_timer_start(void);
int b200mpi_timer_stop(float* ms); /* synchronises; elapsed since timer_start */
/* Measured bound of the host-slice paths: pinned host <-> device-heap copy rates of this rank's GPU,
 * GB/s (1e9): H2D alone, D2H alone, and per direction with both running at once. */
int b200mpi_pcie_probe(size_t bytes, int iters, double* h2d_gbs, double* d2h_gbs, double* bidir_gbs);
/* Raw NVLink rates of the library's copy kernel between neighbouring ranks (collective call).
 * mode 0: every rank pulls `bytes` from rank+1; 1: every rank pushes to rank+1; 2 / 3: only rank 0
 * pulls from / pushes to rank 1; 4: every rank pulls and pushes at once.  *ms: device time per iteration. */
int b200mpi_link_probe(size_t bytes, int mode, int iters, float* ms);
/* Number of kernels this library has launched since init. */
int64_t b200mpi_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* B200MPI_H_ */
