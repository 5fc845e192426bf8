// b200mpi.cu -- host side of libb200mpi.so: the C ABI of include/b200mpi.h.
//
// What it replaces in the reference (/root/reference):
//   Network.Init / Finalize / Rank / Size   network.go:41-65,354-369   -> b200mpi_init & co.
//   Network.Send / Receive (+tagManager)    network.go:449-625         -> b200mpi_send / _recv
//   local (same-rank rendezvous)            network.go:388-446         -> same mailbox, own pair
//   AllReduce stub                          mpi.go:130                 -> b200mpi_allreduce, _bcast, _allgather
// The data never touches a socket: payloads move by sm_100a kernels (kernels.cuh) through the
// peer-mapped heaps (heap.h).  There is no CPU implementation of any data call: without a CUDA
// device every one of them fails with B200MPI_ERR_NO_DEVICE.
#include <cuda.h>
#include <cuda_runtime.h>
#include <ctype.h>
#include <errno.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <unistd.h>

#include <sys/syscall.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <map>
#include <mutex>
#include <thread>
#include <set>
#include <string>
#include <vector>

#include "../../include/b200mpi.h"
#include "ctrl.h"
#include "heap.h"
#include "hostutil.h"
#include "kernels.cuh"

namespace b200 {

// ---------------------------------------------------------------------------------------------
// mailbox: host shared memory (memfd) used for Send/Receive rendezvous, the role of the
// reference's tagManager + per-connection channels (network.go:449-497) and of `local`
// (network.go:388-446).  Control only: payload bytes never pass through it.
// ---------------------------------------------------------------------------------------------
constexpr int kSlotsPerPair = 16;
constexpr int kMsgRegions = 4;
enum SlotState : uint32_t { kFree = 0, kClaimed = 1, kPosted = 2, kMatched = 3, kDone = 4 };

struct alignas(128) MsgSlot {
  std::atomic<uint32_t> state;
  int32_t tag;
  uint32_t dtype;
  int32_t result;          // receiver's verdict, informational
  uint64_t count;          // elements
  uint64_t total_bytes;
  uint64_t chunk_bytes;    // size of one posted region (== total_bytes for a direct post)
  uint32_t nregions;       // ring of staging regions the chunks cycle through (1 for a direct post)
  uint32_t pad_;
  uint64_t region_off[kMsgRegions]; // offsets in the SENDER's heap
  std::atomic<uint64_t> posted; // bytes made available so far
  std::atomic<uint64_t> done;   // bytes consumed so far
};

struct Mailbox {
  MsgSlot slots[B200MPI_MAX_RANKS][B200MPI_MAX_RANKS][kSlotsPerPair]; // [src][dst][k]
};

struct TagSet { // duplicate in-flight {peer,tag} detection: tagManager.Register, network.go:464-472
  std::mutex mu;
  std::set<int> tags;
  bool add(int t) {
    std::lock_guard<std::mutex> g(mu);
    return tags.insert(t).second;
  }
  void remove(int t) {
    std::lock_guard<std::mutex> g(mu);
    tags.erase(t);
  }
};

// NUMA node of a GPU (sysfs entry of its PCI device); -1 unknown.
static int numa_node_of_gpu(int dev) {
  char id[32] = {0};
  if (cudaDeviceGetPCIBusId(id, sizeof id, dev) != cudaSuccess) { (void)cudaGetLastError(); return -1; }
  for (char* p = id; *p; ++p) *p = (char)tolower(*p);
  return read_int_file(std::string("/sys/bus/pci/devices/") + id + "/numa_node", -1);
}
struct Ctx {
  Ctrl ctrl;
  Driver drv;
  Heap heap;
  bool initialised = false;
  bool control_only = false;
  int dev = -1;
  int sm_count = 0;
  bool shared_device = false; // several ranks on one GPU (functional-test mode)
  Comm comm = {};
  uint32_t epoch = 1;
  uint32_t scrub_every = kScrubEvery; // power of two; B200MPI_SCRUB_EVERY lowers it so that a test can reach the scrub
  uint64_t cur_sig = 0; // signature of the collective being launched (checked by sync_start on every rank)
  uint32_t* status_host = nullptr; // mapped pinned
  uint32_t* status_dev = nullptr;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr; // host-slice pipeline
  std::vector<cudaEvent_t> pipe_events;
  size_t pipe_min_bytes = 8u << 20, pipe_chunk_bytes = 16u << 20;
  Mailbox* box = nullptr;
  TagSet sendtags[B200MPI_MAX_RANKS], recvtags[B200MPI_MAX_RANKS];
  std::mutex stream_mu;
  std::vector<cudaStream_t> stream_pool;
  int algo[4] = {0, 0, 0, 0};
  int max_blocks = 0;
  int64_t watchdog_ns = 120ll * 1000000000ll;
  size_t stage_chunk = 32u << 20;
  // staging for collectives on non-heap buffers: [0] send side, [1] recv side
  size_t stage_off[2] = {0, 0}, stage_len[2] = {0, 0};
  std::atomic<int64_t> launches{0};
  static constexpr int kBounceSlots = 4;
  char* bounce = nullptr;      // pinned ring for pageable host slices: [in kBounceSlots][out kBounceSlots] x bounce_chunk
  size_t bounce_chunk = 0, bounce_chunk_bytes = 4u << 20;
  int host_threads = 8;        // helper threads copying between pageable memory and the bounce ring
  CopyPool pool;
  // Host-slice Send/Receive up to kP2PBounceBytes: device-mapped pinned bounce buffers, one copy kernel
  // that reads / writes host memory itself, and a completion word the host spins on (no cudaMemcpy
  // staging by the driver, no stream synchronisation).
  static constexpr int kP2PBounces = 4;
  static constexpr size_t kP2PBounceBytes = 4u << 20;
  struct P2PBounce { char* host = nullptr; char* dev = nullptr; uint32_t seq = 0; bool busy = false; };
  P2PBounce p2p_bounce[kP2PBounces];
  std::mutex p2p_mu;
  int p2p_fast = 1;
  // Opt-in (B200MPI_HOST_REGISTER=1 / "host_register"): pin the caller's pageable buffers in place
  // (cudaHostRegister, cached by address range) instead of bouncing them.  Only for callers whose
  // buffers stay mapped for the life of the cache (a freed and re-mapped range would alias stale
  // pinned pages); off by default.
  int host_register = 0;
  std::vector<std::pair<uintptr_t, size_t>> registered; // LRU, most recent last
  int gpu_numa_node = -1;      // NUMA node of the bound GPU (-1 unknown): host buffers and helper threads go there
  size_t oneshot_max_bytes = 256u << 10;
  int hybrid_p2p_permille = 0; // Allreduce HYBRID: share of the message that goes the P2P way, in 1/1000
  int hybrid_p2p_blocks = 0;   // CTAs given to the P2P part (0 = sm_count - nvls_max_blocks)
  int bcast_nvls2 = 1;         // Bcast NVLS: scatter + multicast (1) or root-only multicast (0)
  size_t bcast_nvls_min = 0, allgather_nvls_min = 0; // AUTO thresholds (bytes), set at init from n
  int twoshot_unroll = 1; // 0: 1/2/4 vectors per thread for n = 8/4/2, 1: 2/4/8
  int nvls_unroll = 2;      // 8 GPUs, 256 MiB: unroll 2 x 64 CTAs 810 GB/s, 4 x 148 CTAs 780 (profiles/r01/sweep_n8_nvls_blocks_unroll_v2.jsonl)
  int nvls_max_blocks = 48; // fewer requests in flight suit the switch reduction better (r02: 48 CTAs 822/845 GB/s at 256 MiB/1 GiB, 64 CTAs 817/840)
  size_t ll_max_bytes = 0; // AUTO uses the barrier-free LL allreduce for messages up to this size (<= 256 KiB)
  uint32_t ll_seq = 0;
  // pinned, device-mapped bounce buffers: a small host-slice Allreduce is a memcpy + ONE kernel that
  // reads its input over PCIe and writes result + completion word to host memory
  char* ll_host = nullptr; char* ll_dev = nullptr; // [in kLLCells*8][out kLLCells*8][done word]
  int copy_variant = 5; // 16 vectors in flight per thread, 256 threads: best of 8 launch shapes at 256 MiB and 1 GiB
  size_t hybrid_min_bytes = 32u << 20;
  int nvls_min_ranks = 4; // below this the fused two-shot moves fewer bytes per link than NVLS
  size_t own_block_bytes = 1u << 20; // interleave granularity of slice ownership (Owner in kernels.cuh)
};

static_assert(kHeapReserved == kCtrlBytes, "heap.h and kernels.cuh disagree about the control region");
static Ctx* g = nullptr;
static std::mutex g_mu;
static thread_local std::string t_err;

static int fail(int code, const std::string& msg) {
  t_err = msg;
  return code;
}
#define CUDA_OK(call)                                                                   \
  do {                                                                                  \
    cudaError_t _e = (call);                                                            \
    if (_e != cudaSuccess) return fail(B200MPI_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e)); \
  } while (0)

// Pinned host memory on the NUMA node of this rank's GPU.
static int numa_host_alloc(size_t bytes, void** p) {
  const cudaError_t e = on_numa_node(g->gpu_numa_node, [&] { return cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocDefault); });
  if (e != cudaSuccess) return fail(B200MPI_ERR_CUDA, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
  return 0;
}

static size_t esize(int dtype) {
  switch (dtype) {
    case B200MPI_U8: return 1;
    case B200MPI_I64: return 8;
    case B200MPI_F32: return 4;
    case B200MPI_F64: return 8;
  }
  return 0;
}

static thread_local int t_bound_dev = -1;
static int need_data_plane() {
  if (!g || !g->initialised) return fail(B200MPI_ERR_NOT_INIT, "mpi: Init has not been called");
  if (g->control_only) return fail(B200MPI_ERR_NO_DEVICE, "no CUDA device bound (control-plane-only init); there is no CPU data path");
  if (t_bound_dev != g->dev) { // the CUDA "current device" is per host thread; Send/Receive come from any goroutine/thread
    if (cudaSetDevice(g->dev) != cudaSuccess) return fail(B200MPI_ERR_CUDA, "cudaSetDevice failed on a caller thread");
    t_bound_dev = g->dev;
  }
  return 0;
}

using Clock = std::chrono::steady_clock;
struct Spinner { // host-side wait with watchdog
  Clock::time_point t0 = Clock::now();
  uint32_t it = 0;
  int64_t limit_ns;
  explicit Spinner(int64_t ns) : limit_ns(ns) {}
  bool step() { // false => timed out
    if (++it < 4096) {
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
      return true;
    }
    if ((it & 63) == 0) {
      if (limit_ns > 0 && std::chrono::duration_cast<std::chrono::nanoseconds>(Clock::now() - t0).count() > limit_ns) return false;
    }
    sched_yield();
    return true;
  }
};

static cudaStream_t borrow_stream() {
  {
    std::lock_guard<std::mutex> l(g->stream_mu);
    if (!g->stream_pool.empty()) {
      cudaStream_t s = g->stream_pool.back();
      g->stream_pool.pop_back();
      return s;
    }
  }
  cudaStream_t s = nullptr;
  cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  return s;
}
static void return_stream(cudaStream_t s) {
  std::lock_guard<std::mutex> l(g->stream_mu);
  g->stream_pool.push_back(s);
}

// The cross-rank barriers pair CTAs by blockIdx and need all of them resident at once: one CTA per
// SM (512 threads, __launch_bounds__(.,1)), so a grid never exceeds the SM count.
static int block_cap() {
  int cap = g->max_blocks > 0 ? std::min(g->max_blocks, g->sm_count) : g->sm_count;
  return std::min(cap, kMaxBlocks);
}

static int grid_for(size_t units_per_rank, int unroll) {
  size_t per_block = (size_t)kThreads * unroll;
  size_t want = (units_per_rank + per_block - 1) / per_block;
  int cap = block_cap();
  if (want < 1) want = 1;
  return (int)(want > (size_t)cap ? cap : want);
}

// Every launch advances the epoch by kEpochStride whatever the algorithm (kernels.cuh), so ranks
// that disagreed about one call still agree about the flag values of the next.
static uint64_t mix64(uint64_t x);
static Comm next_comm() {
  // Every kScrubEvery-th launch (the same one on every rank: epochs move in lockstep) is preceded by
  // scrub_kernel, which rewrites every slot row so that no flag word can fall 2^31 behind the epoch.
  if (g->ctrl.n > 1 && ((g->epoch / kEpochStride) & (g->scrub_every - 1)) == g->scrub_every - 1) {
    Comm s = g->comm;
    s.epoch = g->epoch;
    s.end_epoch = g->epoch + kEpochStride - 1;
    s.sig = mix64(0x5c2b0000ull ^ ((uint64_t)kMaxBlocks << 8 | 0x5u));
    g->epoch += kEpochStride;
    scrub_kernel<<<kMaxBlocks, 32, 0, g->stream>>>(s);
    (void)cudaGetLastError();
  }
  Comm c = g->comm;
  c.epoch = g->epoch;
  c.sig = g->cur_sig;
  c.end_epoch = g->epoch + kEpochStride - 1;
  g->epoch += kEpochStride;
  return c;
}

// After a failed collective the ranks may have issued different numbers of launches (chunked host
// paths) or LL calls: agree on max(epoch), max(ll_seq) over the control plane before the next call.
// Every rank sees the failure of a mismatched call (sync_start compares all signatures on every
// rank), so every rank comes here.
static void resync_after_error(bool peers_alive) {
  if (g->ctrl.n == 1) return;
  struct { uint32_t epoch, ll_seq; } mine = {g->epoch, g->ll_seq}, all[B200MPI_MAX_RANKS];
  std::string err;
  // A mismatched call ends on every rank, so everybody comes here and the exchange is prompt.  After
  // a watchdog expiry a peer may be gone or stuck: do not trade one hang for another.
  const int64_t wait_ns = peers_alive ? 30ll * 1000000000ll : 2ll * 1000000000ll;
  if (g->ctrl.allgather(&mine, sizeof mine, all, err, wait_ns) != 0) return; // nothing to agree with
  for (int r = 0; r < g->ctrl.n; ++r) {
    if ((int32_t)(all[r].epoch - g->epoch) > 0) g->epoch = all[r].epoch;
    if ((int32_t)(all[r].ll_seq - g->ll_seq) > 0) g->ll_seq = all[r].ll_seq;
  }
  g->epoch += kEpochStride;
  g->ll_seq += 2; // keep the parity, skip cells a failed call may have half written
}

static int check_status() {
  if (g->status_host && *(volatile uint32_t*)g->status_host) {
    const uint32_t st = *(volatile uint32_t*)g->status_host;
    *(volatile uint32_t*)g->status_host = 0;
    resync_after_error(st == 2u);
    if (st == 2u) return fail(B200MPI_ERR_PEER, "mismatched collective: ranks disagree on the call (collective, count, dtype, op, root, algorithm or grid); buffers were left untouched");
    return fail(B200MPI_ERR_TIMEOUT, "device-side watchdog: a peer did not reach the collective in time");
  }
  return 0;
}

// What this rank thinks the call is, folded into one word; equal on every rank of a well-formed
// call.  coll: 0 allreduce, 1 bcast, 2 allgather, 3 barrier, 4 reduce_scatter, 5 reduce, 6 alltoall.
// `total` is the element count of the whole user call when the launch is one chunk of a pipelined
// host-slice call (ranks that disagree about the total then disagree from the first chunk on).
static uint64_t mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
  x ^= x >> 27; x *= 0x94d049bb133111ebull;
  return x ^ (x >> 31);
}
static uint64_t make_sig(int coll, int dtype, int extra, int algo, size_t count, size_t total = 0) {
  uint64_t h = mix64((uint64_t)count + 0x9e3779b97f4a7c15ull);
  h = mix64(h ^ ((uint64_t)(coll & 15) << 40 | (uint64_t)(dtype & 15) << 32 | (uint64_t)(extra & 0xffff) << 16 | (uint64_t)(algo & 0xffff)));
  return mix64(h ^ (uint64_t)total);
}
// The grid is part of the signature: the barriers pair CTAs by blockIdx.
static void sig_grid(Comm& c, int blocks) { c.sig = mix64(c.sig ^ ((uint64_t)blocks << 8 | 0x5u)); }

static int launch_check(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(B200MPI_ERR_CUDA, std::string(what) + " launch failed: " + cudaGetErrorString(e));
  g->launches.fetch_add(1, std::memory_order_relaxed);
  return 0;
}

// dst <- src on `s` with the library's own copy kernel (local HBM or peer mapping).
static int launch_copy(void* dst, const void* src, size_t bytes, cudaStream_t s) {
  if (bytes == 0) return 0;
  unsigned char* d = (unsigned char*)dst;
  const unsigned char* c = (const unsigned char*)src;
  const size_t nvec = (bytes + 15) / 16;
  const int sms = g->sm_count;
  auto grid = [&](int threads, int unroll, int per_sm) {
    size_t want = (nvec + (size_t)threads * unroll - 1) / ((size_t)threads * unroll);
    size_t cap = g->max_blocks > 0 ? (size_t)g->max_blocks : (size_t)sms * per_sm;
    return (int)std::max<size_t>(1, std::min(want, cap));
  };
  switch (g->copy_variant) { // launch shapes for the HBM-bound local copy; see profiles/r01/SUMMARY.md
    case 1: copy_bytes_kernel<8, 512, 1><<<grid(512, 8, 1), 512, 0, s>>>(d, c, bytes); break;
    case 2: copy_bytes_kernel<4, 256, 2><<<grid(256, 4, 2), 256, 0, s>>>(d, c, bytes); break;
    case 3: copy_bytes_kernel<8, 256, 2><<<grid(256, 8, 2), 256, 0, s>>>(d, c, bytes); break;
    case 4: copy_bytes_kernel<4, 1024, 1><<<grid(1024, 4, 1), 1024, 0, s>>>(d, c, bytes); break;
    case 5: copy_bytes_kernel<16, 256, 1><<<grid(256, 16, 1), 256, 0, s>>>(d, c, bytes); break;
    case 7: copy_bytes_kernel<4, 512, 2><<<grid(512, 4, 2), 512, 0, s>>>(d, c, bytes); break;
    default: copy_bytes_kernel<4, 512, 1><<<grid(512, 4, 1), 512, 0, s>>>(d, c, bytes); break;
  }
  return launch_check("copy_bytes_kernel");
}

// ---------------------------------------------------------------------------------------------
// kernel dispatch
// ---------------------------------------------------------------------------------------------
// log2 of the ownership block, in 16-byte vectors: at most own_block_bytes, at most the per-rank
// share (so every rank owns something), at least `min_shift`.  Function of (count, n) only.
static uint32_t own_shift(size_t nvec, int n, uint32_t min_shift) {
  size_t per = std::max<size_t>((nvec + n - 1) / n, 1);
  uint32_t sh = 0;
  while (((size_t)2 << sh) <= per) ++sh;                        // floor(log2(per))
  uint32_t cap = 0;
  while (((size_t)32 << cap) <= g->own_block_bytes) ++cap;      // floor(log2(own_block_bytes / 16))
  sh = std::min(sh, cap);
  return std::max(sh, min_shift);
}

// only_dst: -1 allreduce, else the root of a Reduce (twoshot / nvls only)
template <typename T, typename Op>
static int launch_allreduce_body_t(int algo, uint64_t so, uint64_t ro, size_t count, int only_dst, cudaStream_t s) {
  const int n = g->ctrl.n;
  constexpr int EPV = 16 / sizeof(T);
  const size_t nvec = (count + EPV - 1) / EPV;
  const size_t per = (nvec + n - 1) / n;
  const uint32_t sh = own_shift(nvec, n, algo == B200MPI_ALGO_TWOSHOT_SMEM ? 8 : 0);
  Comm c = next_comm();
  switch (algo) {
    case B200MPI_ALGO_TWOSHOT: {
#define B200_TWOSHOT(NRV, UV)                                                                          \
  {                                                                                                    \
    const int blocks = grid_for(per, UV);                                                              \
    sig_grid(c, blocks);                                                                               \
    allreduce_twoshot_kernel<T, Op, NRV, UV><<<blocks, kThreads, 0, s>>>(c, so, ro, count, sh, only_dst); \
  }
      if (g->twoshot_unroll) {
        if (n == 2) B200_TWOSHOT(2, (sizeof(T) == 4 ? 8 : 4)) // 8-byte min/max spilled at 8 vectors per thread
        else if (n == 4) B200_TWOSHOT(4, 4)
        else if (n == 8) B200_TWOSHOT(8, 2)
        else B200_TWOSHOT(0, 1)
      } else {
        if (n == 2) B200_TWOSHOT(2, 4)
        else if (n == 4) B200_TWOSHOT(4, 2)
        else if (n == 8) B200_TWOSHOT(8, 1)
        else B200_TWOSHOT(0, 1)
      }
#undef B200_TWOSHOT
      return launch_check("allreduce_twoshot_kernel");
    }
    case B200MPI_ALGO_RING: {
      const int blocks = grid_for(per, 4);
      sig_grid(c, blocks);
      allreduce_ring_kernel<T, Op><<<blocks, kThreads, 0, s>>>(c, so, ro, count);
      return launch_check("allreduce_ring_kernel");
    }
    case B200MPI_ALGO_TWOSHOT_SMEM: {
      const size_t tiles = (per * 16 + kSmemChunk - 1) / kSmemChunk;
      const int blocks = (int)std::max<size_t>(1, std::min<size_t>(tiles, (size_t)block_cap()));
      sig_grid(c, blocks);
#define B200_SMEM(NRV)                                                                                           \
  {                                                                                                              \
    const size_t smem = (size_t)(kSmemStages * NRV + kSmemOutStages) * kSmemChunk;                               \
    static bool attr_set = false;                                                                                \
    if (!attr_set) {                                                                                             \
      cudaFuncSetAttribute(allreduce_twoshot_smem_kernel<T, Op, NRV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      attr_set = true;                                                                                           \
    }                                                                                                            \
    allreduce_twoshot_smem_kernel<T, Op, NRV><<<blocks, kSmemThreads, smem, s>>>(c, so, ro, count, sh);              \
  }
      if (n == 2) B200_SMEM(2)
      else if (n == 4) B200_SMEM(4)
      else B200_SMEM(8)
#undef B200_SMEM
      return launch_check("allreduce_twoshot_smem_kernel");
    }
  }
  return fail(B200MPI_ERR_UNSUPPORTED, "allreduce: algorithm not available for this dtype/op");
}

// One-shot geometry, shared by pick (which must know whether the call fits) and launch: a function
// of (count, n, element size, grid cap) only.
struct OneshotPlan { bool shfl; int blocks; size_t rounds; };
static OneshotPlan oneshot_plan(size_t count, size_t es) {
  const int n = g->ctrl.n;
  const size_t epv = 16 / es;
  const size_t nvec = (count + epv - 1) / epv;
  OneshotPlan p;
  p.shfl = (n == 2 || n == 4 || n == 8) && nvec <= 4096;
  p.blocks = p.shfl ? grid_for(nvec * n, 1) : grid_for(nvec, 1);
  const size_t per_round = (size_t)p.blocks * kThreads;
  // one mid-barrier value per round (upper bound: the scalar rounds of the unaligned path)
  p.rounds = (count + per_round - 1) / per_round + (p.shfl ? (nvec * n + per_round - 1) / per_round : 0) + 2;
  return p;
}

template <typename T, typename Op>
static int launch_allreduce_t(int algo, uint64_t so, uint64_t ro, size_t count, cudaStream_t s) {
  const int n = g->ctrl.n;
  if (algo != B200MPI_ALGO_ONESHOT) return launch_allreduce_body_t<T, Op>(algo, so, ro, count, -1, s);
  const OneshotPlan p = oneshot_plan(count, sizeof(T));
  if (p.rounds > kMaxMids) return fail(B200MPI_ERR_UNSUPPORTED, "allreduce: message too large for the one-shot algorithm");
  Comm c = next_comm();
  sig_grid(c, p.blocks);
  const int blocks = p.blocks;
  if (p.shfl) {
    if (n == 2) allreduce_oneshot_shfl_kernel<T, Op, 2><<<blocks, kThreads, 0, s>>>(c, so, ro, count);
    else if (n == 4) allreduce_oneshot_shfl_kernel<T, Op, 4><<<blocks, kThreads, 0, s>>>(c, so, ro, count);
    else allreduce_oneshot_shfl_kernel<T, Op, 8><<<blocks, kThreads, 0, s>>>(c, so, ro, count);
    return launch_check("allreduce_oneshot_shfl_kernel");
  }
  if (n == 2) allreduce_oneshot_kernel<T, Op, 2><<<blocks, kThreads, 0, s>>>(c, so, ro, count);
  else if (n == 4) allreduce_oneshot_kernel<T, Op, 4><<<blocks, kThreads, 0, s>>>(c, so, ro, count);
  else if (n == 8) allreduce_oneshot_kernel<T, Op, 8><<<blocks, kThreads, 0, s>>>(c, so, ro, count);
  else allreduce_oneshot_kernel<T, Op, 0><<<blocks, kThreads, 0, s>>>(c, so, ro, count);
  return launch_check("allreduce_oneshot_kernel");
}

static int nvls_blocks(size_t per, int u) {
  int blocks = grid_for(per, u);
  if (g->max_blocks == 0 && blocks > g->nvls_max_blocks) blocks = g->nvls_max_blocks;
  return blocks;
}

template <typename T, typename Op>
static int launch_allreduce_nvls_t(uint64_t so, uint64_t ro, size_t count, int only_dst, cudaStream_t s) {
  Comm c = next_comm();
  constexpr int EPV = 16 / sizeof(T);
  const size_t nvec = (count + EPV - 1) / EPV;
  const size_t per = (nvec + c.n - 1) / c.n;
  const uint32_t sh = own_shift(nvec, c.n, 0);
  int u = g->nvls_unroll;
  if (u != 1 && u != 2 && u != 8) u = 4;
  const int blocks = nvls_blocks(per, u);
  sig_grid(c, blocks);
  switch (u) {
    case 1: allreduce_nvls_kernel<T, Op, 1><<<blocks, kThreads, 0, s>>>(c, so, ro, count, sh, only_dst); break;
    case 2: allreduce_nvls_kernel<T, Op, 2><<<blocks, kThreads, 0, s>>>(c, so, ro, count, sh, only_dst); break;
    case 8: allreduce_nvls_kernel<T, Op, 8><<<blocks, kThreads, 0, s>>>(c, so, ro, count, sh, only_dst); break;
    default: allreduce_nvls_kernel<T, Op, 4><<<blocks, kThreads, 0, s>>>(c, so, ro, count, sh, only_dst); break;
  }
  return launch_check("allreduce_nvls_kernel");
}

// NVLS for [0, split) + fused P2P two-shot for the rest, in one kernel (kernels.cuh).  The split is
// cut on ownership-block boundaries of both parts and depends on (count, n, parameters) only.
template <typename T, typename Op>
static int launch_allreduce_hybrid_t(uint64_t so, uint64_t ro, size_t count, cudaStream_t s) {
  const int n = g->ctrl.n;
  constexpr int EPV = 16 / sizeof(T);
  const size_t nvec = count / EPV; // whole vectors; the tail goes with the P2P part
  size_t p2p_vec = (size_t)((double)nvec * g->hybrid_p2p_permille / 1000.0);
  const size_t gran = (size_t)n << 8; // n blocks of 4 KiB: every rank owns the same share of both parts
  p2p_vec = p2p_vec / gran * gran;
  const size_t split = nvec - p2p_vec;
  if (n != 8 && n != 4 && n != 2) return fail(B200MPI_ERR_UNSUPPORTED, "allreduce: the hybrid algorithm needs 2, 4 or 8 ranks");
  if (split == 0 || p2p_vec == 0) return launch_allreduce_nvls_t<T, Op>(so, ro, count, -1, s);
  Comm c = next_comm();
  const uint32_t sh_n = own_shift(split, n, 0), sh_p = own_shift(p2p_vec, n, 0);
  int u = g->nvls_unroll == 4 ? 4 : 2;
  const int nb_nvls = nvls_blocks((split + n - 1) / n, u);
  int nb_p2p = g->hybrid_p2p_blocks > 0 ? g->hybrid_p2p_blocks : block_cap() - nb_nvls;
  nb_p2p = std::max(1, std::min(nb_p2p, block_cap() - nb_nvls));
  if (nb_nvls + nb_p2p > block_cap()) return launch_allreduce_nvls_t<T, Op>(so, ro, count, -1, s);
  const int blocks = nb_nvls + nb_p2p;
  sig_grid(c, blocks);
#define B200_HYB(NRV, UP)                                                                                                                         \
  if (u == 4) allreduce_hybrid_kernel<T, Op, NRV, 4, UP><<<blocks, kThreads, 0, s>>>(c, so, ro, count, sh_n, sh_p, split, (unsigned)nb_nvls); \
  else allreduce_hybrid_kernel<T, Op, NRV, 2, UP><<<blocks, kThreads, 0, s>>>(c, so, ro, count, sh_n, sh_p, split, (unsigned)nb_nvls);
  if (n == 8) { B200_HYB(8, 2) }
  else if (n == 4) { B200_HYB(4, 4) }
  else { B200_HYB(2, 4) }
#undef B200_HYB
  return launch_check("allreduce_hybrid_kernel");
}

static bool nvls_supports(int dtype, int op) {
  if (op == B200MPI_SUM) return dtype == B200MPI_F32 || dtype == B200MPI_F64 || dtype == B200MPI_I64;
  return dtype == B200MPI_I64; // min/max: integer only in the switch
}

// `total`: element count of the whole user call when this launch is one chunk of it, else 0.
static int launch_allreduce(int algo, int dtype, int op, uint64_t so, uint64_t ro, size_t count, cudaStream_t s, size_t total = 0) {
  g->cur_sig = make_sig(0, dtype, op, algo, count, total);
  if (algo == B200MPI_ALGO_NVLS || algo == B200MPI_ALGO_HYBRID) {
    const bool hyb = algo == B200MPI_ALGO_HYBRID;
#define B200_NVLS(T, OP) return hyb ? launch_allreduce_hybrid_t<T, OP>(so, ro, count, s) : launch_allreduce_nvls_t<T, OP>(so, ro, count, -1, s);
    if (dtype == B200MPI_F32 && op == B200MPI_SUM) { B200_NVLS(float, OpSum) }
    if (dtype == B200MPI_F64 && op == B200MPI_SUM) { B200_NVLS(double, OpSum) }
    if (dtype == B200MPI_I64 && op == B200MPI_SUM) { B200_NVLS(long long, OpSum) }
    if (dtype == B200MPI_I64 && op == B200MPI_MAX) { B200_NVLS(long long, OpMax) }
    if (dtype == B200MPI_I64 && op == B200MPI_MIN) { B200_NVLS(long long, OpMin) }
#undef B200_NVLS
    return fail(B200MPI_ERR_UNSUPPORTED, "allreduce: NVLS supports sum (f32,f64,i64) and min/max (i64) only");
  }
#define B200_DISPATCH(T)                                                                         \
  switch (op) {                                                                                  \
    case B200MPI_SUM: return launch_allreduce_t<T, OpSum>(algo, so, ro, count, s);            \
    case B200MPI_MAX: return launch_allreduce_t<T, OpMax>(algo, so, ro, count, s);            \
    case B200MPI_MIN: return launch_allreduce_t<T, OpMin>(algo, so, ro, count, s);            \
  }                                                                                              \
  break;
  switch (dtype) {
    case B200MPI_F32: B200_DISPATCH(float)
    case B200MPI_F64: B200_DISPATCH(double)
    case B200MPI_I64: B200_DISPATCH(long long)
  }
#undef B200_DISPATCH
  return fail(B200MPI_ERR_UNSUPPORTED, "allreduce: unsupported dtype/op (need f32, f64 or i64 with sum/max/min)");
}

// Reduce to `root`: the owner-reduces bodies with a single destination.
static int launch_reduce(int algo, int dtype, int op, uint64_t so, uint64_t ro, size_t count, int root, cudaStream_t s) {
  g->cur_sig = make_sig(5, dtype, op | (root << 4), algo, count);
  if (algo == B200MPI_ALGO_NVLS) {
    if (dtype == B200MPI_F32 && op == B200MPI_SUM) return launch_allreduce_nvls_t<float, OpSum>(so, ro, count, root, s);
    if (dtype == B200MPI_F64 && op == B200MPI_SUM) return launch_allreduce_nvls_t<double, OpSum>(so, ro, count, root, s);
    if (dtype == B200MPI_I64 && op == B200MPI_SUM) return launch_allreduce_nvls_t<long long, OpSum>(so, ro, count, root, s);
    if (dtype == B200MPI_I64 && op == B200MPI_MAX) return launch_allreduce_nvls_t<long long, OpMax>(so, ro, count, root, s);
    if (dtype == B200MPI_I64 && op == B200MPI_MIN) return launch_allreduce_nvls_t<long long, OpMin>(so, ro, count, root, s);
  }
#define B200_DISPATCH(T)                                                                                          \
  switch (op) {                                                                                                   \
    case B200MPI_SUM: return launch_allreduce_body_t<T, OpSum>(B200MPI_ALGO_TWOSHOT, so, ro, count, root, s);   \
    case B200MPI_MAX: return launch_allreduce_body_t<T, OpMax>(B200MPI_ALGO_TWOSHOT, so, ro, count, root, s);   \
    case B200MPI_MIN: return launch_allreduce_body_t<T, OpMin>(B200MPI_ALGO_TWOSHOT, so, ro, count, root, s);   \
  }                                                                                                               \
  break;
  switch (dtype) {
    case B200MPI_F32: B200_DISPATCH(float)
    case B200MPI_F64: B200_DISPATCH(double)
    case B200MPI_I64: B200_DISPATCH(long long)
  }
#undef B200_DISPATCH
  return fail(B200MPI_ERR_UNSUPPORTED, "reduce: unsupported dtype/op (need f32, f64 or i64 with sum/max/min)");
}

template <typename T, typename Op>
static int launch_reduce_scatter_p2p_t(uint64_t so, uint64_t ro, size_t count, cudaStream_t s) {
  Comm c = next_comm();
  constexpr int EPV = 16 / sizeof(T);
  const int blocks = grid_for((count + EPV - 1) / EPV, 2);
  sig_grid(c, blocks);
  reduce_scatter_kernel<T, Op, 2><<<blocks, kThreads, 0, s>>>(c, so, ro, count);
  return launch_check("reduce_scatter_kernel");
}
template <typename T> // the switch form: sums only
static int launch_reduce_scatter_nvls_t(uint64_t so, uint64_t ro, size_t count, cudaStream_t s) {
  Comm c = next_comm();
  constexpr int EPV = 16 / sizeof(T);
  const int blocks = nvls_blocks((count + EPV - 1) / EPV, 2);
  sig_grid(c, blocks);
  reduce_scatter_nvls_kernel<T, OpSum, 2><<<blocks, kThreads, 0, s>>>(c, so, ro, count);
  return launch_check("reduce_scatter_nvls_kernel");
}

template <typename T>
static int launch_reduce_scatter_op(bool nvls, int op, uint64_t so, uint64_t ro, size_t count, cudaStream_t s) {
  switch (op) {
    case B200MPI_SUM: return nvls ? launch_reduce_scatter_nvls_t<T>(so, ro, count, s) : launch_reduce_scatter_p2p_t<T, OpSum>(so, ro, count, s);
    case B200MPI_MAX: return launch_reduce_scatter_p2p_t<T, OpMax>(so, ro, count, s);
    case B200MPI_MIN: return launch_reduce_scatter_p2p_t<T, OpMin>(so, ro, count, s);
  }
  return fail(B200MPI_ERR_ARG, "reduce_scatter: unknown op");
}

static int launch_reduce_scatter(int algo, int dtype, int op, uint64_t so, uint64_t ro, size_t count, cudaStream_t s) {
  const bool nvls = algo == B200MPI_ALGO_NVLS && op == B200MPI_SUM; // the switch form is instantiated for sums
  g->cur_sig = make_sig(4, dtype, op, nvls ? B200MPI_ALGO_NVLS : B200MPI_ALGO_TWOSHOT, count);
  switch (dtype) {
    case B200MPI_F32: return launch_reduce_scatter_op<float>(nvls, op, so, ro, count, s);
    case B200MPI_F64: return launch_reduce_scatter_op<double>(nvls, op, so, ro, count, s);
    case B200MPI_I64: return launch_reduce_scatter_op<long long>(nvls, op, so, ro, count, s);
  }
  return fail(B200MPI_ERR_UNSUPPORTED, "reduce_scatter: dtype must be f32, f64 or i64");
}

// AUTO for Allreduce.  Thresholds come from the sweeps on 2/4/8 B200s (profiles/r01, profiles/r02):
// best busbw / latency per (n, size) among the variants.
// chunk: the launch is one piece of a pipelined host-slice call, where the barrier-free LL kernel
// (own sequence numbers, no heap offsets) does not apply.
static int pick_allreduce(size_t bytes, int dtype, int op, bool chunk = false) {
  const int n = g->ctrl.n;
  const bool pow2 = n == 2 || n == 4 || n == 8;
  const size_t es = esize(dtype);
  int forced = g->algo[B200MPI_COLL_ALLREDUCE];
  const bool can_nvls = g->heap.mc_base && nvls_supports(dtype, op);
  if (forced == B200MPI_ALGO_TWOSHOT_SMEM && !pow2) forced = B200MPI_ALGO_TWOSHOT;
  if (forced == B200MPI_ALGO_NVLS && !can_nvls) forced = 0;
  if (forced == B200MPI_ALGO_HYBRID && !(can_nvls && pow2 && g->hybrid_p2p_permille > 0)) forced = can_nvls ? B200MPI_ALGO_NVLS : 0;
  if (forced == B200MPI_ALGO_LL && (bytes > ll_cells(g->ctrl.n) * 8 || chunk)) forced = chunk ? B200MPI_ALGO_ONESHOT : 0;
  if (forced == B200MPI_ALGO_ONESHOT && es && oneshot_plan(bytes / es, es).rounds > kMaxMids) forced = B200MPI_ALGO_TWOSHOT;
  if (forced) return forced;
  if (g->ll_max_bytes && bytes <= g->ll_max_bytes && !chunk) return B200MPI_ALGO_LL;
  const bool nvls = can_nvls && n >= g->nvls_min_ranks;
  const int big = pow2 ? B200MPI_ALGO_TWOSHOT_SMEM : B200MPI_ALGO_TWOSHOT;
  if (n >= 8) {
    if (nvls) return (pow2 && g->hybrid_p2p_permille > 0 && bytes >= g->hybrid_min_bytes) ? B200MPI_ALGO_HYBRID : B200MPI_ALGO_NVLS;
    return bytes < (2u << 20) ? B200MPI_ALGO_TWOSHOT : big;
  }
  if (n >= 3) {
    if (nvls && bytes <= (4u << 20)) return B200MPI_ALGO_NVLS;
    if (!nvls && bytes <= g->oneshot_max_bytes) return B200MPI_ALGO_ONESHOT;
    return bytes < (2u << 20) ? B200MPI_ALGO_TWOSHOT : big;
  }
  if (bytes <= g->oneshot_max_bytes) return B200MPI_ALGO_ONESHOT;
  return B200MPI_ALGO_TWOSHOT; // n == 2: LDG and TMA-staged two-shot tie (637 vs 636 GB/s at 256 MiB)
}

// AUTO for Bcast: the switch multicast wins up to a few MiB (one store stream, no second hop); the
// scatter + multicast form keeps winning above that (every byte leaves root once, every ingress
// link fills at the same time); without NVLS the fused pull-slice + push.
static int pick_bcast(size_t bytes) {
  const int n = g->ctrl.n;
  int forced = g->algo[B200MPI_COLL_BCAST];
  if (forced == B200MPI_ALGO_NVLS && !g->heap.mc_base) forced = 0;
  if (forced == B200MPI_ALGO_RING || forced == B200MPI_ALGO_TWOSHOT_SMEM || forced == B200MPI_ALGO_HYBRID || forced == B200MPI_ALGO_LL) forced = B200MPI_ALGO_TWOSHOT;
  if (forced) return forced;
  if (g->heap.mc_base && bytes % 16 == 0 && n >= 3 && (bytes <= (4u << 20) || (g->bcast_nvls2 && bytes >= g->bcast_nvls_min))) return B200MPI_ALGO_NVLS;
  if (n == 2 || bytes <= (64u << 10)) return B200MPI_ALGO_ONESHOT;
  return B200MPI_ALGO_TWOSHOT;
}

static int pick_allgather(size_t bytes_per_rank) {
  int forced = g->algo[B200MPI_COLL_ALLGATHER];
  if (forced == B200MPI_ALGO_RING) return forced;
  if (forced == B200MPI_ALGO_NVLS && g->heap.mc_base) return forced;
  if (forced) return B200MPI_ALGO_ONESHOT;
  if (g->heap.mc_base && g->ctrl.n >= 3 && bytes_per_rank % 16 == 0 && bytes_per_rank >= g->allgather_nvls_min) return B200MPI_ALGO_NVLS;
  return B200MPI_ALGO_ONESHOT; // direct push
}

static int pick_reduce_scatter(size_t, int dtype, int op) {
  int forced = g->algo[B200MPI_COLL_REDUCE_SCATTER];
  const bool can = g->heap.mc_base && op == B200MPI_SUM && nvls_supports(dtype, op);
  if (forced == B200MPI_ALGO_NVLS) return can ? B200MPI_ALGO_NVLS : B200MPI_ALGO_TWOSHOT;
  if (forced) return B200MPI_ALGO_TWOSHOT;
  return can && g->ctrl.n >= g->nvls_min_ranks ? B200MPI_ALGO_NVLS : B200MPI_ALGO_TWOSHOT;
}

// which: 0 allgather push, 1 allgather ring, 2 bcast (extra = root, mode 0 one-shot / 1 two-shot), 3 alltoall
// The grid is a function of the byte count only: ranks may instantiate different access widths
// (their offsets differ in alignment) but must launch the same number of CTAs, because the
// barriers pair CTAs by blockIdx.
template <typename U>
static void launch_units_u(int which, Comm& c, uint64_t a, uint64_t b, size_t bytes, int extra, int mode, cudaStream_t s) {
  constexpr int UNROLL = 4;
  const size_t vecs = bytes / 16 + 1;
  const size_t work = (which == 2 && mode == 1) ? vecs / (size_t)(c.n - 1) + 1 : vecs;
  const int blocks = grid_for(work, UNROLL);
  sig_grid(c, blocks);
  if (which == 0) allgather_push_kernel<U, UNROLL><<<blocks, kThreads, 0, s>>>(c, a, b, bytes);
  else if (which == 1) allgather_ring_kernel<U><<<blocks, kThreads, 0, s>>>(c, a, b, bytes);
  else if (which == 3) alltoall_kernel<U, UNROLL><<<blocks, kThreads, 0, s>>>(c, a, b, bytes);
  else bcast_kernel<U, UNROLL><<<blocks, kThreads, 0, s>>>(c, a, bytes, extra, mode);
}
static int launch_units(int which, Comm& c, uint64_t a, uint64_t b, size_t bytes, int extra, int mode, cudaStream_t s) {
  // The widest access unit that divides the size and THIS rank's offsets.  Peers may be aligned
  // differently: after sync_start every rank knows all offsets and drops to the byte-wide body if
  // some peer's are narrower than its own unit (kernels.cuh: all_aligned_to).
  const uint64_t m = a | b | bytes;
  if ((m & 15) == 0) launch_units_u<uint4>(which, c, a, b, bytes, extra, mode, s);
  else if ((m & 7) == 0) launch_units_u<unsigned long long>(which, c, a, b, bytes, extra, mode, s);
  else if ((m & 3) == 0) launch_units_u<unsigned int>(which, c, a, b, bytes, extra, mode, s);
  else launch_units_u<unsigned char>(which, c, a, b, bytes, extra, mode, s);
  return launch_check(which == 2 ? "bcast_kernel" : which == 3 ? "alltoall_kernel" : "allgather kernel");
}

// Bcast of `bytes` at heap offset `off` on every... (each rank passes its own offset)
static int launch_bcast(int algo, int dtype, size_t count, uint64_t off, size_t bytes, int root, cudaStream_t s, size_t total = 0) {
  g->cur_sig = make_sig(1, dtype, root, algo, count, total);
  Comm c = next_comm();
  if (algo == B200MPI_ALGO_NVLS) {
    const size_t nvec = bytes / 16 + 1;
    if (g->bcast_nvls2) {
      const size_t per = (nvec + c.n - 1) / c.n;
      const int blocks = grid_for(per, 4);
      sig_grid(c, blocks);
      bcast_nvls2_kernel<4><<<blocks, kThreads, 0, s>>>(c, off, bytes, root, own_shift(nvec, c.n, 0));
      return launch_check("bcast_nvls2_kernel");
    }
    const int blocks = grid_for(nvec, 4);
    sig_grid(c, blocks);
    bcast_nvls_kernel<4><<<blocks, kThreads, 0, s>>>(c, off, bytes, root);
    return launch_check("bcast_nvls_kernel");
  }
  return launch_units(2, c, off, off, bytes, root, algo == B200MPI_ALGO_TWOSHOT ? 1 : 0, s);
}

static int launch_allgather(int algo, int dtype, size_t count, uint64_t in_off, uint64_t out_off, size_t bytes, cudaStream_t s, size_t total = 0) {
  g->cur_sig = make_sig(2, dtype, 0, algo, count, total);
  Comm c = next_comm();
  if (algo == B200MPI_ALGO_NVLS) {
    const int blocks = grid_for(bytes / 16 + 1, 4);
    sig_grid(c, blocks);
    allgather_nvls_kernel<4><<<blocks, kThreads, 0, s>>>(c, in_off, out_off, bytes);
    return launch_check("allgather_nvls_kernel");
  }
  return launch_units(algo == B200MPI_ALGO_RING ? 1 : 0, c, in_off, out_off, bytes, 0, 0, s);
}

// ---------------------------------------------------------------------------------------------
// buffer resolution: heap-resident device memory is used in place; anything else (host slices,
// foreign device pointers) is staged through a heap block on the collective stream.
// ---------------------------------------------------------------------------------------------
static int ensure_stage(int which, size_t bytes) {
  if (g->stage_len[which] >= bytes && bytes > 0) return 0;
  if (g->stage_len[which]) {
    // the previous block may still be in use by enqueued work
    cudaStreamSynchronize(g->stream);
    g->heap.free_off(g->stage_off[which]);
    g->stage_len[which] = 0;
  }
  size_t off = 0;
  size_t want = std::max<size_t>(bytes, 1u << 20);
  if (g->heap.alloc(want, off)) return fail(B200MPI_ERR_NOMEM, "symmetric heap exhausted while staging " + std::to_string(bytes) + " bytes (raise B200MPI_HEAP_BYTES)");
  g->stage_off[which] = off;
  g->stage_len[which] = want;
  return 0;
}

struct Buf {
  uint64_t off = 0;
  bool staged = false;
};

static int resolve_in(const void* p, size_t bytes, int memkind, int which, Buf& out) {
  size_t off;
  if (memkind == B200MPI_DEVICE && g->heap.contains(p, bytes, off)) {
    out.off = off;
    out.staged = false;
    return 0;
  }
  int rc = ensure_stage(which, bytes);
  if (rc) return rc;
  out.off = g->stage_off[which];
  out.staged = true;
  if (bytes) {
    char* dst = (char*)g->heap.base[g->ctrl.rank] + out.off;
    if (memkind == B200MPI_HOST) CUDA_OK(cudaMemcpyAsync(dst, p, bytes, cudaMemcpyHostToDevice, g->stream));
    else { rc = launch_copy(dst, p, bytes, g->stream); if (rc) return rc; }
  }
  return 0;
}

static int resolve_out(void* p, size_t bytes, int memkind, int which, Buf& out) {
  size_t off;
  if (memkind == B200MPI_DEVICE && g->heap.contains(p, bytes, off)) {
    out.off = off;
    out.staged = false;
    return 0;
  }
  int rc = ensure_stage(which, bytes);
  if (rc) return rc;
  out.off = g->stage_off[which];
  out.staged = true;
  return 0;
}

static int copy_out(void* p, size_t bytes, int memkind, const Buf& b) {
  if (!b.staged || bytes == 0) return 0;
  const char* src = (const char*)g->heap.base[g->ctrl.rank] + b.off;
  if (memkind == B200MPI_HOST) CUDA_OK(cudaMemcpyAsync(p, src, bytes, cudaMemcpyDeviceToHost, g->stream));
  else return launch_copy(p, src, bytes, g->stream);
  return 0;
}

static int finish(bool async) {
  if (async) return 0;
  CUDA_OK(cudaStreamSynchronize(g->stream));
  return check_status();
}

static int local_copy(void* dst, const void* src, size_t bytes, int memkind, bool async) {
  if (dst == src || bytes == 0) return finish(async);
  if (memkind == B200MPI_HOST) {
    // world of one: the value still makes the round trip through the device (no CPU data path)
    Buf b;
    int rc = resolve_in(src, bytes, memkind, 0, b);
    if (rc) return rc;
    rc = copy_out(dst, bytes, memkind, b);
    if (rc) return rc;
  } else {
    int rc = launch_copy(dst, src, bytes, g->stream);
    if (rc) return rc;
  }
  return finish(async);
}

// Barrier-free LL allreduce for small messages (kernels.cuh): any local device pointers.
template <typename T, typename Op>
static int launch_ll_t(const void* send, void* recv, size_t count, uint32_t* done_host, cudaStream_t s) {
  Comm c = g->comm;
  const uint32_t seq = ++g->ll_seq;
  const size_t ncell = (count * sizeof(T) + 7) / 8;
  const int blocks = (int)std::max<size_t>(1, std::min<size_t>((ncell + 255) / 256, 128)); // one cell per thread up to 256 KiB
  allreduce_ll_kernel<T, Op><<<blocks, 256, 0, s>>>(c, (const T*)send, (T*)recv, count, seq, done_host);
  return launch_check("allreduce_ll_kernel");
}
static int launch_ll(int dtype, int op, const void* send, void* recv, size_t count, uint32_t* done_host, cudaStream_t s) {
#define B200_LL(T)                                                                       \
  switch (op) {                                                                          \
    case B200MPI_SUM: return launch_ll_t<T, OpSum>(send, recv, count, done_host, s);     \
    case B200MPI_MAX: return launch_ll_t<T, OpMax>(send, recv, count, done_host, s);     \
    case B200MPI_MIN: return launch_ll_t<T, OpMin>(send, recv, count, done_host, s);     \
  }                                                                                      \
  break;
  switch (dtype) {
    case B200MPI_F32: B200_LL(float)
    case B200MPI_F64: B200_LL(double)
    case B200MPI_I64: B200_LL(long long)
  }
#undef B200_LL
  return fail(B200MPI_ERR_UNSUPPORTED, "allreduce(LL): unsupported dtype/op");
}

// Small host-slice Allreduce: memcpy into a device-mapped pinned buffer, ONE kernel that loads its
// input over PCIe and stores result and completion word into host memory, spin on that word (no
// H2D / D2H copies, no stream synchronisation), memcpy out.
static int allreduce_ll_host(const void* send, void* recv, size_t count, int dtype, int op) {
  const size_t bytes = count * esize(dtype);
  const size_t half = kLLRegionCells / 4 * 8; // the largest message of any world size (n = 2)
  if (!g->ll_host) {
    CUDA_OK(cudaHostAlloc((void**)&g->ll_host, 2 * half + 64, cudaHostAllocMapped));
    CUDA_OK(cudaHostGetDevicePointer((void**)&g->ll_dev, g->ll_host, 0));
    *(volatile uint32_t*)(g->ll_host + 2 * half) = 0;
  }
  CUDA_OK(cudaStreamSynchronize(g->stream)); // enqueue-only calls issued earlier may still be running
  if (bytes) memcpy(g->ll_host, send, bytes);
  volatile uint32_t* done = (volatile uint32_t*)(g->ll_host + 2 * half);
  int rc = launch_ll(dtype, op, g->ll_dev, g->ll_dev + half, count, (uint32_t*)(g->ll_dev + 2 * half), g->stream);
  if (rc) return rc;
  const uint32_t seq = g->ll_seq;
  Spinner sp(g->watchdog_ns + 5000000000ll);
  while (*done != seq) {
    if (*(volatile uint32_t*)g->status_host) break; // device-side watchdog fired
    if (!sp.step()) break;
  }
  if (*done != seq) {
    cudaStreamSynchronize(g->stream);
    rc = check_status();
    return rc ? rc : fail(B200MPI_ERR_TIMEOUT, "allreduce(LL): kernel did not complete");
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  if (bytes) memcpy(recv, g->ll_host + half, bytes);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Host slices (what an unmodified Go caller passes, bounce.go:120-127): H2D, collective and D2H
// are pipelined in chunks over three streams so PCIe runs in both directions while the GPUs talk.
// Every rank derives the same chunking from (count, dtype, n), so the per-chunk collectives line
// up; the signature of each chunk also carries the total count.
//   pinned (cudaHostAlloc / cudaHostRegister'ed) memory: DMA straight from / to the caller's buffer;
//   pageable memory (a Go slice, a numpy array): through a ring of pinned bounce chunks filled and
//   drained by the helper threads of CopyPool while the DMA engines work on the neighbours.
// ---------------------------------------------------------------------------------------------
static bool is_pinned(const void* p) {
  if (!p) return true;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

static bool ensure_registered(const void* p, size_t bytes) {
  if (!p || bytes == 0) return true;
  const uintptr_t page = 4096, lo = (uintptr_t)p & ~(page - 1), hi = ((uintptr_t)p + bytes + page - 1) & ~(page - 1);
  for (size_t i = 0; i < g->registered.size(); ++i) {
    auto e = g->registered[i];
    if (lo >= e.first && hi <= e.first + e.second) {
      g->registered.erase(g->registered.begin() + i);
      g->registered.push_back(e);
      return true;
    }
  }
  // drop entries the new range overlaps (the caller re-used part of an old buffer) and the oldest beyond 16
  for (size_t i = 0; i < g->registered.size();) {
    auto e = g->registered[i];
    if (lo < e.first + e.second && e.first < hi) { cudaHostUnregister((void*)e.first); g->registered.erase(g->registered.begin() + i); }
    else ++i;
  }
  while (g->registered.size() >= 16) { cudaHostUnregister((void*)g->registered.front().first); g->registered.erase(g->registered.begin()); }
  if (cudaHostRegister((void*)lo, hi - lo, cudaHostRegisterDefault) != cudaSuccess) { (void)cudaGetLastError(); return false; }
  g->registered.push_back({lo, hi - lo});
  return true;
}

struct PipeSpec {
  int coll;          // B200MPI_COLL_ALLREDUCE / _BCAST / _ALLGATHER
  int dtype, op, root;
  size_t count;      // elements (per rank for allgather)
  const char* in;    // host input of this rank (nullptr: none, e.g. a bcast non-root)
  char* out;         // host output of this rank (nullptr: none, e.g. the bcast root)
};

static int ensure_bounce(size_t chunk) {
  if (g->bounce && g->bounce_chunk >= chunk) return 0;
  if (g->bounce) { cudaFreeHost(g->bounce); g->bounce = nullptr; }
  void* p = nullptr;
  int rc = numa_host_alloc(2 * (size_t)Ctx::kBounceSlots * chunk, &p);
  if (rc) return rc;
  g->bounce = (char*)p;
  g->bounce_chunk = chunk;
  if (!g->pool.running()) g->pool.start(g->host_threads, g->gpu_numa_node);
  return 0;
}

static int host_pipeline(const PipeSpec& sp) {
  const size_t es = esize(sp.dtype), B = sp.count * es;
  const int n = g->ctrl.n, R = Ctx::kBounceSlots;
  const bool ag = sp.coll == B200MPI_COLL_ALLGATHER;
  const size_t rows = ag ? (size_t)n : 1;
  bool pinned = is_pinned(sp.in) && is_pinned(sp.out);
  if (!pinned && g->host_register)
    pinned = (is_pinned(sp.in) || ensure_registered(sp.in, sp.in ? B : 0)) && (is_pinned(sp.out) || ensure_registered(sp.out, sp.out ? B * rows : 0));
  const bool bounce = !pinned && g->host_threads > 0;
  size_t chunk_bytes = g->pipe_chunk_bytes;
  if (bounce) chunk_bytes = std::min(chunk_bytes, g->bounce_chunk_bytes);
  size_t chunk_elems = std::max<size_t>(chunk_bytes / rows / es, 4096) / 4096 * 4096;
  const size_t nch = (sp.count + chunk_elems - 1) / chunk_elems;
  int rc = ensure_stage(0, B);
  if (rc) return rc;
  if (ag && n > 1 && (rc = ensure_stage(1, B * n))) return rc;
  if (bounce && (rc = ensure_bounce(chunk_elems * es * rows))) return rc;
  char* heap = (char*)g->heap.base[g->ctrl.rank];
  char* st_in = heap + g->stage_off[0];
  char* st_out = (ag && n > 1) ? heap + g->stage_off[1] : st_in; // world of one: recv block 0 is the send block
  while (g->pipe_events.size() < 3 * nch) {
    cudaEvent_t e;
    CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    g->pipe_events.push_back(e);
  }
  CUDA_OK(cudaStreamSynchronize(g->stream)); // earlier work may still use the staging blocks
  char* b_in = g->bounce;
  char* b_out = g->bounce ? g->bounce + (size_t)R * g->bounce_chunk : nullptr;
  std::atomic<int> pend_in[Ctx::kBounceSlots], pend_out[Ctx::kBounceSlots];
  for (int i = 0; i < R; ++i) { pend_in[i].store(0); pend_out[i].store(0); }
  // queued copy tasks point at these counters: whatever way this function is left (CUDA error paths
  // included), the helper threads are done with them first
  struct Drain {
    CopyPool& pool; std::atomic<int>* a; std::atomic<int>* b; int n; bool on;
    ~Drain() { if (on) for (int i = 0; i < n; ++i) { pool.wait(a[i]); pool.wait(b[i]); } }
  } drain{g->pool, pend_in, pend_out, R, bounce};
  auto lo_of = [&](size_t k) { return k * chunk_elems; };
  auto len_of = [&](size_t k) { return std::min(chunk_elems, sp.count - lo_of(k)); };

  // GPU side of chunk k: H2D, collective, D2H (enqueue only)
  auto gpu_stage = [&](size_t k) -> int {
    const size_t lo = lo_of(k) * es, len = len_of(k) * es;
    cudaEvent_t e_in = g->pipe_events[3 * k], e_coll = g->pipe_events[3 * k + 1], e_out = g->pipe_events[3 * k + 2];
    bool have_in = false;
    if (sp.in) {
      const char* src = sp.in + lo;
      if (bounce) { g->pool.wait(pend_in[k % R]); src = b_in + (k % R) * g->bounce_chunk; }
      CUDA_OK(cudaMemcpyAsync(st_in + lo, src, len, cudaMemcpyHostToDevice, g->h2d_stream));
      CUDA_OK(cudaEventRecord(e_in, g->h2d_stream));
      have_in = true;
    }
    cudaEvent_t ready = e_in;
    if (n > 1) {
      if (have_in) CUDA_OK(cudaStreamWaitEvent(g->stream, e_in, 0));
      const uint64_t off = g->stage_off[0] + lo;
      int r2;
      if (sp.coll == B200MPI_COLL_ALLREDUCE) r2 = launch_allreduce(pick_allreduce(len, sp.dtype, sp.op, true), sp.dtype, sp.op, off, off, len_of(k), g->stream, sp.count);
      else if (sp.coll == B200MPI_COLL_BCAST) r2 = launch_bcast(pick_bcast(len), sp.dtype, len_of(k), off, len, sp.root, g->stream, sp.count);
      else r2 = launch_allgather(pick_allgather(len), sp.dtype, len_of(k), off, g->stage_off[1] + lo * n, len, g->stream, sp.count);
      if (r2) return r2;
      CUDA_OK(cudaEventRecord(e_coll, g->stream));
      ready = e_coll;
    }
    if (sp.out) {
      if (n > 1 || have_in) CUDA_OK(cudaStreamWaitEvent(g->d2h_stream, ready, 0));
      const char* src = st_out + (ag && n > 1 ? lo * n : lo);
      if (bounce) {
        g->pool.wait(pend_out[k % R]); // the copy-out that used this slot R chunks ago
        CUDA_OK(cudaMemcpyAsync(b_out + (k % R) * g->bounce_chunk, src, len * rows, cudaMemcpyDeviceToHost, g->d2h_stream));
      } else if (rows == 1) {
        CUDA_OK(cudaMemcpyAsync(sp.out + lo, src, len, cudaMemcpyDeviceToHost, g->d2h_stream));
      } else { // allgather: n rows of len bytes, B apart in the caller's recv buffer
        CUDA_OK(cudaMemcpy2DAsync(sp.out + lo, B, src, len, len, rows, cudaMemcpyDeviceToHost, g->d2h_stream));
      }
      CUDA_OK(cudaEventRecord(e_out, g->d2h_stream));
    }
    return 0;
  };
  // host side after the D2H of chunk k (bounce only): scatter the chunk into the caller's buffer
  auto host_out = [&](size_t k) -> int {
    if (!bounce || !sp.out) return 0;
    CUDA_OK(cudaEventSynchronize(g->pipe_events[3 * k + 2]));
    const size_t lo = lo_of(k) * es, len = len_of(k) * es;
    for (size_t r = 0; r < rows; ++r) g->pool.submit(sp.out + r * B + lo, b_out + (k % R) * g->bounce_chunk + r * len, len, pend_out[k % R]);
    return 0;
  };
  const size_t lag = 2; // chunks between issuing the D2H of a chunk and draining its bounce slot
  size_t next_gpu = 0, next_out = 0; // next chunk whose GPU stage / host-side drain is due
  for (size_t k = 0; k < nch; ++k) {
    if (bounce && sp.in) {
      // the H2D that read this bounce slot R chunks ago (issued: next_gpu >= k-1 > k-R)
      if (k >= (size_t)R) CUDA_OK(cudaEventSynchronize(g->pipe_events[3 * (k - R)]));
      g->pool.submit(b_in + (k % R) * g->bounce_chunk, sp.in + lo_of(k) * es, len_of(k) * es, pend_in[k % R]);
      while (next_gpu < k) // the helper threads fill chunk k while the GPU works on chunk k-1
        if ((rc = gpu_stage(next_gpu++))) return rc;
    } else {
      if ((rc = gpu_stage(k))) return rc;
      next_gpu = k + 1;
    }
    while (next_out + lag < next_gpu)
      if ((rc = host_out(next_out++))) return rc;
  }
  while (next_gpu < nch)
    if ((rc = gpu_stage(next_gpu++))) return rc;
  while (next_out < nch)
    if ((rc = host_out(next_out++))) return rc;
  if (bounce && sp.out)
    for (int i = 0; i < R; ++i) g->pool.wait(pend_out[i]);
  CUDA_OK(cudaStreamSynchronize(g->d2h_stream));
  CUDA_OK(cudaStreamSynchronize(g->h2d_stream));
  CUDA_OK(cudaStreamSynchronize(g->stream));
  return check_status();
}

static int do_allreduce(const void* send, void* recv, size_t count, int dtype, int op, int memkind, bool async) {
  int rc = need_data_plane();
  if (rc) return rc;
  const size_t es = esize(dtype);
  if (es == 0 || dtype == B200MPI_U8) return fail(B200MPI_ERR_UNSUPPORTED, "allreduce: dtype must be f32, f64 or i64");
  if (op < B200MPI_SUM || op > B200MPI_MIN) return fail(B200MPI_ERR_ARG, "allreduce: unknown op");
  if (count && (!send || !recv)) return fail(B200MPI_ERR_ARG, "allreduce: NULL buffer with count > 0");
  if (memkind != B200MPI_HOST && memkind != B200MPI_DEVICE) return fail(B200MPI_ERR_ARG, "allreduce: bad memkind");
  const size_t bytes = count * es;
  if (memkind == B200MPI_HOST && !async && bytes >= g->pipe_min_bytes) {
    PipeSpec sp = {B200MPI_COLL_ALLREDUCE, dtype, op, 0, count, (const char*)send, (char*)recv};
    return host_pipeline(sp);
  }
  if (g->ctrl.n == 1) return local_copy(recv, send, bytes, memkind, async);
  const int algo = pick_allreduce(bytes, dtype, op);
  if (algo == B200MPI_ALGO_LL) {
    // any local device pointer works (peers never read it)
    if (memkind == B200MPI_DEVICE) {
      rc = launch_ll(dtype, op, send, recv, count, nullptr, g->stream);
      return rc ? rc : finish(async);
    }
    return allreduce_ll_host(send, recv, count, dtype, op);
  }
  Buf in, out;
  rc = resolve_in(send, bytes, memkind, 0, in);
  if (rc) return rc;
  if (send == recv) out = in;
  else { rc = resolve_out(recv, bytes, memkind, 1, out); if (rc) return rc; }
  rc = launch_allreduce(algo, dtype, op, in.off, out.off, count, g->stream);
  if (rc) return rc;
  rc = copy_out(recv, bytes, memkind, out);
  if (rc) return rc;
  return finish(async);
}

static int do_bcast(void* buf, size_t count, int dtype, int root, int memkind, bool async) {
  int rc = need_data_plane();
  if (rc) return rc;
  const size_t es = esize(dtype);
  if (es == 0) return fail(B200MPI_ERR_ARG, "bcast: unknown dtype");
  if (root < 0 || root >= g->ctrl.n) return fail(B200MPI_ERR_ARG, "bcast: root out of range");
  if (count && !buf) return fail(B200MPI_ERR_ARG, "bcast: NULL buffer with count > 0");
  if (memkind != B200MPI_HOST && memkind != B200MPI_DEVICE) return fail(B200MPI_ERR_ARG, "bcast: bad memkind");
  if (g->ctrl.n == 1) return finish(async);
  const size_t bytes = count * es;
  const bool is_root = g->ctrl.rank == root;
  if (memkind == B200MPI_HOST && !async && bytes >= g->pipe_min_bytes) {
    PipeSpec sp = {B200MPI_COLL_BCAST, dtype, 0, root, count, is_root ? (const char*)buf : nullptr, is_root ? nullptr : (char*)buf};
    return host_pipeline(sp);
  }
  Buf b;
  if (is_root) rc = resolve_in(buf, bytes, memkind, 0, b);
  else rc = resolve_out(buf, bytes, memkind, 0, b);
  if (rc) return rc;
  rc = launch_bcast(pick_bcast(bytes), dtype, count, b.off, bytes, root, g->stream);
  if (rc) return rc;
  if (!is_root) { rc = copy_out(buf, bytes, memkind, b); if (rc) return rc; }
  return finish(async);
}

static int do_allgather(const void* send, void* recv, size_t count, int dtype, int memkind, bool async) {
  int rc = need_data_plane();
  if (rc) return rc;
  const size_t es = esize(dtype);
  if (es == 0) return fail(B200MPI_ERR_ARG, "allgather: unknown dtype");
  if (count && (!send || !recv)) return fail(B200MPI_ERR_ARG, "allgather: NULL buffer with count > 0");
  if (memkind != B200MPI_HOST && memkind != B200MPI_DEVICE) return fail(B200MPI_ERR_ARG, "allgather: bad memkind");
  const size_t bytes = count * es;
  const int n = g->ctrl.n;
  if (memkind == B200MPI_HOST && !async && bytes * n >= g->pipe_min_bytes) {
    PipeSpec sp = {B200MPI_COLL_ALLGATHER, dtype, 0, 0, count, (const char*)send, (char*)recv};
    return host_pipeline(sp);
  }
  if (n == 1) return local_copy(recv, send, bytes, memkind, async);
  Buf out, in;
  rc = resolve_out(recv, bytes * n, memkind, 1, out);
  if (rc) return rc;
  const bool inplace = (const char*)send == (const char*)recv + (size_t)g->ctrl.rank * bytes;
  if (inplace && !out.staged) { in.off = out.off + (size_t)g->ctrl.rank * bytes; }
  else { rc = resolve_in(send, bytes, memkind, 0, in); if (rc) return rc; }
  rc = launch_allgather(pick_allgather(bytes), dtype, count, in.off, out.off, bytes, g->stream);
  if (rc) return rc;
  rc = copy_out(recv, bytes * n, memkind, out);
  if (rc) return rc;
  return finish(async);
}

// ReduceScatter: send holds size()*count elements (block j is reduced onto rank j), recv holds count.
static int do_reduce_scatter(const void* send, void* recv, size_t count, int dtype, int op, int memkind, bool async) {
  int rc = need_data_plane();
  if (rc) return rc;
  const size_t es = esize(dtype);
  if (es == 0 || dtype == B200MPI_U8) return fail(B200MPI_ERR_UNSUPPORTED, "reduce_scatter: dtype must be f32, f64 or i64");
  if (op < B200MPI_SUM || op > B200MPI_MIN) return fail(B200MPI_ERR_ARG, "reduce_scatter: unknown op");
  if (count && (!send || !recv)) return fail(B200MPI_ERR_ARG, "reduce_scatter: NULL buffer with count > 0");
  if (memkind != B200MPI_HOST && memkind != B200MPI_DEVICE) return fail(B200MPI_ERR_ARG, "reduce_scatter: bad memkind");
  const size_t bytes = count * es;
  const int n = g->ctrl.n;
  if (n == 1) return local_copy(recv, send, bytes, memkind, async);
  Buf in, out;
  rc = resolve_in(send, bytes * n, memkind, 0, in);
  if (rc) return rc;
  const bool inplace = (const char*)recv == (const char*)send + (size_t)g->ctrl.rank * bytes;
  if (inplace) { out.off = in.off + (size_t)g->ctrl.rank * bytes; out.staged = in.staged; }
  else { rc = resolve_out(recv, bytes, memkind, 1, out); if (rc) return rc; }
  rc = launch_reduce_scatter(pick_reduce_scatter(bytes, dtype, op), dtype, op, in.off, out.off, count, g->stream);
  if (rc) return rc;
  rc = copy_out(recv, bytes, memkind, out);
  if (rc) return rc;
  return finish(async);
}

// Reduce: like Allreduce but only `root`'s recv is written (recv may be NULL elsewhere).
static int do_reduce(const void* send, void* recv, size_t count, int dtype, int op, int root, int memkind, bool async) {
  int rc = need_data_plane();
  if (rc) return rc;
  const size_t es = esize(dtype);
  const int n = g->ctrl.n;
  if (es == 0 || dtype == B200MPI_U8) return fail(B200MPI_ERR_UNSUPPORTED, "reduce: dtype must be f32, f64 or i64");
  if (op < B200MPI_SUM || op > B200MPI_MIN) return fail(B200MPI_ERR_ARG, "reduce: unknown op");
  if (root < 0 || root >= n) return fail(B200MPI_ERR_ARG, "reduce: root out of range");
  const bool is_root = g->ctrl.rank == root;
  if (count && (!send || (is_root && !recv))) return fail(B200MPI_ERR_ARG, "reduce: NULL buffer with count > 0");
  if (memkind != B200MPI_HOST && memkind != B200MPI_DEVICE) return fail(B200MPI_ERR_ARG, "reduce: bad memkind");
  const size_t bytes = count * es;
  if (n == 1) return local_copy(recv, send, bytes, memkind, async);
  Buf in, out;
  rc = resolve_in(send, bytes, memkind, 0, in);
  if (rc) return rc;
  out = in; // non-roots announce their send offset as recv offset; nobody writes there
  if (is_root && recv != send) { rc = resolve_out(recv, bytes, memkind, 1, out); if (rc) return rc; }
  int algo = g->algo[B200MPI_COLL_ALLREDUCE] == B200MPI_ALGO_TWOSHOT ? B200MPI_ALGO_TWOSHOT
             : (g->heap.mc_base && nvls_supports(dtype, op) && n >= g->nvls_min_ranks) ? B200MPI_ALGO_NVLS : B200MPI_ALGO_TWOSHOT;
  rc = launch_reduce(algo, dtype, op, in.off, out.off, count, root, g->stream);
  if (rc) return rc;
  if (is_root) { rc = copy_out(recv, bytes, memkind, out); if (rc) return rc; }
  return finish(async);
}

// Alltoall: send and recv hold size()*count elements; block j of send goes to rank j's block rank().
static int do_alltoall(const void* send, void* recv, size_t count, int dtype, int memkind, bool async) {
  int rc = need_data_plane();
  if (rc) return rc;
  const size_t es = esize(dtype);
  if (es == 0) return fail(B200MPI_ERR_ARG, "alltoall: unknown dtype");
  if (count && (!send || !recv)) return fail(B200MPI_ERR_ARG, "alltoall: NULL buffer with count > 0");
  if (send == recv && count) return fail(B200MPI_ERR_ARG, "alltoall: in place is not supported");
  if (memkind != B200MPI_HOST && memkind != B200MPI_DEVICE) return fail(B200MPI_ERR_ARG, "alltoall: bad memkind");
  const size_t bytes = count * es;
  const int n = g->ctrl.n;
  if (n == 1) return local_copy(recv, send, bytes, memkind, async);
  Buf in, out;
  rc = resolve_in(send, bytes * n, memkind, 0, in);
  if (rc) return rc;
  rc = resolve_out(recv, bytes * n, memkind, 1, out);
  if (rc) return rc;
  g->cur_sig = make_sig(6, dtype, 0, 0, count);
  Comm c = next_comm();
  rc = launch_units(3, c, in.off, out.off, bytes, 0, 0, g->stream);
  if (rc) return rc;
  rc = copy_out(recv, bytes * n, memkind, out);
  if (rc) return rc;
  return finish(async);
}

// ---------------------------------------------------------------------------------------------
// point to point
// ---------------------------------------------------------------------------------------------
static Ctx::P2PBounce* acquire_bounce() {
  std::lock_guard<std::mutex> l(g->p2p_mu);
  for (auto& b : g->p2p_bounce) {
    if (b.busy) continue;
    if (!b.host) {
      void* p = nullptr;
      const size_t len = Ctx::kP2PBounceBytes + 64;
      cudaError_t e = on_numa_node(g->gpu_numa_node, [&] { return cudaHostAlloc(&p, len, cudaHostAllocMapped); });
      if (e != cudaSuccess || cudaHostGetDevicePointer((void**)&b.dev, p, 0) != cudaSuccess) { (void)cudaGetLastError(); if (p) cudaFreeHost(p); return nullptr; }
      b.host = (char*)p;
      *(volatile uint32_t*)(b.host + Ctx::kP2PBounceBytes) = 0;
    }
    b.busy = true;
    return &b;
  }
  return nullptr;
}
static void release_bounce(Ctx::P2PBounce* b) {
  std::lock_guard<std::mutex> l(g->p2p_mu);
  b->busy = false;
}
// dst <- src with the copy kernel on `s`, then spin on the bounce's completion word.  One of dst / src
// is the bounce's device mapping.  0, or an error (watchdog / launch failure).
static int copy_and_wait(Ctx::P2PBounce* b, void* dst, const void* src, size_t bytes, cudaStream_t s) {
  int rc = launch_copy(dst, src, bytes, s);
  if (rc) return rc;
  const uint32_t seq = ++b->seq;
  flag_kernel<<<1, 1, 0, s>>>((uint32_t*)(b->dev + Ctx::kP2PBounceBytes), seq);
  if ((rc = launch_check("flag_kernel"))) return rc;
  volatile uint32_t* done = (volatile uint32_t*)(b->host + Ctx::kP2PBounceBytes);
  Spinner sp(g->watchdog_ns);
  while (*done != seq)
    if (!sp.step()) { cudaStreamSynchronize(s); return fail(B200MPI_ERR_TIMEOUT, "p2p: copy kernel did not complete"); }
  std::atomic_thread_fence(std::memory_order_acquire);
  return 0;
}

static size_t p2p_chunk(size_t bytes, int memkind) {
  // Host slices: small chunks so the sender's H2D of chunk k+1 overlaps the receiver's pull + D2H of
  // chunk k (both block their host thread for pageable memory).  Device buffers: one big chunk.
  if (memkind == B200MPI_HOST) return std::min(g->stage_chunk, std::max<size_t>(64u << 10, (bytes / 4 + 4095) / 4096 * 4096));
  return g->stage_chunk;
}

static MsgSlot* claim_slot(int dest) {
  const int me = g->ctrl.rank;
  Spinner sp(g->watchdog_ns);
  for (;;) {
    for (int k = 0; k < kSlotsPerPair; ++k) {
      uint32_t expect = kFree;
      MsgSlot* s = &g->box->slots[me][dest][k];
      if (s->state.compare_exchange_strong(expect, kClaimed, std::memory_order_acq_rel)) return s;
    }
    if (!sp.step()) return nullptr;
  }
}

// The message could not be delivered (timeout / staging error).  Withdraw the post if nobody
// matched it; if a receiver is in the middle of pulling, the staging block must outlive it.
// Returns true when the slot and the staging block may be reused.
static bool retire_failed_send(MsgSlot* slot) {
  uint32_t expect = kPosted;
  if (slot->state.compare_exchange_strong(expect, kFree, std::memory_order_acq_rel)) return true;
  if (expect == kClaimed) { slot->state.store(kFree, std::memory_order_release); return true; }
  Spinner sp(g->watchdog_ns);
  while (slot->state.load(std::memory_order_acquire) != kDone)
    if (!sp.step()) return false; // receiver is stuck mid-pull: leak slot and staging rather than free memory it reads
  slot->state.store(kFree, std::memory_order_release);
  return true;
}

struct PendingSend { MsgSlot* slot; size_t stage; bool staged; };
static std::mutex g_pending_mu;
static std::map<std::pair<int, int>, PendingSend> g_pending; // {dest, tag} -> posted, not yet acknowledged (Isend)

// wait_ack = true: mpi.Send (network.go:518-572), returns after the receiver's acknowledgement.
// wait_ack = false: the Send of the design sketched in mpi.go:132-152 -- returns once the data has
// left the caller's buffer ("sent on connection"); b200mpi_wait collects the acknowledgement.
static int do_send(const void* buf, size_t count, int dtype, int dest, int tag, int memkind, bool wait_ack) {
  int rc = need_data_plane();
  if (rc) return rc;
  const size_t es = esize(dtype);
  if (es == 0) return fail(B200MPI_ERR_ARG, "send: unknown dtype");
  if (dest < 0 || dest >= g->ctrl.n) return fail(B200MPI_ERR_ARG, "send: destination " + std::to_string(dest) + " out of range");
  if (count && !buf) return fail(B200MPI_ERR_ARG, "send: NULL buffer with count > 0");
  if (memkind != B200MPI_HOST && memkind != B200MPI_DEVICE) return fail(B200MPI_ERR_ARG, "send: bad memkind");
  if (!g->sendtags[dest].add(tag)) return fail(B200MPI_ERR_TAG_EXISTS, "Tag " + std::to_string(tag) + " already in use sending"); // mpi.go:180-182
  const size_t bytes = count * es;
  const int me = g->ctrl.rank;
  MsgSlot* slot = claim_slot(dest);
  if (!slot) {
    g->sendtags[dest].remove(tag);
    return fail(B200MPI_ERR_TIMEOUT, "send: no free mailbox slot");
  }
  slot->tag = tag;
  slot->dtype = (uint32_t)dtype;
  slot->result = 0;
  slot->count = count;
  slot->total_bytes = bytes;
  slot->done.store(0, std::memory_order_relaxed);
  size_t off = 0;
  size_t stage = 0;
  bool staged = false;
  cudaStream_t s = nullptr;
  Ctx::P2PBounce* bounce = nullptr;
  rc = 0;
  if (bytes == 0 || (wait_ack && memkind == B200MPI_DEVICE && g->heap.contains(buf, bytes, off))) {
    // heap-resident: the receiver pulls straight out of the caller's buffer
    slot->chunk_bytes = bytes;
    slot->nregions = 1;
    slot->region_off[0] = off;
    slot->posted.store(bytes, std::memory_order_relaxed);
    slot->state.store(kPosted, std::memory_order_release);
  } else if (memkind == B200MPI_HOST && g->p2p_fast && bytes <= Ctx::kP2PBounceBytes && (bounce = acquire_bounce()) != nullptr) {
    // small host slice: memcpy into a mapped pinned buffer, one kernel moves it into the heap (the
    // SMs read host memory over PCIe), the host spins on the completion word, then the whole
    // message is posted at once
    staged = true;
    if (g->heap.alloc(bytes, stage)) {
      release_bounce(bounce);
      slot->state.store(kFree, std::memory_order_release);
      g->sendtags[dest].remove(tag);
      return fail(B200MPI_ERR_NOMEM, "send: symmetric heap exhausted while staging (raise B200MPI_HEAP_BYTES)");
    }
    memcpy(bounce->host, buf, bytes);
    s = borrow_stream();
    rc = copy_and_wait(bounce, (char*)g->heap.base[me] + stage, bounce->dev, bytes, s);
    release_bounce(bounce);
    slot->chunk_bytes = bytes;
    slot->nregions = 1;
    slot->region_off[0] = stage;
    slot->posted.store(rc == 0 ? bytes : 0, std::memory_order_relaxed);
    slot->state.store(kPosted, std::memory_order_release);
  } else {
    staged = true;
    // Isend keeps the whole message in the staging block until Wait; Send streams it through a ring
    const size_t chunk = wait_ack ? std::min(bytes, p2p_chunk(bytes, memkind)) : bytes;
    const size_t nchunks = (bytes + chunk - 1) / chunk;
    const size_t nreg = std::min<size_t>(nchunks, kMsgRegions);
    if (g->heap.alloc(chunk * nreg, stage)) {
      slot->state.store(kFree, std::memory_order_release);
      g->sendtags[dest].remove(tag);
      return fail(B200MPI_ERR_NOMEM, "send: symmetric heap exhausted while staging (raise B200MPI_HEAP_BYTES)");
    }
    slot->chunk_bytes = chunk;
    slot->nregions = (uint32_t)nreg;
    for (size_t r = 0; r < nreg; ++r) slot->region_off[r] = stage + r * chunk;
    slot->posted.store(0, std::memory_order_relaxed);
    slot->state.store(kPosted, std::memory_order_release);
    s = borrow_stream();
    char* base = (char*)g->heap.base[me];
    for (size_t k = 0; k < nchunks && rc == 0; ++k) {
      if (k >= nreg) { // the region was used by chunk k-nreg: wait until the receiver drained it
        Spinner sp(g->watchdog_ns);
        while (slot->done.load(std::memory_order_acquire) < (k - nreg + 1) * chunk && slot->state.load(std::memory_order_acquire) != kDone)
          if (!sp.step()) { rc = fail(B200MPI_ERR_TIMEOUT, "send: receiver stalled"); break; }
        if (rc) break;
      }
      if (slot->state.load(std::memory_order_acquire) == kDone) break; // receiver gave up (truncate)
      const size_t lo = k * chunk, len = std::min(chunk, bytes - lo);
      cudaError_t e;
      if (memkind == B200MPI_HOST) e = cudaMemcpyAsync(base + slot->region_off[k % nreg], (const char*)buf + lo, len, cudaMemcpyHostToDevice, s);
      else { e = cudaSuccess; rc = launch_copy(base + slot->region_off[k % nreg], (const char*)buf + lo, len, s); }
      if (e == cudaSuccess && rc == 0) e = cudaStreamSynchronize(s);
      if (e != cudaSuccess) { rc = fail(B200MPI_ERR_CUDA, std::string("send staging: ") + cudaGetErrorString(e)); break; }
      slot->posted.store(lo + len, std::memory_order_release);
    }
  }
  if (s) return_stream(s);
  if (rc == 0 && !wait_ack) { // Isend: the acknowledgement is collected by b200mpi_wait
    std::lock_guard<std::mutex> l(g_pending_mu);
    g_pending[{dest, tag}] = {slot, stage, staged};
    return 0;
  }
  // wait for the receiver's acknowledgement (network.go:569)
  if (rc == 0) {
    Spinner sp(g->watchdog_ns);
    while (slot->state.load(std::memory_order_acquire) != kDone)
      if (!sp.step()) { rc = fail(B200MPI_ERR_TIMEOUT, "send: no matching receive within the watchdog time"); break; }
  }
  if (rc == 0) {
    if (staged) g->heap.free_off(stage);
    slot->state.store(kFree, std::memory_order_release);
  } else {
    const std::string why = t_err;
    if (retire_failed_send(slot) && staged) g->heap.free_off(stage);
    t_err = why;
  }
  g->sendtags[dest].remove(tag);
  return rc;
}

// mpi.go:146-152 (commented design): block until `destination` confirmed the message sent with
// `tag`, then free the {destination, tag} pair.
static int do_wait(int dest, int tag) {
  int rc = need_data_plane();
  if (rc) return rc;
  PendingSend ps;
  {
    std::lock_guard<std::mutex> l(g_pending_mu);
    auto it = g_pending.find({dest, tag});
    if (it == g_pending.end()) return fail(B200MPI_ERR_ARG, "wait: no unacknowledged send to " + std::to_string(dest) + " with tag " + std::to_string(tag));
    ps = it->second;
    g_pending.erase(it);
  }
  Spinner sp(g->watchdog_ns);
  while (ps.slot->state.load(std::memory_order_acquire) != kDone)
    if (!sp.step()) { rc = fail(B200MPI_ERR_TIMEOUT, "wait: no matching receive within the watchdog time"); break; }
  if (rc == 0) {
    if (ps.staged) g->heap.free_off(ps.stage);
    ps.slot->state.store(kFree, std::memory_order_release);
  } else {
    const std::string why = t_err;
    if (retire_failed_send(ps.slot) && ps.staged) g->heap.free_off(ps.stage);
    t_err = why;
  }
  g->sendtags[dest].remove(tag);
  return rc;
}

static int do_recv(void* buf, size_t capacity, size_t* count_out, int dtype, int src, int tag, int memkind) {
  int rc = need_data_plane();
  if (rc) return rc;
  const size_t es = esize(dtype);
  if (es == 0) return fail(B200MPI_ERR_ARG, "recv: unknown dtype");
  if (src < 0 || src >= g->ctrl.n) return fail(B200MPI_ERR_ARG, "recv: source " + std::to_string(src) + " out of range");
  if (capacity && !buf) return fail(B200MPI_ERR_ARG, "recv: NULL buffer with capacity > 0");
  if (memkind != B200MPI_HOST && memkind != B200MPI_DEVICE) return fail(B200MPI_ERR_ARG, "recv: bad memkind");
  if (!g->recvtags[src].add(tag)) return fail(B200MPI_ERR_TAG_EXISTS, "Tag " + std::to_string(tag) + " already in use receiving");
  const int me = g->ctrl.rank;
  MsgSlot* slot = nullptr;
  {
    Spinner sp(g->watchdog_ns);
    for (;;) {
      for (int k = 0; k < kSlotsPerPair && !slot; ++k) {
        MsgSlot* s = &g->box->slots[src][me][k];
        if (s->state.load(std::memory_order_acquire) == kPosted && s->tag == tag) {
          uint32_t expect = kPosted;
          if (s->state.compare_exchange_strong(expect, kMatched, std::memory_order_acq_rel)) {
            if (s->tag == tag) slot = s;
            else s->state.store(kPosted, std::memory_order_release); // the slot was recycled between the look and the CAS
          }
        }
      }
      if (slot) break;
      if (!sp.step()) {
        g->recvtags[src].remove(tag);
        return fail(B200MPI_ERR_TIMEOUT, "recv: no matching send within the watchdog time");
      }
    }
  }
  const size_t count = slot->count, bytes = slot->total_bytes;
  if (count_out) *count_out = count;
  rc = 0;
  if (slot->dtype != (uint32_t)dtype) {
    rc = fail(B200MPI_ERR_ARG, "recv: sender used dtype " + std::to_string(slot->dtype) + ", receiver asked for " + std::to_string(dtype));
  } else if (count > capacity) {
    // Nothing is consumed: the message stays posted so the caller can retry with a buffer of
    // *count_out elements (what gob's slice resize does for the reference, network.go:597).
    slot->state.store(kPosted, std::memory_order_release);
    g->recvtags[src].remove(tag);
    return fail(B200MPI_ERR_TRUNCATE, "recv: message has " + std::to_string(count) + " elements, capacity is " + std::to_string(capacity));
  }
  size_t stage = 0, stage_len = 0;
  cudaStream_t s = nullptr;
  Ctx::P2PBounce* bounce = nullptr;
  const bool whole = slot->chunk_bytes * (slot->nregions ? slot->nregions : 1) >= bytes; // every chunk has its own region, contiguous in the sender's heap
  if (rc == 0 && bytes && memkind == B200MPI_HOST && g->p2p_fast && whole && bytes <= Ctx::kP2PBounceBytes && (bounce = acquire_bounce()) != nullptr) {
    // small message into a host slice: wait until all of it is posted, ONE kernel pulls it over
    // NVLink straight into a mapped pinned buffer (the SMs write host memory), spin, memcpy out
    Spinner sp(g->watchdog_ns);
    while (slot->posted.load(std::memory_order_acquire) < bytes)
      if (!sp.step()) { rc = fail(B200MPI_ERR_TIMEOUT, "recv: sender stalled"); break; }
    if (rc == 0) {
      s = borrow_stream();
      rc = copy_and_wait(bounce, bounce->dev, (const char*)g->heap.base[src] + slot->region_off[0], bytes, s);
      if (rc == 0) memcpy(buf, bounce->host, bytes);
      slot->done.store(bytes, std::memory_order_release);
    }
    release_bounce(bounce);
  } else if (rc == 0 && bytes) {
    s = borrow_stream();
    const size_t chunk = slot->chunk_bytes;
    const size_t nreg = slot->nregions ? slot->nregions : 1;
    if (memkind == B200MPI_HOST) {
      stage_len = std::min(bytes, std::max<size_t>(chunk, 1));
      if (g->heap.alloc(stage_len, stage)) { rc = fail(B200MPI_ERR_NOMEM, "recv: symmetric heap exhausted while staging"); stage_len = 0; }
    }
    const char* peer = (const char*)g->heap.base[src];
    char* mine = (char*)g->heap.base[me];
    size_t consumed = 0;
    while (rc == 0 && consumed < bytes) {
      size_t posted;
      Spinner sp(g->watchdog_ns);
      while ((posted = slot->posted.load(std::memory_order_acquire)) <= consumed)
        if (!sp.step()) { rc = fail(B200MPI_ERR_TIMEOUT, "recv: sender stalled"); break; }
      if (rc) break;
      const size_t k = consumed / chunk;
      const size_t in_chunk = consumed - k * chunk;
      size_t len = std::min(posted, (k + 1) * chunk) - consumed;
      const char* from = peer + slot->region_off[k % nreg] + in_chunk;
      cudaError_t e = cudaSuccess;
      if (memkind == B200MPI_DEVICE) {
        rc = launch_copy((char*)buf + consumed, from, len, s);
      } else {
        len = std::min(len, stage_len);
        rc = launch_copy(mine + stage, from, len, s);
        if (rc == 0) e = cudaMemcpyAsync((char*)buf + consumed, mine + stage, len, cudaMemcpyDeviceToHost, s);
      }
      if (rc == 0 && e == cudaSuccess) e = cudaStreamSynchronize(s);
      if (rc == 0 && e != cudaSuccess) rc = fail(B200MPI_ERR_CUDA, std::string("recv pull: ") + cudaGetErrorString(e));
      if (rc) break;
      consumed += len;
      slot->done.store(consumed, std::memory_order_release);
    }
  }
  if (s) return_stream(s);
  if (stage_len) g->heap.free_off(stage);
  slot->result = rc;
  slot->state.store(kDone, std::memory_order_release); // the ack (network.go:617-621)
  g->recvtags[src].remove(tag);
  return rc;
}

} // namespace b200

// =============================================================================================
// C ABI
// =============================================================================================
using namespace b200;

extern "C" {

int b200mpi_version(void) { return B200MPI_VERSION; }
const char* b200mpi_last_error(void) { return t_err.c_str(); }
int b200mpi_rank(void) { return (g && g->initialised) ? g->ctrl.rank : -1; }
int b200mpi_size(void) { return (g && g->initialised) ? g->ctrl.n : 0; }
int b200mpi_device(void) { return (g && g->initialised && !g->control_only) ? g->dev : -1; }
int64_t b200mpi_launch_count(void) { return g ? g->launches.load() : 0; }

// Size thresholds of the AUTO policies, from the 2/4/8-GPU sweeps (profiles/r02/SUMMARY.md).
// Functions of n (and of whether the switch path exists) only, so every rank agrees.
static void apply_defaults() {
  const int n = g->ctrl.n;
  const bool nvls = g->heap.mc_base != 0;
  g->ll_max_bytes = 0;
  // device time, LL vs the best barrier kernel (profiles/r02): 2 GPUs 256 KiB 12.3 us vs 14.5; 4 GPUs 256 KiB 9.5 vs 13.2;
  // 8 GPUs: see SUMMARY.md section 6
  // 8 GPUs (two-phase kernel): 1 KiB 6.9 us vs 12.7, 32 KiB 8.6 vs 12.7, 256 KiB 12.9 vs 15.3 -> LL over its whole range
  if (!g->shared_device) g->ll_max_bytes = std::min<size_t>(256u << 10, ll_cells(n) * 8);
  // Measured on 8 B200s (profiles/r02/sweep_n8_bcag_v1.jsonl, sweep_n8_hybrid_v1.jsonl): multicast
  // delivery tops out near 480-510 GB/s of ingress per GPU, below what plain P2P stores reach
  // (620-650), so above a few MiB Bcast and Allgather stay on the P2P kernels; mixing P2P traffic
  // into the NVLS allreduce only slows it (817 -> 793 GB/s at 10 %), so the hybrid stays off.
  g->bcast_nvls_min = SIZE_MAX;
  g->allgather_nvls_min = SIZE_MAX;
  g->hybrid_p2p_permille = 0;
  // ranks time-slicing one GPU (functional-test mode): every extra kernel costs a time slice, keep the
  // host-slice Send/Receive on the copy-engine path there
  g->p2p_fast = g->shared_device ? 0 : 1;
  if (const char* w = getenv("B200MPI_P2P_FAST")) g->p2p_fast = atoi(w) != 0;
  (void)nvls;
  if (const char* w = getenv("B200MPI_LL_MAX")) g->ll_max_bytes = std::min<size_t>(strtoull(w, nullptr, 0), ll_cells(g->ctrl.n) * 8);
  if (const char* w = getenv("B200MPI_HYBRID_PERMILLE")) g->hybrid_p2p_permille = std::max(0, std::min(900, atoi(w)));
}

int b200mpi_init(const char* addr, const char* alladdr_csv, const char* password, int64_t timeout_ns, int gpu) {
  std::lock_guard<std::mutex> lock(g_mu);
  if (g && g->initialised) return fail(B200MPI_ERR_ARG, "mpi: Init called twice");
  delete g;
  g = new Ctx();
  std::string err;
  int rc = g->ctrl.init(addr, alladdr_csv, password, timeout_ns, err);
  if (rc) { delete g; g = nullptr; return fail(rc, err); }
  if (const char* w = getenv("B200MPI_WATCHDOG_S")) g->watchdog_ns = (int64_t)(atof(w) * 1e9);
  if (const char* w = getenv("B200MPI_STAGE_CHUNK")) g->stage_chunk = std::max<size_t>(strtoull(w, nullptr, 0), 4096);
  if (const char* w = getenv("B200MPI_ONESHOT_MAX")) g->oneshot_max_bytes = strtoull(w, nullptr, 0);
  auto bail = [&](int code, const std::string& m) {
    // tell nobody: peers notice through their own control-plane errors / timeouts
    if (g->box) munmap(g->box, sizeof(Mailbox));
    if (g->status_host) cudaFreeHost(g->status_host);
    for (cudaStream_t st : {g->own_stream, g->h2d_stream, g->d2h_stream})
      if (st) cudaStreamDestroy(st);
    if (g->ev0) cudaEventDestroy(g->ev0);
    if (g->ev1) cudaEventDestroy(g->ev1);
    if (g->drv.MemUnmap) g->heap.destroy(g->drv); // no-op when the heap was never created
    g->ctrl.shutdown();
    delete g;
    g = nullptr;
    return fail(code, m);
  };
  if (gpu == -2) {
    g->control_only = true;
    g->initialised = true;
    return 0;
  }
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0) {
    (void)cudaGetLastError();
    return bail(B200MPI_ERR_NO_DEVICE, std::string("no usable CUDA device: ") + (ce == cudaSuccess ? "device count is 0" : cudaGetErrorString(ce)) + " (libb200mpi has no CPU fallback)");
  }
  g->dev = gpu >= 0 ? gpu : g->ctrl.rank % ndev;
  if (g->dev >= ndev) return bail(B200MPI_ERR_ARG, "gpu ordinal " + std::to_string(g->dev) + " out of range (" + std::to_string(ndev) + " devices)");
  if ((ce = cudaSetDevice(g->dev)) != cudaSuccess || (ce = cudaFree(0)) != cudaSuccess)
    return bail(B200MPI_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(ce));
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, g->dev);
  g->sm_count = prop.multiProcessorCount;
  if (prop.major < 10) return bail(B200MPI_ERR_NO_DEVICE, std::string("device '") + prop.name + "' is not sm_100-class; this library ships sm_100a code only");
  if (!g->drv.load(err)) return bail(B200MPI_ERR_CUDA, err);
  if (!(getenv("B200MPI_NUMA") && atoi(getenv("B200MPI_NUMA")) == 0)) g->gpu_numa_node = numa_node_of_gpu(g->dev);
  if (const char* w = getenv("B200MPI_SCRUB_EVERY")) {
    uint32_t v = (uint32_t)std::max(2, atoi(w)), p2 = 2;
    while (p2 * 2 <= v) p2 *= 2;
    g->scrub_every = p2;
  }
  if (const char* w = getenv("B200MPI_HOST_THREADS")) g->host_threads = std::max(0, atoi(w));
  if (const char* w = getenv("B200MPI_HOST_REGISTER")) g->host_register = atoi(w) != 0;

  // who sits where: ranks sharing a device (functional-test mode) rule out NVLS
  struct Hello { unsigned char uuid[16]; int32_t dev; int32_t mc; } mine = {}, all[B200MPI_MAX_RANKS];
  memcpy(mine.uuid, &prop.uuid, 16);
  mine.dev = g->dev;
  int mc = 0;
  g->drv.DeviceGetAttribute(&mc, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, g->dev);
  mine.mc = mc;
  rc = g->ctrl.allgather(&mine, sizeof mine, all, err);
  if (rc) return bail(rc, err);
  bool want_nvls = g->ctrl.n > 1;
  if (const char* w = getenv("B200MPI_NVLS")) want_nvls = want_nvls && atoi(w) != 0;
  for (int a = 0; a < g->ctrl.n; ++a) {
    if (!all[a].mc) want_nvls = false;
    for (int b = a + 1; b < g->ctrl.n; ++b)
      if (memcmp(all[a].uuid, all[b].uuid, 16) == 0) g->shared_device = true;
  }
  if (g->shared_device) want_nvls = false;
  if (g->shared_device && !getenv("B200MPI_WATCHDOG_S")) g->watchdog_ns = 600ll * 1000000000ll; // time-sliced flags are slow

  size_t heap_bytes = 2ull << 30;
  if (const char* w = getenv("B200MPI_HEAP_BYTES")) heap_bytes = strtoull(w, nullptr, 0);
  if (heap_bytes < (64u << 20)) heap_bytes = 64u << 20;
  rc = g->heap.create(g->drv, g->ctrl, g->dev, heap_bytes, want_nvls, err);
  if (rc) return bail(rc, err);

  // mailbox shared memory
  {
    int mfd = -1, got = -1;
    if (g->ctrl.rank == 0) {
      mfd = memfd_create("b200mpi-mailbox", 0);
      if (mfd < 0 || ftruncate(mfd, sizeof(Mailbox)) != 0) return bail(B200MPI_ERR_BOOTSTRAP, std::string("mailbox memfd: ") + strerror(errno));
    }
    if (g->ctrl.n > 1) {
      rc = g->ctrl.bcast_fd(0, mfd, got, err);
      if (rc) return bail(rc, err);
    } else got = dup(mfd);
    void* m = mmap(nullptr, sizeof(Mailbox), PROT_READ | PROT_WRITE, MAP_SHARED, got, 0);
    if (mfd >= 0) close(mfd);
    close(got);
    if (m == MAP_FAILED) return bail(B200MPI_ERR_BOOTSTRAP, std::string("mailbox mmap: ") + strerror(errno));
    g->box = (Mailbox*)m; // zero-filled by ftruncate == every slot kFree
  }
  if (cudaHostAlloc((void**)&g->status_host, 64, cudaHostAllocMapped) != cudaSuccess ||
      cudaHostGetDevicePointer((void**)&g->status_dev, g->status_host, 0) != cudaSuccess)
    return bail(B200MPI_ERR_CUDA, "status word allocation failed");
  *g->status_host = 0;
  cudaStreamCreateWithFlags(&g->own_stream, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&g->h2d_stream, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&g->d2h_stream, cudaStreamNonBlocking);
  g->stream = g->own_stream;
  cudaEventCreate(&g->ev0);
  cudaEventCreate(&g->ev1);
  for (int r = 0; r < g->ctrl.n; ++r) g->comm.base[r] = (char*)g->heap.base[r];
  g->comm.mc = (char*)g->heap.mc_base;
  g->comm.status = g->status_dev;
  g->comm.timeout_ns = (unsigned long long)g->watchdog_ns;
  g->comm.rank = g->ctrl.rank;
  g->comm.n = g->ctrl.n;
  apply_defaults();
  rc = g->ctrl.barrier(err);
  if (rc) return bail(rc, err);
  g->initialised = true;
  return 0;
}

int b200mpi_finalize(void) {
  std::lock_guard<std::mutex> lock(g_mu);
  if (!g || !g->initialised) return fail(B200MPI_ERR_NOT_INIT, "mpi: Finalize without Init");
  std::string err;
  if (!g->control_only) {
    cudaSetDevice(g->dev);
    cudaDeviceSynchronize();
  }
  g->ctrl.barrier(err); // nobody unmaps while a peer may still read
  if (!g->control_only) {
    for (cudaStream_t s : g->stream_pool) cudaStreamDestroy(s);
    if (g->own_stream) cudaStreamDestroy(g->own_stream);
    if (g->h2d_stream) cudaStreamDestroy(g->h2d_stream);
    if (g->d2h_stream) cudaStreamDestroy(g->d2h_stream);
    for (cudaEvent_t e : g->pipe_events) cudaEventDestroy(e);
    if (g->ev0) cudaEventDestroy(g->ev0);
    if (g->ev1) cudaEventDestroy(g->ev1);
    if (g->status_host) cudaFreeHost(g->status_host);
    if (g->ll_host) cudaFreeHost(g->ll_host);
    for (auto& b : g->p2p_bounce) if (b.host) cudaFreeHost(b.host);
    if (g->pool.running()) g->pool.shutdown();
    if (g->bounce) cudaFreeHost(g->bounce);
    for (auto& e : g->registered) cudaHostUnregister((void*)e.first);
    if (g->box) munmap(g->box, sizeof(Mailbox));
    {
      std::lock_guard<std::mutex> l(g_pending_mu);
      g_pending.clear();
    }
    g->heap.destroy(g->drv);
  }
  g->ctrl.shutdown();
  delete g;
  g = nullptr;
  return 0;
}

int b200mpi_alloc(size_t bytes, void** dptr) {
  int rc = need_data_plane();
  if (rc) return rc;
  if (!dptr) return fail(B200MPI_ERR_ARG, "alloc: NULL result pointer");
  size_t off;
  if (g->heap.alloc(bytes, off)) return fail(B200MPI_ERR_NOMEM, "symmetric heap exhausted (" + std::to_string(g->heap.used()) + " of " + std::to_string(g->heap.size) + " bytes in use; raise B200MPI_HEAP_BYTES)");
  *dptr = (char*)g->heap.base[g->ctrl.rank] + off;
  return 0;
}

int b200mpi_free(void* dptr) {
  int rc = need_data_plane();
  if (rc) return rc;
  if (!dptr) return 0;
  size_t off = (char*)dptr - (char*)g->heap.base[g->ctrl.rank];
  if (g->heap.free_off(off)) return fail(B200MPI_ERR_ARG, "free: pointer was not returned by b200mpi_alloc");
  return 0;
}

int b200mpi_host_alloc(size_t bytes, void** hptr) {
  int rc = need_data_plane();
  if (rc) return rc;
  if (!hptr) return fail(B200MPI_ERR_ARG, "host_alloc: NULL result pointer");
  return numa_host_alloc(bytes, hptr);
}
int b200mpi_host_free(void* hptr) {
  int rc = need_data_plane();
  if (rc) return rc;
  CUDA_OK(cudaFreeHost(hptr));
  return 0;
}
int b200mpi_memcpy(void* dst, const void* src, size_t bytes, int kind) {
  int rc = need_data_plane();
  if (rc) return rc;
  cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice : kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  CUDA_OK(cudaMemcpyAsync(dst, src, bytes, k, g->stream));
  CUDA_OK(cudaStreamSynchronize(g->stream));
  return 0;
}
// Measured bound of the host-slice paths: plain pinned copies between a NUMA-local pinned buffer and
// the device heap, no collective.  bidir = H2D and D2H of `bytes` each running at the same time
// (what a pipelined host-slice Allreduce needs), reported per direction.
int b200mpi_pcie_probe(size_t bytes, int iters, double* h2d_gbs, double* d2h_gbs, double* bidir_gbs) {
  int rc = need_data_plane();
  if (rc) return rc;
  if (bytes == 0 || iters < 1) return fail(B200MPI_ERR_ARG, "pcie_probe: bytes and iters must be positive");
  void *ha = nullptr, *hb = nullptr;
  if ((rc = numa_host_alloc(bytes, &ha)) || (rc = numa_host_alloc(bytes, &hb))) { if (ha) cudaFreeHost(ha); return rc; }
  memset(ha, 1, bytes);
  size_t oa = 0, ob = 0;
  if (g->heap.alloc(bytes, oa) || g->heap.alloc(bytes, ob)) {
    cudaFreeHost(ha); cudaFreeHost(hb);
    return fail(B200MPI_ERR_NOMEM, "pcie_probe: symmetric heap exhausted");
  }
  char* da = (char*)g->heap.base[g->ctrl.rank] + oa;
  char* db = (char*)g->heap.base[g->ctrl.rank] + ob;
  auto timed = [&](bool up, bool down) {
    cudaStreamSynchronize(g->h2d_stream);
    cudaStreamSynchronize(g->d2h_stream);
    const auto t0 = Clock::now();
    for (int i = 0; i < iters; ++i) {
      if (up) cudaMemcpyAsync(da, ha, bytes, cudaMemcpyHostToDevice, g->h2d_stream);
      if (down) cudaMemcpyAsync(hb, db, bytes, cudaMemcpyDeviceToHost, g->d2h_stream);
    }
    cudaStreamSynchronize(g->h2d_stream);
    cudaStreamSynchronize(g->d2h_stream);
    const double sec = std::chrono::duration<double>(Clock::now() - t0).count();
    return (double)bytes * iters / sec / 1e9;
  };
  timed(true, true); // warm-up
  const double up = timed(true, false), down = timed(false, true), both = timed(true, true);
  if (h2d_gbs) *h2d_gbs = up;
  if (d2h_gbs) *d2h_gbs = down;
  if (bidir_gbs) *bidir_gbs = both;
  g->heap.free_off(oa);
  g->heap.free_off(ob);
  cudaFreeHost(ha);
  cudaFreeHost(hb);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(B200MPI_ERR_CUDA, std::string("pcie_probe: ") + cudaGetErrorString(e));
  return 0;
}
// Raw NVLink rates seen by this library's copy kernel between neighbouring ranks (the roofline the
// P2P collectives are held against).  Collective call.  mode 0: every rank pulls `bytes` from rank+1;
// 1: every rank pushes to rank+1; 2: only rank 0 pulls from rank 1; 3: only rank 0 pushes to rank 1;
// 4: every rank pulls from and pushes to rank+1 at the same time (two streams).
// *ms = device time of one iteration on this rank (0 for idle ranks).
int b200mpi_link_probe(size_t bytes, int mode, int iters, float* ms) {
  int rc = need_data_plane();
  if (rc) return rc;
  const int n = g->ctrl.n, me = g->ctrl.rank;
  if (n < 2 || bytes == 0 || iters < 1 || mode < 0 || mode > 4 || !ms) return fail(B200MPI_ERR_ARG, "link_probe: needs >= 2 ranks, bytes, iters >= 1, mode 0..4");
  size_t a = 0, b = 0;
  if (g->heap.alloc(bytes, a) || g->heap.alloc(bytes, b)) return fail(B200MPI_ERR_NOMEM, "link_probe: symmetric heap exhausted");
  struct Offs { uint64_t a, b; } mine = {a, b}, all[B200MPI_MAX_RANKS];
  std::string err;
  if ((rc = g->ctrl.allgather(&mine, sizeof mine, all, err))) return fail(rc, err);
  const int peer = (me + 1) % n;
  char* my_a = (char*)g->heap.base[me] + a;
  char* my_b = (char*)g->heap.base[me] + b;
  char* peer_a = (char*)g->heap.base[peer] + all[peer].a;
  char* peer_b = (char*)g->heap.base[peer] + all[peer].b;
  const bool active = mode == 0 || mode == 1 || mode == 4 || me == 0;
  cudaStream_t s2 = g->h2d_stream;
  CUDA_OK(cudaStreamSynchronize(g->stream));
  if ((rc = g->ctrl.barrier(err))) return fail(rc, err);
  *ms = 0.f;
  if (active) {
    for (int pass = 0; pass < 2; ++pass) { // pass 0 warms up
      const int k = pass ? iters : 2;
      CUDA_OK(cudaEventRecord(g->ev0, g->stream));
      if (mode == 4) CUDA_OK(cudaStreamWaitEvent(s2, g->ev0, 0));
      for (int i = 0; i < k && rc == 0; ++i) {
        if (mode == 0 || mode == 2 || mode == 4) rc = launch_copy(my_a, peer_a, bytes, g->stream);          // pull: loads cross the link
        if (rc == 0 && (mode == 1 || mode == 3)) rc = launch_copy(peer_b, my_b, bytes, g->stream);           // push: stores cross the link
        if (rc == 0 && mode == 4) rc = launch_copy(peer_b, my_b, bytes, s2);
      }
      if (rc) break;
      if (mode == 4) {
        cudaEvent_t e = g->pipe_events.empty() ? nullptr : g->pipe_events[0];
        if (!e) { CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); g->pipe_events.push_back(e); }
        CUDA_OK(cudaEventRecord(e, s2));
        CUDA_OK(cudaStreamWaitEvent(g->stream, e, 0));
      }
      CUDA_OK(cudaEventRecord(g->ev1, g->stream));
      CUDA_OK(cudaEventSynchronize(g->ev1));
      float t = 0;
      CUDA_OK(cudaEventElapsedTime(&t, g->ev0, g->ev1));
      *ms = t / k;
    }
  }
  std::string err2;
  g->ctrl.barrier(err2); // nobody frees while a neighbour still copies
  g->heap.free_off(a);
  g->heap.free_off(b);
  return rc;
}
int b200mpi_numa_node(void) { return (g && g->initialised && !g->control_only) ? g->gpu_numa_node : -1; }
int b200mpi_heap_info(size_t* total, size_t* used, int* nvls) {
  int rc = need_data_plane();
  if (rc) return rc;
  if (total) *total = g->heap.size;
  if (used) *used = g->heap.used();
  if (nvls) *nvls = g->heap.mc_base ? 1 : 0;
  return 0;
}

int b200mpi_send(const void* buf, size_t count, int dtype, int dest, int tag, int memkind) { return do_send(buf, count, dtype, dest, tag, memkind, true); }
int b200mpi_isend(const void* buf, size_t count, int dtype, int dest, int tag, int memkind) { return do_send(buf, count, dtype, dest, tag, memkind, false); }
int b200mpi_wait(int dest, int tag) { return do_wait(dest, tag); }
int b200mpi_recv(void* buf, size_t capacity, size_t* count_out, int dtype, int src, int tag, int memkind) { return do_recv(buf, capacity, count_out, dtype, src, tag, memkind); }

int b200mpi_bcast(void* buf, size_t count, int dtype, int root, int memkind) { return do_bcast(buf, count, dtype, root, memkind, false); }
int b200mpi_allreduce(const void* send, void* recv, size_t count, int dtype, int op, int memkind) { return do_allreduce(send, recv, count, dtype, op, memkind, false); }
int b200mpi_allgather(const void* send, void* recv, size_t count, int dtype, int memkind) { return do_allgather(send, recv, count, dtype, memkind, false); }
int b200mpi_reduce_scatter(const void* send, void* recv, size_t count_per_rank, int dtype, int op, int memkind) { return do_reduce_scatter(send, recv, count_per_rank, dtype, op, memkind, false); }
int b200mpi_reduce(const void* send, void* recv, size_t count, int dtype, int op, int root, int memkind) { return do_reduce(send, recv, count, dtype, op, root, memkind, false); }
int b200mpi_alltoall(const void* send, void* recv, size_t count_per_rank, int dtype, int memkind) { return do_alltoall(send, recv, count_per_rank, dtype, memkind, false); }
int b200mpi_reduce_scatter_async(const void* send, void* recv, size_t count_per_rank, int dtype, int op) { return do_reduce_scatter(send, recv, count_per_rank, dtype, op, B200MPI_DEVICE, true); }
int b200mpi_bcast_async(void* buf, size_t count, int dtype, int root) { return do_bcast(buf, count, dtype, root, B200MPI_DEVICE, true); }
int b200mpi_allreduce_async(const void* send, void* recv, size_t count, int dtype, int op) { return do_allreduce(send, recv, count, dtype, op, B200MPI_DEVICE, true); }
int b200mpi_allgather_async(const void* send, void* recv, size_t count, int dtype) { return do_allgather(send, recv, count, dtype, B200MPI_DEVICE, true); }

int b200mpi_stream_sync(void) {
  int rc = need_data_plane();
  if (rc) return rc;
  CUDA_OK(cudaStreamSynchronize(g->stream));
  return check_status();
}

int b200mpi_barrier(void) {
  if (!g || !g->initialised) return fail(B200MPI_ERR_NOT_INIT, "mpi: Init has not been called");
  if (g->control_only) {
    std::string err;
    int rc = g->ctrl.barrier(err);
    return rc ? fail(rc, err) : 0;
  }
  if (g->ctrl.n == 1) return finish(false);
  g->cur_sig = make_sig(3, 0, 0, 0, 0);
  Comm c = next_comm();
  sig_grid(c, 1);
  barrier_kernel<<<1, 32, 0, g->stream>>>(c);
  int rc = launch_check("barrier_kernel");
  if (rc) return rc;
  return finish(false);
}

int b200mpi_set_algo(int coll, int algo) {
  if (!g || !g->initialised) return fail(B200MPI_ERR_NOT_INIT, "mpi: Init has not been called");
  if (coll < 0 || coll > B200MPI_COLL_REDUCE_SCATTER || algo < 0 || algo > B200MPI_ALGO_HYBRID) return fail(B200MPI_ERR_ARG, "set_algo: bad collective or algorithm id");
  g->algo[coll] = algo;
  return 0;
}
int b200mpi_get_algo(int coll, size_t count, int dtype) {
  int rc = need_data_plane();
  if (rc) return rc;
  const size_t bytes = count * esize(dtype);
  if (coll == B200MPI_COLL_ALLREDUCE) return pick_allreduce(bytes, dtype, B200MPI_SUM);
  if (coll == B200MPI_COLL_BCAST) return pick_bcast(bytes);
  if (coll == B200MPI_COLL_ALLGATHER) return pick_allgather(bytes);
  if (coll == B200MPI_COLL_REDUCE_SCATTER) return pick_reduce_scatter(bytes, dtype, B200MPI_SUM);
  return fail(B200MPI_ERR_ARG, "get_algo: bad collective id");
}
int b200mpi_set_param(const char* name, int64_t value) {
  if (!g || !g->initialised) return fail(B200MPI_ERR_NOT_INIT, "mpi: Init has not been called");
  std::string k = name ? name : "";
  if (k == "twoshot_unroll") g->twoshot_unroll = value ? 1 : 0;
  else if (k == "nvls_unroll") g->nvls_unroll = (int)value;
  else if (k == "nvls_min_ranks") g->nvls_min_ranks = (int)value;
  else if (k == "copy_variant") g->copy_variant = (int)value;
  else if (k == "ll_max_bytes") g->ll_max_bytes = (size_t)std::min<int64_t>(std::max<int64_t>(value, 0), (int64_t)(ll_cells(g->ctrl.n) * 8));
  else if (k == "nvls_max_blocks") g->nvls_max_blocks = (int)std::max<int64_t>(1, value);
  else if (k == "oneshot_max_bytes") g->oneshot_max_bytes = (size_t)value;
  else if (k == "pipe_min_bytes") g->pipe_min_bytes = (size_t)value;
  else if (k == "pipe_chunk_bytes") g->pipe_chunk_bytes = (size_t)std::max<int64_t>(value, 65536);
  else if (k == "own_block_bytes") g->own_block_bytes = (size_t)std::max<int64_t>(value, 4096);
  else if (k == "stage_chunk") g->stage_chunk = (size_t)std::max<int64_t>(value, 4096);
  else if (k == "watchdog_ms") { g->watchdog_ns = std::max<int64_t>(value, 1) * 1000000ll; g->comm.timeout_ns = (unsigned long long)g->watchdog_ns; }
  else if (k == "hybrid_p2p_permille") g->hybrid_p2p_permille = (int)std::min<int64_t>(std::max<int64_t>(value, 0), 900);
  else if (k == "hybrid_p2p_blocks") g->hybrid_p2p_blocks = (int)std::max<int64_t>(value, 0);
  else if (k == "hybrid_min_bytes") g->hybrid_min_bytes = (size_t)std::max<int64_t>(value, 0);
  else if (k == "bcast_nvls2") g->bcast_nvls2 = value ? 1 : 0;
  else if (k == "bcast_nvls_min") g->bcast_nvls_min = (size_t)std::max<int64_t>(value, 0);
  else if (k == "allgather_nvls_min") g->allgather_nvls_min = (size_t)std::max<int64_t>(value, 0);
  else if (k == "host_register") {
    g->host_register = value ? 1 : 0;
    if (!value) { // leaving the mode drops every cached registration: the caller may free those buffers now
      cudaStreamSynchronize(g->h2d_stream);
      cudaStreamSynchronize(g->d2h_stream);
      for (auto& e : g->registered) cudaHostUnregister((void*)e.first);
      g->registered.clear();
      (void)cudaGetLastError();
    }
  }
  else if (k == "p2p_fast") g->p2p_fast = value ? 1 : 0;
  else if (k == "bounce_chunk_bytes") g->bounce_chunk_bytes = (size_t)std::max<int64_t>(value, 65536);
  else if (k == "host_threads") { if (g->pool.running()) return fail(B200MPI_ERR_ARG, "set_param: host_threads must be set before the first pageable host-slice call"); g->host_threads = (int)std::max<int64_t>(value, 0); }
  else return fail(B200MPI_ERR_ARG, "set_param: unknown parameter '" + k + "'");
  return 0;
}
int b200mpi_get_param(const char* name, int64_t* value) {
  if (!g || !g->initialised) return fail(B200MPI_ERR_NOT_INIT, "mpi: Init has not been called");
  if (!value) return fail(B200MPI_ERR_ARG, "get_param: NULL result pointer");
  const std::string k = name ? name : "";
  if (k == "twoshot_unroll") *value = g->twoshot_unroll;
  else if (k == "nvls_unroll") *value = g->nvls_unroll;
  else if (k == "nvls_min_ranks") *value = g->nvls_min_ranks;
  else if (k == "nvls_max_blocks") *value = g->nvls_max_blocks;
  else if (k == "copy_variant") *value = g->copy_variant;
  else if (k == "ll_max_bytes") *value = (int64_t)g->ll_max_bytes;
  else if (k == "oneshot_max_bytes") *value = (int64_t)g->oneshot_max_bytes;
  else if (k == "pipe_min_bytes") *value = (int64_t)g->pipe_min_bytes;
  else if (k == "pipe_chunk_bytes") *value = (int64_t)g->pipe_chunk_bytes;
  else if (k == "bounce_chunk_bytes") *value = (int64_t)g->bounce_chunk_bytes;
  else if (k == "host_threads") *value = g->host_threads;
  else if (k == "host_register") *value = g->host_register;
  else if (k == "own_block_bytes") *value = (int64_t)g->own_block_bytes;
  else if (k == "stage_chunk") *value = (int64_t)g->stage_chunk;
  else if (k == "watchdog_ms") *value = g->watchdog_ns / 1000000ll;
  else if (k == "hybrid_p2p_permille") *value = g->hybrid_p2p_permille;
  else if (k == "hybrid_p2p_blocks") *value = g->hybrid_p2p_blocks;
  else if (k == "hybrid_min_bytes") *value = (int64_t)g->hybrid_min_bytes;
  else if (k == "bcast_nvls2") *value = g->bcast_nvls2;
  else if (k == "bcast_nvls_min") *value = (int64_t)g->bcast_nvls_min;
  else if (k == "allgather_nvls_min") *value = (int64_t)g->allgather_nvls_min;
  else if (k == "sm_count") *value = g->sm_count;
  else if (k == "shared_device") *value = g->shared_device ? 1 : 0;
  else return fail(B200MPI_ERR_ARG, "get_param: unknown parameter '" + k + "'");
  return 0;
}
int b200mpi_set_max_blocks(int blocks) {
  if (!g || !g->initialised) return fail(B200MPI_ERR_NOT_INIT, "mpi: Init has not been called");
  if (blocks < 0 || blocks > kMaxBlocks) return fail(B200MPI_ERR_ARG, "set_max_blocks: out of range");
  g->max_blocks = blocks; // block_cap() clamps to the SM count: the cross-rank barriers need every CTA resident
  return 0;
}
int b200mpi_get_stream(void** stream) {
  int rc = need_data_plane();
  if (rc) return rc;
  *stream = (void*)g->stream;
  return 0;
}
int b200mpi_set_stream(void* stream) {
  int rc = need_data_plane();
  if (rc) return rc;
  CUDA_OK(cudaStreamSynchronize(g->stream));
  g->stream = stream ? (cudaStream_t)stream : g->own_stream;
  return 0;
}
int b200mpi_timer_start(void) {
  int rc = need_data_plane();
  if (rc) return rc;
  CUDA_OK(cudaEventRecord(g->ev0, g->stream));
  return 0;
}
int b200mpi_timer_stop(float* ms) {
  int rc = need_data_plane();
  if (rc) return rc;
  CUDA_OK(cudaEventRecord(g->ev1, g->stream));
  CUDA_OK(cudaEventSynchronize(g->ev1));
  CUDA_OK(cudaEventElapsedTime(ms, g->ev0, g->ev1));
  return check_status();
}

} // extern "C"
