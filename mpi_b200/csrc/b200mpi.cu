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
#include <errno.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "../../include/b200mpi.h"
#include "ctrl.h"
#include "heap.h"
#include "kernels.cuh"

namespace b200 {

// ---------------------------------------------------------------------------------------------
// mailbox: host shared memory (memfd) used for Send/Receive rendezvous, the role of the
// reference's tagManager + per-connection channels (network.go:449-497) and of `local`
// (network.go:388-446).  Control only: payload bytes never pass through it.
// ---------------------------------------------------------------------------------------------
constexpr int kSlotsPerPair = 16;
enum SlotState : uint32_t { kFree = 0, kClaimed = 1, kPosted = 2, kMatched = 3, kDone = 4 };

struct alignas(128) MsgSlot {
  std::atomic<uint32_t> state;
  int32_t tag;
  uint32_t dtype;
  int32_t result;          // receiver's verdict, informational
  uint64_t count;          // elements
  uint64_t total_bytes;
  uint64_t chunk_bytes;    // size of one posted region (== total_bytes for a direct post)
  uint64_t region_off[2];  // offsets in the SENDER's heap
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
  int algo[3] = {0, 0, 0};
  int max_blocks = 0;
  int64_t watchdog_ns = 120ll * 1000000000ll;
  size_t stage_chunk = 32u << 20;
  // staging for collectives on non-heap buffers: [0] send side, [1] recv side
  size_t stage_off[2] = {0, 0}, stage_len[2] = {0, 0};
  std::atomic<int64_t> launches{0};
  size_t oneshot_max_bytes = 256u << 10;
  int twoshot_unroll = 1; // 0: 1/2/4 vectors per thread for n = 8/4/2, 1: 2/4/8
  int nvls_unroll = 2;      // 8 GPUs, 256 MiB: unroll 2 x 64 CTAs 810 GB/s, 4 x 148 CTAs 780 (profiles/r01/sweep_n8_nvls_blocks_unroll_v2.jsonl)
  int nvls_max_blocks = 64; // fewer requests in flight suit the switch reduction better
  size_t ll_max_bytes = 0; // > 0 enables the experimental LL allreduce for messages up to this size (<= 32 KiB)
  uint32_t ll_seq = 0;
  int copy_variant = 5; // 16 vectors in flight per thread, 256 threads: best of 8 launch shapes at 256 MiB and 1 GiB
  int nvls_min_ranks = 4; // below this the fused two-shot moves fewer bytes per link than NVLS
  size_t own_block_bytes = 1u << 20; // interleave granularity of slice ownership (Owner in kernels.cuh)
};

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

static int grid_for(size_t units_per_rank, int unroll) {
  size_t per_block = (size_t)kThreads * unroll;
  size_t want = (units_per_rank + per_block - 1) / per_block;
  int cap = g->max_blocks > 0 ? g->max_blocks : g->sm_count;
  if (cap > kMaxBlocks) cap = kMaxBlocks;
  if (want < 1) want = 1;
  return (int)(want > (size_t)cap ? cap : want);
}

// mids: number of flag values reserved between the start and the end barrier (in-place one-shot
// rounds).  Every rank computes it from (count, grid) only, so epochs stay in step.
static Comm next_comm(uint32_t mids = 0) {
  Comm c = g->comm;
  c.epoch = g->epoch;
  c.sig = g->cur_sig;
  c.end_epoch = g->epoch + 1 + mids;
  g->epoch += 2 + mids;
  return c;
}

static int check_status() {
  if (g->status_host && *(volatile uint32_t*)g->status_host) {
    const uint32_t st = *(volatile uint32_t*)g->status_host;
    *(volatile uint32_t*)g->status_host = 0;
    if (st == 2u) return fail(B200MPI_ERR_PEER, "mismatched collective: ranks disagree on the call (collective, count, dtype, op, root or algorithm); buffers were left untouched");
    return fail(B200MPI_ERR_TIMEOUT, "device-side watchdog: a peer did not reach the collective in time");
  }
  return 0;
}

// (count, collective, dtype, op/root, algorithm) folded into one word; equal on every rank of a
// well-formed call.  coll: 0 allreduce, 1 bcast, 2 allgather, 3 barrier.
static uint64_t make_sig(int coll, int dtype, int extra, int algo, size_t count) {
  return ((uint64_t)count << 16) ^ ((uint64_t)(coll & 3) << 14) ^ ((uint64_t)(dtype & 3) << 12) ^ ((uint64_t)(extra & 15) << 8) ^ ((uint64_t)(algo & 15) << 4) ^ 0x5u;
}

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
    case 6: copy_bytes_kernel<2, 1024, 2><<<grid(1024, 2, 2), 1024, 0, s>>>(d, c, bytes); break;
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

template <typename T, typename Op>
static int launch_allreduce_body_t(int algo, const Comm& c, uint64_t so, uint64_t ro, size_t count, cudaStream_t s) {
  const int n = c.n;
  constexpr int EPV = 16 / sizeof(T);
  const size_t nvec = (count + EPV - 1) / EPV;
  const size_t per = (nvec + n - 1) / n;
  const uint32_t sh = own_shift(nvec, n, algo == B200MPI_ALGO_TWOSHOT_SMEM ? 8 : 0);
  switch (algo) {
    case B200MPI_ALGO_TWOSHOT: {
      if (g->twoshot_unroll) {
        if (n == 2) { allreduce_twoshot_kernel<T, Op, 2, 8><<<grid_for(per, 8), kThreads, 0, s>>>(c, so, ro, count, sh); }
        else if (n == 4) { allreduce_twoshot_kernel<T, Op, 4, 4><<<grid_for(per, 4), kThreads, 0, s>>>(c, so, ro, count, sh); }
        else if (n == 8) { allreduce_twoshot_kernel<T, Op, 8, 2><<<grid_for(per, 2), kThreads, 0, s>>>(c, so, ro, count, sh); }
        else { allreduce_twoshot_kernel<T, Op, 0, 2><<<grid_for(per, 2), kThreads, 0, s>>>(c, so, ro, count, sh); }
      } else {
        if (n == 2) { allreduce_twoshot_kernel<T, Op, 2, 4><<<grid_for(per, 4), kThreads, 0, s>>>(c, so, ro, count, sh); }
        else if (n == 4) { allreduce_twoshot_kernel<T, Op, 4, 2><<<grid_for(per, 2), kThreads, 0, s>>>(c, so, ro, count, sh); }
        else if (n == 8) { allreduce_twoshot_kernel<T, Op, 8, 1><<<grid_for(per, 1), kThreads, 0, s>>>(c, so, ro, count, sh); }
        else { allreduce_twoshot_kernel<T, Op, 0, 1><<<grid_for(per, 1), kThreads, 0, s>>>(c, so, ro, count, sh); }
      }
      return launch_check("allreduce_twoshot_kernel");
    }
    case B200MPI_ALGO_RING: {
      allreduce_ring_kernel<T, Op><<<grid_for(per, 4), kThreads, 0, s>>>(c, so, ro, count);
      return launch_check("allreduce_ring_kernel");
    }
    case B200MPI_ALGO_TWOSHOT_SMEM: {
      const size_t tiles = (per * 16 + kSmemChunk - 1) / kSmemChunk;
      int cap = g->max_blocks > 0 ? g->max_blocks : g->sm_count;
      const int blocks = (int)std::max<size_t>(1, std::min<size_t>(tiles, (size_t)cap));
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

template <typename T, typename Op>
static int launch_allreduce_t(int algo, uint64_t so, uint64_t ro, size_t count, cudaStream_t s) {
  const int n = g->ctrl.n;
  constexpr int EPV = 16 / sizeof(T);
  const size_t nvec = (count + EPV - 1) / EPV;
  if (algo != B200MPI_ALGO_ONESHOT) {
    const Comm c = next_comm();
    return launch_allreduce_body_t<T, Op>(algo, c, so, ro, count, s);
  }
  // one-shot: reserve one mid-barrier value per round (upper bound: scalar rounds of the
  // unaligned path, which has the most), see allreduce_oneshot_kernel
  const bool shfl = (n == 2 || n == 4 || n == 8) && nvec <= 4096;
  const int blocks = shfl ? grid_for(nvec * n, 1) : grid_for(nvec, 1);
  const size_t per_round = (size_t)blocks * kThreads;
  const size_t rounds = (count + per_round - 1) / per_round + (shfl ? (nvec * n + per_round - 1) / per_round : 0) + 2;
  const Comm c = next_comm((uint32_t)rounds);
  if (shfl) {
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

template <typename T, typename Op>
static int launch_allreduce_nvls_t(uint64_t so, uint64_t ro, size_t count, cudaStream_t s) {
  const Comm c = next_comm();
  constexpr int EPV = 16 / sizeof(T);
  const size_t nvec = (count + EPV - 1) / EPV;
  const size_t per = (nvec + c.n - 1) / c.n;
  const uint32_t sh = own_shift(nvec, c.n, 0);
  const int u = g->nvls_unroll;
  int blocks = grid_for(per, u == 1 || u == 2 || u == 8 ? u : 4);
  if (g->max_blocks == 0 && blocks > g->nvls_max_blocks) blocks = g->nvls_max_blocks;
  switch (u) {
    case 1: allreduce_nvls_kernel<T, Op, 1><<<blocks, kThreads, 0, s>>>(c, so, ro, count, sh); break;
    case 2: allreduce_nvls_kernel<T, Op, 2><<<blocks, kThreads, 0, s>>>(c, so, ro, count, sh); break;
    case 8: allreduce_nvls_kernel<T, Op, 8><<<blocks, kThreads, 0, s>>>(c, so, ro, count, sh); break;
    default: allreduce_nvls_kernel<T, Op, 4><<<blocks, kThreads, 0, s>>>(c, so, ro, count, sh); break;
  }
  return launch_check("allreduce_nvls_kernel");
}

static bool nvls_supports(int dtype, int op) {
  if (op == B200MPI_SUM) return dtype == B200MPI_F32 || dtype == B200MPI_F64 || dtype == B200MPI_I64;
  return dtype == B200MPI_I64; // min/max: integer only in the switch
}

static int launch_allreduce(int algo, int dtype, int op, uint64_t so, uint64_t ro, size_t count, cudaStream_t s) {
  g->cur_sig = make_sig(0, dtype, op, algo, count);
  if (algo == B200MPI_ALGO_NVLS) {
    if (dtype == B200MPI_F32 && op == B200MPI_SUM) return launch_allreduce_nvls_t<float, OpSum>(so, ro, count, s);
    if (dtype == B200MPI_F64 && op == B200MPI_SUM) return launch_allreduce_nvls_t<double, OpSum>(so, ro, count, s);
    if (dtype == B200MPI_I64 && op == B200MPI_SUM) return launch_allreduce_nvls_t<long long, OpSum>(so, ro, count, s);
    if (dtype == B200MPI_I64 && op == B200MPI_MAX) return launch_allreduce_nvls_t<long long, OpMax>(so, ro, count, s);
    if (dtype == B200MPI_I64 && op == B200MPI_MIN) return launch_allreduce_nvls_t<long long, OpMin>(so, ro, count, s);
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

// AUTO for Allreduce.  Thresholds come from the round-1 sweeps on 2/4/8 B200s
// (profiles/r01/sweep_n*_v1.jsonl): best busbw per (n, size) among the variants.
static int pick_allreduce(size_t bytes, int dtype, int op) {
  const int n = g->ctrl.n;
  const bool pow2 = n == 2 || n == 4 || n == 8;
  int forced = g->algo[B200MPI_COLL_ALLREDUCE];
  if (forced == B200MPI_ALGO_TWOSHOT_SMEM && !pow2) forced = B200MPI_ALGO_TWOSHOT;
  if (forced == B200MPI_ALGO_NVLS && !(g->heap.mc_base && nvls_supports(dtype, op))) forced = 0;
  if (forced == B200MPI_ALGO_LL && bytes > kLLCells * 8) forced = 0;
  if (forced) return forced;
  if (g->ll_max_bytes && bytes <= g->ll_max_bytes) return B200MPI_ALGO_LL;
  const bool nvls = g->heap.mc_base && nvls_supports(dtype, op) && n >= g->nvls_min_ranks;
  const int big = pow2 ? B200MPI_ALGO_TWOSHOT_SMEM : B200MPI_ALGO_TWOSHOT;
  if (n >= 8) {
    if (nvls) return B200MPI_ALGO_NVLS; // fastest at every size from 1 KiB (15 us) to 1 GiB (834 GB/s)
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

// AUTO for Bcast: the switch multicast wins up to a few MiB (one store stream, no second hop);
// above that the fused pull-slice + push keeps every link busy in both directions
// (8 GPUs, 256 MiB: 620 GB/s vs 397 NVLS vs 111 everyone-pulls-from-root).
static int pick_bcast(size_t bytes) {
  const int n = g->ctrl.n;
  int forced = g->algo[B200MPI_COLL_BCAST];
  if (forced == B200MPI_ALGO_NVLS && !g->heap.mc_base) forced = 0;
  if (forced == B200MPI_ALGO_RING || forced == B200MPI_ALGO_TWOSHOT_SMEM) forced = B200MPI_ALGO_TWOSHOT;
  if (forced) return forced;
  if (g->heap.mc_base && bytes % 16 == 0 && bytes <= (4u << 20) && n >= 3) return B200MPI_ALGO_NVLS;
  if (n == 2 || bytes <= (64u << 10)) return B200MPI_ALGO_ONESHOT;
  return B200MPI_ALGO_TWOSHOT;
}

static int pick_allgather(size_t) {
  int forced = g->algo[B200MPI_COLL_ALLGATHER];
  if (forced == B200MPI_ALGO_RING) return forced;
  return B200MPI_ALGO_ONESHOT; // direct push
}

// which: 0 allgather push, 1 allgather ring, 2 bcast (extra = root, mode 0 one-shot / 1 two-shot)
// The grid is a function of the byte count only: ranks may instantiate different access widths
// (their offsets differ in alignment) but must launch the same number of CTAs, because the
// barriers pair CTAs by blockIdx.
template <typename U>
static void launch_units_u(int which, const Comm& c, uint64_t a, uint64_t b, size_t bytes, int extra, int mode, cudaStream_t s) {
  constexpr int UNROLL = 4;
  const size_t vecs = bytes / 16 + 1;
  if (which == 0) allgather_push_kernel<U, UNROLL><<<grid_for(vecs, UNROLL), kThreads, 0, s>>>(c, a, b, bytes);
  else if (which == 1) allgather_ring_kernel<U><<<grid_for(vecs, UNROLL), kThreads, 0, s>>>(c, a, b, bytes);
  else {
    const size_t work = mode == 1 ? vecs / (size_t)(c.n - 1) + 1 : vecs;
    bcast_kernel<U, UNROLL><<<grid_for(work, UNROLL), kThreads, 0, s>>>(c, a, bytes, extra, mode);
  }
}
static int launch_units(int which, const Comm& c, uint64_t a, uint64_t b, size_t bytes, int extra, int mode, cudaStream_t s) {
  // The widest access unit that divides the size and THIS rank's offsets.  Peers may be aligned
  // differently: after sync_start every rank knows all offsets and drops to the byte-wide body if
  // some peer's are narrower than its own unit (kernels.cuh: all_aligned_to).
  const uint64_t m = a | b | bytes;
  if ((m & 15) == 0) launch_units_u<uint4>(which, c, a, b, bytes, extra, mode, s);
  else if ((m & 7) == 0) launch_units_u<unsigned long long>(which, c, a, b, bytes, extra, mode, s);
  else if ((m & 3) == 0) launch_units_u<unsigned int>(which, c, a, b, bytes, extra, mode, s);
  else launch_units_u<unsigned char>(which, c, a, b, bytes, extra, mode, s);
  return launch_check(which == 2 ? "bcast_kernel" : "allgather kernel");
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

// EXPERIMENTAL (functionally tested, latency unmeasured): barrier-free LL allreduce for <= 32 KiB, any local device pointers.
template <typename T, typename Op>
static int launch_ll_t(const void* send, void* recv, size_t count, cudaStream_t s) {
  Comm c = g->comm;
  const uint32_t seq = ++g->ll_seq;
  const size_t ncell = (count * sizeof(T) + 7) / 8;
  const int blocks = (int)std::max<size_t>(1, std::min<size_t>((ncell + 255) / 256, 16));
  allreduce_ll_kernel<T, Op><<<blocks, 256, 0, s>>>(c, (const T*)send, (T*)recv, count, seq);
  return launch_check("allreduce_ll_kernel");
}
static int launch_ll(int dtype, int op, const void* send, void* recv, size_t count, cudaStream_t s) {
#define B200_LL(T)                                                            \
  switch (op) {                                                               \
    case B200MPI_SUM: return launch_ll_t<T, OpSum>(send, recv, count, s);     \
    case B200MPI_MAX: return launch_ll_t<T, OpMax>(send, recv, count, s);     \
    case B200MPI_MIN: return launch_ll_t<T, OpMin>(send, recv, count, s);     \
  }                                                                           \
  break;
  switch (dtype) {
    case B200MPI_F32: B200_LL(float)
    case B200MPI_F64: B200_LL(double)
    case B200MPI_I64: B200_LL(long long)
  }
#undef B200_LL
  return fail(B200MPI_ERR_UNSUPPORTED, "allreduce(LL): unsupported dtype/op");
}

// Host slices (what an unmodified Go caller passes): H2D, collective and D2H are pipelined in
// chunks over three streams so PCIe runs in both directions while the GPUs reduce.  Every rank
// derives the same chunking from (count, dtype), so the per-chunk collectives line up.
static int allreduce_host_pipelined(const void* send, void* recv, size_t count, int dtype, int op) {
  const size_t es = esize(dtype), bytes = count * es;
  const int n = g->ctrl.n;
  int rc = ensure_stage(0, bytes);
  if (rc) return rc;
  char* stage = (char*)g->heap.base[g->ctrl.rank] + g->stage_off[0];
  size_t chunk_elems = std::max<size_t>(g->pipe_chunk_bytes / es, 4096) / 4096 * 4096;
  const size_t nchunks = (count + chunk_elems - 1) / chunk_elems;
  while (g->pipe_events.size() < 2 * nchunks) {
    cudaEvent_t e;
    CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    g->pipe_events.push_back(e);
  }
  CUDA_OK(cudaStreamSynchronize(g->stream)); // earlier work may still use the staging block
  for (size_t k = 0; k < nchunks; ++k) {
    const size_t lo = k * chunk_elems, len = std::min(chunk_elems, count - lo);
    cudaEvent_t in = g->pipe_events[2 * k], out = g->pipe_events[2 * k + 1];
    CUDA_OK(cudaMemcpyAsync(stage + lo * es, (const char*)send + lo * es, len * es, cudaMemcpyHostToDevice, g->h2d_stream));
    CUDA_OK(cudaEventRecord(in, g->h2d_stream));
    if (n > 1) {
      CUDA_OK(cudaStreamWaitEvent(g->stream, in, 0));
      const uint64_t off = g->stage_off[0] + lo * es;
      rc = launch_allreduce(pick_allreduce(len * es, dtype, op), dtype, op, off, off, len, g->stream);
      if (rc) return rc;
      CUDA_OK(cudaEventRecord(out, g->stream));
    } else {
      out = in; // world of one: the D2H only waits for its own H2D
    }
    CUDA_OK(cudaStreamWaitEvent(g->d2h_stream, out, 0));
    CUDA_OK(cudaMemcpyAsync((char*)recv + lo * es, stage + lo * es, len * es, cudaMemcpyDeviceToHost, g->d2h_stream));
  }
  CUDA_OK(cudaStreamSynchronize(g->d2h_stream));
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
  if (memkind == B200MPI_HOST && !async && bytes >= g->pipe_min_bytes) return allreduce_host_pipelined(send, recv, count, dtype, op);
  if (g->ctrl.n == 1) return local_copy(recv, send, bytes, memkind, async);
  if (pick_allreduce(bytes, dtype, op) == B200MPI_ALGO_LL) {
    // any local device pointer works (peers never read it); host slices go through the staging block
    if (memkind == B200MPI_DEVICE) {
      rc = launch_ll(dtype, op, send, recv, count, g->stream);
      return rc ? rc : finish(async);
    }
    Buf b;
    rc = resolve_in(send, bytes, memkind, 0, b);
    if (rc) return rc;
    char* p = (char*)g->heap.base[g->ctrl.rank] + b.off;
    rc = launch_ll(dtype, op, p, p, count, g->stream);
    if (rc) return rc;
    rc = copy_out(recv, bytes, memkind, b);
    return rc ? rc : finish(async);
  }
  Buf in, out;
  rc = resolve_in(send, bytes, memkind, 0, in);
  if (rc) return rc;
  if (send == recv && in.staged) out = in;
  else if (send == recv) out = in;
  else { rc = resolve_out(recv, bytes, memkind, 1, out); if (rc) return rc; }
  const int algo = pick_allreduce(bytes, dtype, op);
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
  Buf b;
  if (g->ctrl.rank == root) rc = resolve_in(buf, bytes, memkind, 0, b);
  else rc = resolve_out(buf, bytes, memkind, 0, b);
  if (rc) return rc;
  const int algo = pick_bcast(bytes);
  g->cur_sig = make_sig(1, dtype, root, algo, count);
  Comm c = next_comm();
  if (algo == B200MPI_ALGO_NVLS) {
    bcast_nvls_kernel<4><<<grid_for(bytes / 16 + 1, 4), kThreads, 0, g->stream>>>(c, b.off, bytes, root);
    rc = launch_check("bcast_nvls_kernel");
  } else {
    rc = launch_units(2, c, b.off, b.off, bytes, root, algo == B200MPI_ALGO_TWOSHOT ? 1 : 0, g->stream);
  }
  if (rc) return rc;
  if (g->ctrl.rank != root) { rc = copy_out(buf, bytes, memkind, b); if (rc) return rc; }
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
  if (n == 1) return local_copy(recv, send, bytes, memkind, async);
  Buf out, in;
  rc = resolve_out(recv, bytes * n, memkind, 1, out);
  if (rc) return rc;
  const bool inplace = (const char*)send == (const char*)recv + (size_t)g->ctrl.rank * bytes;
  if (inplace && !out.staged) { in.off = out.off + (size_t)g->ctrl.rank * bytes; }
  else { rc = resolve_in(send, bytes, memkind, 0, in); if (rc) return rc; }
  const int algo = pick_allgather(bytes);
  g->cur_sig = make_sig(2, dtype, 0, algo, count);
  Comm c = next_comm();
  rc = launch_units(algo == B200MPI_ALGO_RING ? 1 : 0, c, in.off, out.off, bytes, 0, 0, g->stream);
  if (rc) return rc;
  rc = copy_out(recv, bytes * n, memkind, out);
  if (rc) return rc;
  return finish(async);
}

// ---------------------------------------------------------------------------------------------
// point to point
// ---------------------------------------------------------------------------------------------
static int do_send(const void* buf, size_t count, int dtype, int dest, int tag, int memkind) {
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
  MsgSlot* slot = nullptr;
  {
    Spinner sp(g->watchdog_ns);
    for (;;) {
      for (int k = 0; k < kSlotsPerPair && !slot; ++k) {
        uint32_t expect = kFree;
        MsgSlot* s = &g->box->slots[me][dest][k];
        if (s->state.compare_exchange_strong(expect, kClaimed, std::memory_order_acq_rel)) slot = s;
      }
      if (slot) break;
      if (!sp.step()) {
        g->sendtags[dest].remove(tag);
        return fail(B200MPI_ERR_TIMEOUT, "send: no free mailbox slot");
      }
    }
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
  rc = 0;
  if (bytes == 0 || (memkind == B200MPI_DEVICE && g->heap.contains(buf, bytes, off))) {
    slot->chunk_bytes = bytes;
    slot->region_off[0] = slot->region_off[1] = off;
    slot->posted.store(bytes, std::memory_order_relaxed);
    slot->state.store(kPosted, std::memory_order_release);
  } else {
    staged = true;
    const size_t chunk = std::min(bytes, g->stage_chunk);
    const size_t nchunks = (bytes + chunk - 1) / chunk;
    if (g->heap.alloc(chunk * (nchunks > 1 ? 2 : 1), stage)) {
      slot->state.store(kFree, std::memory_order_release);
      g->sendtags[dest].remove(tag);
      return fail(B200MPI_ERR_NOMEM, "send: symmetric heap exhausted while staging (raise B200MPI_HEAP_BYTES)");
    }
    slot->chunk_bytes = chunk;
    slot->region_off[0] = stage;
    slot->region_off[1] = nchunks > 1 ? stage + chunk : stage;
    slot->posted.store(0, std::memory_order_relaxed);
    slot->state.store(kPosted, std::memory_order_release);
    s = borrow_stream();
    char* base = (char*)g->heap.base[me];
    for (size_t k = 0; k < nchunks && rc == 0; ++k) {
      if (k >= 2) { // region k&1 was used by chunk k-2: wait until the receiver drained it
        Spinner sp(g->watchdog_ns);
        while (slot->done.load(std::memory_order_acquire) < (k - 1) * chunk && slot->state.load(std::memory_order_acquire) != kDone)
          if (!sp.step()) { rc = fail(B200MPI_ERR_TIMEOUT, "send: receiver stalled"); break; }
        if (rc) break;
      }
      if (slot->state.load(std::memory_order_acquire) == kDone) break; // receiver gave up (truncate)
      const size_t lo = k * chunk, len = std::min(chunk, bytes - lo);
      cudaError_t e;
      if (memkind == B200MPI_HOST) e = cudaMemcpyAsync(base + slot->region_off[k & 1], (const char*)buf + lo, len, cudaMemcpyHostToDevice, s);
      else { e = cudaSuccess; rc = launch_copy(base + slot->region_off[k & 1], (const char*)buf + lo, len, s); }
      if (e == cudaSuccess && rc == 0) e = cudaStreamSynchronize(s);
      if (e != cudaSuccess) { rc = fail(B200MPI_ERR_CUDA, std::string("send staging: ") + cudaGetErrorString(e)); break; }
      slot->posted.store(lo + len, std::memory_order_release);
    }
  }
  // wait for the receiver's acknowledgement (network.go:569)
  if (rc == 0) {
    Spinner sp(g->watchdog_ns);
    while (slot->state.load(std::memory_order_acquire) != kDone)
      if (!sp.step()) { rc = fail(B200MPI_ERR_TIMEOUT, "send: no matching receive within the watchdog time"); break; }
  }
  if (s) return_stream(s);
  if (staged) g->heap.free_off(stage);
  if (rc == 0) slot->state.store(kFree, std::memory_order_release);
  // on timeout the slot stays claimed: the peer may still touch it
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
          if (s->state.compare_exchange_strong(expect, kMatched, std::memory_order_acq_rel)) slot = s;
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
  if (rc == 0 && bytes) {
    s = borrow_stream();
    if (memkind == B200MPI_HOST) {
      stage_len = std::min(bytes, g->stage_chunk);
      if (g->heap.alloc(stage_len, stage)) { rc = fail(B200MPI_ERR_NOMEM, "recv: symmetric heap exhausted while staging"); stage_len = 0; }
    }
    const char* peer = (const char*)g->heap.base[src];
    char* mine = (char*)g->heap.base[me];
    const size_t chunk = slot->chunk_bytes;
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
      const char* from = peer + slot->region_off[k & 1] + in_chunk;
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
    if (g->box) munmap(g->box, sizeof(Mailbox));
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
  CUDA_OK(cudaHostAlloc(hptr, bytes ? bytes : 1, cudaHostAllocDefault));
  return 0;
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
int b200mpi_heap_info(size_t* total, size_t* used, int* nvls) {
  int rc = need_data_plane();
  if (rc) return rc;
  if (total) *total = g->heap.size;
  if (used) *used = g->heap.used();
  if (nvls) *nvls = g->heap.mc_base ? 1 : 0;
  return 0;
}

int b200mpi_send(const void* buf, size_t count, int dtype, int dest, int tag, int memkind) { return do_send(buf, count, dtype, dest, tag, memkind); }
int b200mpi_recv(void* buf, size_t capacity, size_t* count_out, int dtype, int src, int tag, int memkind) { return do_recv(buf, capacity, count_out, dtype, src, tag, memkind); }

int b200mpi_bcast(void* buf, size_t count, int dtype, int root, int memkind) { return do_bcast(buf, count, dtype, root, memkind, false); }
int b200mpi_allreduce(const void* send, void* recv, size_t count, int dtype, int op, int memkind) { return do_allreduce(send, recv, count, dtype, op, memkind, false); }
int b200mpi_allgather(const void* send, void* recv, size_t count, int dtype, int memkind) { return do_allgather(send, recv, count, dtype, memkind, false); }
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
  barrier_kernel<<<1, 32, 0, g->stream>>>(c);
  int rc = launch_check("barrier_kernel");
  if (rc) return rc;
  return finish(false);
}

int b200mpi_set_algo(int coll, int algo) {
  if (!g || !g->initialised) return fail(B200MPI_ERR_NOT_INIT, "mpi: Init has not been called");
  if (coll < 0 || coll > 2 || algo < 0 || algo > B200MPI_ALGO_LL) return fail(B200MPI_ERR_ARG, "set_algo: bad collective or algorithm id");
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
  return fail(B200MPI_ERR_ARG, "get_algo: bad collective id");
}
int b200mpi_set_param(const char* name, int64_t value) {
  if (!g || !g->initialised) return fail(B200MPI_ERR_NOT_INIT, "mpi: Init has not been called");
  std::string k = name ? name : "";
  if (k == "twoshot_unroll") g->twoshot_unroll = value ? 1 : 0;
  else if (k == "nvls_unroll") g->nvls_unroll = (int)value;
  else if (k == "nvls_min_ranks") g->nvls_min_ranks = (int)value;
  else if (k == "copy_variant") g->copy_variant = (int)value;
  else if (k == "ll_max_bytes") g->ll_max_bytes = (size_t)std::min<int64_t>(std::max<int64_t>(value, 0), (int64_t)(kLLCells * 8));
  else if (k == "nvls_max_blocks") g->nvls_max_blocks = (int)std::max<int64_t>(1, value);
  else if (k == "oneshot_max_bytes") g->oneshot_max_bytes = (size_t)value;
  else if (k == "pipe_min_bytes") g->pipe_min_bytes = (size_t)value;
  else if (k == "pipe_chunk_bytes") g->pipe_chunk_bytes = (size_t)std::max<int64_t>(value, 65536);
  else if (k == "own_block_bytes") g->own_block_bytes = (size_t)std::max<int64_t>(value, 4096);
  else if (k == "stage_chunk") g->stage_chunk = (size_t)std::max<int64_t>(value, 4096);
  else return fail(B200MPI_ERR_ARG, "set_param: unknown parameter '" + k + "'");
  return 0;
}
int b200mpi_set_max_blocks(int blocks) {
  if (!g || !g->initialised) return fail(B200MPI_ERR_NOT_INIT, "mpi: Init has not been called");
  if (blocks < 0 || blocks > kMaxBlocks) return fail(B200MPI_ERR_ARG, "set_max_blocks: out of range");
  g->max_blocks = blocks;
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
