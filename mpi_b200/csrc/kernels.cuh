// kernels.cuh -- sm_100a device code of libb200mpi: the data plane that replaces the
// reference's gob-over-TCP Send/Receive (/root/reference/network.go:518-625) and implements the
// collectives the reference only stubs (/root/reference/mpi.go:130).
//
// Memory model.  Every rank owns one cuMem heap; all heaps are mapped into every rank's address
// space (Comm::base[r]) so a kernel reaches rank r's memory with plain ld/st.global that the
// GPU routes over NVLink 5 / NVSwitch (or stays local when r is this rank or shares the device).
// The first kCtrlBytes of each heap are control words:
//     Slot slots[kMaxBlocks][kMaxRanks]   cross-rank flag + per-call buffer descriptor
//     u32  ring[kMaxBlocks]               ring step counters
// A collective kernel is   sync_start -> body -> sync_end   where both syncs are per-CTA
// barriers between the CTAs with the same blockIdx on every rank (flag writes with
// st.release.sys into the peer's slot, ld.acquire.sys spins on the own slot).  sync_start also
// carries this rank's {send,recv} heap offsets for the call, so user buffers may sit at
// different offsets on different ranks.  Epochs only grow; comparisons are wrap-safe.
//
// Reduction order (what oracle/collectives.c restates, bit for bit):
//   one-shot / two-shot : acc = x_0; acc = op(acc, x_r) for r = 1..n-1          ("rank order")
//   one-shot shuffle    : pairwise tree ((x0+x1)+(x2+x3))+((x4+x5)+(x6+x7))      ("tree order")
//   ring                : chunk c: ((x_c + x_{c+1}) + ...) + x_{c-1}  (cyclic from c) ("ring order")
//   nvls                : order chosen by the switch (tolerance-checked for floats, exact for ints)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200 {

constexpr int kMaxRanks = 8;
constexpr int kMaxBlocks = 592;            // 148 SMs x 4
constexpr size_t kCtrlBytes = 16u << 20;   // control region at the start of every heap (== kHeapReserved, heap.h)
constexpr size_t kRingFlagsOff = 512u << 10;
constexpr size_t kLLStateOff = 768u << 10;   // LL bookkeeping words (device-resident sequence number, CTA counter)
constexpr size_t kLLOff = 1u << 20;          // LL cells: [parity 2][src rank n][ll_cells(n)] x 16 B, 8 MiB in all
constexpr size_t kLLRegionCells = 2 * 8 * 32768;
// cells per lane: 8 payload bytes each -> 256 KiB per call at n = 8, 512 KiB at n = 4, 1 MiB at n = 2
__host__ __device__ constexpr size_t ll_cells(int n) { return kLLRegionCells / (2 * (size_t)(n < 1 ? 1 : n)); }
constexpr int kThreads = 512;
// Every collective launch consumes exactly kEpochStride flag values, whatever the algorithm:
// start barrier = epoch, mid barriers = epoch+1 .. epoch+kEpochStride-2, end barrier =
// epoch+kEpochStride-1.  Ranks that disagreed about a call (B200MPI_ERR_PEER) therefore still
// agree about the epoch of the next one.
constexpr uint32_t kEpochStride = 4096;
constexpr uint32_t kMaxMids = kEpochStride - 2;
constexpr uint32_t kScrubEvery = 1u << 14;  // launches between two scrub_kernel runs (see there): barrier flags then lag by at most
                                            // 2^14 * 4096 = 2^26, ring step counters (epoch * 8) by 2^29, both far below 2^31

// One slot per (CTA, source rank).  The start barrier needs no release fence: the three cells are
// plain 16-byte stores that carry their own flag next to every 8 payload bytes (the LL idea: only
// 8-byte store atomicity is assumed), so a peer's {send offset, recv offset, signature} are valid
// as soon as all six flag words equal this call's epoch -- one NVLink one-way latency instead of a
// round trip (release = wait for the acks of the payload stores) plus a one-way.  `flag` is the
// release/acquire word of the mid and end barriers, where the body's stores must have landed.
struct __align__(64) Slot {
  uint4 cell[3];    // {lo32, epoch, hi32, epoch} of: a = writer's send offset, b = recv offset, c = signature
  uint32_t flag;
  uint32_t pad[3];
};
static_assert(sizeof(Slot) == 64, "slot size");
struct Comm {
  char* base[kMaxRanks];     // heap of rank r as mapped in this process
  char* mc;                  // multicast (NVLS) mapping of all heaps, or nullptr
  uint32_t* status;          // host-mapped word: != 0 after a device-side watchdog timeout
  unsigned long long timeout_ns;
  int rank, n;
  uint32_t epoch;            // start barrier value
  uint32_t end_epoch;        // end barrier value (epoch + 1 + number of reserved mid barriers)
  uint64_t sig;              // what this rank thinks the call is: (count, collective, dtype, op, algorithm)
};

// ---------------------------------------------------------------------------------------------
// flag primitives
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys_u64(uint64_t* p, uint64_t v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_relaxed_sys_u64(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Spin until *p >= want (wrap-safe).  On watchdog expiry raises the host-visible status word
// and gives up so the kernel can drain; the host then reports B200MPI_ERR_TIMEOUT.
__device__ __forceinline__ bool wait_flag(const uint32_t* p, uint32_t want, const Comm& c) {
  unsigned long long t0 = 0;
  uint32_t it = 0;
  while ((int32_t)(ld_acquire_sys(p) - want) < 0) {
    if ((++it & 0xfffu) == 0) {
      unsigned long long now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (c.timeout_ns && now - t0 > c.timeout_ns) {
        *(volatile uint32_t*)c.status = 1u;
        __threadfence_system();
        return false;
      }
    }
  }
  return true;
}

__device__ __forceinline__ Slot* slot_of(const Comm& c, int owner, int block, int src) {
  return reinterpret_cast<Slot*>(c.base[owner]) + (size_t)block * kMaxRanks + src;
}

// Per-CTA barrier #1: announce {a,b,sig} to every rank, wait for everyone's, publish the offsets
// in smem.  Returns false when the body must not run:
//   * a peer's signature differs from ours (status 2 -> B200MPI_ERR_PEER): ranks called different
//     collectives / counts / types / grids, or
//   * a peer did not show up within the watchdog time (status 1 -> B200MPI_ERR_TIMEOUT).
// sync_end still runs, so nobody hangs, and no buffer is touched.
// The signature covers the grid size.  CTAs other than 0 first look at what every peer's CTA 0
// announced (slot row 0 of this rank's heap): if the peers launched a different grid, a CTA
// without a partner learns it there instead of waiting for a flag that never comes.  (While the
// grids match the row cannot be overwritten early: a peer's next kernel starts only after its
// current one has finished, which needs this CTA to have passed sync_end.)
__device__ __forceinline__ void st_cell16(uint4* p, uint4 v) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_cell16(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
// Spin until the three cells of `s` carry `epoch`; values out through a/b/sig.  false on watchdog expiry.
__device__ __forceinline__ bool wait_cells(const Slot* s, uint32_t epoch, const Comm& c, uint64_t& a, uint64_t& b, uint64_t& sig) {
  unsigned long long t0 = 0;
  uint32_t it = 0;
  for (;;) {
    const uint4 x = ld_cell16(&s->cell[0]), y = ld_cell16(&s->cell[1]), z = ld_cell16(&s->cell[2]);
    // >= (wrap-safe), not ==: in normal operation a peer is never ahead (it waits for this CTA at its end
    // barrier), so the two are the same; after a failed call a peer may already be announcing its next
    // one, and a late CTA must then see "newer, different signature" at once instead of spinning.
    if ((int32_t)(x.y - epoch) >= 0 && (int32_t)(x.w - epoch) >= 0 && (int32_t)(y.y - epoch) >= 0 && (int32_t)(y.w - epoch) >= 0 &&
        (int32_t)(z.y - epoch) >= 0 && (int32_t)(z.w - epoch) >= 0) {
      a = ((uint64_t)x.z << 32) | x.x;
      b = ((uint64_t)y.z << 32) | y.x;
      sig = ((uint64_t)z.z << 32) | z.x;
      return true;
    }
    if ((++it & 0xfffu) == 0) {
      const unsigned long long now = globaltimer_ns();
      if (t0 == 0) t0 = now;
      else if (c.timeout_ns && now - t0 > c.timeout_ns) {
        *(volatile uint32_t*)c.status = 1u;
        __threadfence_system();
        return false;
      }
    }
  }
}

// Why sync_start said no, for sync_end_failed: 0 fine, 1 this CTA's partner announced a different
// signature (the partner exists and fails the same way), 2 a peer's CTA 0 announced a different
// signature (grids may differ: this CTA may have no partner at all), 3 watchdog.
__device__ __forceinline__ int* start_verdict() {
  __shared__ int verdict;
  return &verdict;
}

__device__ __forceinline__ bool sync_start(const Comm& c, uint64_t a, uint64_t b, uint64_t* s_a,
                                           uint64_t* s_b) {
  int* verdict = start_verdict();
  const int t = threadIdx.x;
  if (t == 0) *verdict = 0;
  __syncthreads();
  if (t < c.n) {
    Slot* theirs = slot_of(c, t, blockIdx.x, c.rank);
    st_cell16(&theirs->cell[0], make_uint4((uint32_t)a, c.epoch, (uint32_t)(a >> 32), c.epoch));
    st_cell16(&theirs->cell[1], make_uint4((uint32_t)b, c.epoch, (uint32_t)(b >> 32), c.epoch));
    st_cell16(&theirs->cell[2], make_uint4((uint32_t)c.sig, c.epoch, (uint32_t)(c.sig >> 32), c.epoch));
    bool ok = true;
    uint64_t pa, pb, psig;
    if (blockIdx.x != 0) {
      if (!wait_cells(slot_of(c, c.rank, 0, t), c.epoch, c, pa, pb, psig)) { atomicMax(verdict, 3); ok = false; }
      else if (psig != c.sig) { atomicMax(verdict, 2); ok = false; }
    }
    if (ok) {
      if (!wait_cells(slot_of(c, c.rank, blockIdx.x, t), c.epoch, c, pa, pb, psig)) atomicMax(verdict, 3);
      else {
        s_a[t] = pa;
        s_b[t] = pb;
        if (psig != c.sig) atomicMax(verdict, 1);
      }
    }
  }
  __syncthreads();
  const int v = *verdict;
  if (v) {
    if (t == 0 && v != 3) { // a watchdog expiry has already raised status 1
      *(volatile uint32_t*)c.status = 2u;
      __threadfence_system();
    }
    return false;
  }
  return true;
}

// Per-CTA barrier #2: all of this CTA's stores (local, peer and multicast) are released to every
// rank, and every rank's matching CTA has finished reading/writing this rank's buffers.
__device__ __forceinline__ void sync_end(const Comm& c) {
  __syncthreads();
  const int t = threadIdx.x;
  if (t < c.n) {
    st_release_sys(&slot_of(c, t, blockIdx.x, c.rank)->flag, c.end_epoch);
    wait_flag(&slot_of(c, c.rank, blockIdx.x, t)->flag, c.end_epoch, c);
  }
}

// The way out of a kernel whose sync_start said no (the body was skipped, nothing of this CTA is in
// flight).  The end flag is always signalled, so a partner that is waiting gets released; this CTA
// itself waits only when it knows a partner exists and left the same way (verdict 1) -- a CTA that
// learnt about the mismatch from the peers' CTA 0, or that timed out, may have no partner, and
// waiting for one would stall until the watchdog.
__device__ __forceinline__ void sync_end_failed(const Comm& c) {
  __syncthreads();
  const int t = threadIdx.x;
  const bool wait = *start_verdict() == 1;
  if (t < c.n) {
    st_release_sys(&slot_of(c, t, blockIdx.x, c.rank)->flag, c.end_epoch);
    if (wait) wait_flag(&slot_of(c, c.rank, blockIdx.x, t)->flag, c.end_epoch, c);
  }
}

// Per-CTA barrier between a read phase and a write phase of the same addresses (in-place
// one-shot): `value` must lie strictly between c.epoch and c.end_epoch and grow from call to call.
__device__ __forceinline__ void sync_mid(const Comm& c, uint32_t value) {
  __syncthreads();
  const int t = threadIdx.x;
  if (t < c.n) {
    st_release_sys(&slot_of(c, t, blockIdx.x, c.rank)->flag, value);
    wait_flag(&slot_of(c, c.rank, blockIdx.x, t)->flag, value, c);
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// 16-byte packs
// ---------------------------------------------------------------------------------------------
template <typename T>
struct alignas(16) Pack {
  static constexpr int N = 16 / sizeof(T);
  T v[N];
};

__device__ __forceinline__ uint4 ldg16(const void* p) { // streaming 128-bit load (local or peer)
  uint4 r;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ void stg16(void* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}
template <typename T>
__device__ __forceinline__ Pack<T> ld_pack(const void* p) {
  union { uint4 u; Pack<T> k; } x;
  x.u = ldg16(p);
  return x.k;
}
__device__ __forceinline__ uint4 ldg16_sys(const void* p) { // coherent at system scope: never served from L1
  uint4 r;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
template <typename T>
__device__ __forceinline__ Pack<T> ld_pack_sys(const void* p) {
  union { uint4 u; Pack<T> k; } x;
  x.u = ldg16_sys(p);
  return x.k;
}
template <typename T>
__device__ __forceinline__ void st_pack(void* p, const Pack<T>& v) {
  union { uint4 u; Pack<T> k; } x;
  x.k = v;
  stg16(p, x.u);
}

struct OpSum {
  template <typename T> __device__ __forceinline__ static T apply(T a, T b) { return a + b; }
};
template <> __device__ __forceinline__ long long OpSum::apply<long long>(long long a, long long b) {
  return (long long)((unsigned long long)a + (unsigned long long)b); // Go int64 wrap-around
}
struct OpMax {
  template <typename T> __device__ __forceinline__ static T apply(T a, T b) { return b > a ? b : a; }
};
struct OpMin {
  template <typename T> __device__ __forceinline__ static T apply(T a, T b) { return b < a ? b : a; }
};

template <typename T, typename Op>
__device__ __forceinline__ Pack<T> combine(const Pack<T>& a, const Pack<T>& b) {
  Pack<T> r;
#pragma unroll
  for (int i = 0; i < Pack<T>::N; ++i) r.v[i] = Op::template apply<T>(a.v[i], b.v[i]);
  return r;
}

__device__ __forceinline__ bool all_aligned16(const uint64_t* s_a, const uint64_t* s_b, int n) {
  uint64_t m = 0;
  for (int r = 0; r < n; ++r) m |= s_a[r] | s_b[r];
  return (m & 15) == 0;
}

// ---------------------------------------------------------------------------------------------
// Allreduce, one-shot: every rank reads all n buffers and keeps the whole result (latency path).
// The work is cut into rounds of one unit per thread.  When any rank runs in place
// (send == recv) a round is  read+reduce -> sync_mid -> write : nobody overwrites a value that a
// peer has yet to read.  Out of place the mid barriers are skipped.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool any_inplace(const uint64_t* s_a, const uint64_t* s_b, int n) {
  bool ip = false;
  for (int r = 0; r < n; ++r) ip = ip || s_a[r] == s_b[r];
  return ip;
}

template <typename T, typename Op, int NR>
__global__ void __launch_bounds__(kThreads, 1)
allreduce_oneshot_kernel(Comm c, uint64_t send_off, uint64_t recv_off, size_t count) {
  __shared__ uint64_t s_a[kMaxRanks], s_b[kMaxRanks];
  const bool call_ok = sync_start(c, send_off, recv_off, s_a, s_b);
  if (!call_ok) { // ranks disagree about this call: touch nothing, leave through the end barrier
    sync_end_failed(c);
    return;
  }
  const int n = NR ? NR : c.n;
  const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t gstride = (size_t)gridDim.x * blockDim.x;
  char* out = c.base[c.rank] + s_b[c.rank];
  const bool inplace = any_inplace(s_a, s_b, n);
  uint32_t mid = c.epoch + 1;
  constexpr int EPV = Pack<T>::N;
  const bool al = all_aligned16(s_a, s_b, n);
  const size_t nvec = al ? count / EPV : 0;
  for (size_t base = 0; base < nvec; base += gstride) {
    const size_t i = base + gtid;
    Pack<T> acc;
    if (i < nvec) {
      Pack<T> v[NR ? NR : kMaxRanks];
#pragma unroll
      for (int r = 0; r < (NR ? NR : kMaxRanks); ++r)
        if (r < n) v[r] = ld_pack<T>(c.base[r] + s_a[r] + i * 16);
      acc = v[0];
#pragma unroll
      for (int r = 1; r < (NR ? NR : kMaxRanks); ++r)
        if (r < n) acc = combine<T, Op>(acc, v[r]);
    }
    if (inplace) sync_mid(c, mid++);
    if (i < nvec) st_pack<T>(out + i * 16, acc);
  }
  // scalar rounds: the count % EPV tail, or everything when some buffer is not 16-byte aligned
  for (size_t base = nvec * EPV; base < count; base += gstride) {
    const size_t e = base + gtid;
    T acc = T(0);
    if (e < count) {
      acc = reinterpret_cast<const volatile T*>(c.base[0] + s_a[0])[e];
      for (int r = 1; r < n; ++r) acc = Op::template apply<T>(acc, reinterpret_cast<const volatile T*>(c.base[r] + s_a[r])[e]);
    }
    if (inplace) sync_mid(c, mid++);
    if (e < count) reinterpret_cast<T*>(out)[e] = acc;
  }
  sync_end(c);
}

// One-shot with a lane per peer and warp-shuffle partial sums (n = 2, 4, 8; tree order).
// 32/NR vectors per warp step; all NR peer loads of one vector are issued by different lanes in
// the same instruction, so a small message costs one NVLink round trip.
template <typename T, typename Op, int NR>
__global__ void __launch_bounds__(kThreads, 1)
allreduce_oneshot_shfl_kernel(Comm c, uint64_t send_off, uint64_t recv_off, size_t count) {
  __shared__ uint64_t s_a[kMaxRanks], s_b[kMaxRanks];
  __shared__ const char* s_src[kMaxRanks];
  const bool call_ok = sync_start(c, send_off, recv_off, s_a, s_b);
  if (!call_ok) { // ranks disagree about this call: touch nothing, leave through the end barrier
    sync_end_failed(c);
    return;
  }
  if (threadIdx.x < NR) s_src[threadIdx.x] = c.base[threadIdx.x] + s_a[threadIdx.x];
  __syncthreads();
  constexpr int EPV = Pack<T>::N;
  constexpr int VPW = 32 / NR; // vectors per warp step
  const int lane = threadIdx.x & 31;
  const int peer = lane % NR, sub = lane / NR;
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  char* out = c.base[c.rank] + s_b[c.rank];
  const bool al = all_aligned16(s_a, s_b, NR);
  const bool inplace = any_inplace(s_a, s_b, NR);
  uint32_t mid = c.epoch + 1;
  const size_t nvec = al ? count / EPV : 0;
  const char* src = s_src[peer];
  for (size_t round = 0; round < nvec; round += nwarps * VPW) { // uniform across the CTA
    const size_t i = round + warp * VPW + sub;
    const bool valid = i < nvec;
    Pack<T> v;
    if (valid) v = ld_pack<T>(src + i * 16);
    else
      for (int k = 0; k < EPV; ++k) v.v[k] = T(0);
#pragma unroll
    for (int m = 1; m < NR; m <<= 1) {
      union { Pack<T> k; uint32_t w[4]; } a, b;
      a.k = v;
#pragma unroll
      for (int w = 0; w < 4; ++w) b.w[w] = __shfl_xor_sync(0xffffffffu, a.w[w], m);
      v = (peer & m) ? combine<T, Op>(b.k, v) : combine<T, Op>(v, b.k); // lower ranks on the left
    }
    if (inplace) sync_mid(c, mid++);
    if (peer == 0 && valid) st_pack<T>(out + i * 16, v);
  }
  // scalar rounds (tail, or the whole message when unaligned), tree order too
  const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t gstride = (size_t)gridDim.x * blockDim.x;
  for (size_t base = nvec * EPV; base < count; base += gstride) {
    const size_t e = base + gtid;
    T x[NR];
    if (e < count) {
#pragma unroll
      for (int r = 0; r < NR; ++r) x[r] = reinterpret_cast<const volatile T*>(s_src[r])[e];
#pragma unroll
      for (int m = 1; m < NR; m <<= 1)
#pragma unroll
        for (int r = 0; r < NR; r += 2 * m) x[r] = Op::template apply<T>(x[r], x[r + m]);
    }
    if (inplace) sync_mid(c, mid++);
    if (e < count) reinterpret_cast<T*>(out)[e] = x[0];
  }
  sync_end(c);
}

// ---------------------------------------------------------------------------------------------
// Ownership of the message for the owner-reduces kernels (two-shot, smem two-shot, NVLS).
// The message is cut into blocks of 2^shift 16-byte vectors; block k belongs to rank k % n, so at
// any moment all n ranks work inside the same few-MiB window of every heap (fine interleave keeps
// DRAM pages, GPU TLBs and the switch's multicast tables hot; with one contiguous slice per rank
// NVLS lost 30% beyond 256 MiB).  The host picks `shift` from (count, n) only, so every rank
// agrees.  Results do not depend on it: each element is reduced exactly once, in rank order.
// ---------------------------------------------------------------------------------------------
struct Owner {
  size_t nvec;   // whole 16-byte vectors in the message
  size_t slots;  // upper bound of this rank's local vector slots: (blocks owned) << shift
  uint32_t shift;
  int n, rank;
  __host__ __device__ __forceinline__ Owner(size_t nvec_, uint32_t shift_, int n_, int rank_) : nvec(nvec_), shift(shift_), n(n_), rank(rank_) {
    const size_t nblk = (nvec_ + ((size_t)1 << shift_) - 1) >> shift_;
    const size_t mine = nblk > (size_t)rank_ ? (nblk - rank_ + n_ - 1) / n_ : 0;
    slots = mine << shift_;
  }
  // local slot -> global vector index (may be >= nvec in the last, partial block)
  __host__ __device__ __forceinline__ size_t global(size_t slot) const {
    const size_t q = slot >> shift, w = slot & (((size_t)1 << shift) - 1);
    return ((q * n + rank) << shift) + w;
  }
};

// ---------------------------------------------------------------------------------------------
// Allreduce, two-shot fused in one pass: the owner of a vector loads it from all n ranks
// (7 concurrent NVLink read streams + 1 local), reduces in rank order and stores the result into
// every rank's recv buffer (7 NVLink write streams + 1 local).  Reduce-scatter and all-gather
// traffic therefore overlap in both link directions and no mid barrier is needed.
// In place is safe: a vector is read and then written only by its owner.
// ---------------------------------------------------------------------------------------------
// bid / nb: this CTA's index and the number of CTAs working on the message (a fused kernel may give
// only part of its grid to this body).  only_dst >= 0: the reduced value is stored into that
// rank's recv buffer only (Reduce); -1: into every rank's (Allreduce).
template <typename T, typename Op, int NR, int UNROLL>
__device__ __forceinline__ void twoshot_body(const Comm& c, const uint64_t* s_a, const uint64_t* s_b, size_t count, uint32_t shift,
                                             unsigned bid, unsigned nb, int only_dst = -1) {
  const int n = NR ? NR : c.n;
  constexpr int R = NR ? NR : kMaxRanks;
  constexpr int EPV = Pack<T>::N;
  const size_t tid = threadIdx.x;
  if (all_aligned16(s_a, s_b, n)) {
    const size_t nvec = count / EPV;
    const Owner own(nvec, shift, n, c.rank);
    // base pointers live in shared memory, not in 2 x R registers per thread (ptxas spilled the
    // 8-rank and generic instances with register arrays)
    __shared__ const char* src[kMaxRanks];
    __shared__ char* dst[kMaxRanks];
    if (tid < (size_t)R) {
      const int r = (int)tid;
      const int q = r < n ? r : 0;
      src[r] = c.base[q] + s_a[q];
      // stores start at the next rank so the ranks do not all hit the same target at once
      const int w = only_dst >= 0 ? only_dst : (c.rank + 1 + r) % n;
      dst[r] = c.base[w] + s_b[w];
    }
    __syncthreads();
    const int ndst = only_dst >= 0 ? 1 : n;
    const size_t step = (size_t)nb * blockDim.x * UNROLL;
    for (size_t l0 = (size_t)bid * blockDim.x * UNROLL + tid; l0 < own.slots; l0 += step) {
      Pack<T> v[UNROLL][R];
      size_t gi[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const size_t l = l0 + (size_t)u * blockDim.x;
        gi[u] = l < own.slots ? own.global(l) : nvec;
        if (gi[u] < nvec) {
#pragma unroll
          for (int r = 0; r < R; ++r)
            if (r < n) v[u][r] = ld_pack<T>(src[r] + gi[u] * 16);
        }
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        if (gi[u] < nvec) {
          Pack<T> acc = v[u][0];
#pragma unroll
          for (int r = 1; r < R; ++r)
            if (r < n) acc = combine<T, Op>(acc, v[u][r]);
#pragma unroll
          for (int r = 0; r < R; ++r)
            if (r < ndst) st_pack<T>(dst[r] + gi[u] * 16, acc);
        }
      }
    }
    // elements past the last whole vector: reduced and pushed by the last rank
    if (c.rank == n - 1) {
      const size_t gtid = (size_t)bid * blockDim.x + tid;
      for (size_t e = nvec * EPV + gtid; e < count; e += (size_t)nb * blockDim.x) {
        T acc = reinterpret_cast<const volatile T*>(c.base[0] + s_a[0])[e];
        for (int r = 1; r < n; ++r) acc = Op::template apply<T>(acc, reinterpret_cast<const volatile T*>(c.base[r] + s_a[r])[e]);
        for (int r = 0; r < n; ++r)
          if (only_dst < 0 || r == only_dst) reinterpret_cast<volatile T*>(c.base[r] + s_b[r])[e] = acc;
      }
    }
  } else {
    // unaligned buffers: contiguous ownership by element, scalar accesses
    const size_t per = (count + n - 1) / n;
    const size_t lo = per * c.rank < count ? per * c.rank : count;
    const size_t hi = lo + per < count ? lo + per : count;
    for (size_t e = lo + (size_t)bid * blockDim.x + tid; e < hi; e += (size_t)nb * blockDim.x) {
      T acc = reinterpret_cast<const volatile T*>(c.base[0] + s_a[0])[e];
      for (int r = 1; r < n; ++r) acc = Op::template apply<T>(acc, reinterpret_cast<const volatile T*>(c.base[r] + s_a[r])[e]);
      for (int r = 0; r < n; ++r)
        if (only_dst < 0 || r == only_dst) reinterpret_cast<volatile T*>(c.base[r] + s_b[r])[e] = acc;
    }
  }
}

template <typename T, typename Op, int NR, int UNROLL>
__global__ void __launch_bounds__(kThreads, 1)
allreduce_twoshot_kernel(Comm c, uint64_t send_off, uint64_t recv_off, size_t count, uint32_t shift, int only_dst) {
  __shared__ uint64_t s_a[kMaxRanks], s_b[kMaxRanks];
  const bool call_ok = sync_start(c, send_off, recv_off, s_a, s_b);
  if (!call_ok) { // ranks disagree about this call: touch nothing, leave through the end barrier
    sync_end_failed(c);
    return;
  }
  twoshot_body<T, Op, NR, UNROLL>(c, s_a, s_b, count, shift, blockIdx.x, gridDim.x, only_dst);
  sync_end(c);
}

// ---------------------------------------------------------------------------------------------
// Allreduce, two-shot with shared-memory staging (TMA bulk copies).  Same ownership, traffic and
// rank-order arithmetic as allreduce_twoshot_kernel, but the SM's load/store units never touch
// NVLink: one producer thread per CTA streams kChunk-byte tiles of slice j from all NR heaps into
// a ring of shared-memory stages with cp.async.bulk (completion on an mbarrier); 8 consumer warps
// reduce the NR tiles out of shared memory into an output tile, and one thread writes that tile
// to all NR recv buffers with cp.async.bulk shared->global.  In flight per SM: kStages*NR*kChunk.
// ---------------------------------------------------------------------------------------------
constexpr int kSmemChunk = 4096;       // bytes per tile per rank (256 consumer threads x 16 B)
constexpr int kSmemStages = 4;
constexpr int kSmemOutStages = 2;
constexpr int kSmemConsumers = 256;
constexpr int kSmemThreads = kSmemConsumers + 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <typename T, typename Op, int NR>
__global__ void __launch_bounds__(kSmemThreads, 1)
allreduce_twoshot_smem_kernel(Comm c, uint64_t send_off, uint64_t recv_off, size_t count, uint32_t shift) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint64_t s_a[kMaxRanks], s_b[kMaxRanks];
  __shared__ __align__(8) uint64_t full_bar[kSmemStages], empty_bar[kSmemStages];
  unsigned char* in_tiles = smem_raw;                                         // [stage][rank][chunk]
  unsigned char* out_tiles = smem_raw + (size_t)kSmemStages * NR * kSmemChunk; // [ostage][chunk]
  const bool call_ok = sync_start(c, send_off, recv_off, s_a, s_b);
  if (!call_ok) { // ranks disagree about this call: touch nothing, leave through the end barrier
    sync_end_failed(c);
    return;
  }
  constexpr int EPV = Pack<T>::N;
  const int tid = threadIdx.x;
  const bool al = all_aligned16(s_a, s_b, NR);
  const size_t nvec = al ? count / EPV : 0;
  // shift >= 8 (a block is a whole number of 4 KiB tiles); tile t covers local slots [t*256, t*256+256)
  const Owner own(nvec, shift, NR, c.rank);
  constexpr size_t kTileVecs = kSmemChunk / 16;
  const size_t ntiles = own.slots / kTileVecs;
  if (tid == 0) {
    for (int s = 0; s < kSmemStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kSmemConsumers / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid >= kSmemConsumers) {
    // ---- producer warp: one lane drives the TMA loads ----
    if (tid == kSmemConsumers) {
      uint32_t it = 0;
      for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        if (own.global(t * kTileVecs) >= nvec) continue;
        const int s = it % kSmemStages;
        const uint32_t use = it / kSmemStages;
        ++it;
        if (use > 0) mbar_wait(&empty_bar[s], (use - 1) & 1);
        const size_t g0 = own.global(t * kTileVecs);
        const size_t off = g0 * 16;
        const uint32_t bytes = (uint32_t)((nvec - g0) < kTileVecs ? (nvec - g0) * 16 : (size_t)kSmemChunk);
        mbar_expect_tx(&full_bar[s], bytes * NR);
#pragma unroll
        for (int r = 0; r < NR; ++r)
          bulk_g2s(in_tiles + ((size_t)s * NR + r) * kSmemChunk, c.base[r] + s_a[r] + off, bytes, &full_bar[s]);
      }
    }
  } else {
    // ---- consumer warps ----
    uint32_t nit = 0;
    for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const size_t g0 = own.global(t * kTileVecs);
      if (g0 >= nvec) continue;
      const uint32_t it = nit++;
      const int s = it % kSmemStages;
      const int os = it % kSmemOutStages;
      const uint32_t bytes = (uint32_t)((nvec - g0) < kTileVecs ? (nvec - g0) * 16 : (size_t)kSmemChunk);
      mbar_wait(&full_bar[s], (it / kSmemStages) & 1);
      if (it >= (uint32_t)kSmemOutStages) { // the stores that read out_tiles[os] two tiles ago must have drained it
        if (tid == 0) bulk_wait_read<kSmemOutStages - 1>();
        asm volatile("bar.sync 1, %0;" ::"n"(kSmemConsumers) : "memory");
      }
      if ((uint32_t)tid * 16 < bytes) {
        const unsigned char* src = in_tiles + (size_t)s * NR * kSmemChunk + (size_t)tid * 16;
        Pack<T> acc = *reinterpret_cast<const Pack<T>*>(src);
#pragma unroll
        for (int r = 1; r < NR; ++r) acc = combine<T, Op>(acc, *reinterpret_cast<const Pack<T>*>(src + (size_t)r * kSmemChunk));
        *reinterpret_cast<Pack<T>*>(out_tiles + (size_t)os * kSmemChunk + (size_t)tid * 16) = acc;
      }
      fence_proxy_async_smem(); // generic-proxy writes to out_tiles -> visible to the bulk-copy engine
      asm volatile("bar.sync 1, %0;" ::"n"(kSmemConsumers) : "memory");
      if ((tid & 31) == 0) mbar_arrive(&empty_bar[s]); // this warp is done reading in_tiles[s]
      if (tid == 0) {
        const size_t off = g0 * 16;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          const int w = (c.rank + 1 + r) % NR;
          bulk_s2g(c.base[w] + s_b[w] + off, out_tiles + (size_t)os * kSmemChunk, bytes);
        }
        bulk_commit();
      }
    }
    if (tid == 0) bulk_wait<0>(); // every peer write of this CTA has been performed
  }
  // elements outside whole 16-byte vectors, or everything when unaligned: scalar, last rank / by ownership
  {
    const size_t gtid = (size_t)blockIdx.x * blockDim.x + tid;
    const size_t gstride = (size_t)gridDim.x * blockDim.x;
    size_t elo, ehi;
    if (al) { elo = c.rank == NR - 1 ? nvec * EPV : count; ehi = count; }
    else {
      const size_t pe = (count + NR - 1) / NR;
      elo = pe * c.rank < count ? pe * c.rank : count;
      ehi = elo + pe < count ? elo + pe : count;
    }
    for (size_t e = elo + gtid; e < ehi; e += gstride) {
      T acc = reinterpret_cast<const volatile T*>(c.base[0] + s_a[0])[e];
      for (int r = 1; r < NR; ++r) acc = Op::template apply<T>(acc, reinterpret_cast<const volatile T*>(c.base[r] + s_a[r])[e]);
      for (int r = 0; r < NR; ++r) reinterpret_cast<volatile T*>(c.base[r] + s_b[r])[e] = acc;
    }
  }
  sync_end(c);
}

// ---------------------------------------------------------------------------------------------
// Allreduce, ring: n-1 reduce-scatter steps then n-1 all-gather steps; each rank only ever loads
// from its predecessor.  CTA b of rank r depends only on CTA b of rank r-1 (same vector subset on
// every rank), signalled through ring[b] in the successor's heap.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t* ring_flag(const Comm& c, int owner, int block) {
  return reinterpret_cast<uint32_t*>(c.base[owner] + kRingFlagsOff) + block;
}

template <typename T, typename Op>
__global__ void __launch_bounds__(kThreads, 1)
allreduce_ring_kernel(Comm c, uint64_t send_off, uint64_t recv_off, size_t count) {
  __shared__ uint64_t s_a[kMaxRanks], s_b[kMaxRanks];
  const bool call_ok = sync_start(c, send_off, recv_off, s_a, s_b);
  if (!call_ok) { // ranks disagree about this call: touch nothing, leave through the end barrier
    sync_end_failed(c);
    return;
  }
  const int n = c.n, r = c.rank;
  const int prev = (r + n - 1) % n, next = (r + 1) % n;
  constexpr int EPV = Pack<T>::N;
  const bool al = all_aligned16(s_a, s_b, n);
  // Chunks are made of whole 16-byte groups whatever the alignment, so the summation order is a
  // function of (count, n, dtype) only; the count % EPV tail is reduced in rank order.
  const size_t groups = count / EPV;
  const size_t per = (groups + n - 1) / n;
  const char* my_send = c.base[r] + s_a[r];
  char* my_recv = c.base[r] + s_b[r];
  const char* prev_send = c.base[prev] + s_a[prev];
  const char* prev_recv = c.base[prev] + s_b[prev];
  const uint32_t fbase = c.epoch * 8u;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const int steps = 2 * (n - 1);
  for (int g = 0; g < steps; ++g) {
    const bool rs = g < n - 1;
    const int s = rs ? g : g - (n - 1);
    const int chunk = rs ? (r - s - 1 + 2 * n) % n : (r - s + 2 * n) % n;
    if (g > 0) {
      if (threadIdx.x == 0) wait_flag(ring_flag(c, r, blockIdx.x), fbase + (uint32_t)g, c);
      __syncthreads();
    }
    const size_t lo = per * chunk < groups ? per * chunk : groups;
    const size_t hi = lo + per < groups ? lo + per : groups;
    const char* in = (g == 0) ? prev_send : prev_recv;
    if (al) {
      for (size_t i = lo + tid; i < hi; i += stride) {
        Pack<T> p = ld_pack_sys<T>(in + i * 16);
        if (rs) {
          Pack<T> m = ld_pack<T>(my_send + i * 16);
          p = combine<T, Op>(p, m);
        }
        st_pack<T>(my_recv + i * 16, p);
      }
    } else {
      for (size_t e = lo * EPV + tid; e < hi * EPV; e += stride) {
        T p = reinterpret_cast<const volatile T*>(in)[e];
        if (rs) p = Op::template apply<T>(p, reinterpret_cast<const volatile T*>(my_send)[e]);
        reinterpret_cast<volatile T*>(my_recv)[e] = p;
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) st_release_sys(ring_flag(c, next, blockIdx.x), fbase + (uint32_t)g + 1u);
  }
  // tail: at most EPV-1 elements, one per thread of CTA 0, kept in a register across the barrier
  const size_t te = groups * EPV + tid;
  T tail = T(0);
  if (te < count) {
    tail = reinterpret_cast<const volatile T*>(c.base[0] + s_a[0])[te];
    for (int q = 1; q < n; ++q) tail = Op::template apply<T>(tail, reinterpret_cast<const volatile T*>(c.base[q] + s_a[q])[te]);
  }
  sync_end(c);
  __syncthreads(); // sync_end's waits are done by threads 0..n-1 only
  // after the barrier nobody reads this rank's send buffer any more (in-place safe)
  if (te < count) reinterpret_cast<T*>(my_recv)[te] = tail;
}

// ---------------------------------------------------------------------------------------------
// Allreduce through the switch (NVLS): rank j reduces slice j with multimem.ld_reduce (the
// NVSwitch fetches the slice from all n heaps and returns the reduced value) and multicasts the
// result into every rank's recv buffer with multimem.st.  Needs every rank to use the same heap
// offsets (checked after sync_start; otherwise falls back to the two-shot body by returning 0).
// ---------------------------------------------------------------------------------------------
template <typename T, typename Op> struct Multimem;
template <> struct Multimem<float, OpSum> {
  __device__ __forceinline__ static Pack<float> ld_reduce(const void* p) {
    Pack<float> r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]) : "l"(p) : "memory");
    return r;
  }
};
template <> struct Multimem<double, OpSum> {
  __device__ __forceinline__ static Pack<double> ld_reduce(const void* p) {
    Pack<double> r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f64 %0, [%1];" : "=d"(r.v[0]) : "l"(p) : "memory");
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f64 %0, [%1];" : "=d"(r.v[1]) : "l"((const char*)p + 8) : "memory");
    return r;
  }
};
template <> struct Multimem<long long, OpSum> {
  __device__ __forceinline__ static Pack<long long> ld_reduce(const void* p) {
    Pack<long long> r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.u64 %0, [%1];" : "=l"(r.v[0]) : "l"(p) : "memory");
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.u64 %0, [%1];" : "=l"(r.v[1]) : "l"((const char*)p + 8) : "memory");
    return r;
  }
};
template <> struct Multimem<long long, OpMax> {
  __device__ __forceinline__ static Pack<long long> ld_reduce(const void* p) {
    Pack<long long> r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.max.s64 %0, [%1];" : "=l"(r.v[0]) : "l"(p) : "memory");
    asm volatile("multimem.ld_reduce.relaxed.sys.global.max.s64 %0, [%1];" : "=l"(r.v[1]) : "l"((const char*)p + 8) : "memory");
    return r;
  }
};
template <> struct Multimem<long long, OpMin> {
  __device__ __forceinline__ static Pack<long long> ld_reduce(const void* p) {
    Pack<long long> r;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.min.s64 %0, [%1];" : "=l"(r.v[0]) : "l"(p) : "memory");
    asm volatile("multimem.ld_reduce.relaxed.sys.global.min.s64 %0, [%1];" : "=l"(r.v[1]) : "l"((const char*)p + 8) : "memory");
    return r;
  }
};
__device__ __forceinline__ void multimem_st16(void* p, uint4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p),
               "f"(__uint_as_float(v.x)), "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)),
               "f"(__uint_as_float(v.w))
               : "memory");
}

// The switch part of a message: vectors [0, nvec) of the buffers at heap offsets send_off /
// recv_off (the same on every rank).  only_dst < 0: the reduced vector is multicast into every
// rank's recv buffer; >= 0: stored into that rank's recv buffer only (Reduce).
template <typename T, typename Op, int UNROLL>
__device__ __forceinline__ void nvls_body(const Comm& c, uint64_t send_off, uint64_t recv_off, size_t nvec, uint32_t shift,
                                          unsigned bid, unsigned nb, int only_dst = -1) {
  const Owner own(nvec, shift, c.n, c.rank);
  const char* src = c.mc + send_off;
  char* dst = only_dst < 0 ? c.mc + recv_off : c.base[only_dst] + recv_off;
  const size_t tid = threadIdx.x;
  const size_t step = (size_t)nb * blockDim.x * UNROLL;
  for (size_t l0 = (size_t)bid * blockDim.x * UNROLL + tid; l0 < own.slots; l0 += step) {
    Pack<T> v[UNROLL];
    size_t gi[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const size_t l = l0 + (size_t)u * blockDim.x;
      gi[u] = l < own.slots ? own.global(l) : nvec;
      if (gi[u] < nvec) v[u] = Multimem<T, Op>::ld_reduce(src + gi[u] * 16);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if (gi[u] < nvec) {
        union { uint4 u4; Pack<T> k; } x;
        x.k = v[u];
        if (only_dst < 0) multimem_st16(dst + gi[u] * 16, x.u4);
        else stg16(dst + gi[u] * 16, x.u4);
      }
    }
  }
}

// count % EPV trailing elements of an aligned message: last rank, rank order, P2P.
template <typename T, typename Op>
__device__ __forceinline__ void tail_body(const Comm& c, const uint64_t* s_a, const uint64_t* s_b, size_t first, size_t count,
                                          unsigned bid, unsigned nb, int only_dst = -1) {
  const int n = c.n;
  if (c.rank != n - 1) return;
  for (size_t e = first + (size_t)bid * blockDim.x + threadIdx.x; e < count; e += (size_t)nb * blockDim.x) {
    T acc = reinterpret_cast<const volatile T*>(c.base[0] + s_a[0])[e];
    for (int r = 1; r < n; ++r) acc = Op::template apply<T>(acc, reinterpret_cast<const volatile T*>(c.base[r] + s_a[r])[e]);
    for (int r = 0; r < n; ++r)
      if (only_dst < 0 || r == only_dst) reinterpret_cast<volatile T*>(c.base[r] + s_b[r])[e] = acc;
  }
}

__device__ __forceinline__ bool symmetric_offsets(const uint64_t* s_a, const uint64_t* s_b, uint64_t send_off, uint64_t recv_off, int n) {
  bool sym = ((send_off | recv_off) & 15) == 0;
  for (int r = 0; r < n; ++r) sym = sym && s_a[r] == send_off && s_b[r] == recv_off;
  return sym;
}

template <typename T, typename Op, int UNROLL>
__global__ void __launch_bounds__(kThreads, 1)
allreduce_nvls_kernel(Comm c, uint64_t send_off, uint64_t recv_off, size_t count, uint32_t shift, int only_dst) {
  __shared__ uint64_t s_a[kMaxRanks], s_b[kMaxRanks];
  const bool call_ok = sync_start(c, send_off, recv_off, s_a, s_b);
  if (!call_ok) { // ranks disagree about this call: touch nothing, leave through the end barrier
    sync_end_failed(c);
    return;
  }
  constexpr int EPV = Pack<T>::N;
  // Reduce (only_dst >= 0): recv matters on the root only; the other ranks' recv offsets are ignored
  bool symmetric;
  if (only_dst < 0) symmetric = symmetric_offsets(s_a, s_b, send_off, recv_off, c.n);
  else {
    symmetric = ((send_off | s_b[only_dst]) & 15) == 0;
    for (int r = 0; r < c.n; ++r) symmetric = symmetric && s_a[r] == send_off;
  }
  if (symmetric) {
    const size_t nvec = count / EPV;
    nvls_body<T, Op, UNROLL>(c, send_off, only_dst < 0 ? recv_off : s_b[only_dst], nvec, shift, blockIdx.x, gridDim.x, only_dst);
    tail_body<T, Op>(c, s_a, s_b, nvec * EPV, count, blockIdx.x, gridDim.x, only_dst);
  } else {
    // the multicast address needs the same offsets on every rank; otherwise the P2P two-shot body
    twoshot_body<T, Op, 0, 1>(c, s_a, s_b, count, shift, blockIdx.x, gridDim.x, only_dst);
  }
  sync_end(c);
}

// ---------------------------------------------------------------------------------------------
// Allreduce, hybrid: the switch reduction (multimem) saturates below the link rate
// (profiles/r01: 256 MiB..1 GiB x 8 GPUs flat at ~480 GB/s algbw whatever the launch shape), so a
// part of the message goes the P2P way at the same time: CTAs [0, nb_nvls) run nvls_body on
// vectors [0, split), the remaining CTAs run the fused P2P two-shot body on [split, nvec) and the
// tail.  Every element is still reduced exactly once by its owner and the same bits reach every
// rank.  split and nb_nvls are functions of (count, n, tuning parameters) only.
// ---------------------------------------------------------------------------------------------
template <typename T, typename Op, int NR, int UN, int UP>
__global__ void __launch_bounds__(kThreads, 1)
allreduce_hybrid_kernel(Comm c, uint64_t send_off, uint64_t recv_off, size_t count, uint32_t shift_nvls, uint32_t shift_p2p,
                        size_t split_vec, unsigned nb_nvls) {
  __shared__ uint64_t s_a[kMaxRanks], s_b[kMaxRanks], s_a2[kMaxRanks], s_b2[kMaxRanks];
  const bool call_ok = sync_start(c, send_off, recv_off, s_a, s_b);
  if (!call_ok) {
    sync_end_failed(c);
    return;
  }
  constexpr int EPV = Pack<T>::N;
  if (symmetric_offsets(s_a, s_b, send_off, recv_off, c.n)) {
    if (blockIdx.x < nb_nvls) {
      nvls_body<T, Op, UN>(c, send_off, recv_off, split_vec, shift_nvls, blockIdx.x, nb_nvls);
    } else {
      if (threadIdx.x < (unsigned)c.n) {
        s_a2[threadIdx.x] = s_a[threadIdx.x] + split_vec * 16;
        s_b2[threadIdx.x] = s_b[threadIdx.x] + split_vec * 16;
      }
      __syncthreads();
      twoshot_body<T, Op, NR, UP>(c, s_a2, s_b2, count - split_vec * EPV, shift_p2p, blockIdx.x - nb_nvls, gridDim.x - nb_nvls);
    }
  } else {
    twoshot_body<T, Op, NR, UP>(c, s_a, s_b, count, shift_p2p, blockIdx.x, gridDim.x);
  }
  sync_end(c);
}

// ---------------------------------------------------------------------------------------------
// ReduceScatter: rank j ends up with op over r of block j of rank r's send buffer (n blocks of
// `count` elements), in rank order.  Only the owner reads block j, so recv may alias the owner's
// own block of send.  P2P form: n concurrent read streams.  NVLS form: multimem.ld_reduce.
// This is the first half of the two-shot allreduce exposed on its own.
// ---------------------------------------------------------------------------------------------
template <typename T, typename Op, int UNROLL>
__global__ void __launch_bounds__(kThreads, 1)
reduce_scatter_kernel(Comm c, uint64_t send_off, uint64_t recv_off, size_t count) {
  __shared__ uint64_t s_a[kMaxRanks], s_b[kMaxRanks];
  __shared__ const char* src[kMaxRanks];
  const bool call_ok = sync_start(c, send_off, recv_off, s_a, s_b);
  if (!call_ok) {
    sync_end_failed(c);
    return;
  }
  const int n = c.n;
  constexpr int EPV = Pack<T>::N;
  const size_t blk_bytes = count * sizeof(T);
  if (threadIdx.x < (unsigned)n) src[threadIdx.x] = c.base[threadIdx.x] + s_a[threadIdx.x] + (size_t)c.rank * blk_bytes;
  __syncthreads();
  char* out = c.base[c.rank] + recv_off;
  uint64_t m = recv_off | blk_bytes; // block j of every rank must be 16-byte aligned for the vector path
  for (int r = 0; r < n; ++r) m |= s_a[r];
  const size_t nvec = (m & 15) == 0 ? count / EPV : 0;
  const size_t tid = threadIdx.x;
  const size_t step = (size_t)gridDim.x * blockDim.x * UNROLL;
  for (size_t i0 = (size_t)blockIdx.x * blockDim.x * UNROLL + tid; i0 < nvec; i0 += step) {
    Pack<T> acc[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const size_t i = i0 + (size_t)u * blockDim.x;
      if (i < nvec) {
        acc[u] = ld_pack<T>(src[0] + i * 16);
        for (int r = 1; r < n; ++r) acc[u] = combine<T, Op>(acc[u], ld_pack<T>(src[r] + i * 16));
      }
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const size_t i = i0 + (size_t)u * blockDim.x;
      if (i < nvec) st_pack<T>(out + i * 16, acc[u]);
    }
  }
  for (size_t e = nvec * EPV + (size_t)blockIdx.x * blockDim.x + tid; e < count; e += (size_t)gridDim.x * blockDim.x) {
    T acc = reinterpret_cast<const volatile T*>(src[0])[e];
    for (int r = 1; r < n; ++r) acc = Op::template apply<T>(acc, reinterpret_cast<const volatile T*>(src[r])[e]);
    reinterpret_cast<T*>(out)[e] = acc;
  }
  sync_end(c);
}

template <typename T, typename Op, int UNROLL>
__global__ void __launch_bounds__(kThreads, 1)
reduce_scatter_nvls_kernel(Comm c, uint64_t send_off, uint64_t recv_off, size_t count) {
  __shared__ uint64_t s_a[kMaxRanks], s_b[kMaxRanks];
  const bool call_ok = sync_start(c, send_off, recv_off, s_a, s_b);
  if (!call_ok) {
    sync_end_failed(c);
    return;
  }
  const int n = c.n;
  constexpr int EPV = Pack<T>::N;
  const size_t blk_bytes = count * sizeof(T);
  bool symmetric = ((send_off | recv_off | blk_bytes) & 15) == 0;
  for (int r = 0; r < n; ++r) symmetric = symmetric && s_a[r] == send_off && (s_b[r] & 15) == 0;
  char* out = c.base[c.rank] + recv_off;
  const size_t tid = threadIdx.x;
  if (symmetric) {
    const size_t nvec = count / EPV; // blk_bytes % 16 == 0: no tail
    const char* src = c.mc + send_off + (size_t)c.rank * blk_bytes;
    const size_t step = (size_t)gridDim.x * blockDim.x * UNROLL;
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x * UNROLL + tid; i0 < nvec; i0 += step) {
      Pack<T> v[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const size_t i = i0 + (size_t)u * blockDim.x;
        if (i < nvec) v[u] = Multimem<T, Op>::ld_reduce(src + i * 16);
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const size_t i = i0 + (size_t)u * blockDim.x;
        if (i < nvec) st_pack<T>(out + i * 16, v[u]);
      }
    }
  } else { // P2P, scalar, rank order (every rank takes this branch: the decision uses exchanged offsets only)
    for (size_t e = (size_t)blockIdx.x * blockDim.x + tid; e < count; e += (size_t)gridDim.x * blockDim.x) {
      T acc = reinterpret_cast<const volatile T*>(c.base[0] + s_a[0] + (size_t)c.rank * blk_bytes)[e];
      for (int r = 1; r < n; ++r) acc = Op::template apply<T>(acc, reinterpret_cast<const volatile T*>(c.base[r] + s_a[r] + (size_t)c.rank * blk_bytes)[e]);
      reinterpret_cast<T*>(out)[e] = acc;
    }
  }
  sync_end(c);
}

// Coherent (system-scope) loads of one access unit, for data that a peer rewrites during the
// same kernel (ring steps): never served from a stale L1 line.
__device__ __forceinline__ uint4 ld_unit_sys(const volatile uint4* p) { return ldg16_sys((const void*)p); }
__device__ __forceinline__ unsigned long long ld_unit_sys(const volatile unsigned long long* p) { return *p; }
__device__ __forceinline__ unsigned int ld_unit_sys(const volatile unsigned int* p) { return *p; }
__device__ __forceinline__ unsigned char ld_unit_sys(const volatile unsigned char* p) { return *p; }

// ---------------------------------------------------------------------------------------------
// Allgather / Bcast bodies.  U is the access unit the host picked from ITS offsets and the size;
// after sync_start every rank's offsets are known, and a rank that sees a less aligned peer
// drops to the byte-wide instance of the same body.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ bool all_aligned_to(const uint64_t* s_a, const uint64_t* s_b, int n, unsigned unit) {
  uint64_t m = 0;
  for (int r = 0; r < n; ++r) m |= s_a[r] | s_b[r];
  return (m & (unit - 1)) == 0;
}

// Direct push: every rank streams its block once from local HBM and stores it into all n recv
// buffers (own copy skipped when the call is in place).
template <typename U, int UNROLL>
__device__ __forceinline__ void allgather_push_body(const Comm& c, const uint64_t* s_a, const uint64_t* s_b,
                                                    size_t bytes_per_rank) {
  const int n = c.n;
  const size_t units = bytes_per_rank / sizeof(U);
  const U* src = reinterpret_cast<const U*>(c.base[c.rank] + s_a[c.rank]);
  const bool inplace = s_a[c.rank] == s_b[c.rank] + (uint64_t)c.rank * bytes_per_rank;
  U* dst[kMaxRanks];
#pragma unroll
  for (int r = 0; r < kMaxRanks; ++r) {
    const int w = (c.rank + 1 + r) % n; // w == rank when r == n-1
    dst[r] = reinterpret_cast<U*>(c.base[w] + s_b[w] + (size_t)c.rank * bytes_per_rank);
  }
  const int ntargets = inplace ? n - 1 : n;
  const size_t step = (size_t)gridDim.x * blockDim.x * UNROLL;
  for (size_t i0 = (size_t)blockIdx.x * blockDim.x * UNROLL + threadIdx.x; i0 < units; i0 += step) {
    U v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const size_t i = i0 + (size_t)u * blockDim.x;
      if (i < units) v[u] = src[i];
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const size_t i = i0 + (size_t)u * blockDim.x;
      if (i < units) {
#pragma unroll
        for (int r = 0; r < kMaxRanks; ++r)
          if (r < ntargets) dst[r][i] = v[u];
      }
    }
  }
}

template <typename U, int UNROLL>
__global__ void __launch_bounds__(kThreads, 1)
allgather_push_kernel(Comm c, uint64_t send_off, uint64_t recv_off, size_t bytes_per_rank) {
  __shared__ uint64_t s_a[kMaxRanks], s_b[kMaxRanks];
  const bool call_ok = sync_start(c, send_off, recv_off, s_a, s_b);
  if (!call_ok) { // ranks disagree about this call: touch nothing, leave through the end barrier
    sync_end_failed(c);
    return;
  }
  if (sizeof(U) == 1 || all_aligned_to(s_a, s_b, c.n, sizeof(U))) allgather_push_body<U, UNROLL>(c, s_a, s_b, bytes_per_rank);
  else allgather_push_body<unsigned char, 1>(c, s_a, s_b, bytes_per_rank);
  sync_end(c);
}

// Ring allgather: step s pulls block (r - s - 1) from the predecessor.
template <typename U>
__device__ __forceinline__ void allgather_ring_body(const Comm& c, const uint64_t* s_a, const uint64_t* s_b,
                                                    size_t bytes_per_rank) {
  const int n = c.n, r = c.rank;
  const int prev = (r + n - 1) % n, next = (r + 1) % n;
  const size_t units = bytes_per_rank / sizeof(U);
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  char* my_recv = c.base[r] + s_b[r];
  const uint32_t fbase = c.epoch * 8u;
  if (s_a[r] != s_b[r] + (uint64_t)r * bytes_per_rank) { // own block, out of place
    const U* s = reinterpret_cast<const U*>(c.base[r] + s_a[r]);
    U* d = reinterpret_cast<U*>(my_recv + (size_t)r * bytes_per_rank);
    for (size_t i = tid; i < units; i += stride) d[i] = s[i];
  }
  for (int s = 0; s < n - 1; ++s) {
    const int blk = (r - s - 1 + 2 * n) % n;
    if (s > 0) {
      if (threadIdx.x == 0) wait_flag(ring_flag(c, r, blockIdx.x), fbase + (uint32_t)s, c);
      __syncthreads();
    }
    const volatile U* in = (s == 0) ? reinterpret_cast<const volatile U*>(c.base[prev] + s_a[prev])
                                    : reinterpret_cast<const volatile U*>(c.base[prev] + s_b[prev] + (size_t)blk * bytes_per_rank);
    U* out = reinterpret_cast<U*>(my_recv + (size_t)blk * bytes_per_rank);
    for (size_t i = tid; i < units; i += stride) out[i] = ld_unit_sys(in + i);
    __syncthreads();
    if (threadIdx.x == 0) st_release_sys(ring_flag(c, next, blockIdx.x), fbase + (uint32_t)s + 1u);
  }
}

template <typename U>
__global__ void __launch_bounds__(kThreads, 1)
allgather_ring_kernel(Comm c, uint64_t send_off, uint64_t recv_off, size_t bytes_per_rank) {
  __shared__ uint64_t s_a[kMaxRanks], s_b[kMaxRanks];
  const bool call_ok = sync_start(c, send_off, recv_off, s_a, s_b);
  if (!call_ok) { // ranks disagree about this call: touch nothing, leave through the end barrier
    sync_end_failed(c);
    return;
  }
  if (sizeof(U) == 1 || all_aligned_to(s_a, s_b, c.n, sizeof(U))) allgather_ring_body<U>(c, s_a, s_b, bytes_per_rank);
  else allgather_ring_body<unsigned char>(c, s_a, s_b, bytes_per_rank);
  sync_end(c);
}

// ---------------------------------------------------------------------------------------------
// Bcast
// ---------------------------------------------------------------------------------------------
// mode 0 (one-shot): every non-root pulls the whole buffer from root (root egress (n-1)*S).
// mode 1 (two-shot, fused): non-root k pulls slice k from root and stores it locally and into
// the other non-roots, so root sends every byte once and all links run in both directions.
template <typename U, int UNROLL>
__device__ __forceinline__ void bcast_body(const Comm& c, const uint64_t* s_a, size_t bytes, int root, int mode) {
  const int n = c.n;
  if (c.rank == root) return;
  const size_t units = bytes / sizeof(U);
  const U* src = reinterpret_cast<const U*>(c.base[root] + s_a[root]);
  size_t lo = 0, hi = units;
  int ntargets = 1;
  U* dst[kMaxRanks];
  dst[0] = reinterpret_cast<U*>(c.base[c.rank] + s_a[c.rank]);
  if (mode == 1) {
    // slices are cut in bytes on 16-byte granules so that ranks running different access widths
    // (their offsets differ in alignment) still cover the buffer exactly once
    const int m = n - 1;
    const int k = c.rank < root ? c.rank : c.rank - 1;
    const size_t per_b = ((bytes + 15) / 16 + m - 1) / m * 16;
    const size_t lo_b = per_b * k < bytes ? per_b * k : bytes;
    const size_t hi_b = lo_b + per_b < bytes ? lo_b + per_b : bytes;
    lo = lo_b / sizeof(U);
    hi = hi_b / sizeof(U);
    for (int j = 1; j < n; ++j) { // other non-roots, starting after me
      const int w = (c.rank + j) % n;
      if (w == root) continue;
      dst[ntargets++] = reinterpret_cast<U*>(c.base[w] + s_a[w]);
    }
  }
  for (int j = ntargets; j < kMaxRanks; ++j) dst[j] = dst[0];
  const size_t step = (size_t)gridDim.x * blockDim.x * UNROLL;
  for (size_t i0 = lo + (size_t)blockIdx.x * blockDim.x * UNROLL + threadIdx.x; i0 < hi; i0 += step) {
    U v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const size_t i = i0 + (size_t)u * blockDim.x;
      if (i < hi) v[u] = src[i];
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const size_t i = i0 + (size_t)u * blockDim.x;
      if (i < hi) {
#pragma unroll
        for (int r = 0; r < kMaxRanks; ++r)
          if (r < ntargets) dst[r][i] = v[u];
      }
    }
  }
}

template <typename U, int UNROLL>
__global__ void __launch_bounds__(kThreads, 1)
bcast_kernel(Comm c, uint64_t buf_off, size_t bytes, int root, int mode) {
  __shared__ uint64_t s_a[kMaxRanks], s_b[kMaxRanks];
  const bool call_ok = sync_start(c, buf_off, buf_off, s_a, s_b);
  if (!call_ok) { // ranks disagree about this call: touch nothing, leave through the end barrier
    sync_end_failed(c);
    return;
  }
  if (sizeof(U) == 1 || all_aligned_to(s_a, s_b, c.n, sizeof(U))) bcast_body<U, UNROLL>(c, s_a, bytes, root, mode);
  else bcast_body<unsigned char, 1>(c, s_a, bytes, root, mode);
  sync_end(c);
}

// Bcast through the switch: root streams its buffer once into the multicast address.
template <int UNROLL>
__global__ void __launch_bounds__(kThreads, 1)
bcast_nvls_kernel(Comm c, uint64_t buf_off, size_t bytes, int root) {
  __shared__ uint64_t s_a[kMaxRanks], s_b[kMaxRanks];
  const bool call_ok = sync_start(c, buf_off, buf_off, s_a, s_b);
  if (!call_ok) { // ranks disagree about this call: touch nothing, leave through the end barrier
    sync_end_failed(c);
    return;
  }
  const int n = c.n;
  bool symmetric = (buf_off & 15) == 0 && (bytes & 15) == 0;
  for (int r = 0; r < n; ++r) symmetric = symmetric && s_a[r] == buf_off;
  if (symmetric) {
    if (c.rank == root) {
      const size_t nvec = bytes / 16;
      const char* src = c.base[root] + buf_off;
      char* dst = c.mc + buf_off;
      const size_t step = (size_t)gridDim.x * blockDim.x * UNROLL;
      for (size_t i0 = (size_t)blockIdx.x * blockDim.x * UNROLL + threadIdx.x; i0 < nvec; i0 += step) {
        uint4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          const size_t i = i0 + (size_t)u * blockDim.x;
          if (i < nvec) v[u] = ldg16(src + i * 16);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
          const size_t i = i0 + (size_t)u * blockDim.x;
          if (i < nvec) multimem_st16(dst + i * 16, v[u]);
        }
      }
    }
  } else if ((bytes & 15) == 0 && all_aligned_to(s_a, s_b, n, 16)) { // different heap offsets per rank: pull from root
    bcast_body<uint4, UNROLL>(c, s_a, bytes, root, 0);
  } else {
    bcast_body<unsigned char, 1>(c, s_a, bytes, root, 0);
  }
  sync_end(c);
}

// Bcast, scatter + multicast: the buffer is cut into ownership blocks (Owner, as in the allreduce);
// the owner of a block fetches it from root with P2P loads (root owns blocks too and reads them
// locally) and multicasts it into every rank's buffer with multimem.st.  Root's link carries every
// byte once (read responses), every other rank's egress is S/n, and all ingress links fill at the
// same time -- the root-only bcast_nvls_kernel is limited by one GPU's multimem.st rate instead.
// Multicast stores also land in root's own buffer: identical bytes over identical bytes.
template <int UNROLL>
__global__ void __launch_bounds__(kThreads, 1)
bcast_nvls2_kernel(Comm c, uint64_t buf_off, size_t bytes, int root, uint32_t shift) {
  __shared__ uint64_t s_a[kMaxRanks], s_b[kMaxRanks];
  const bool call_ok = sync_start(c, buf_off, buf_off, s_a, s_b);
  if (!call_ok) {
    sync_end_failed(c);
    return;
  }
  const int n = c.n;
  bool symmetric = (buf_off & 15) == 0 && (bytes & 15) == 0;
  for (int r = 0; r < n; ++r) symmetric = symmetric && s_a[r] == buf_off;
  if (symmetric) {
    const size_t nvec = bytes / 16;
    const Owner own(nvec, shift, n, c.rank);
    const char* src = c.base[root] + buf_off;
    char* dst = c.mc + buf_off;
    const size_t tid = threadIdx.x;
    const size_t step = (size_t)gridDim.x * blockDim.x * UNROLL;
    for (size_t l0 = (size_t)blockIdx.x * blockDim.x * UNROLL + tid; l0 < own.slots; l0 += step) {
      uint4 v[UNROLL];
      size_t gi[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const size_t l = l0 + (size_t)u * blockDim.x;
        gi[u] = l < own.slots ? own.global(l) : nvec;
        if (gi[u] < nvec) v[u] = ldg16(src + gi[u] * 16);
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
        if (gi[u] < nvec) multimem_st16(dst + gi[u] * 16, v[u]);
    }
  } else if ((bytes & 15) == 0 && all_aligned_to(s_a, s_b, n, 16)) { // different heap offsets per rank: fused P2P two-shot
    bcast_body<uint4, UNROLL>(c, s_a, bytes, root, 1);
  } else {
    bcast_body<unsigned char, 1>(c, s_a, bytes, root, 1);
  }
  sync_end(c);
}

// Allgather through the switch: every rank streams its block once from local HBM into the
// multicast address of its place in the recv buffer; per-GPU egress is one block instead of n-1.
template <int UNROLL>
__global__ void __launch_bounds__(kThreads, 1)
allgather_nvls_kernel(Comm c, uint64_t send_off, uint64_t recv_off, size_t bytes_per_rank) {
  __shared__ uint64_t s_a[kMaxRanks], s_b[kMaxRanks];
  const bool call_ok = sync_start(c, send_off, recv_off, s_a, s_b);
  if (!call_ok) {
    sync_end_failed(c);
    return;
  }
  const int n = c.n;
  bool symmetric = ((recv_off | bytes_per_rank) & 15) == 0;
  for (int r = 0; r < n; ++r) symmetric = symmetric && s_b[r] == recv_off && (s_a[r] & 15) == 0;
  if (symmetric) {
    const size_t nvec = bytes_per_rank / 16;
    const char* src = c.base[c.rank] + send_off;
    char* dst = c.mc + recv_off + (size_t)c.rank * bytes_per_rank;
    const size_t step = (size_t)gridDim.x * blockDim.x * UNROLL;
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x * UNROLL + threadIdx.x; i0 < nvec; i0 += step) {
      uint4 v[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const size_t i = i0 + (size_t)u * blockDim.x;
        if (i < nvec) v[u] = ldg16(src + i * 16);
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const size_t i = i0 + (size_t)u * blockDim.x;
        if (i < nvec) multimem_st16(dst + i * 16, v[u]);
      }
    }
  } else if ((bytes_per_rank & 15) == 0 && all_aligned_to(s_a, s_b, n, 16)) { // different heap offsets per rank: P2P push
    allgather_push_body<uint4, UNROLL>(c, s_a, s_b, bytes_per_rank);
  } else {
    allgather_push_body<unsigned char, 1>(c, s_a, s_b, bytes_per_rank);
  }
  sync_end(c);
}

// ---------------------------------------------------------------------------------------------
// Alltoall: block j of rank r's send buffer becomes block r of rank j's recv buffer.  Direct
// push, n write streams per GPU, targets rotated so that the ranks do not all hit one peer.
// ---------------------------------------------------------------------------------------------
template <typename U, int UNROLL>
__device__ __forceinline__ void alltoall_body(const Comm& c, const uint64_t* s_a, const uint64_t* s_b, size_t bytes_per_block) {
  const int n = c.n;
  const size_t units = bytes_per_block / sizeof(U);
  const size_t step = (size_t)gridDim.x * blockDim.x * UNROLL;
  for (int k = 0; k < n; ++k) {
    const int j = (c.rank + 1 + k) % n; // own block last
    const U* src = reinterpret_cast<const U*>(c.base[c.rank] + s_a[c.rank] + (size_t)j * bytes_per_block);
    U* dst = reinterpret_cast<U*>(c.base[j] + s_b[j] + (size_t)c.rank * bytes_per_block);
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x * UNROLL + threadIdx.x; i0 < units; i0 += step) {
      U v[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const size_t i = i0 + (size_t)u * blockDim.x;
        if (i < units) v[u] = src[i];
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const size_t i = i0 + (size_t)u * blockDim.x;
        if (i < units) dst[i] = v[u];
      }
    }
  }
}

template <typename U, int UNROLL>
__global__ void __launch_bounds__(kThreads, 1)
alltoall_kernel(Comm c, uint64_t send_off, uint64_t recv_off, size_t bytes_per_block) {
  __shared__ uint64_t s_a[kMaxRanks], s_b[kMaxRanks];
  const bool call_ok = sync_start(c, send_off, recv_off, s_a, s_b);
  if (!call_ok) {
    sync_end_failed(c);
    return;
  }
  if (sizeof(U) == 1 || all_aligned_to(s_a, s_b, c.n, sizeof(U))) alltoall_body<U, UNROLL>(c, s_a, s_b, bytes_per_block);
  else alltoall_body<unsigned char, 1>(c, s_a, s_b, bytes_per_block);
  sync_end(c);
}

// ---------------------------------------------------------------------------------------------
// Point-to-point and local copies: dst <- src, `bytes` bytes, any alignment.  src may be a peer
// mapping (receiver pulls the sender's posted region over NVLink).
// ---------------------------------------------------------------------------------------------
template <int UNROLL, int THREADS = kThreads, int MINB = 1>
__global__ void __launch_bounds__(THREADS, MINB)
copy_bytes_kernel(unsigned char* __restrict__ dst, const unsigned char* __restrict__ src, size_t bytes) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const uintptr_t da = reinterpret_cast<uintptr_t>(dst), sa = reinterpret_cast<uintptr_t>(src);
  if (((da ^ sa) & 15) == 0) {
    size_t head = (16 - (da & 15)) & 15;
    if (head > bytes) head = bytes;
    for (size_t i = tid; i < head; i += stride) dst[i] = src[i];
    const size_t nvec = (bytes - head) / 16;
    const unsigned char* s = src + head;
    unsigned char* d = dst + head;
    const size_t step = stride * UNROLL;
    for (size_t i0 = (size_t)blockIdx.x * blockDim.x * UNROLL + threadIdx.x; i0 < nvec; i0 += step) {
      uint4 v[UNROLL];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const size_t i = i0 + (size_t)u * blockDim.x;
        if (i < nvec) v[u] = ldg16(s + i * 16);
      }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) {
        const size_t i = i0 + (size_t)u * blockDim.x;
        if (i < nvec) stg16(d + i * 16, v[u]);
      }
    }
    for (size_t i = head + nvec * 16 + tid; i < bytes; i += stride) dst[i] = src[i];
  } else {
    for (size_t i = tid; i < bytes; i += stride) dst[i] = src[i];
  }
}

// ---------------------------------------------------------------------------------------------
// Allreduce, LL (low latency) -- EXPERIMENTAL, written at the end of round 1: bit-exact against the
// oracle in worlds of 2 and 4 ranks sharing one B200 (tests/test_gpu_worlds.py::test_ll_...), but
// its latency over NVLink has not been measured yet, so it is off unless "ll_max_bytes" > 0 or
// the algorithm is forced.
// One kernel, no barrier at all.  Every rank pushes its whole (<= 32 KiB) message into a private
// lane of every peer's heap as 16-byte cells {data32, seq, data32, seq} (the NCCL "LL" layout: each
// 8-byte half carries its own flag, so only 8-byte store atomicity is assumed), then reduces its
// own lane set in rank order, spinning per cell until both flags equal this call's sequence number.
// Lanes are double buffered by seq parity: a peer can only be one call ahead, because finishing
// call k+1 needs this rank's k+1 cells, which are sent after this rank finished reading call k.
// Peers never read this rank's send buffer, so there is no end barrier and in place is free; send
// and recv may be any local device pointers.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4* ll_lane(const Comm& c, int owner, uint32_t parity, int src) {
  return reinterpret_cast<uint4*>(c.base[owner] + kLLOff) + ((size_t)parity * c.n + src) * ll_cells(c.n);
}
__device__ __forceinline__ void st_cell(uint4* p, uint4 v) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_cell(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
template <typename T> struct LLCodec; // 8 payload bytes <-> elements
template <> struct LLCodec<float> {
  static constexpr int EPC = 2;
  __device__ __forceinline__ static void pack(const float* e, uint32_t& w0, uint32_t& w1) { w0 = __float_as_uint(e[0]); w1 = __float_as_uint(e[1]); }
  __device__ __forceinline__ static void unpack(uint32_t w0, uint32_t w1, float* e) { e[0] = __uint_as_float(w0); e[1] = __uint_as_float(w1); }
};
template <> struct LLCodec<double> {
  static constexpr int EPC = 1;
  __device__ __forceinline__ static void pack(const double* e, uint32_t& w0, uint32_t& w1) { unsigned long long b = (unsigned long long)__double_as_longlong(e[0]); w0 = (uint32_t)b; w1 = (uint32_t)(b >> 32); }
  __device__ __forceinline__ static void unpack(uint32_t w0, uint32_t w1, double* e) { e[0] = __longlong_as_double((long long)(((unsigned long long)w1 << 32) | w0)); }
};
template <> struct LLCodec<long long> {
  static constexpr int EPC = 1;
  __device__ __forceinline__ static void pack(const long long* e, uint32_t& w0, uint32_t& w1) { unsigned long long b = (unsigned long long)e[0]; w0 = (uint32_t)b; w1 = (uint32_t)(b >> 32); }
  __device__ __forceinline__ static void unpack(uint32_t w0, uint32_t w1, long long* e) { e[0] = (long long)(((unsigned long long)w1 << 32) | w0); }
};

template <typename T, typename Op>
__global__ void __launch_bounds__(256, 1)
allreduce_ll_kernel(Comm c, const T* __restrict__ send, T* recv, size_t count, uint32_t seq, uint32_t* done_host) {
  constexpr int EPC = LLCodec<T>::EPC;
  const size_t ncell = (count + EPC - 1) / EPC;
  const uint32_t parity = seq & 1u;
  const int n = c.n;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  // phase 1: every cell of this thread goes out to every peer before anything is waited for, so a
  // message of many cells costs one exchange latency, not one per cell
  for (size_t i = tid; i < ncell; i += stride) {
    T mine[EPC];
#pragma unroll
    for (int k = 0; k < EPC; ++k) mine[k] = i * EPC + k < count ? send[i * EPC + k] : T(0);
    uint32_t w0, w1;
    LLCodec<T>::pack(mine, w0, w1);
    const uint4 cell = make_uint4(w0, seq, w1, seq);
    for (int j = 1; j < n; ++j) { // push to every peer, starting after this rank
      const int r = (c.rank + j) % n;
      st_cell(ll_lane(c, r, parity, c.rank) + i, cell);
    }
  }
  // phase 2: reduce the own lanes in rank order.  The n-1 cells of one element are loaded together
  // (independent loads in flight: one L2 latency, not n-1) and re-loaded only while their flags lag.
  for (size_t i = tid; i < ncell; i += stride) {
    T mine[EPC];
#pragma unroll
    for (int k = 0; k < EPC; ++k) mine[k] = i * EPC + k < count ? send[i * EPC + k] : T(0);
    uint4 got[kMaxRanks];
    uint32_t pending = 0;
#pragma unroll
    for (int r = 0; r < kMaxRanks; ++r)
      if (r < n && r != c.rank) {
        got[r] = ld_cell(ll_lane(c, c.rank, parity, r) + i);
        pending |= 1u << r;
      }
    unsigned long long t0 = 0;
    uint32_t it = 0;
    for (;;) {
#pragma unroll
      for (int r = 0; r < kMaxRanks; ++r)
        if ((pending >> r) & 1u) {
          if (got[r].y == seq && got[r].w == seq) pending &= ~(1u << r);
        }
      if (!pending) break;
#pragma unroll
      for (int r = 0; r < kMaxRanks; ++r)
        if ((pending >> r) & 1u) got[r] = ld_cell(ll_lane(c, c.rank, parity, r) + i);
      if ((++it & 0xfffu) == 0) {
        const unsigned long long now = globaltimer_ns();
        if (t0 == 0) t0 = now;
        else if (c.timeout_ns && now - t0 > c.timeout_ns) {
          *(volatile uint32_t*)c.status = 1u;
          __threadfence_system();
          break;
        }
      }
    }
    T acc[EPC];
#pragma unroll
    for (int r = 0; r < kMaxRanks; ++r)
      if (r < n) {
        T v[EPC];
        if (r == c.rank) {
#pragma unroll
          for (int k = 0; k < EPC; ++k) v[k] = mine[k];
        } else {
          LLCodec<T>::unpack(got[r].x, got[r].z, v);
        }
#pragma unroll
        for (int k = 0; k < EPC; ++k) acc[k] = r == 0 ? v[k] : Op::template apply<T>(acc[k], v[k]);
      }
#pragma unroll
    for (int k = 0; k < EPC; ++k)
      if (i * EPC + k < count) recv[i * EPC + k] = acc[k];
  }
  // Host-slice calls: send/recv are device-mapped pinned host buffers and the caller spins on
  // *done_host instead of paying a stream synchronisation.  The last CTA to finish publishes seq
  // after every CTA's result stores have been fenced at system scope.
  if (done_host) {
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned* ctas = reinterpret_cast<unsigned*>(c.base[c.rank] + kLLStateOff);
      __threadfence_system();
      if (atomicAdd(ctas, 1u) == gridDim.x - 1) {
        *ctas = 0;
        __threadfence_system();
        *(volatile uint32_t*)done_host = seq;
      }
    }
  }
}

// Completion word for host-polled copies: enqueued right behind a copy kernel on the same stream, so
// the copy's stores (possibly into mapped host memory) are complete when the word becomes visible.
__global__ void flag_kernel(uint32_t* done, uint32_t value) {
  __threadfence_system();
  *(volatile uint32_t*)done = value;
}

// Flag words are compared wrap-safe ((int32_t)(flag - want) >= 0), which is only sound while no word lags
// the epoch by 2^31.  A slot row of a block index that has not been used for 2^31 / kEpochStride =
// 524,288 launches would look "ahead" (65,536 for the ring step counters, which count epoch * 8).
// Every kScrubEvery launches the host therefore runs this
// kernel with the largest grid: its start and end barriers rewrite the cells and the flag of every
// row on every rank, and each CTA refreshes its ring step counter.
__global__ void scrub_kernel(Comm c) {
  __shared__ uint64_t s_a[kMaxRanks], s_b[kMaxRanks];
  const bool ok = sync_start(c, 0, 0, s_a, s_b);
  if (threadIdx.x == 0) *ring_flag(c, c.rank, blockIdx.x) = c.epoch * 8u;
  if (ok) sync_end(c);
  else sync_end_failed(c);
}

// A device-wide rendezvous with nothing in between (b200mpi_barrier).
__global__ void barrier_kernel(Comm c) {
  __shared__ uint64_t s_a[kMaxRanks], s_b[kMaxRanks];
  if (sync_start(c, 0, 0, s_a, s_b)) sync_end(c);
  else sync_end_failed(c);
}

} // namespace b200
