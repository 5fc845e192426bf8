// hostutil.h -- host-only helpers of libb200mpi (no CUDA in this file, so they are unit-tested and
// run under ThreadSanitizer on a CPU box: tests/cpp/hostutil_test.cpp).
#pragma once
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace b200 {

// ---------------------------------------------------------------------------------------------
// NUMA placement.  The DMA engines of GPU g read and write host memory fastest on the socket the
// GPU hangs off (GPU0-3 / GPU4-7 of an HGX box sit on different sockets); rank processes started
// by torchrun / gompirun float across both.  Pinned buffers the library allocates and the helper
// threads that fill them are therefore placed on the GPU's node.
// ---------------------------------------------------------------------------------------------
static inline int read_int_file(const std::string& path, int dflt) {
  FILE* f = fopen(path.c_str(), "r");
  if (!f) return dflt;
  int v = dflt;
  if (fscanf(f, "%d", &v) != 1) v = dflt;
  fclose(f);
  return v;
}
// cpulist ("0-31,64-95") of a node -> cpu set; false when unknown
static inline bool cpus_of_node(int node, cpu_set_t& set) {
  CPU_ZERO(&set);
  if (node < 0) return false;
  FILE* f = fopen(("/sys/devices/system/node/node" + std::to_string(node) + "/cpulist").c_str(), "r");
  if (!f) return false;
  char buf[4096] = {0};
  const bool ok = fgets(buf, sizeof buf, f) != nullptr;
  fclose(f);
  if (!ok) return false;
  int count = 0;
  for (char* p = buf; *p;) {
    char* e;
    long a = strtol(p, &e, 10);
    if (e == p) break;
    long b = a;
    if (*e == '-') { p = e + 1; b = strtol(p, &e, 10); }
    for (long c = a; c <= b && c < CPU_SETSIZE; ++c) { CPU_SET((int)c, &set); ++count; }
    p = (*e == ',') ? e + 1 : e;
    if (*e != ',') break;
  }
  return count > 0;
}
// Runs fn with the calling thread's memory policy preferring `node` and its affinity on the node's
// CPUs (driver-side page allocation follows the calling thread), then restores both.
template <typename F>
static auto on_numa_node(int node, F fn) -> decltype(fn()) {
  cpu_set_t want, old;
  const bool have = cpus_of_node(node, want) && sched_getaffinity(0, sizeof old, &old) == 0;
  unsigned long mask[16] = {0}, old_mask[16] = {0};
  int old_mode = 0;
  bool have_policy = false;
  if (have) {
    have_policy = syscall(SYS_get_mempolicy, &old_mode, old_mask, sizeof old_mask * 8, nullptr, 0) == 0;
    if (node < (int)(sizeof mask * 8)) mask[node / (8 * sizeof(long))] |= 1ul << (node % (8 * sizeof(long)));
    syscall(SYS_set_mempolicy, 1 /* MPOL_PREFERRED */, mask, sizeof mask * 8);
    // only narrow the affinity: a caller pinned elsewhere (taskset, cgroup cpuset) stays where it is
    cpu_set_t both;
    CPU_AND(&both, &want, &old);
    if (CPU_COUNT(&both) > 0) sched_setaffinity(0, sizeof both, &both);
  }
  auto r = fn();
  if (have) {
    // back to what the caller had (MPOL_DEFAULT when it could not be read)
    if (have_policy && old_mode != 0) syscall(SYS_set_mempolicy, old_mode, old_mask, sizeof old_mask * 8);
    else syscall(SYS_set_mempolicy, 0 /* MPOL_DEFAULT */, nullptr, 0);
    sched_setaffinity(0, sizeof old, &old);
  }
  return r;
}

// Helper threads that move bytes between a caller's pageable buffers and pinned bounce chunks.
// This is staging for the DMA engines, not a data path between ranks: every byte still travels
// host -> GPU -> NVLink -> GPU -> host.
struct CopyPool {
  struct Task { char* dst; const char* src; size_t n; std::atomic<int>* pending; };
  std::vector<std::thread> threads;
  std::mutex mu;
  std::condition_variable cv;
  std::deque<Task> q;
  bool stop = false;
  int nthreads = 0;
  bool running() const { return !threads.empty(); }
  void start(int n, int numa_node) {
    nthreads = std::max(n, 1);
    for (int i = 0; i < nthreads; ++i)
      threads.emplace_back([this, numa_node] {
        cpu_set_t set;
        if (cpus_of_node(numa_node, set)) sched_setaffinity(0, sizeof set, &set);
        for (;;) {
          Task t;
          {
            std::unique_lock<std::mutex> l(mu);
            cv.wait(l, [this] { return stop || !q.empty(); });
            if (q.empty()) return;
            t = q.front();
            q.pop_front();
          }
          memcpy(t.dst, t.src, t.n);
          t.pending->fetch_sub(1, std::memory_order_release);
        }
      });
  }
  void submit(char* dst, const char* src, size_t n, std::atomic<int>& pending) {
    if (n == 0) return;
    const size_t piece = std::max<size_t>(256u << 10, ((n + nthreads) / (nthreads + 1) + 4095) / 4096 * 4096);
    std::lock_guard<std::mutex> l(mu);
    for (size_t o = 0; o < n; o += piece) {
      pending.fetch_add(1, std::memory_order_relaxed);
      q.push_back({dst + o, src + o, std::min(piece, n - o), &pending});
    }
    cv.notify_all();
  }
  bool help_one() {
    Task t;
    {
      std::lock_guard<std::mutex> l(mu);
      if (q.empty()) return false;
      t = q.front();
      q.pop_front();
    }
    memcpy(t.dst, t.src, t.n);
    t.pending->fetch_sub(1, std::memory_order_release);
    return true;
  }
  void wait(std::atomic<int>& pending) { // the caller copies too while it waits
    while (pending.load(std::memory_order_acquire) > 0)
      if (!help_one()) sched_yield();
  }
  void shutdown() {
    {
      std::lock_guard<std::mutex> l(mu);
      stop = true;
    }
    cv.notify_all();
    for (auto& t : threads) t.join();
    threads.clear();
  }
};

} // namespace b200
