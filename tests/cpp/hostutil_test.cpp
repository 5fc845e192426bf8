// Host-only helpers of libb200mpi (mpi_b200/csrc/hostutil.h): the copy pool that stages pageable host
// slices into pinned bounce chunks, and the NUMA helpers.  Built with -fsanitize=thread / address by
// tests/test_sanitizers.py.
#include <cstdio>
#include <cstring>
#include <random>
#include <vector>

#include "../../mpi_b200/csrc/hostutil.h"

using namespace b200;

static int fail(const char* what) {
  printf("FAIL: %s\n", what);
  return 1;
}

int main() {
  // ---- CopyPool: many overlapping submissions, several counters, the waiter helps -----------------
  CopyPool pool;
  pool.start(4, -1);
  if (!pool.running()) return fail("pool not running");
  std::mt19937_64 rng(12345);
  const size_t kBuf = 6u << 20;
  std::vector<char> src(kBuf), dst(kBuf);
  for (size_t i = 0; i < kBuf; ++i) src[i] = (char)(rng() >> 7);
  for (int round = 0; round < 40; ++round) {
    std::fill(dst.begin(), dst.end(), 0);
    std::vector<char> want(kBuf, 0);
    std::atomic<int> pend[4];
    for (auto& p : pend) p.store(0);
    // four disjoint regions, each cut into pieces by submit(); sizes from 0 to > 1 MiB, odd offsets
    size_t off = 0;
    for (int k = 0; k < 4; ++k) {
      size_t len = (size_t)(rng() % (1500u << 10));
      if (round == 0 && k == 0) len = 0;
      if (off + len > kBuf) len = kBuf - off;
      pool.submit(dst.data() + off, src.data() + off, len, pend[k]);
      memcpy(want.data() + off, src.data() + off, len);
      off = std::min(kBuf, off + len + (size_t)(rng() % 3));
    }
    for (int k = 3; k >= 0; --k) pool.wait(pend[k]); // any order
    for (auto& p : pend)
      if (p.load() != 0) return fail("counter not drained");
    if (memcmp(dst.data(), want.data(), kBuf) != 0) return fail("copied bytes differ from the submitted regions");
  }
  // exact check of one big copy
  {
    std::fill(dst.begin(), dst.end(), 0);
    std::atomic<int> p(0);
    pool.submit(dst.data() + 3, src.data() + 3, kBuf - 7, p);
    pool.wait(p);
    if (memcmp(dst.data() + 3, src.data() + 3, kBuf - 7) != 0 || dst[0] || dst[1] || dst[2] || dst[kBuf - 1]) return fail("big copy differs");
  }
  pool.shutdown();
  if (pool.running()) return fail("pool still running after shutdown");

  // ---- NUMA helpers ---------------------------------------------------------------------------------
  cpu_set_t before, after, node0;
  sched_getaffinity(0, sizeof before, &before);
  int ran = on_numa_node(0, [&] { return 41 + 1; });
  if (ran != 42) return fail("on_numa_node did not run the function");
  ran = on_numa_node(-1, [&] { return 7; }); // unknown node: just runs
  if (ran != 7) return fail("on_numa_node(-1)");
  sched_getaffinity(0, sizeof after, &after);
  if (!CPU_EQUAL(&before, &after)) return fail("affinity not restored");
  if (cpus_of_node(0, node0)) {
    if (CPU_COUNT(&node0) < 1) return fail("node 0 without CPUs");
  }
  if (cpus_of_node(4096, node0)) return fail("node 4096 exists?");
  if (read_int_file("/nonexistent/file", -5) != -5) return fail("read_int_file default");
  printf("hostutil ok\n");
  return 0;
}
