// Host unit test of the device-heap sub-allocator (mpi_b200/csrc/heap.cpp: first fit, 512-byte
// granules, coalescing on free) -- no CUDA call is made.
#include <cassert>
#include <cstdio>
#include <random>
#include <vector>

#include "../../include/b200mpi.h"
#include "../../mpi_b200/csrc/heap.h"

int main() {
  b200::Heap h;
  const size_t total = 64u << 20, reserved = 2u << 20;
  char* fake = reinterpret_cast<char*>(0x100000000ull);
  h.reset_allocator(total, reserved, fake);
  size_t a, b, c, off;
  assert(h.alloc(1, a) == 0 && a == reserved);            // first block right after the control region
  assert(h.alloc(1000, b) == 0 && b == a + 512);          // 512-byte granules
  assert(h.alloc(4096, c) == 0 && c == b + 1024);
  assert(h.used() == 512 + 1024 + 4096);
  assert(h.contains(fake + a, 1, off) && off == a);
  assert(!h.contains(fake + 100, 1, off));                // control region is not user memory
  assert(!h.contains(fake + total - 8, 16, off));         // crosses the end
  assert(h.free_off(b) == 0);
  assert(h.free_off(b) == B200MPI_ERR_ARG);               // double free
  assert(h.free_off(12345) == B200MPI_ERR_ARG);           // never allocated
  size_t d;
  assert(h.alloc(900, d) == 0 && d == b);                 // hole is reused (first fit)
  assert(h.free_off(a) == 0 && h.free_off(d) == 0 && h.free_off(c) == 0);
  assert(h.used() == 0);
  size_t big;
  assert(h.alloc(total - reserved, big) == 0 && big == reserved); // everything coalesced back into one block
  size_t none;
  assert(h.alloc(1, none) == B200MPI_ERR_NOMEM);
  assert(h.free_off(big) == 0);
  // random stress: never overlapping, always aligned, everything returns
  std::mt19937 rng(7);
  std::vector<std::pair<size_t, size_t>> live;
  for (int it = 0; it < 20000; ++it) {
    if (live.empty() || rng() % 3) {
      size_t want = 1 + rng() % (256 << 10), o;
      if (h.alloc(want, o) == 0) {
        assert(o % 512 == 0 && o >= reserved && o + want <= total);
        for (auto& l : live) assert(o + want <= l.first || l.first + l.second <= o);
        live.push_back({o, want});
      }
    } else {
      size_t k = rng() % live.size();
      assert(h.free_off(live[k].first) == 0);
      live.erase(live.begin() + k);
    }
  }
  for (auto& l : live) assert(h.free_off(l.first) == 0);
  assert(h.used() == 0 && h.alloc(total - reserved, big) == 0);
  printf("heap allocator ok\n");
  return 0;
}
