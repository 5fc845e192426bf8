// Host-side check of the ownership arithmetic the owner-reduces kernels share (kernels.cuh: Owner):
// for every (vector count, block shift, world size) each 16-byte vector of the message is owned by
// exactly one rank exactly once, and a rank's surplus slots (last, partial block) map past the end.
// Compiled with nvcc, runs on the CPU: "every element is reduced exactly once" without a GPU.
#include <cstdio>
#include <vector>

#include "../../mpi_b200/csrc/kernels.cuh"

int main() {
  using b200::Owner;
  const size_t sizes[] = {0, 1, 2, 3, 7, 8, 9, 255, 256, 257, 1000, 4095, 4096, 4097, 65535, 65536, 65537, 100003, 1u << 20, (1u << 20) + 5};
  long checked = 0;
  for (int n = 1; n <= 8; ++n)
    for (uint32_t shift = 0; shift <= 17; ++shift)
      for (size_t nvec : sizes) {
        std::vector<unsigned char> hits(nvec, 0);
        size_t surplus = 0;
        for (int r = 0; r < n; ++r) {
          Owner own(nvec, shift, n, r);
          if (own.slots & (((size_t)1 << shift) - 1)) { printf("slots not a whole number of blocks\n"); return 1; }
          for (size_t l = 0; l < own.slots; ++l) {
            const size_t g = own.global(l);
            if (g < nvec) {
              if (++hits[g] != 1) { printf("vector %zu owned twice (n=%d shift=%u nvec=%zu)\n", g, n, shift, nvec); return 1; }
              if ((g >> shift) % n != (size_t)r) { printf("vector %zu in the wrong rank's block\n", g); return 1; }
            } else ++surplus;
          }
        }
        for (size_t g = 0; g < nvec; ++g)
          if (hits[g] != 1) { printf("vector %zu not owned (n=%d shift=%u nvec=%zu)\n", g, n, shift, nvec); return 1; }
        if (surplus >= ((size_t)1 << shift) * (size_t)n + ((size_t)1 << shift)) { printf("too many surplus slots\n"); return 1; }
        ++checked;
      }
  // LL lanes: every world size fits the region, and the per-call capacity is what DESIGN.md says
  for (int n = 1; n <= 8; ++n)
    if (2 * (size_t)n * b200::ll_cells(n) > b200::kLLRegionCells) { printf("LL lanes overflow the region at n=%d\n", n); return 1; }
  if (b200::ll_cells(8) * 8 != (256u << 10) || b200::ll_cells(2) * 8 != (1u << 20)) { printf("LL capacity changed\n"); return 1; }
  printf("owner ok (%ld cases)\n", checked);
  return 0;
}
