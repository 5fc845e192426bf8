// heap.h -- the peer-mapped device heap of one rank, built on the CUDA virtual-memory-management
// driver API (cuMemCreate / export as POSIX fd / import / map) so that the same allocation can be
//   (a) mapped by every other rank's process (plain P2P ld/st over NVLink), and
//   (b) bound to a multicast object (NVLS: multimem.ld_reduce / multimem.st through the switch).
// Replaces the reference's per-pair net.Conn table (/root/reference/network.go:501-506): after
// init "talking to rank r" means dereferencing base[r] + offset.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "ctrl.h"

namespace b200 {

// Control region at the start of every heap: barrier slots, ring flags, LL cells (kernels.cuh: kCtrlBytes).
constexpr size_t kHeapReserved = 16u << 20;

// Driver entry points resolved through cudaGetDriverEntryPoint (no link-time libcuda dependency,
// so the library still loads on a box without a driver and fails loudly at init instead).
struct Driver {
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice) = nullptr;
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags) = nullptr;
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long) = nullptr;
  CUresult (*MemRelease)(CUmemGenericAllocationHandle) = nullptr;
  CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long) = nullptr;
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType) = nullptr;
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long) = nullptr;
  CUresult (*MemAddressFree)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long) = nullptr;
  CUresult (*MemUnmap)(CUdeviceptr, size_t) = nullptr;
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t) = nullptr;
  CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*) = nullptr;
  CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice) = nullptr;
  CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long) = nullptr;
  CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t) = nullptr;
  CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags) = nullptr;
  bool load(std::string& err);
  std::string errstr(CUresult r) const;
};

struct Heap {
  int dev = -1;
  int rank = -1, n = 0;
  size_t size = 0;                        // bytes per rank (same on every rank)
  CUmemGenericAllocationHandle handle = 0;
  std::vector<CUmemGenericAllocationHandle> peer_handle; // imported
  std::vector<CUdeviceptr> base;          // [r] mapping of rank r's heap in this process
  CUmemGenericAllocationHandle mc_handle = 0;
  CUdeviceptr mc_base = 0;                // multicast mapping (0 = NVLS unavailable)
  bool mc_bound = false;

  // Collective over the control plane.  want_nvls: try to set up the multicast mapping.
  int create(Driver& drv, Ctrl& ctrl, int dev, size_t bytes, bool want_nvls, std::string& err);
  void destroy(Driver& drv);

  // Local sub-allocator over [reserved, size).
  int alloc(size_t bytes, size_t& off);
  int free_off(size_t off);
  bool contains(const void* p, size_t bytes, size_t& off) const;
  size_t used() const;
  // Allocator-only initialisation (no device memory): host unit tests of alloc/free_off/contains.
  void reset_allocator(size_t total, size_t reserved_bytes, void* fake_base);

  size_t reserved = 0; // control region
 private:
  mutable std::mutex mu_;
  std::map<size_t, size_t> free_; // offset -> length
  std::map<size_t, size_t> live_; // offset -> length
};

} // namespace b200
