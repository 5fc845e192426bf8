// heap.cpp -- see heap.h.
#include "heap.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#include <algorithm>

#include "../../include/b200mpi.h"

namespace b200 {

namespace {
template <typename F>
bool entry(F& f, const char* name, std::string& err) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || p == nullptr || q != cudaDriverEntryPointSuccess) {
    err = std::string("driver entry point ") + name + " unavailable: " + cudaGetErrorString(e);
    (void)cudaGetLastError();
    return false;
  }
  f = reinterpret_cast<F>(p);
  return true;
}
} // namespace

bool Driver::load(std::string& err) {
  return entry(GetErrorString, "cuGetErrorString", err) &&
         entry(DeviceGetAttribute, "cuDeviceGetAttribute", err) &&
         entry(MemGetAllocationGranularity, "cuMemGetAllocationGranularity", err) &&
         entry(MemCreate, "cuMemCreate", err) && entry(MemRelease, "cuMemRelease", err) &&
         entry(MemExportToShareableHandle, "cuMemExportToShareableHandle", err) &&
         entry(MemImportFromShareableHandle, "cuMemImportFromShareableHandle", err) &&
         entry(MemAddressReserve, "cuMemAddressReserve", err) &&
         entry(MemAddressFree, "cuMemAddressFree", err) && entry(MemMap, "cuMemMap", err) &&
         entry(MemUnmap, "cuMemUnmap", err) && entry(MemSetAccess, "cuMemSetAccess", err) &&
         entry(MulticastCreate, "cuMulticastCreate", err) &&
         entry(MulticastAddDevice, "cuMulticastAddDevice", err) &&
         entry(MulticastBindMem, "cuMulticastBindMem", err) &&
         entry(MulticastUnbind, "cuMulticastUnbind", err) &&
         entry(MulticastGetGranularity, "cuMulticastGetGranularity", err);
}

std::string Driver::errstr(CUresult r) const {
  const char* s = nullptr;
  if (GetErrorString) GetErrorString(r, &s);
  return std::string(s ? s : "unknown") + " (" + std::to_string((int)r) + ")";
}

#define DRV(call)                                                          \
  do {                                                                     \
    CUresult _r = (call);                                                  \
    if (_r != CUDA_SUCCESS) {                                              \
      err = std::string(#call) + ": " + drv.errstr(_r);                    \
      return B200MPI_ERR_CUDA;                                             \
    }                                                                      \
  } while (0)

int Heap::create(Driver& drv, Ctrl& ctrl, int device, size_t bytes, bool want_nvls, std::string& err) {
  dev = device;
  rank = ctrl.rank;
  n = ctrl.n;
  CUmemAllocationProp prop = {};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = dev;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  size_t gran = 0;
  DRV(drv.MemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  CUmulticastObjectProp mprop = {};
  size_t mgran = 0;
  if (want_nvls) {
    mprop.numDevices = (unsigned)n;
    mprop.size = bytes;
    mprop.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    if (drv.MulticastGetGranularity(&mgran, &mprop, CU_MULTICAST_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS) {
      want_nvls = false;
      mgran = 0;
    }
  }
  // every rank must agree on size and on whether NVLS is attempted
  struct Plan { uint64_t gran, mgran, want; } mine = {gran, mgran, want_nvls ? 1u : 0u}, all[B200MPI_MAX_RANKS];
  int rc = ctrl.allgather(&mine, sizeof mine, all, err);
  if (rc) return rc;
  for (int r = 0; r < n; ++r) {
    gran = std::max<size_t>(gran, all[r].gran);
    mgran = std::max<size_t>(mgran, all[r].mgran);
    want_nvls = want_nvls && all[r].want;
  }
  size_t align = want_nvls ? std::max(gran, mgran) : gran;
  size = (bytes + align - 1) / align * align;

  DRV(drv.MemCreate(&handle, size, &prop, 0));
  int fd = -1;
  DRV(drv.MemExportToShareableHandle(&fd, handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  std::vector<int> fds;
  rc = ctrl.alltoall_fd(fd, fds, err);
  ::close(fd);
  if (rc) return rc;

  CUmemAccessDesc access = {};
  access.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  access.location.id = dev;
  access.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  base.assign(n, 0);
  peer_handle.assign(n, 0);
  for (int r = 0; r < n; ++r) {
    CUmemGenericAllocationHandle h = handle;
    if (r != rank) {
      DRV(drv.MemImportFromShareableHandle(&h, (void*)(uintptr_t)fds[r], CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
      peer_handle[r] = h;
    }
    ::close(fds[r]);
    DRV(drv.MemAddressReserve(&base[r], size, align, 0, 0));
    DRV(drv.MemMap(base[r], size, 0, h, 0));
    CUresult ar = drv.MemSetAccess(base[r], size, &access, 1);
    if (ar != CUDA_SUCCESS) {
      err = "cannot map rank " + std::to_string(r) + "'s heap on device " + std::to_string(dev) +
            " (no P2P path?): " + drv.errstr(ar);
      return B200MPI_ERR_CUDA;
    }
  }
  reserved = kHeapReserved;
  if (cudaMemset((void*)base[rank], 0, reserved) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
    err = std::string("heap control region clear failed: ") + cudaGetErrorString(cudaGetLastError());
    return B200MPI_ERR_CUDA;
  }
  {
    std::lock_guard<std::mutex> g(mu_);
    free_.clear();
    live_.clear();
    free_[reserved] = size - reserved;
  }
  rc = ctrl.barrier(err);
  if (rc) return rc;

  // ---- multicast (NVLS) mapping: best effort, all-or-nothing across ranks --------------------
  uint32_t ok = want_nvls ? 1u : 0u;
  std::string why;
  if (want_nvls) {
    mprop.size = size;
    int mfd = -1, got = -1;
    if (rank == 0) {
      CUresult r1 = drv.MulticastCreate(&mc_handle, &mprop);
      if (r1 != CUDA_SUCCESS) { ok = 0; why = "cuMulticastCreate: " + drv.errstr(r1); }
      else if ((r1 = drv.MemExportToShareableHandle(&mfd, mc_handle, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0)) != CUDA_SUCCESS) {
        ok = 0; why = "multicast export: " + drv.errstr(r1);
      }
    }
    // rank 0 tells everyone whether there is an fd to expect
    uint32_t flags[B200MPI_MAX_RANKS];
    rc = ctrl.allgather(&ok, sizeof ok, flags, err);
    if (rc) return rc;
    if (!flags[0]) ok = 0;
    if (flags[0]) {
      rc = ctrl.bcast_fd(0, mfd, got, err);
      if (rc) return rc;
      if (rank != 0) {
        CUresult r2 = drv.MemImportFromShareableHandle(&mc_handle, (void*)(uintptr_t)got, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
        if (r2 != CUDA_SUCCESS) { ok = 0; why = "multicast import: " + drv.errstr(r2); mc_handle = 0; }
      }
      if (got >= 0) ::close(got);
      if (mfd >= 0) ::close(mfd);
      if (ok) {
        CUresult r3 = drv.MulticastAddDevice(mc_handle, dev);
        if (r3 != CUDA_SUCCESS) { ok = 0; why = "cuMulticastAddDevice: " + drv.errstr(r3); }
      }
      rc = ctrl.allgather(&ok, sizeof ok, flags, err); // all devices added before anyone binds
      if (rc) return rc;
      for (int r = 0; r < n; ++r) ok = ok && flags[r];
      if (ok) {
        CUresult r4 = drv.MulticastBindMem(mc_handle, 0, handle, 0, size, 0);
        if (r4 != CUDA_SUCCESS) { ok = 0; why = "cuMulticastBindMem: " + drv.errstr(r4); }
        else mc_bound = true;
      }
      if (ok) {
        CUresult r5 = drv.MemAddressReserve(&mc_base, size, align, 0, 0);
        if (r5 == CUDA_SUCCESS) r5 = drv.MemMap(mc_base, size, 0, mc_handle, 0);
        if (r5 == CUDA_SUCCESS) r5 = drv.MemSetAccess(mc_base, size, &access, 1);
        if (r5 != CUDA_SUCCESS) { ok = 0; why = "multicast map: " + drv.errstr(r5); }
      }
      rc = ctrl.allgather(&ok, sizeof ok, flags, err);
      if (rc) return rc;
      for (int r = 0; r < n; ++r) ok = ok && flags[r];
    }
    if (!ok) {
      if (!why.empty() && getenv("B200MPI_DEBUG")) fprintf(stderr, "[b200mpi %d] NVLS disabled: %s\n", rank, why.c_str());
      mc_base = 0; // mapping (if any) is torn down in destroy()
    }
  }
  return 0;
}

void Heap::destroy(Driver& drv) {
  if (size == 0) return;
  if (mc_handle) {
    if (mc_bound) drv.MulticastUnbind(mc_handle, dev, 0, size);
    // mc_base may be 0 when setup failed half way; an unmapped reservation is simply leaked
    if (mc_base) {
      drv.MemUnmap(mc_base, size);
      drv.MemAddressFree(mc_base, size);
    }
    drv.MemRelease(mc_handle);
  }
  for (int r = 0; r < (int)base.size(); ++r) {
    if (!base[r]) continue;
    drv.MemUnmap(base[r], size);
    drv.MemAddressFree(base[r], size);
    if (r != rank && peer_handle[r]) drv.MemRelease(peer_handle[r]);
  }
  if (handle) drv.MemRelease(handle);
  handle = 0;
  mc_handle = 0;
  mc_base = 0;
  mc_bound = false;
  base.clear();
  peer_handle.clear();
  size = 0;
  std::lock_guard<std::mutex> g(mu_);
  free_.clear();
  live_.clear();
}

int Heap::alloc(size_t bytes, size_t& off) {
  const size_t a = 512; // also keeps every block 16-byte aligned for vector and multimem access
  size_t need = (std::max<size_t>(bytes, 1) + a - 1) / a * a;
  std::lock_guard<std::mutex> g(mu_);
  for (auto it = free_.begin(); it != free_.end(); ++it) {
    if (it->second >= need) {
      off = it->first;
      size_t rest = it->second - need;
      free_.erase(it);
      if (rest) free_[off + need] = rest;
      live_[off] = need;
      return 0;
    }
  }
  return B200MPI_ERR_NOMEM;
}

int Heap::free_off(size_t off) {
  std::lock_guard<std::mutex> g(mu_);
  auto it = live_.find(off);
  if (it == live_.end()) return B200MPI_ERR_ARG;
  size_t len = it->second;
  live_.erase(it);
  auto nx = free_.lower_bound(off);
  if (nx != free_.end() && off + len == nx->first) { // merge with the following hole
    len += nx->second;
    nx = free_.erase(nx);
  }
  if (nx != free_.begin()) { // merge with the preceding hole
    auto pv = std::prev(nx);
    if (pv->first + pv->second == off) {
      pv->second += len;
      return 0;
    }
  }
  free_[off] = len;
  return 0;
}

bool Heap::contains(const void* p, size_t bytes, size_t& off) const {
  if (size == 0) return false;
  uintptr_t a = (uintptr_t)p, b = (uintptr_t)base[rank];
  if (a < b + reserved || a + bytes > b + size) return false;
  off = a - b;
  return true;
}

void Heap::reset_allocator(size_t total, size_t reserved_bytes, void* fake_base) {
  std::lock_guard<std::mutex> g(mu_);
  size = total;
  reserved = reserved_bytes;
  rank = 0;
  n = 1;
  base.assign(1, (CUdeviceptr)(uintptr_t)fake_base);
  peer_handle.assign(1, 0);
  free_.clear();
  live_.clear();
  free_[reserved] = size - reserved;
}

size_t Heap::used() const {
  std::lock_guard<std::mutex> g(mu_);
  size_t u = 0;
  for (auto& kv : live_) u += kv.second;
  return u;
}

} // namespace b200
