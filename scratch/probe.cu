// Scratch probe: facts about the GPU box that shape the bootstrap design.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <unistd.h>
#include <sys/socket.h>
#include <sys/wait.h>
#include <sys/un.h>
#include <chrono>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA err %s at %d: %s\n", #x, __LINE__, cudaGetErrorString(e)); exit(1);} } while(0)
#define DK(x) do { CUresult e = (x); if (e != CUDA_SUCCESS) { const char* s=""; p_cuGetErrorString(e,&s); printf("[pid %d] DRV err %s at %d: %d %s\n", getpid(), #x, __LINE__, (int)e, s); exit(1);} } while(0)

typedef CUresult (*fn_cuGetErrorString)(CUresult, const char**);
typedef CUresult (*fn_cuMemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
typedef CUresult (*fn_cuMemExport)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
typedef CUresult (*fn_cuMemImport)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType);
typedef CUresult (*fn_cuMemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
typedef CUresult (*fn_cuMemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
typedef CUresult (*fn_cuMemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
typedef CUresult (*fn_cuMemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
typedef CUresult (*fn_cuDeviceGetAttribute)(int*, CUdevice_attribute, CUdevice);
typedef CUresult (*fn_cuMulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags);
fn_cuGetErrorString p_cuGetErrorString; fn_cuMemCreate p_cuMemCreate; fn_cuMemExport p_cuMemExport; fn_cuMemImport p_cuMemImport;
fn_cuMemAddressReserve p_cuMemAddressReserve; fn_cuMemMap p_cuMemMap; fn_cuMemSetAccess p_cuMemSetAccess; fn_cuMemGetAllocationGranularity p_gran;
fn_cuDeviceGetAttribute p_attr; fn_cuMulticastGetGranularity p_mcgran;

template<typename T> void load(T& f, const char* name) {
  cudaDriverEntryPointQueryResult q; void* p = nullptr;
  cudaError_t e = cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q);
  if (e != cudaSuccess || !p) { printf("no entry %s (%s)\n", name, cudaGetErrorString(e)); }
  f = (T)p;
}

static int send_fd(int sock, int fd) {
  struct msghdr msg = {}; char buf[CMSG_SPACE(sizeof(int))]; memset(buf,0,sizeof buf);
  char dummy = 'x'; struct iovec io = { &dummy, 1 };
  msg.msg_iov = &io; msg.msg_iovlen = 1; msg.msg_control = buf; msg.msg_controllen = sizeof buf;
  struct cmsghdr* c = CMSG_FIRSTHDR(&msg); c->cmsg_level = SOL_SOCKET; c->cmsg_type = SCM_RIGHTS; c->cmsg_len = CMSG_LEN(sizeof(int));
  memcpy(CMSG_DATA(c), &fd, sizeof(int));
  return (int)sendmsg(sock, &msg, 0);
}
static int recv_fd(int sock) {
  struct msghdr msg = {}; char buf[CMSG_SPACE(sizeof(int))]; char dummy; struct iovec io = { &dummy, 1 };
  msg.msg_iov = &io; msg.msg_iovlen = 1; msg.msg_control = buf; msg.msg_controllen = sizeof buf;
  if (recvmsg(sock, &msg, 0) <= 0) return -1;
  struct cmsghdr* c = CMSG_FIRSTHDR(&msg); int fd; memcpy(&fd, CMSG_DATA(c), sizeof(int)); return fd;
}

__global__ void pingpong(volatile unsigned* mine, volatile unsigned* peer, int iters, int who, unsigned long long* cycles) {
  unsigned long long t0 = clock64();
  for (int i = 1; i <= iters; ++i) {
    if (who == 0) { *peer = i; __threadfence_system(); while (*mine < (unsigned)i) {} }
    else { while (*mine < (unsigned)i) {} *peer = i; __threadfence_system(); }
  }
  *cycles = clock64() - t0;
}

int run_rank(int who, int sock) {
  CK(cudaSetDevice(0)); CK(cudaFree(0));
  load(p_cuGetErrorString, "cuGetErrorString"); load(p_cuMemCreate, "cuMemCreate"); load(p_cuMemExport, "cuMemExportToShareableHandle");
  load(p_cuMemImport, "cuMemImportFromShareableHandle"); load(p_cuMemAddressReserve, "cuMemAddressReserve"); load(p_cuMemMap, "cuMemMap");
  load(p_cuMemSetAccess, "cuMemSetAccess"); load(p_gran, "cuMemGetAllocationGranularity"); load(p_attr, "cuDeviceGetAttribute"); load(p_mcgran, "cuMulticastGetGranularity");
  if (who == 0) {
    int n; CK(cudaGetDeviceCount(&n)); printf("device count %d\n", n);
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0)); printf("name %s sms %d mem %zu MiB l2 %d\n", pr.name, pr.multiProcessorCount, pr.totalGlobalMem>>20, pr.l2CacheSize);
    int v;
    p_attr(&v, CU_DEVICE_ATTRIBUTE_VIRTUAL_MEMORY_MANAGEMENT_SUPPORTED, 0); printf("vmm %d\n", v);
    p_attr(&v, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, 0); printf("posix_fd %d\n", v);
    p_attr(&v, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_FABRIC_SUPPORTED, 0); printf("fabric %d\n", v);
    p_attr(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, 0); printf("multicast %d\n", v);
    p_attr(&v, CU_DEVICE_ATTRIBUTE_COMPUTE_PREEMPTION_SUPPORTED, 0); printf("preempt %d\n", v);
    p_attr(&v, CU_DEVICE_ATTRIBUTE_COMPUTE_MODE, 0); printf("compute_mode %d\n", v);
    int drv, rt; cudaDriverGetVersion(&drv); cudaRuntimeGetVersion(&rt); printf("driver %d runtime %d\n", drv, rt);
  }
  CUmemAllocationProp prop = {}; prop.type = CU_MEM_ALLOCATION_TYPE_PINNED; prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = 0;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  size_t gran = 0; DK(p_gran(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED)); if (who==0) printf("granularity %zu\n", gran);
  size_t sz = gran; CUmemGenericAllocationHandle h; DK(p_cuMemCreate(&h, sz, &prop, 0));
  int fd = -1; DK(p_cuMemExport(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0));
  CUdeviceptr mine; DK(p_cuMemAddressReserve(&mine, sz, gran, 0, 0)); DK(p_cuMemMap(mine, sz, 0, h, 0));
  CUmemAccessDesc ad = {}; ad.location.type = CU_MEM_LOCATION_TYPE_DEVICE; ad.location.id = 0; ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  DK(p_cuMemSetAccess(mine, sz, &ad, 1));
  CK(cudaMemset((void*)mine, 0, sz)); CK(cudaDeviceSynchronize());
  send_fd(sock, fd); int pfd = recv_fd(sock); 
  CUmemGenericAllocationHandle ph; DK(p_cuMemImport(&ph, (void*)(uintptr_t)pfd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR));
  CUdeviceptr peer; DK(p_cuMemAddressReserve(&peer, sz, gran, 0, 0)); DK(p_cuMemMap(peer, sz, 0, ph, 0)); DK(p_cuMemSetAccess(peer, sz, &ad, 1));
  printf("[%d] vmm fd import ok mine=%p peer=%p\n", who, (void*)mine, (void*)peer);
  // handshake so both are mapped before running
  char c='r'; write(sock,&c,1); read(sock,&c,1);
  unsigned long long* cyc; CK(cudaMalloc(&cyc, 8));
  for (int iters : {1, 10, 100}) {
    // reset flags: each resets own, then sync
    CK(cudaMemset((void*)mine, 0, 64)); CK(cudaDeviceSynchronize()); write(sock,&c,1); read(sock,&c,1);
    auto t0 = std::chrono::steady_clock::now();
    pingpong<<<1,1>>>((volatile unsigned*)mine, (volatile unsigned*)peer, iters, who, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    printf("[%d] pingpong iters=%d: %s, %.3f ms wall (%.3f ms/iter)\n", who, iters, cudaGetErrorString(e), ms, ms/iters);
    write(sock,&c,1); read(sock,&c,1);
  }
  // legacy IPC check
  void* legacy; CK(cudaMalloc(&legacy, 1<<20)); cudaIpcMemHandle_t ih; cudaError_t e = cudaIpcGetMemHandle(&ih, legacy);
  printf("[%d] cudaIpcGetMemHandle: %s\n", who, cudaGetErrorString(e));
  write(sock, &ih, sizeof ih); cudaIpcMemHandle_t oh; read(sock, &oh, sizeof oh);
  void* op = nullptr; e = cudaIpcOpenMemHandle(&op, oh, cudaIpcMemLazyEnablePeerAccess);
  printf("[%d] cudaIpcOpenMemHandle: %s\n", who, cudaGetErrorString(e));
  write(sock,&c,1); read(sock,&c,1);
  return 0;
}

int main() {
  int sv[2]; socketpair(AF_UNIX, SOCK_STREAM, 0, sv);
  printf("nproc %ld\n", sysconf(_SC_NPROCESSORS_ONLN)); fflush(stdout);
  pid_t pid = fork();
  if (pid == 0) { close(sv[0]); int r = run_rank(1, sv[1]); fflush(stdout); _exit(r); }
  close(sv[1]); int r = run_rank(0, sv[0]); int st; waitpid(pid, &st, 0); printf("child status %d\n", st);
  return r;
}
