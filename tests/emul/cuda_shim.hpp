// TEST-ONLY host stand-in for the part of the CUDA runtime (and of cub::DeviceScan) that the fistr_b200 library uses.
// "Device memory" is host memory, streams execute immediately, events are wall-clock stamps, one "device" with two
// "SMs" (so that launch grids stay small).  Together with kernel_emul.hpp (kernels as fibers) and the launch-syntax
// rewrite of tools/build_hostemu.py it lets the WHOLE library — C-ABI, caches, option handling, launch sequences —
// be compiled by g++ into frontistr_b200/libfistr_b200_hostemu.so, which tests/test_gpu_suite_on_host_api.py loads in
// place of the product library to run the GPU test bodies on the CPU.  Never part of the product build.
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <type_traits>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorInvalidValue = 1, cudaErrorMemoryAllocation = 2, cudaErrorNotSupported = 801 };
typedef struct CUstream_st_emul *cudaStream_t;
struct cudaEventEmul { double t; };
typedef cudaEventEmul *cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaHostRegisterPortable = 1, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
struct cudaIpcMemHandle_t { char reserved[64]; };
struct cudaDeviceProp {
  char name[256];
  int multiProcessorCount;
  size_t totalGlobalMem;
};

inline const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : (e == cudaErrorMemoryAllocation ? "out of memory" : "not supported by the host emulation"); }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaDeviceReset() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
  memset(p, 0, sizeof *p);
  strcpy(p->name, "host emulation (TEST ONLY)");
  p->multiProcessorCount = 2;
  p->totalGlobalMem = (size_t)8 << 30;
  return cudaSuccess;
}
// "Device" allocations are electric-fence style: every buffer ENDS (up to 16-byte rounding) at an inaccessible guard
// page and begins right after another one, so a kernel that reads or writes out of bounds dies with SIGSEGV on the spot
// instead of silently touching a neighbour - every host run of a kernel is also a bounds check of its addressing.
// The memory is NOT zeroed, like device memory: it is filled with 0xFF bytes (NaN as a double, -1 as an int), so that a
// value that is used without having been written shows up in the results.
// The pages come from a memfd mapped MAP_SHARED, so that another PROCESS (= another emulated GPU) can map the same
// buffer: cudaIpcGetMemHandle hands out (pid, fd, length, offset of the user pointer) and cudaIpcOpenMemHandle maps
// /proc/<pid>/fd/<fd>.
struct CudaShimAlloc { void *base; size_t len; int fd; size_t user_off; size_t bytes; };
struct CudaShimAllocs {
  std::mutex mu;
  std::map<void *, CudaShimAlloc> live;      // user pointer -> mapping
  std::map<void *, size_t> peer;             // pointers returned by cudaIpcOpenMemHandle -> mapping length (base = ptr - off)
  std::map<void *, void *> peer_base;
  static CudaShimAllocs &get() { static CudaShimAllocs a; return a; }
};
struct CudaShimIpc { int pid, fd; unsigned long long len, off; };
// FB_EMUL_DEVICE_BYTES=<n>: the emulated device holds at most n bytes of live allocations; a cudaMalloc beyond that fails
// with cudaErrorMemoryAllocation (tests of the out-of-memory paths: scratch fall-backs, bench.py's workload fall-back)
inline size_t cuda_shim_capacity() {
  const char *e = getenv("FB_EMUL_DEVICE_BYTES");
  return e ? (size_t)strtoull(e, nullptr, 10) : ~(size_t)0;
}
inline size_t &cuda_shim_live_bytes() { static size_t b = 0; return b; }
inline cudaError_t cudaMalloc(void **p, size_t bytes) {
  {
    CudaShimAllocs &A = CudaShimAllocs::get();
    std::lock_guard<std::mutex> g(A.mu);
    if (cuda_shim_live_bytes() + bytes > cuda_shim_capacity()) { *p = nullptr; return cudaErrorMemoryAllocation; }
    cuda_shim_live_bytes() += bytes;
  }
  const size_t page = (size_t)sysconf(_SC_PAGESIZE);
  const size_t user = (std::max<size_t>(bytes, 1) + 15) & ~(size_t)15;
  const size_t body = (user + page - 1) / page * page;
  const size_t len = body + 2 * page;
  const int fd = (int)syscall(SYS_memfd_create, "fb_cuda_shim", 0u);
  if (fd < 0 || ftruncate(fd, (off_t)len) != 0) { if (fd >= 0) close(fd); *p = nullptr; return cudaErrorMemoryAllocation; }
  char *base = (char *)mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  if (base == (char *)MAP_FAILED) { close(fd); *p = nullptr; return cudaErrorMemoryAllocation; }
  memset(base + page, 0xFF, body);   // doubles read as NaN, ints as -1
  mprotect(base, page, PROT_NONE);
  mprotect(base + page + body, page, PROT_NONE);
  char *q = base + page + body - user;
  {
    CudaShimAllocs &A = CudaShimAllocs::get();
    std::lock_guard<std::mutex> g(A.mu);
    A.live[q] = CudaShimAlloc{(void *)base, len, fd, (size_t)(q - base), bytes};
  }
  *p = q;
  return cudaSuccess;
}
template <class T>
inline cudaError_t cudaMalloc(T **p, size_t bytes) { return cudaMalloc((void **)p, bytes); }
inline cudaError_t cudaFree(void *p) {
  if (!p) return cudaSuccess;
  CudaShimAllocs &A = CudaShimAllocs::get();
  std::lock_guard<std::mutex> g(A.mu);
  auto it = A.live.find(p);
  if (it == A.live.end()) { fprintf(stderr, "cuda_shim: cudaFree of an unknown pointer\n"); abort(); }
  munmap(it->second.base, it->second.len);
  close(it->second.fd);
  cuda_shim_live_bytes() -= it->second.bytes;
  A.live.erase(it);
  return cudaSuccess;
}
inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { if (n) memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemset(void *d, int v, size_t n) { if (n) memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = nullptr) { if (n) memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline double cuda_shim_now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new cudaEventEmul{0.0}; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) { e->t = cuda_shim_now_ms(); return cudaSuccess; }
inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return cudaSuccess; }
template <class F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return cudaSuccess; }
inline cudaError_t cudaHostRegister(void *, size_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaHostUnregister(void *) { return cudaSuccess; }
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) {
  CudaShimAllocs &A = CudaShimAllocs::get();
  std::lock_guard<std::mutex> g(A.mu);
  auto it = A.live.find(p);
  if (it == A.live.end()) return cudaErrorNotSupported;
  CudaShimIpc k{(int)getpid(), it->second.fd, (unsigned long long)it->second.len, (unsigned long long)it->second.user_off};
  memset(h, 0, sizeof *h);
  memcpy(h, &k, sizeof k);
  return cudaSuccess;
}
inline cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned) {
  // FB_EMUL_IPC_FAIL_RANK=<r> (test of the mapping-failure path): the process whose RANK is r cannot map its peers
  if (const char *fr = getenv("FB_EMUL_IPC_FAIL_RANK"))
    if (const char *me = getenv("RANK"))
      if (atoi(fr) == atoi(me)) return cudaErrorNotSupported;
  CudaShimIpc k;
  memcpy(&k, &h, sizeof k);
  char path[64];
  snprintf(path, sizeof path, "/proc/%d/fd/%d", k.pid, k.fd);
  const int fd = open(path, O_RDWR);
  if (fd < 0) return cudaErrorNotSupported;
  char *base = (char *)mmap(nullptr, (size_t)k.len, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (base == (char *)MAP_FAILED) return cudaErrorNotSupported;
  CudaShimAllocs &A = CudaShimAllocs::get();
  std::lock_guard<std::mutex> g(A.mu);
  A.peer[base + k.off] = (size_t)k.len;
  A.peer_base[base + k.off] = base;
  *p = base + k.off;
  return cudaSuccess;
}
inline cudaError_t cudaIpcCloseMemHandle(void *p) {
  CudaShimAllocs &A = CudaShimAllocs::get();
  std::lock_guard<std::mutex> g(A.mu);
  auto it = A.peer.find(p);
  if (it == A.peer.end()) return cudaSuccess;
  munmap(A.peer_base[p], it->second);
  A.peer_base.erase(p); A.peer.erase(it);
  return cudaSuccess;
}

namespace cub {
struct DeviceScan {
  // two-phase API as in CUB: d_temp == nullptr only reports the scratch size
  template <class In, class Out>
  static cudaError_t ExclusiveSum(void *d_temp, size_t &temp_bytes, In in, Out out, int n, cudaStream_t = nullptr) {
    if (!d_temp) { temp_bytes = 64; return cudaSuccess; }
    typename std::remove_reference<decltype(out[0])>::type run = 0;
    for (int i = 0; i < n; i++) { const auto v = in[i]; out[i] = run; run += v; }
    return cudaSuccess;
  }
};
struct DeviceRadixSort {
  // ascending on the bits [begin_bit, end_bit) of the key, stable - as CUB's
  template <class K>
  static cudaError_t SortKeys(void *d_temp, size_t &temp_bytes, const K *in, K *out, int n, int begin_bit = 0,
                              int end_bit = (int)sizeof(K) * 8, cudaStream_t = nullptr) {
    if (!d_temp) { temp_bytes = 64; return cudaSuccess; }
    const K mask = (end_bit - begin_bit >= (int)sizeof(K) * 8) ? ~(K)0 : (((K)1 << (end_bit - begin_bit)) - 1);
    for (int i = 0; i < n; i++) out[i] = in[i];
    std::stable_sort(out, out + n, [=](K a, K b) { return ((a >> begin_bit) & mask) < ((b >> begin_bit) & mask); });
    return cudaSuccess;
  }
};
}  // namespace cub
