// Common device/host helpers for the fistr_b200 kernels (sm_100a only).
#pragma once
#ifdef FB_EMUL_KERNELS
#include "kernel_emul.hpp"   // TEST-ONLY host stand-in for the CUDA execution model (libfistr_b200_emul.so)
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace fb {

#ifndef FB_EMUL_KERNELS
// dynamic shared memory of a kernel (the TEST-ONLY host stand-in kernel_emul.hpp maps it to a host buffer)
#define FB_DYNAMIC_SMEM(name) extern __shared__ __align__(16) unsigned char name[]

constexpr int kWarp = 32;
#endif
constexpr int kNumSM = 148;  // B200: 2 dies x 74 SMs (queried at init; this is the sizing default)

extern thread_local std::string g_last_error;
extern int64_t g_launch_count;

#define FB_CUDA(expr)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      fb::g_last_error = std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" +        \
                         __FILE__ + ":" + std::to_string(__LINE__) + ")";                  \
      cudaGetLastError(); /* reported here; do not let a later launch check see it again */ \
      return FISTR_B200_ERR_CUDA;                                                          \
    }                                                                                      \
  } while (0)

#define FB_LAUNCH_CHECK()                                                                  \
  do {                                                                                     \
    ++fb::g_launch_count;                                                                  \
    cudaError_t _e = cudaGetLastError(); /* returns AND resets a non-sticky error */        \
    if (_e != cudaSuccess) {                                                               \
      fb::g_last_error = std::string("kernel launch: ") + cudaGetErrorString(_e) + " (" +  \
                         __FILE__ + ":" + std::to_string(__LINE__) + ")";                  \
      return FISTR_B200_ERR_CUDA;                                                          \
    }                                                                                      \
  } while (0)

#ifndef FB_EMUL_KERNELS   // (host versions in kernel_emul.hpp)
// ---- streaming loads: values/indices are touched once per SpMV, keep them out of L1 so the
// gathered x vector owns it.
__device__ __forceinline__ double ld_stream(const double *p) {
  double v;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ double2 ld_stream2(const double2 *p) {
  double2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_stream(const int *p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int2 ld_stream2(const int2 *p) {
  int2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ int4 ld_stream4(const int4 *p) {
  int4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
  return v;
}
#endif

// ---- reductions -------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic grid reduction of NV doubles per thread.  Every block writes its partial to
// partial[blockIdx.x*NV + k]; the last block to arrive (atomic ticket) sums the partials in a
// fixed order and calls `fin(sums)` from thread 0.  counter must be zero on entry; it is reset.
// WARP_FIN = true: `fin(sums)` is called by all 32 lanes of warp 0 of the last block instead (the
// peer-memory publication uses one lane per destination rank).
template <int NV, int BLOCK, bool WARP_FIN = false, class Fin>
__device__ __forceinline__ void grid_reduce(double (&v)[NV], double *partial, unsigned int *counter, Fin fin) {
  __shared__ double sm[NV][BLOCK / kWarp];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; k++) {
    double s = warp_sum(v[k]);
    if (lane == 0) sm[k][w] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < NV; k++) {
      double s = 0.0;
      for (int i = 0; i < BLOCK / kWarp; i++) s += sm[k][i];
      partial[(size_t)blockIdx.x * NV + k] = s;
    }
    __threadfence();
    unsigned int t = atomicAdd(counter, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    double acc[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) acc[k] = 0.0;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += BLOCK) {
#pragma unroll
      for (int k = 0; k < NV; k++) acc[k] += __ldcg(&partial[(size_t)b * NV + k]);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; k++) {
      double s = warp_sum(acc[k]);
      if (lane == 0) sm[k][w] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double out[NV];
#pragma unroll
      for (int k = 0; k < NV; k++) {
        double s = 0.0;
        for (int i = 0; i < BLOCK / kWarp; i++) s += sm[k][i];
        out[k] = s;
      }
      *counter = 0u;
      if (!WARP_FIN) fin(out);
      else {
#pragma unroll
        for (int k = 0; k < NV; k++) sm[k][0] = out[k];
      }
    }
    if (WARP_FIN) {
      __syncthreads();
      if (threadIdx.x < 32) {
        double out[NV];
#pragma unroll
        for (int k = 0; k < NV; k++) out[k] = sm[k][0];
        fin(out);
      }
    }
  }
}

// Owning device buffer: frees its memory when it goes out of scope (early error returns included) or is re-assigned;
// move-only.  view() makes a non-owning alias (memory that lives in the peer-mapped arena).
template <class T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  bool owned = true;
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n), owned(o.owned) { o.p = nullptr; o.n = 0; o.owned = true; }
  DevBuf &operator=(DevBuf &&o) noexcept {
    if (this != &o) {
      release();
      p = o.p; n = o.n; owned = o.owned;
      o.p = nullptr; o.n = 0; o.owned = true;
    }
    return *this;
  }
  ~DevBuf() { release(); }
  void view(T *ptr, size_t count) {   // non-owning alias (memory lives in the peer-mapped arena)
    release();
    p = ptr; n = count; owned = false;
  }
  int alloc(size_t count) {
    if (count <= n && p) return 0;
    release();
    if (count == 0) count = 1;
    cudaError_t e = cudaMalloc((void **)&p, count * sizeof(T));
    if (e != cudaSuccess) {
      g_last_error = std::string("cudaMalloc(") + std::to_string(count * sizeof(T)) + "): " + cudaGetErrorString(e);
      cudaGetLastError();   // reported through the return value: a later launch check must not see it again
      p = nullptr;
      n = 0;
      return -1;
    }
    n = count;
    return 0;
  }
  void release() {
    if (p && owned) cudaFree(p);
    p = nullptr;
    n = 0;
    owned = true;
  }
};

// CUDA events that are destroyed on every exit path of the function that made them
struct EventSet {
  static const int kMax = 24;
  cudaEvent_t ev[kMax];
  int n = 0;
  EventSet() = default;
  EventSet(const EventSet &) = delete;
  EventSet &operator=(const EventSet &) = delete;
  cudaError_t make(cudaEvent_t *out) {
    if (n >= kMax) return cudaErrorInvalidValue;
    cudaError_t e = cudaEventCreate(out);
    if (e == cudaSuccess) ev[n++] = *out;
    return e;
  }
  ~EventSet() { for (int i = 0; i < n; i++) cudaEventDestroy(ev[i]); }
};

}  // namespace fb
