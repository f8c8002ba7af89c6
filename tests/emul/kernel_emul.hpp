// TEST-ONLY stand-in for the CUDA execution model, used when assembly.cuh is compiled by g++ into
// libfistr_b200_emul.so (FB_EMUL_KERNELS).  Two launchers:
//   emul_launch       : grid-stride kernels without intra-block cooperation (slot / gather maps, node incidence, nodal
//                       stress ...) run UNCHANGED, one "thread" after the other in a scrambled order;
//   emul_launch_warps : the one-element-per-warp kernels (stiff / update / self-test) run UNCHANGED, every warp by 32
//                       fibers in lock step (__syncwarp / shuffles = switch points, kernel_emul_impl.hpp), warps one after
//                       the other, dynamic shared memory = a host buffer per block
//   emul_launch_blocks: block-cooperative kernels (SpMV, CG, grid reductions; __syncthreads, static __shared__, warp
//                       shuffles / ballots) run UNCHANGED: the threads of a block are cooperative fibers of ONE host
//                       thread that switch at every barrier / collective, blocks one after the other.  Several host
//                       threads can each run their own launches concurrently = several "GPUs" (ranks), which is how
//                       the peer-memory protocol of p2p.cuh is exercised with 2 / 4 / 8 ranks;
// so that the kernels' own index arithmetic, staging, store patterns and inter-rank protocol are exercised by the CPU
// unit tests without a GPU.  The product library never sees this header.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <algorithm>
#include <numeric>
#include <random>
#include <vector>

#include <type_traits>

#include "cuda_shim.hpp"   // host stand-in of the CUDA runtime subset the library uses

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__
#define __grid_constant__
#define __shared__ static thread_local   /* one block at a time per host thread (= per emulated GPU) */

struct int2 { int x, y; };
struct int4 { int x, y, z, w; };

// CUDA's 16-byte vector of two doubles.  Every copy checks that both addresses are 16-byte aligned (a misaligned
// double2 access is a hardware fault on the GPU: "misaligned address"); violations are counted, never ignored.
inline long &emul_misaligned_count() { static long n = 0; return n; }
struct alignas(16) double2 {
  double x, y;
  double2() = default;
  double2(const double2 &o) { check(&o); x = o.x; y = o.y; }
  double2 &operator=(const double2 &o) { check(this); check(&o); x = o.x; y = o.y; return *this; }
  static void check(const void *p) { if (reinterpret_cast<uintptr_t>(p) % 16 != 0) __atomic_fetch_add(&emul_misaligned_count(), 1L, __ATOMIC_RELAXED); }
};

namespace fb {
// Every warp collective carries the call site it was issued from (file / line hash): when the lanes of a warp are released
// the scheduler checks that all of them arrived at the SAME shuffle (a __shfl_sync whose participants sit at different
// instructions, or some at a shuffle and others at a __syncwarp, exchanges garbage or hangs on hardware - in an emulator that
// merely counts arrivals it would pass unnoticed).  __syncwarp may legitimately be reached from different places.
constexpr unsigned long long emul_site(const char *file, int line) {
  unsigned long long h = 1469598103934665603ull;
  for (const char *c = file; *c; ++c) h = (h ^ (unsigned char)*c) * 1099511628211ull;
  return ((h ^ (unsigned long long)line * 0x9E3779B97F4A7C15ull) << 2) >> 2;   // two low bits are the kind of collective
}
#define FB_EMUL_SITE fb::emul_site(__FILE__, __LINE__)
namespace emulwarp {   // warp / block collectives, implemented in kernel_emul_impl.hpp (fibers)
double shfl(double v, int src, unsigned long long site = 0);
int shfl(int v, int src, unsigned long long site = 0);
void syncwarp(unsigned long long site = 0);
void syncthreads();
}
constexpr int kWarp = 32;
struct EmulDim { unsigned x = 1, y = 1, z = 1; };
inline thread_local EmulDim threadIdx, blockIdx, blockDim, gridDim;

// real atomics: the lanes of an emulated warp are concurrent host threads
inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline double atomicAdd(double *p, double v) {
  uint64_t *q = reinterpret_cast<uint64_t *>(p), o = __atomic_load_n(q, __ATOMIC_RELAXED), n;
  double od, nd;
  do { __builtin_memcpy(&od, &o, 8); nd = od + v; __builtin_memcpy(&n, &nd, 8); }
  while (!__atomic_compare_exchange_n(q, &o, n, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
  return od;
}
inline int atomicMax(int *p, int v) {
  int o = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (o < v && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return o;
}
inline int atomicExch(int *p, int v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
inline unsigned char *&emul_smem() { static unsigned char *p = nullptr; return p; }   // dynamic shared memory of the running block
#define FB_DYNAMIC_SMEM(name) unsigned char *name = fb::emul_smem()
inline void __syncthreads() { emulwarp::syncthreads(); }
inline unsigned emul_ballot(bool pred, unsigned long long site) {
  unsigned bits = 0;
  for (int l = 0; l < 32; l++) bits |= (unsigned)(emulwarp::shfl(pred ? 1 : 0, l, site) & 1) << l;
  return bits;
}
// (macros, so that the call site is the kernel's own line; the product uses full masks only - checked by the static_assert-like
// runtime test in emul_check_mask)
inline void emul_check_mask(unsigned m) { if (m != 0xffffffffu) { fprintf(stderr, "emul: warp collective with partial mask 0x%08x is not modelled\n", m); abort(); } }
#define __syncwarp() fb::emulwarp::syncwarp(FB_EMUL_SITE)
#define __shfl_xor_sync(m, v, o) (fb::emul_check_mask(m), fb::emulwarp::shfl((v), (int)((fb::threadIdx.x & 31) ^ (unsigned)(o)), FB_EMUL_SITE))
#define __shfl_sync(m, v, s) (fb::emul_check_mask(m), fb::emulwarp::shfl((v), (s), FB_EMUL_SITE))
#define __ballot_sync(m, p) (fb::emul_check_mask(m), fb::emul_ballot((p), FB_EMUL_SITE))
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline double __ldcg(const double *p) { return *(const volatile double *)p; }
inline double ld_stream(const double *p) { return *p; }
inline int ld_stream(const int *p) { return *p; }
inline double2 ld_stream2(const double2 *p) { return *p; }
inline int2 ld_stream2(const int2 *p) { return *p; }
inline int4 ld_stream4(const int4 *p) { return *p; }

// kernel<<<grid, block>>>(args...): every (block, thread) pair runs to completion, in a seeded random order
template <class K, class... A>
inline void emul_launch(unsigned grid, unsigned block, unsigned seed, K kernel, A... args) {
  std::vector<unsigned> order((size_t)grid * block);
  std::iota(order.begin(), order.end(), 0u);
  std::mt19937 rng(seed);
  std::shuffle(order.begin(), order.end(), rng);
  gridDim.x = grid; blockDim.x = block;
  for (unsigned id : order) {
    blockIdx.x = id / block; threadIdx.x = id % block;
    kernel(args...);
  }
}
}  // namespace fb
