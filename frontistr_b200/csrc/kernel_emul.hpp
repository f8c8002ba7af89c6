// TEST-ONLY stand-in for the CUDA execution model, used when assembly.cuh is compiled by g++ into
// libfistr_b200_emul.so (FB_EMUL_KERNELS): lets the grid-stride kernels that have no intra-block cooperation
// (slot / gather maps, node incidence, nodal stress ...) run UNCHANGED on the host, one "thread" after the other in
// a scrambled order, so their index arithmetic is exercised by the CPU unit tests without a GPU.  The product
// library never sees this header.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <algorithm>
#include <numeric>
#include <random>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__

namespace fb {
constexpr int kWarp = 32;
struct EmulDim { unsigned x = 1, y = 1, z = 1; };
inline thread_local EmulDim threadIdx, blockIdx, blockDim, gridDim;

inline unsigned atomicAdd(unsigned *p, unsigned v) { const unsigned o = *p; *p = o + v; return o; }
inline int atomicAdd(int *p, int v) { const int o = *p; *p = o + v; return o; }
inline double atomicAdd(double *p, double v) { const double o = *p; *p = o + v; return o; }
inline int atomicExch(int *p, int v) { const int o = *p; *p = v; return o; }
inline double ld_stream(const double *p) { return *p; }
inline int ld_stream(const int *p) { return *p; }

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
