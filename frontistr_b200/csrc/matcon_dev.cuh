// hecmw_mat_con (hecmw1/src/solver/matrix/hecmw_mat_con.f90:23-268) on the device: the node-node graph of the element
// connectivity as the reference's indexL / itemL / indexU / itemU (rows ascending, 1-based node ids, over all NP rows).
//
// Where the host version (matcon.hpp) walks a node -> element adjacency row by row, this one is three streaming passes
// that do not care how many elements meet in a node:
//   1. every ordered node pair (i, j), i != j, of every element becomes one 64-bit key  i << s | j, 2^s >= NP, so that only
//      2 s bits have to be sorted (matcon_keys_kernel),
//   2. the keys are radix-sorted (cub) - equal pairs become neighbours, rows and the columns inside a row come out
//      ascending, the lower-triangle entries of a row before its upper-triangle ones,
//   3. the first key of every run is an entry of the graph: two exclusive scans of the "first and lower" / "first and
//      upper" flags give its position in itemL / itemU, and a binary search per row gives indexL / indexU.
// Bytes: 8 per pair written, sorted (2 passes per radix digit) and read twice; H150 has 2.2e8 pairs (1.7 GB of keys).
#pragma once
#include "common.cuh"

namespace fb {

// keys of the elements [e0, e1) of one type (nn nodes each); `out` starts at this type's first pair.  A node id outside
// 1..NP raises *bad (the host version's "node id out of range").
__global__ void __launch_bounds__(256)
matcon_keys_kernel(int e0, int e1, int nn, int NP, int shift, const int *__restrict__ elem_node_index,
                   const int *__restrict__ elem_node_item, unsigned long long *__restrict__ out, int *bad) {
  const size_t per = (size_t)nn * (nn - 1), total = (size_t)(e1 - e0) * per;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int e = e0 + (int)(t / per);
    const int r = (int)(t % per), a = r / (nn - 1);
    int b = r % (nn - 1);
    b += (b >= a);
    const int *c = elem_node_item + elem_node_index[e];
    const int i = c[a], j = c[b];
    if (i < 1 || i > NP || j < 1 || j > NP) { atomicExch(bad, 1); out[t] = ~0ull; continue; }
    // a node named twice by one element (degenerate element) would give i == j: not a graph entry
    out[t] = (i == j) ? ~0ull : (((unsigned long long)(i - 1) << shift) | (unsigned)(j - 1));
  }
}

// flags of the sorted keys: first of its run and below / above the diagonal
__global__ void __launch_bounds__(256)
matcon_flags_kernel(size_t n, int shift, const unsigned long long *__restrict__ keys, unsigned *__restrict__ isL, unsigned *__restrict__ isU) {
  for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p <= n; p += (size_t)gridDim.x * blockDim.x) {
    unsigned l = 0, u = 0;
    if (p < n) {
      const unsigned long long k = keys[p];
      if (k != ~0ull && (p == 0 || keys[p - 1] != k)) {
        const unsigned i = (unsigned)(k >> shift), j = (unsigned)(k & ((1ull << shift) - 1));
        l = j < i; u = j > i;
      }
    }
    isL[p] = l; isU[p] = u;   // entry n: the scans' grand totals land there
  }
}

// itemL / itemU from the scanned flags
__global__ void __launch_bounds__(256)
matcon_items_kernel(size_t n, int shift, const unsigned long long *__restrict__ keys, const unsigned *__restrict__ posL,
                    const unsigned *__restrict__ posU, int *__restrict__ itemL, int *__restrict__ itemU) {
  for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x) {
    const unsigned long long k = keys[p];
    if (k == ~0ull || (p > 0 && keys[p - 1] == k)) continue;
    const unsigned i = (unsigned)(k >> shift), j = (unsigned)(k & ((1ull << shift) - 1));
    if (j < i) itemL[posL[p]] = (int)j + 1;
    else itemU[posU[p]] = (int)j + 1;
  }
}

// indexL(0:NP), indexU(0:NP): entries in front of row i = scanned flags at the first key of a row >= i
__global__ void __launch_bounds__(256)
matcon_index_kernel(int NP, size_t n, int shift, const unsigned long long *__restrict__ keys, const unsigned *__restrict__ posL,
                    const unsigned *__restrict__ posU, int *__restrict__ indexL, int *__restrict__ indexU) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i <= (size_t)NP; i += (size_t)gridDim.x * blockDim.x) {
    const unsigned long long want = (unsigned long long)i << shift;
    size_t lo = 0, hi = n;
    while (lo < hi) {
      const size_t mid = lo + (hi - lo) / 2;
      if (keys[mid] < want) lo = mid + 1; else hi = mid;
    }
    indexL[i] = (int)posL[lo];
    indexU[i] = (int)posU[lo];
  }
}

}  // namespace fb
