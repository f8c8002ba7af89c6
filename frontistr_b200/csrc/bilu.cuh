// PRECOND = 10: block ILU(0) of hecmw_precond_BILU_33 (precond/33/hecmw_precond_BILU_33.f90) on the device.
//
// The reference factorises row after row (FORM_ILU0_33, :183-362) and solves row after row (apply, :91-158).  Both
// recurrences only reach back through the L pattern (forward) / forward through the U pattern (backward), so rows whose
// dependencies are complete form LEVELS that can run side by side: the host sorts the rows by level once per matrix
// graph (bilu_levels below), and every level is one launch with one thread per row.  Inside a row the operations keep the
// reference's order (k ascending over L(i), j ascending over U(k); forward sum ascending, backward sum descending; and,
// as in the reference, the pivot blocks never receive an update - see the guard in the factor kernel) with
// the products and sums written as explicit fma chains - the same chains as the oracle's restatement; what is left to the
// compilers is whether "x -= a * t" after a block solve becomes one fma (nvcc and gcc both do it; the host-compiled test
// build does not), a last-bit matter.
//
// Layout: the factor lives in one array shaped like the matrix values, [Dlu | ALlu | AUlu] (9 doubles per block, row-major
// as in hecMAT), and uses the matrix's own index arrays; Dlu holds the 3x3 LU of the pivot block with reciprocal pivots
// (ILU1a33, :1492-1530).  The path is launch-bound on meshes with many levels (a structured hexahedral block has about
// 7 n of them, n^2 / 7 rows each on average): runs of small levels share one single-block launch (bilu_plan, option
// "bilu_chain_rows", default 256; 0 = one launch per level).  Not measured on hardware yet.
#pragma once
#include <algorithm>
#include <vector>

#include "common.cuh"

namespace fb {

struct Bilu {
  bool structure_valid = false, values_valid = false;
  int N = 0, NP = 0;
  std::vector<int> fac_ptr, fwd_ptr, bwd_ptr;   // level offsets into the three row lists
  DevBuf<int> fac_rows, fwd_rows, bwd_rows;     // 0-based rows sorted by level (ascending row inside a level)
  DevBuf<int> d_fac_ptr, d_fwd_ptr, d_bwd_ptr;  // the level offsets on the device (for the chained launches)
  // launch plan: runs of consecutive small levels go to ONE single-block launch that walks them with __syncthreads()
  // in between; {first level, one past the last level, chained?}
  struct Seg { int l0, l1; bool chain; };
  std::vector<Seg> fac_plan, fwd_plan, bwd_plan;
  DevBuf<double> lu;                            // [9 NP | 9 NPL | 9 NPU]
  DevBuf<int> err;                              // [0] = 1-based row of a zero pivot (0: none)
};

// host: rows -> levels.  Factorisation (all NP rows) and forward solve (rows < N) share level(i) = 1 + max level(k),
// k in L(i); the backward solve uses level(i) = 1 + max level(k), k in U(i), k < N (halo rows of the vector are zero and
// never written).
inline void bilu_levels(int N, int NP, const int *indexL, const int *itemL, const int *indexU, const int *itemU,
                        std::vector<int> &fac_rows, std::vector<int> &fac_ptr, std::vector<int> &fwd_rows,
                        std::vector<int> &fwd_ptr, std::vector<int> &bwd_rows, std::vector<int> &bwd_ptr) {
  auto bucket = [](const std::vector<int> &lev, int n, std::vector<int> &rows, std::vector<int> &ptr) {
    int nlev = 0;
    for (int i = 0; i < n; i++) nlev = std::max(nlev, lev[i]);
    ptr.assign((size_t)nlev + 1, 0);
    for (int i = 0; i < n; i++) ptr[lev[i]]++;
    for (int l = 1; l <= nlev; l++) ptr[l] += ptr[l - 1];
    rows.resize((size_t)n);
    std::vector<int> cur(ptr.begin(), ptr.end() - (nlev > 0 ? 1 : 0));
    for (int i = 0; i < n; i++) rows[cur[lev[i] - 1]++] = i;
  };
  std::vector<int> lev((size_t)std::max(NP, 1), 0);
  for (int i = 0; i < NP; i++) {
    int l = 0;
    for (int k = indexL[i]; k < indexL[i + 1]; k++) l = std::max(l, lev[itemL[k] - 1]);
    lev[i] = l + 1;
  }
  bucket(lev, NP, fac_rows, fac_ptr);
  bucket(lev, N, fwd_rows, fwd_ptr);
  for (int i = N - 1; i >= 0; i--) {
    int l = 0;
    for (int k = indexU[i]; k < indexU[i + 1]; k++)
      if (itemU[k] - 1 < N) l = std::max(l, lev[itemU[k] - 1]);
    lev[i] = l + 1;
  }
  bucket(lev, N, bwd_rows, bwd_ptr);
}

// A level with few rows costs a launch for a microsecond of work.  Consecutive levels of at most `chain_rows` rows are
// therefore handed to one block (kBiluChainThreads threads) that runs them one after the other with a block barrier in
// between - no inter-block synchronisation anywhere; larger levels keep one grid-wide launch each.
constexpr int kBiluChainThreads = 256;
inline void bilu_plan(const std::vector<int> &ptr, int chain_rows, std::vector<Bilu::Seg> &plan) {
  plan.clear();
  const int nlev = (int)ptr.size() - 1;
  for (int l = 0; l < nlev;) {
    if (ptr[l + 1] - ptr[l] <= chain_rows) {
      int e = l + 1;
      while (e < nlev && ptr[e + 1] - ptr[e] <= chain_rows) e++;
      plan.push_back({l, e, e - l > 1});   // a lone small level gains nothing from the chained kernel
      l = e;
    } else {
      plan.push_back({l, l + 1, false});
      l++;
    }
  }
}

// forward / back substitution with the 3x3 LU of a pivot block (the tail of the reference's apply, :118-125)
__device__ __forceinline__ void bilu_solve3(const double *__restrict__ a, double &X1, double &X2, double &X3) {
  X2 = fma(-a[3], X1, X2);
  X3 = fma(-a[7], X2, fma(-a[6], X1, X3));
  X3 = a[8] * X3;
  X2 = a[4] * fma(-a[5], X3, X2);
  X1 = a[0] * fma(-a[1], X2, fma(-a[2], X3, X1));
}

// ILU1b33 (:1539-1596) and the subtraction that follows it: Aij -= Aik * (Dk^-1 Akj)
__device__ __forceinline__ void bilu_update_block(const double *Dk, const double *Aik, const double *Akj, double *Aij) {
#pragma unroll
  for (int c = 0; c < 3; c++) {
    double X1 = Akj[c], X2 = Akj[3 + c], X3 = Akj[6 + c];
    bilu_solve3(Dk, X1, X2, X3);
#pragma unroll
    for (int r = 0; r < 3; r++) Aij[3 * r + c] -= fma(Aik[3 * r + 2], X3, fma(Aik[3 * r + 1], X2, Aik[3 * r] * X1));
  }
}

// the factor starts as a copy of the matrix with SIGMA_DIAG on the main diagonal (:208-243)
__global__ void __launch_bounds__(256)
bilu_init_kernel(size_t total, size_t nd, const double *__restrict__ vals, double sigma_diag, double *__restrict__ lu) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    double v = vals[i];
    if (i < nd && (i % 9) % 4 == 0) v *= sigma_diag;
    lu[i] = v;
  }
}

// FORM_ILU0_33 for one row: row i takes the updates of every k in L(i) (complete: lower levels), then its pivot block is
// factorised (ILU1a33).  A row writes only its own blocks.
__device__ __forceinline__ void bilu_factor_row(int i, int NP, int NPL, const int *__restrict__ indexL, const int *__restrict__ itemL,
                                                const int *__restrict__ indexU, const int *__restrict__ itemU, double *lu, int *err) {
  double *Dlu = lu, *ALlu = lu + 9 * (size_t)NP, *AUlu = lu + 9 * ((size_t)NP + NPL);
  const int l0 = indexL[i], l1 = indexL[i + 1], u0 = indexU[i], u1 = indexU[i + 1];
  double *Di = Dlu + 9 * (size_t)i;
  for (int kk = l0; kk < l1; kk++) {
    const int k = itemL[kk] - 1;
    const double *Dk = Dlu + 9 * (size_t)k;
    const double *Aik = ALlu + 9 * (size_t)kk;   // as updated so far by the smaller k of this row
    for (int jj = indexU[k]; jj < indexU[k + 1]; jj++) {
      const int j = itemU[jj] - 1;
      double *Aij = nullptr;
      if (j == i) continue;   // the reference's guard (:291) drops j == i with every other j outside L(i), U(i): pivot blocks get no update
      if (j < i) {
        for (int q = l0; q < l1; q++)
          if (itemL[q] - 1 == j) { Aij = ALlu + 9 * (size_t)q; break; }
      } else {
        for (int q = u0; q < u1; q++)
          if (itemU[q] - 1 == j) { Aij = AUlu + 9 * (size_t)q; break; }
      }
      if (Aij) bilu_update_block(Dk, Aik, AUlu + 9 * (size_t)jj, Aij);
    }
  }
  double a[9];
#pragma unroll
  for (int q = 0; q < 9; q++) a[q] = Di[q];
  bool bad = false;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    if (a[4 * k] == 0.0) bad = true;
    a[4 * k] = 1.0 / a[4 * k];
#pragma unroll
    for (int ii = k + 1; ii < 3; ii++) {
      a[3 * ii + k] = a[3 * ii + k] * a[4 * k];
#pragma unroll
      for (int j = k + 1; j < 3; j++) a[3 * ii + j] = fma(-a[3 * ii + k], a[3 * k + j], a[3 * ii + j]);
    }
  }
#pragma unroll
  for (int q = 0; q < 9; q++) Di[q] = a[q];
  if (bad) atomicMax(err, i + 1);
}

// one (large) level: rows [r0, r1) of the level-sorted list, one thread per row
__global__ void __launch_bounds__(128)
bilu_factor_level_kernel(int r0, int r1, const int *__restrict__ rows, int NP, int NPL, const int *__restrict__ indexL,
                         const int *__restrict__ itemL, const int *__restrict__ indexU, const int *__restrict__ itemU,
                         double *lu, int *err) {
  for (int r = r0 + blockIdx.x * blockDim.x + threadIdx.x; r < r1; r += gridDim.x * blockDim.x)
    bilu_factor_row(rows[r], NP, NPL, indexL, itemL, indexU, itemU, lu, err);
}

// a run of small levels [l0, l1) in ONE block: a block barrier separates the levels
__global__ void __launch_bounds__(kBiluChainThreads)
bilu_factor_chain_kernel(int l0, int l1, const int *__restrict__ ptr, const int *__restrict__ rows, int NP, int NPL,
                         const int *__restrict__ indexL, const int *__restrict__ itemL, const int *__restrict__ indexU,
                         const int *__restrict__ itemU, double *lu, int *err) {
  for (int l = l0; l < l1; l++) {
    for (int r = ptr[l] + threadIdx.x; r < ptr[l + 1]; r += blockDim.x)
      bilu_factor_row(rows[r], NP, NPL, indexL, itemL, indexU, itemU, lu, err);
    __syncthreads();
  }
}

// one row of the forward (FWD, :96-126) or backward (:131-157) substitution, in place on zp
template <bool FWD>
__device__ __forceinline__ void bilu_sweep_row(size_t i, const int *__restrict__ index, const int *__restrict__ item,
                                               const double *__restrict__ blocks, const double *__restrict__ Dlu, double *zp) {
  const int j0 = index[i], j1 = index[i + 1];
  double s1, s2, s3;
  if (FWD) {
    s1 = zp[3 * i]; s2 = zp[3 * i + 1]; s3 = zp[3 * i + 2];
    for (int j = j0; j < j1; j++) {
      const size_t k = (size_t)(item[j] - 1);
      const double *a = blocks + 9 * (size_t)j;
      const double x1 = zp[3 * k], x2 = zp[3 * k + 1], x3 = zp[3 * k + 2];
      s1 = fma(-a[2], x3, fma(-a[1], x2, fma(-a[0], x1, s1)));
      s2 = fma(-a[5], x3, fma(-a[4], x2, fma(-a[3], x1, s2)));
      s3 = fma(-a[8], x3, fma(-a[7], x2, fma(-a[6], x1, s3)));
    }
  } else {
    s1 = s2 = s3 = 0.0;
    for (int j = j1 - 1; j >= j0; j--) {
      const size_t k = (size_t)(item[j] - 1);
      const double *a = blocks + 9 * (size_t)j;
      const double x1 = zp[3 * k], x2 = zp[3 * k + 1], x3 = zp[3 * k + 2];
      s1 = fma(a[2], x3, fma(a[1], x2, fma(a[0], x1, s1)));
      s2 = fma(a[5], x3, fma(a[4], x2, fma(a[3], x1, s2)));
      s3 = fma(a[8], x3, fma(a[7], x2, fma(a[6], x1, s3)));
    }
  }
  bilu_solve3(Dlu + 9 * i, s1, s2, s3);
  if (FWD) { zp[3 * i] = s1; zp[3 * i + 1] = s2; zp[3 * i + 2] = s3; }
  else { zp[3 * i] -= s1; zp[3 * i + 1] -= s2; zp[3 * i + 2] -= s3; }
}

// one (large) level of a sweep
template <bool FWD>
__global__ void __launch_bounds__(128)
bilu_sweep_level_kernel(int r0, int r1, const int *__restrict__ rows, const int *__restrict__ index,
                        const int *__restrict__ item, const double *__restrict__ blocks, const double *__restrict__ Dlu,
                        double *zp, const int *phase) {
  if (phase && *phase != 0) return;
  for (int r = r0 + blockIdx.x * blockDim.x + threadIdx.x; r < r1; r += gridDim.x * blockDim.x)
    bilu_sweep_row<FWD>((size_t)rows[r], index, item, blocks, Dlu, zp);
}

// a run of small levels [l0, l1) of a sweep in ONE block (zp is read and written through plain pointers: what one level
// stores is what the next one loads after the block barrier)
template <bool FWD>
__global__ void __launch_bounds__(kBiluChainThreads)
bilu_sweep_chain_kernel(int l0, int l1, const int *__restrict__ ptr, const int *__restrict__ rows,
                        const int *__restrict__ index, const int *__restrict__ item, const double *__restrict__ blocks,
                        const double *__restrict__ Dlu, double *zp, const int *phase) {
  if (phase && *phase != 0) return;
  for (int l = l0; l < l1; l++) {
    for (int r = ptr[l] + threadIdx.x; r < ptr[l + 1]; r += blockDim.x)
      bilu_sweep_row<FWD>((size_t)rows[r], index, item, blocks, Dlu, zp);
    __syncthreads();
  }
}

}  // namespace fb
