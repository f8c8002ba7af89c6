// !SOLVER SCALING=YES: symmetric diagonal scaling around the Krylov loop
// (hecmw1/src/solver/las/hecmw_solver_scaling_33.f90: hecmw_solver_scaling_fw_33 :20-124, _bk_33 :126-206; called from
// hecmw_solve_CG, iterative/hecmw_solver_CG.f90:99 and :272).
//   fw: s(3i+c) = 1 / sqrt(|D_i(c,c)|) for the N owned nodes, halo-exchanged (hecmw_update_3_R) so that the columns of
//       external nodes are known; every matrix entry a(ip, iq) <- a * s(ip) * s(iq) (two multiplications, left to right,
//       as the reference writes them); B <- B * s on the owned rows.
//   bk: X <- X * s, B <- B / s on the owned rows; every entry a <- a / (s(ip) * s(iq)).
// The matrix is scaled IN PLACE in the [D | AL | AU] allocation, like the reference scales hecMAT%D/AL/AU in place; the
// round trip leaves every entry within an ulp or two of what it was (the reference has the same property).
#pragma once
#include "common.cuh"

namespace fb {

__global__ void __launch_bounds__(256)
scaling_vector_kernel(int N, const double *__restrict__ D, double *__restrict__ scale) {
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < 3 * (size_t)N; t += (size_t)gridDim.x * blockDim.x) {
    const size_t i = t / 3;
    const int c = (int)(t - 3 * i);
    scale[t] = 1.0 / sqrt(fabs(D[9 * i + 4 * c]));
  }
}

// One warp per block row i < NP: the diagonal block and the row's L and U blocks are contiguous in memory, so the lanes
// stride over their scalar entries (coalesced).  BACK = false: a * s_ip * s_iq; BACK = true: a / (s_ip * s_iq).
template <bool BACK>
__global__ void __launch_bounds__(256)
scaling_matrix_kernel(int NP, const int *__restrict__ indexL, const int *__restrict__ itemL,
                      const int *__restrict__ indexU, const int *__restrict__ itemU, const double *__restrict__ scale,
                      double *__restrict__ D, double *__restrict__ AL, double *__restrict__ AU) {
  const int lane = threadIdx.x & 31;
  const size_t wg = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t nw = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t i = wg; i < (size_t)NP; i += nw) {
    const double *si = scale + 3 * i;
    if (lane < 9) {
      const double sp = si[lane / 3], sq = si[lane % 3];
      double &a = D[9 * i + lane];
      a = BACK ? a / (sp * sq) : a * sp * sq;
    }
    for (int part = 0; part < 2; part++) {
      const int *index = part ? indexU : indexL;
      const int *item = part ? itemU : itemL;
      double *A = part ? AU : AL;
      const size_t k0 = (size_t)index[i], k1 = (size_t)index[i + 1];
      for (size_t t = 9 * k0 + lane; t < 9 * k1; t += 32) {
        const size_t k = t / 9;
        const int e = (int)(t - 9 * k);
        const size_t inod = (size_t)(item[k] - 1);
        const double sp = si[e / 3], sq = scale[3 * inod + e % 3];
        A[t] = BACK ? A[t] / (sp * sq) : A[t] * sp * sq;
      }
    }
  }
}

// fw: B <- B * s (owned rows).  bk: X <- X * s, B <- B / s (owned rows)
template <bool BACK>
__global__ void __launch_bounds__(256)
scaling_rhs_kernel(size_t n3, const double *__restrict__ scale, double *__restrict__ B, double *__restrict__ X) {
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n3; t += (size_t)gridDim.x * blockDim.x) {
    if (BACK) { X[t] = X[t] * scale[t]; B[t] = B[t] / scale[t]; }
    else B[t] = B[t] * scale[t];
  }
}

}  // namespace fb
