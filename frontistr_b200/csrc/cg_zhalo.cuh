// "z-halo" schedule of the collective (NCCL) transport: two communication steps per CG iteration instead of three.
//
// hecmw_solve_CG (hecmw1/src/solver/iterative/hecmw_solver_CG.f90:153-271) communicates three times per iteration: the
// halo update of p inside hecmw_matvec (hecmw_update_3_R, las_33.f90:255-262), the all-reduce of p.q (:211) and the
// all-reduces of r.z / r.r (:168,240; fused into one here).  The halo update of p can only start once beta = rho/rho1
// is known, i.e. AFTER the (r.r, r.z) all-reduce: two dependent communication steps back to back.
//
// p = z + beta p holds row by row, so a rank can update the halo rows of p itself if it has the halo rows of z - and z
// does not depend on beta.  The schedule therefore sends the export rows of z TOGETHER with the rank's partial sums of
// (r.r, r.z) in one grouped send/recv step: every rank receives its neighbours' z rows and every rank's partials, sums the
// partials in rank order (the same order on every rank: bit-identical scalars everywhere, no all-reduce kernel) and runs
// p = z + beta p over all NP rows.  Per iteration: {z halo + partials} and the p.q all-reduce - the minimum for a CG that
// keeps the reference's recurrences (two dependent reductions per iteration).  The halo rows of p are the same FMA of the
// same operands as on their owner, so they are bit-identical to what hecmw_update_3_R would have delivered.
//
// Only the fused DIAG iteration uses it (api_solver.inc, option "comm_zhalo"); the peer-memory transport (p2p.cuh) needs
// no NCCL call at all and does not.
#pragma once
#include "cg_kernels.cuh"

namespace fb {

// (r.r, r.z) of all ranks = sum of the gathered partials in rank order; tab = [nranks][2] of the parity this iteration reads
__device__ __forceinline__ void cg_sum_partials(const double *tab, int nranks, double &rr, double &rz) {
  double s0 = 0.0, s1 = 0.0;
  for (int r = 0; r < nranks; r++) { s0 += __ldcg(tab + 2 * r); s1 += __ldcg(tab + 2 * r + 1); }
  rr = s0; rz = s1;
}

// p = z + beta p over ALL rows (owned + halo: z's halo rows were received with the partials); block 0 leaves the summed
// scalars in red[] for the kernels that follow (alpha = rho / p.q, rho1 of the next iteration)
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
cg_update_p_zh_kernel(int k, size_t np3, const double *__restrict__ z, double *__restrict__ p, double *red, CgCtl *ctl,
                      double *hist, int hist_len, const double *tab, int nranks) {
  if (ctl->phase != 0) return;
  double beta, rr, rz;
  cg_sum_partials(tab, nranks, rr, rz);
  const bool leader = blockIdx.x == 0 && threadIdx.x == 0;
  if (leader) { red[2 * (k & 1)] = rr; red[2 * (k & 1) + 1] = rz; }
  if (!cg_iter_prologue(k, red, ctl, hist, hist_len, leader, rr, rz, beta)) return;
  const size_t stride = (size_t)gridDim.x * BLOCK;
  if (k == 1) {
    for (size_t i = (size_t)blockIdx.x * BLOCK + threadIdx.x; i < np3; i += stride) p[i] = z[i];
  } else {
    for (size_t i = (size_t)blockIdx.x * BLOCK + threadIdx.x; i < np3; i += stride) p[i] = fma(beta, p[i], z[i]);
  }
}

// check after the last iteration of a launch batch (cg_check_kernel of the other transports)
__global__ void cg_check_zh_kernel(int k, double *red, CgCtl *ctl, double *hist, int hist_len, const double *tab, int nranks) {
  if (ctl->phase != 0) return;
  if (threadIdx.x == 0) {
    double beta, rr, rz;
    cg_sum_partials(tab, nranks, rr, rz);
    red[2 * (k & 1)] = rr; red[2 * (k & 1) + 1] = rz;
    cg_iter_prologue(k, red, ctl, hist, hist_len, true, rr, rz, beta);
  }
}

// cg_fused_update_kernel whose partial sums go to this rank's slot of the table the next iteration reads
// (x += alpha p ; r -= alpha q ; r.r ; z = Minv r ; r.z: CG.f90:219-240,160-168 + DIAG_33.f90:136-148)
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
cg_fused_update_zh_kernel(int k, int N, const double *__restrict__ p, const double *__restrict__ q,
                          double *__restrict__ x, double *__restrict__ r, double *__restrict__ z,
                          const double *__restrict__ alu, size_t ld, double *red, CgCtl *ctl,
                          double *partial, unsigned int *counter, double *my_slot) {
  if (ctl->phase != 0) return;
  double alpha;
  if (!cg_alpha_prologue(k, red, ctl, blockIdx.x == 0 && threadIdx.x == 0, red[RED_PQ], alpha)) return;
  double acc[2] = {0.0, 0.0};
  for (size_t i = (size_t)blockIdx.x * BLOCK + threadIdx.x; i < (size_t)N; i += (size_t)gridDim.x * BLOCK) {
    const size_t j = 3 * i;
    x[j] = fma(alpha, p[j], x[j]); x[j + 1] = fma(alpha, p[j + 1], x[j + 1]); x[j + 2] = fma(alpha, p[j + 2], x[j + 2]);
    const double r0 = fma(-alpha, q[j], r[j]), r1 = fma(-alpha, q[j + 1], r[j + 1]), r2 = fma(-alpha, q[j + 2], r[j + 2]);
    r[j] = r0; r[j + 1] = r1; r[j + 2] = r2;
    acc[0] = fma(r0, r0, acc[0]); acc[0] = fma(r1, r1, acc[0]); acc[0] = fma(r2, r2, acc[0]);
    double a = r0, b = r1, c = r2;
    diag_solve(alu, ld, i, a, b, c);
    z[j] = a; z[j + 1] = b; z[j + 2] = c;
    acc[1] = fma(r0, a, acc[1]); acc[1] = fma(r1, b, acc[1]); acc[1] = fma(r2, c, acc[1]);
  }
  grid_reduce<2, BLOCK>(acc, partial, counter, [=](double(&o)[2]) { my_slot[0] = o[0]; my_slot[1] = o[1]; });
}

}  // namespace fb
