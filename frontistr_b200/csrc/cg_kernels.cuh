// Fused vector kernels of hecmw_solve_CG (hecmw1/src/solver/iterative/hecmw_solver_CG.f90:153-271)
// and the 3x3 block-Jacobi "DIAG" preconditioner (precond/33/hecmw_precond_DIAG_33.f90).
//
// All CG scalars live on the device; no kernel waits for the host.  Control decisions of the
// reference loop (RESID<=TOL, RHO==0, sign changes of RHO, C1<=0) are evaluated in the prologue of
// the kernel that needs the scalar; a decision to stop is published in `phase`, which turns every
// later kernel of the speculative launch queue into a no-op.  The host inspects `phase` every few
// iterations only.
//
// red[] slot map (all doubles, device):
//   red[2*par+0] = r.r of iteration k   (DNRM2, CG.f90:240)      par = (k+1)&1
//   red[2*par+1] = r.z for iteration k+1 (RHO,  CG.f90:168)
//   red[4]       = p.q (C1, CG.f90:211)   red[5] = b.b (BNRM2, :123)   red[6] = scratch r.r
#pragma once
#include "common.cuh"
#include "p2p.cuh"

namespace fb {

struct CgCtl {          // device-resident control block
  int phase;            // 0 run | 1 RESID<=TOL candidate | 2 RHO==0 | 3002 | 3003
  int event_iter;       // iteration at which phase was raised
  int n_indef[2];       // sign-change counter (CG.f90:173-179), parity double-buffered
  double event_resid;
  double tol;
};

enum { RED_PQ = 4, RED_BNRM2 = 5, RED_TMP = 6, RED_COUNT = 8 };

// Per-launch communication plan of the peer-memory mode (p2p.cuh).  All zero = the scalars are in red[]
// (single rank, or NCCL all-reduce between the kernels) and no halo traffic is issued by the kernel.
struct P2PArgs {
  unsigned long long consume_seq = 0;   // > 0: take the reduced scalars from the peer table, not from red[]
  unsigned long long publish_seq = 0;   // > 0: publish this kernel's partial sums to all ranks, not to red[]
  unsigned long long halo_seq = 0;      // > 0: update_p pushes the export values + signals; SpMV waits
};

// ALU is stored SoA: alu[c*ld + node], c = 0..8 (reference ALU(9*i-8+c)), coalesced per lane.
__device__ __forceinline__ void diag_solve(const double *__restrict__ alu, size_t ld, size_t i,
                                           double &X1, double &X2, double &X3) {
  // forward/back substitution, precond/33/hecmw_precond_DIAG_33.f90:136-148
  const double a1 = alu[1 * ld + i], a2 = alu[2 * ld + i], a3 = alu[3 * ld + i];
  const double a5 = alu[5 * ld + i], a6 = alu[6 * ld + i], a7 = alu[7 * ld + i];
  const double a0 = alu[0 * ld + i], a4 = alu[4 * ld + i], a8 = alu[8 * ld + i];
  X2 = fma(-a3, X1, X2);
  X3 = fma(-a7, X2, fma(-a6, X1, X3));
  X3 = a8 * X3;
  X2 = a4 * fma(-a5, X3, X2);
  X1 = a0 * fma(-a1, X2, fma(-a2, X3, X1));
}

// hecmw_precond_DIAG_33_setup, DIAG_33.f90:82-117 (3x3 LU, no pivoting, reciprocal pivots)
__global__ void __launch_bounds__(256)
diag_setup_kernel(int N, const double *__restrict__ D, double sigma, double *__restrict__ alu, size_t ld) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)N) return;
  double T[3][3];
  const double *d = D + 9 * i;
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int b = 0; b < 3; b++) T[a][b] = d[3 * a + b] * ((a == b) ? sigma : 1.0);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    T[k][k] = 1.0 / T[k][k];
#pragma unroll
    for (int a = k + 1; a < 3; a++) {
      T[a][k] = T[a][k] * T[k][k];
#pragma unroll
      for (int b = k + 1; b < 3; b++) T[a][b] = fma(-T[a][k], T[k][b], T[a][b]);
    }
  }
#pragma unroll
  for (int a = 0; a < 3; a++)
#pragma unroll
    for (int b = 0; b < 3; b++) alu[(3 * a + b) * ld + i] = T[a][b];
}

// standalone hecmw_precond_DIAG_33_apply on a vector (generic preconditioner path / C-ABI apply)
__global__ void __launch_bounds__(256)
diag_apply_kernel(int N, const double *__restrict__ alu, size_t ld, double *__restrict__ zp, const int *phase) {
  if (phase && *phase != 0) return;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)N; i += (size_t)gridDim.x * blockDim.x) {
    double a = zp[3 * i], b = zp[3 * i + 1], c = zp[3 * i + 2];
    diag_solve(alu, ld, i, a, b, c);
    zp[3 * i] = a; zp[3 * i + 1] = b; zp[3 * i + 2] = c;
  }
}

// z = Minv r  and  rz = r.z   (first iteration, after a true-residual recompute)
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
cg_precond_dot_kernel(int N, const double *__restrict__ alu, size_t ld, const double *__restrict__ r,
                      double *__restrict__ z, double *partial, unsigned int *counter, double *rz_out,
                      const int *phase) {
  if (phase && *phase != 0) return;
  double acc[1] = {0.0};
  for (size_t i = (size_t)blockIdx.x * BLOCK + threadIdx.x; i < (size_t)N; i += (size_t)gridDim.x * BLOCK) {
    const double r0 = r[3 * i], r1 = r[3 * i + 1], r2 = r[3 * i + 2];
    double a = r0, b = r1, c = r2;
    diag_solve(alu, ld, i, a, b, c);
    z[3 * i] = a; z[3 * i + 1] = b; z[3 * i + 2] = c;
    acc[0] = fma(r0, a, acc[0]); acc[0] = fma(r1, b, acc[0]); acc[0] = fma(r2, c, acc[0]);
  }
  grid_reduce<1, BLOCK>(acc, partial, counter, [=](double(&o)[1]) { *rz_out = o[0]; });
}

// generic dot product over the first n entries: out = sum a_i b_i  (hecmw_InnerProduct_R,
// communication/hecmw_solver_misc.f90:46-70, local part)
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
dot_kernel(size_t n, const double *__restrict__ a, const double *__restrict__ b, double *partial,
           unsigned int *counter, double *out, const int *phase) {
  if (phase && *phase != 0) return;
  double acc[1] = {0.0};
  for (size_t i = (size_t)blockIdx.x * BLOCK + threadIdx.x; i < n; i += (size_t)gridDim.x * BLOCK)
    acc[0] = fma(a[i], b[i], acc[0]);
  grid_reduce<1, BLOCK>(acc, partial, counter, [=](double(&o)[1]) { *out = o[0]; });
}

// Prologue of iteration k (k>=1): convergence test of iteration k-1, RHO tests of iteration k.
// Returns false when the iteration must not run.  `leader` (one thread of block 0) records events.
__device__ __forceinline__ bool cg_iter_prologue(int k, const double *red, CgCtl *ctl, double *hist,
                                                 int hist_len, bool leader, double rr, double rz, double &beta) {
  const int par = k & 1;
  if (k > 1) {
    const double resid = sqrt(rr / red[RED_BNRM2]);            // RESID of iteration k-1, CG.f90:242
    if (leader && hist && k - 2 < hist_len) hist[k - 2] = resid;
    if (resid <= ctl->tol) {                                   // CG.f90:259
      if (leader) { ctl->event_iter = k - 1; ctl->event_resid = resid; __threadfence(); ctl->phase = 1; }
      return false;
    }
  }
  if (rz == 0.0) {                                             // CG.f90:170-172
    if (leader) { ctl->event_iter = k; __threadfence(); ctl->phase = 2; }
    return false;
  }
  int ni = ctl->n_indef[par];
  const double rz1 = red[2 * (par ^ 1) + 1];
  if (k > 1 && rz * rz1 <= 0) {                                // CG.f90:173-179
    ni++;
    if (ni >= 3) {
      if (leader) { ctl->event_iter = k; __threadfence(); ctl->phase = 3003; }
      return false;
    }
  }
  if (leader) ctl->n_indef[par ^ 1] = ni;
  beta = (k == 1) ? 0.0 : rz / rz1;                            // CG.f90:188-197
  return true;
}

// (r.r, r.z) that iteration k starts from: red[] or, in peer-memory mode, the sum over all ranks' partials.
// Block-uniform result; block 0 also stores the sums in red[] for the kernels that follow in the stream.
__device__ __forceinline__ void cg_fetch_rr_rz(int k, double *red, const P2PView &pv, const P2PArgs &a,
                                               double &rr, double &rz) {
  const int par = k & 1;
  if (a.consume_seq == 0) { rr = red[2 * par]; rz = red[2 * par + 1]; return; }
  __shared__ double s_rrrz[2];
  if (threadIdx.x < 32) {
    double s0, s1;
    p2p_consume_warp(pv, P2P_Q_RRRZ, a.consume_seq, s0, s1);
    if (threadIdx.x == 0) {
      s_rrrz[0] = s0; s_rrrz[1] = s1;
      if (blockIdx.x == 0) { red[2 * par] = s0; red[2 * par + 1] = s1; }
    }
  }
  __syncthreads();
  rr = s_rrrz[0]; rz = s_rrrz[1];
}

// p = z + beta p  (p = z on the first iteration) + gather of the export halo values of p.
// Peer-memory mode (a.halo_seq > 0): the export values are also stored straight into the halo part of the
// neighbours' p vectors and the last block signals exchange number a.halo_seq to every neighbour.
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
cg_update_p_kernel(int k, size_t n3, const double *__restrict__ z, double *__restrict__ p,
                   double *red, CgCtl *ctl, double *hist, int hist_len,
                   int n_export, const int *__restrict__ export_item, double *__restrict__ sendbuf,
                   P2PView pv, P2PNeighbors nb, P2PArgs a, const int *__restrict__ exp_peer,
                   const int *__restrict__ exp_remote, unsigned int *ticket) {
  if (ctl->phase != 0) return;
  double beta, rr, rz;
  cg_fetch_rr_rz(k, red, pv, a, rr, rz);
  if (!cg_iter_prologue(k, red, ctl, hist, hist_len, blockIdx.x == 0 && threadIdx.x == 0, rr, rz, beta)) return;
  const size_t stride = (size_t)gridDim.x * BLOCK;
  // export pack first (hecmw_solver_SR_33.F90:72-81) so the peer stores are in flight while p is updated;
  // recomputed from z and the previous export value, so no grid-wide ordering with the update below is needed.
  for (size_t e = (size_t)blockIdx.x * BLOCK + threadIdx.x; e < 3 * (size_t)n_export; e += stride) {
    const size_t en = e / 3, c = e % 3;
    const size_t i = 3 * (size_t)(export_item[en] - 1) + c;
    const double v = (k == 1) ? z[i] : fma(beta, sendbuf[e], z[i]);
    sendbuf[e] = v;
    if (a.halo_seq) st_relaxed_sys(pv.vecP[exp_peer[en]] + 3 * (size_t)exp_remote[en] + c, v);
  }
  if (k == 1) {
    for (size_t i = (size_t)blockIdx.x * BLOCK + threadIdx.x; i < n3; i += stride) p[i] = z[i];
  } else {
    for (size_t i = (size_t)blockIdx.x * BLOCK + threadIdx.x; i < n3; i += stride) p[i] = fma(beta, p[i], z[i]);
  }
  if (a.halo_seq) {
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();   // cumulative: orders the peer stores of the whole block (bar.sync above) before the ticket
      const unsigned int t = atomicAdd(ticket, 1u);
      if (t == gridDim.x - 1) {
        *ticket = 0u;
        p2p_signal_halo(pv, nb, a.halo_seq);
      }
    }
  }
}

// standalone check used after the last iteration of a launch batch
__global__ void cg_check_kernel(int k, double *red, CgCtl *ctl, double *hist, int hist_len, P2PView pv, P2PArgs a) {
  if (ctl->phase != 0) return;
  double beta, rr, rz;
  cg_fetch_rr_rz(k, red, pv, a, rr, rz);
  if (threadIdx.x == 0) cg_iter_prologue(k, red, ctl, hist, hist_len, true, rr, rz, beta);
}

__device__ __forceinline__ bool cg_alpha_prologue(int k, const double *red, CgCtl *ctl, bool leader, double c1, double &alpha) {
  if (c1 <= 0) {                                               // CG.f90:213-217
    if (leader) { ctl->event_iter = k; __threadfence(); ctl->phase = 3002; }
    return false;
  }
  alpha = red[2 * (k & 1) + 1] / c1;                           // ALPHA = RHO / C1
  return true;
}

// Fused tail of iteration k for PRECOND=3, iterPREmax=1:
//   x += alpha p ; r -= alpha q ; rr = r.r ; z = Minv r ; rz = r.z        (CG.f90:219-240 + :160-168)
// PEER = true: the instantiation of the peer-memory transport (p.q taken from / (r.r, r.z) published to the peer tables
// when the launch's P2PArgs say so).  PEER = false (one GPU, NCCL) contains none of it - in particular not the 240-byte
// copy of the P2PView into every thread's stack frame that the by-value capture of the publication costs.
template <int BLOCK, bool PEER = false>
__global__ void __launch_bounds__(BLOCK)
cg_fused_update_kernel(int k, int N, const double *__restrict__ p, const double *__restrict__ q,
                       double *__restrict__ x, double *__restrict__ r, double *__restrict__ z,
                       const double *__restrict__ alu, size_t ld, double *red, CgCtl *ctl,
                       double *partial, unsigned int *counter, const __grid_constant__ P2PView pv, P2PArgs a) {
  if (ctl->phase != 0) return;
  double alpha, c1 = 0.0;
  bool from_table = false;
  if constexpr (PEER) {
    if (a.consume_seq) {        // p.q = sum over all ranks' partials (peer table), block-uniform
      __shared__ double s_pq;
      if (threadIdx.x < 32) {
        double s0, s1;
        p2p_consume_warp(pv, P2P_Q_PQ, a.consume_seq, s0, s1);
        if (threadIdx.x == 0) s_pq = s0;
      }
      __syncthreads();
      c1 = s_pq;
      from_table = true;
    }
  }
  if (!from_table) c1 = red[RED_PQ];
  if (!cg_alpha_prologue(k, red, ctl, blockIdx.x == 0 && threadIdx.x == 0, c1, alpha)) return;
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
  double *out = red + 2 * ((k + 1) & 1);
  if constexpr (PEER) {
    const P2PView *ppv = &pv;   // (__grid_constant__: an address in the constant bank, no per-thread copy of the table)
    grid_reduce<2, BLOCK, true>(acc, partial, counter, [=](double(&o)[2]) {
      if (a.publish_seq) p2p_publish_warp(*ppv, P2P_Q_RRRZ, a.publish_seq, o[0], o[1]);
      else if (threadIdx.x == 0) { out[0] = o[0]; out[1] = o[1]; }
    });
  } else {
    grid_reduce<2, BLOCK>(acc, partial, counter, [=](double(&o)[2]) { out[0] = o[0]; out[1] = o[1]; });
  }
}

// x += alpha p only (iterations where r is recomputed from b - A x, CG.f90:227-233), or
// x += alpha p ; r -= alpha q ; rr (generic preconditioner path) when r != nullptr
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
cg_update_xr_kernel(int k, size_t n3, const double *__restrict__ p, const double *__restrict__ q,
                    double *__restrict__ x, double *__restrict__ r, double *red, CgCtl *ctl,
                    double *partial, unsigned int *counter) {
  if (ctl->phase != 0) return;
  double alpha;
  if (!cg_alpha_prologue(k, red, ctl, blockIdx.x == 0 && threadIdx.x == 0, red[RED_PQ], alpha)) return;
  double acc[1] = {0.0};
  for (size_t i = (size_t)blockIdx.x * BLOCK + threadIdx.x; i < n3; i += (size_t)gridDim.x * BLOCK) {
    x[i] = fma(alpha, p[i], x[i]);
    if (r) {
      const double ri = fma(-alpha, q[i], r[i]);
      r[i] = ri;
      acc[0] = fma(ri, ri, acc[0]);
    }
  }
  if (r) {
    double *out = red + 2 * ((k + 1) & 1);
    grid_reduce<1, BLOCK>(acc, partial, counter, [=](double(&o)[1]) { out[0] = o[0]; });
  }
}

// x += alpha p ; r -= alpha q ; rr ; z = r: the tail of an iteration whose preconditioner then works on z in place (fused
// SSOR path, api_solver.inc) - the copy ZP = R of hecmw_precond_apply (precond.f90:91-96) rides on the update instead of
// being a pass of its own
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK)
cg_update_xrz_kernel(int k, size_t n3, const double *__restrict__ p, const double *__restrict__ q,
                     double *__restrict__ x, double *__restrict__ r, double *__restrict__ z, double *red, CgCtl *ctl,
                     double *partial, unsigned int *counter) {
  if (ctl->phase != 0) return;
  double alpha;
  if (!cg_alpha_prologue(k, red, ctl, blockIdx.x == 0 && threadIdx.x == 0, red[RED_PQ], alpha)) return;
  double acc[1] = {0.0};
  for (size_t i = (size_t)blockIdx.x * BLOCK + threadIdx.x; i < n3; i += (size_t)gridDim.x * BLOCK) {
    x[i] = fma(alpha, p[i], x[i]);
    const double ri = fma(-alpha, q[i], r[i]);
    r[i] = ri;
    z[i] = ri;
    acc[0] = fma(ri, ri, acc[0]);
  }
  double *out = red + 2 * ((k + 1) & 1);
  grid_reduce<1, BLOCK>(acc, partial, counter, [=](double(&o)[1]) { out[0] = o[0]; });
}

// ---- small vector helpers for the generic preconditioner path (precond.f90:75-123, precond_33.f90:86-114)
__global__ void vec_copy_zero_tail_kernel(size_t n_head, size_t n_all, const double *__restrict__ src,
                                          double *__restrict__ dst, const int *phase) {
  if (phase && *phase != 0) return;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_all; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = (i < n_head) ? src[i] : 0.0;
}
__global__ void vec_add_kernel(size_t n, double *__restrict__ z, const double *__restrict__ zp, int first,
                               const int *phase) {
  if (phase && *phase != 0) return;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    z[i] = first ? zp[i] : z[i] + zp[i];
}
__global__ void gather3_kernel(int n, const int *__restrict__ item, const double *__restrict__ x,
                               double *__restrict__ buf) {
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < 3 * (size_t)n; e += (size_t)gridDim.x * blockDim.x)
    buf[e] = x[3 * (size_t)(item[e / 3] - 1) + e % 3];
}
__global__ void scatter3_kernel(int n, const int *__restrict__ item, const double *__restrict__ buf,
                                double *__restrict__ x) {
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < 3 * (size_t)n; e += (size_t)gridDim.x * blockDim.x)
    x[3 * (size_t)(item[e / 3] - 1) + e % 3] = buf[e];
}

}  // namespace fb
