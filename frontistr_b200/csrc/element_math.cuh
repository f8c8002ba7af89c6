// Element stiffness integration for STF_C3D8IC / STF_C3 / STF_C3D8Bbar
// (fistr1/src/lib/static_LIB_3dIC.f90:21-215, static_LIB_3d.f90:47-205, static_LIB_C3D8.f90:23-198),
// written as per-lane phase functions over a per-warp scratch block.  In the CUDA kernel a phase is
// executed by the 32 lanes of one warp with __syncwarp() between phases; the host emulation used by
// the CPU unit tests (tests/test_element_emul.py via csrc/emul_test.cpp) runs the same functions in
// a lane loop.  One element per warp; FP64 throughout.
//
// For an isotropic elastic D (ElasticLinear.f90:48-59: D11, D12, D44) the 3x3 block coupling
// "nodes" i and j at one Gauss point is evaluated in closed form,
//     K_ij[a][a] += w (D11 gi[a] gj[a] + D44 sum_{c!=a} gi[c] gj[c])
//     K_ij[a][b] += w (D12 gi[a] gj[b] + D44 gi[b] gj[a])            a != b
// which is B_i^T D B_j of static_LIB_3d.f90:120-167 with the structural zeros of B removed.
// The incompatible-mode element treats its 3 enhanced modes as nodes 8..10 and condenses them
// with the reference's own pivoted Gauss-Jordan (lib/utilities/utilities.f90:247-316).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define FB_HD __host__ __device__ __forceinline__
#else
#define FB_HD inline
#endif

// Warp collectives used by the element code.  In the CUDA kernels they are the hardware shuffles; the
// TEST-ONLY host emulation (emul_test.cpp) defines FB_EMUL_WARP and supplies lock-step versions (32 host
// threads + barriers), so the very same source is what the CPU unit tests execute.
#if defined(FB_EMUL_WARP)
namespace fb { namespace emulwarp {   // (declared with the call-site tag in tests/emul/kernel_emul.hpp)
double shfl(double v, int src, unsigned long long site);
int shfl(int v, int src, unsigned long long site);
void syncwarp(unsigned long long site);
} }
#define FB_WARP_FN inline
#define FB_SHFL(v, src) fb::emulwarp::shfl((v), (src), FB_EMUL_SITE)
#define FB_SYNCWARP() fb::emulwarp::syncwarp(FB_EMUL_SITE)
#elif defined(__CUDACC__)
#define FB_WARP_FN __device__ __forceinline__
#define FB_SHFL(v, src) __shfl_sync(0xffffffffu, (v), (src))
#define FB_SYNCWARP() __syncwarp()
#endif

namespace fb {
namespace elem {

enum Form { FORM_FI = 1, FORM_BBAR = 2, FORM_IC = 3, FORM_TET10 = 10 };

constexpr double kGaussHex = 0.577350269189626;     // quadrature.f90:83-91
constexpr double kTetA = 0.138196601125011;         // quadrature.f90:124-128
constexpr double kTetB = 0.585410196624968;
constexpr double kTetW = 0.041666666666667;         // quadrature.f90:236-238

// per-warp scratch (doubles).  MAXN = nodes + enhanced modes, NGP Gauss points
struct alignas(16) Scratch {
  double stage[100 * 9];    // gather-mode output staging: nn*nn blocks of 9 (tet10: 100 blocks); 16-byte aligned
  double xc[10 * 3];        // element coordinates, node-major
  double gd[8 * 11 * 3];    // gderiv[g][n][3]  (hex: 8 GP x 11; tet10: 4 GP x 10 fits as well)
  double bb[8 * 8 * 3];     // B-bar correction (Bbar - gderiv)/3 per GP,node (B-bar only)
  double wdet[8];           // weight * det per GP
  double Kda[24 * 10];      // IC: K(d, alpha), row stride 10 (16-byte aligned rows)
  double Kaa[9 * 9];        // IC: K(alpha, alpha) -> inverse, row-major [r][c]
  double W[24 * 10];        // IC: Kda * Kaa^-1, row stride 10
  double piv[12];           // IC, shared-memory inverse only: IP(0..8), LR, PIVOT
  // ---- UpdateST_C3D8IC only (strain / stress recovery of the IC element)
  double uc[10 * 3];        // nodal displacement unode + dunode
  double tc[10], t0c[10];   // nodal temperature / initial temperature
  double cd[12];            // enhanced-mode amplitudes -Kaa^-1 Kad u (9 used)
  double tf[12];            // Kad u, then Kaa^-1 tv(24:32) (9 used)
  double tv[36];            // thermal load vector incl. the enhanced modes (33 used)
  double eth[8];            // thermal strain per Gauss point
};

FB_HD void hex_shape_deriv(double x, double y, double z, double (*d)[3]) {
  // ShapeDeriv_hex8n, lib/element/hex8n.f90:24-53
  const double sx[8] = {-1, 1, 1, -1, -1, 1, 1, -1};
  const double sy[8] = {-1, -1, 1, 1, -1, -1, 1, 1};
  const double sz[8] = {-1, -1, -1, -1, 1, 1, 1, 1};
  for (int n = 0; n < 8; n++) {
    d[n][0] = (0.125 * sx[n]) * (1.0 + sy[n] * y) * (1.0 + sz[n] * z);
    d[n][1] = (0.125 * sy[n]) * (1.0 + sx[n] * x) * (1.0 + sz[n] * z);
    d[n][2] = (0.125 * sz[n]) * (1.0 + sx[n] * x) * (1.0 + sy[n] * y);
  }
}

FB_HD void tet10_shape_deriv(double xi, double et, double ze, double (*s)[3]) {
  // ShapeDeriv_tet10n, lib/element/tet10n.f90:29-75
  const double a = 1.0 - xi - et - ze;
  s[0][0] = 1.0 - 4.0 * a; s[1][0] = 4.0 * xi - 1.0; s[2][0] = 0.0; s[3][0] = 0.0;
  s[4][0] = 4.0 * (1.0 - 2.0 * xi - et - ze); s[5][0] = 4.0 * et; s[6][0] = -4.0 * et;
  s[7][0] = -4.0 * ze; s[8][0] = 4.0 * ze; s[9][0] = 0.0;
  s[0][1] = 1.0 - 4.0 * a; s[1][1] = 0.0; s[2][1] = 4.0 * et - 1.0; s[3][1] = 0.0;
  s[4][1] = -4.0 * xi; s[5][1] = 4.0 * xi; s[6][1] = 4.0 * (1.0 - xi - 2.0 * et - ze);
  s[7][1] = -4.0 * ze; s[8][1] = 0.0; s[9][1] = 4.0 * ze;
  s[0][2] = 1.0 - 4.0 * a; s[1][2] = 0.0; s[2][2] = 0.0; s[3][2] = 4.0 * ze - 1.0;
  s[4][2] = -4.0 * xi; s[5][2] = 0.0; s[6][2] = -4.0 * et;
  s[7][2] = 4.0 * (1.0 - xi - et - 2.0 * ze); s[8][2] = 4.0 * xi; s[9][2] = 4.0 * et;
}

// J = X^T dN, det, cofactor inverse (getGlobalDeriv / getJacobian, element.f90:693-741,772-819)
FB_HD void jacobian(int nn, const double *xc, const double (*d)[3], double &det, double (*inv)[3]) {
  double J[3][3];
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0.0;
      for (int n = 0; n < nn; n++) s += xc[3 * n + i] * d[n][j];
      J[i][j] = s;
    }
  det = J[0][0] * J[1][1] * J[2][2] + J[1][0] * J[2][1] * J[0][2] + J[2][0] * J[0][1] * J[1][2] -
        J[2][0] * J[1][1] * J[0][2] - J[1][0] * J[0][1] * J[2][2] - J[0][0] * J[2][1] * J[1][2];
  const double dum = 1.0 / det;
  inv[0][0] = dum * (J[1][1] * J[2][2] - J[2][1] * J[1][2]);
  inv[0][1] = dum * (-J[0][1] * J[2][2] + J[2][1] * J[0][2]);
  inv[0][2] = dum * (J[0][1] * J[1][2] - J[1][1] * J[0][2]);
  inv[1][0] = dum * (-J[1][0] * J[2][2] + J[2][0] * J[1][2]);
  inv[1][1] = dum * (J[0][0] * J[2][2] - J[2][0] * J[0][2]);
  inv[1][2] = dum * (-J[0][0] * J[1][2] + J[1][0] * J[0][2]);
  inv[2][0] = dum * (J[1][0] * J[2][1] - J[2][0] * J[1][1]);
  inv[2][1] = dum * (-J[0][0] * J[2][1] + J[2][0] * J[0][1]);
  inv[2][2] = dum * (J[0][0] * J[1][1] - J[1][0] * J[0][1]);
}

// ---- phase G: lane g < NGP computes the geometry of Gauss point g --------------------------
template <int FORM, class SCR>
FB_HD void phase_geometry(int lane, SCR &S) {
  constexpr int NN = (FORM == FORM_TET10) ? 10 : 8;
  constexpr int NGP = (FORM == FORM_TET10) ? 4 : 8;
  constexpr int NT = (FORM == FORM_IC) ? 11 : NN;
  if (lane >= NGP) return;
  double d[10][3], inv[3][3], det;
  double nc[3];
  if (FORM == FORM_TET10) {
    nc[0] = (lane == 1) ? kTetB : kTetA; nc[1] = (lane == 2) ? kTetB : kTetA; nc[2] = (lane == 3) ? kTetB : kTetA;
    tet10_shape_deriv(nc[0], nc[1], nc[2], d);
  } else {
    nc[0] = (lane & 1) ? kGaussHex : -kGaussHex; nc[1] = (lane & 2) ? kGaussHex : -kGaussHex;
    nc[2] = (lane & 4) ? kGaussHex : -kGaussHex;
    hex_shape_deriv(nc[0], nc[1], nc[2], d);
  }
  jacobian(NN, S.xc, d, det, inv);
  double *gd = S.gd + lane * NT * 3;
  for (int n = 0; n < NN; n++)
    for (int j = 0; j < 3; j++) {
      double s = 0.0;
      for (int k = 0; k < 3; k++) s += d[n][k] * inv[k][j];
      gd[3 * n + j] = s;
    }
  S.wdet[lane] = ((FORM == FORM_TET10) ? kTetW : 1.0) * det;
  if (FORM == FORM_IC || FORM == FORM_BBAR) {
    double d0[10][3], inv0[3][3], det0;
    hex_shape_deriv(0.0, 0.0, 0.0, d0);
    jacobian(8, S.xc, d0, det0, inv0);
    if (FORM == FORM_IC) {
      // gderiv(nn+m,:) = -2 xi_m (J0^-1 det0)(m,:) / det, static_LIB_3dIC.f90:79-81,120-122
      for (int m = 0; m < 3; m++)
        for (int j = 0; j < 3; j++) gd[3 * (8 + m) + j] = -2.0 * nc[m] * (inv0[m][j] * det0) / det;
    } else {
      // B4,B6,B8 = (Bbar - gderiv)/3 with Bbar the centroid derivative, static_LIB_C3D8.f90:72-74,104-106
      double *bb = S.bb + lane * 8 * 3;
      for (int n = 0; n < 8; n++)
        for (int j = 0; j < 3; j++) {
          double s = 0.0;
          for (int k = 0; k < 3; k++) s += d0[n][k] * inv0[k][j];
          bb[3 * n + j] = (s - gd[3 * n + j]) / 3.0;
        }
    }
  }
}

// closed-form isotropic pair block, accumulated over the Gauss points
template <int NT, int NGP>
FB_HD void pair_block_iso(const Scratch &S, int i, int j, double D11, double D12, double D44, double *k9) {
  for (int a = 0; a < 9; a++) k9[a] = 0.0;
  for (int g = 0; g < NGP; g++) {
    const double *gi = S.gd + (g * NT + i) * 3, *gj = S.gd + (g * NT + j) * 3;
    const double w = S.wdet[g];
    const double gi0 = gi[0], gi1 = gi[1], gi2 = gi[2];
    const double h0 = w * gj[0], h1 = w * gj[1], h2 = w * gj[2];
    const double p0 = gi0 * h0, p1 = gi1 * h1, p2 = gi2 * h2;
    k9[0] += D11 * p0 + D44 * (p1 + p2);
    k9[4] += D11 * p1 + D44 * (p0 + p2);
    k9[8] += D11 * p2 + D44 * (p0 + p1);
    k9[1] += D12 * (gi0 * h1) + D44 * (gi1 * h0);
    k9[2] += D12 * (gi0 * h2) + D44 * (gi2 * h0);
    k9[3] += D12 * (gi1 * h0) + D44 * (gi0 * h1);
    k9[5] += D12 * (gi1 * h2) + D44 * (gi2 * h1);
    k9[6] += D12 * (gi2 * h0) + D44 * (gi0 * h2);
    k9[7] += D12 * (gi2 * h1) + D44 * (gi1 * h2);
  }
}

// B-bar pair block: B rows 1-3 carry the volumetric correction (static_LIB_C3D8.f90:103-125)
FB_HD void pair_block_bbar(const Scratch &S, int i, int j, double D11, double D12, double D44, double *k9) {
  for (int a = 0; a < 9; a++) k9[a] = 0.0;
  for (int g = 0; g < 8; g++) {
    const double *gi = S.gd + (g * 8 + i) * 3, *gj = S.gd + (g * 8 + j) * 3;
    const double *bi = S.bb + (g * 8 + i) * 3, *bj = S.bb + (g * 8 + j) * 3;
    const double w = S.wdet[g];
    double Bi[6][3], Bj[6][3];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) {
        Bi[r][c] = bi[c] + ((r == c) ? gi[c] : 0.0);
        Bj[r][c] = bj[c] + ((r == c) ? gj[c] : 0.0);
      }
    for (int c = 0; c < 3; c++) { Bi[3][c] = Bi[4][c] = Bi[5][c] = 0.0; Bj[3][c] = Bj[4][c] = Bj[5][c] = 0.0; }
    Bi[3][0] = gi[1]; Bi[3][1] = gi[0]; Bi[4][1] = gi[2]; Bi[4][2] = gi[1]; Bi[5][0] = gi[2]; Bi[5][2] = gi[0];
    Bj[3][0] = gj[1]; Bj[3][1] = gj[0]; Bj[4][1] = gj[2]; Bj[4][2] = gj[1]; Bj[5][0] = gj[2]; Bj[5][2] = gj[0];
    for (int b = 0; b < 3; b++) {
      double DB[6];
      DB[0] = D11 * Bj[0][b] + D12 * Bj[1][b] + D12 * Bj[2][b];
      DB[1] = D12 * Bj[0][b] + D11 * Bj[1][b] + D12 * Bj[2][b];
      DB[2] = D12 * Bj[0][b] + D12 * Bj[1][b] + D11 * Bj[2][b];
      DB[3] = D44 * Bj[3][b]; DB[4] = D44 * Bj[4][b]; DB[5] = D44 * Bj[5][b];
      for (int a = 0; a < 3; a++) {
        double s = 0.0;
        for (int r = 0; r < 6; r++) s += Bi[r][a] * DB[r];
        k9[3 * a + b] += s * w;
      }
    }
  }
}

// ---- IC condensation ----------------------------------------------------------------------
// round 2 of the IC task list: lanes 0..23 -> K(d_i, alpha_m) blocks, lanes 24..29 -> K(alpha,alpha)
FB_HD void ic_phase_coupling(int lane, Scratch &S, double D11, double D12, double D44) {
  double k9[9];
  if (lane < 24) {
    const int i = lane / 3, m = lane % 3;
    pair_block_iso<11, 8>(S, i, 8 + m, D11, D12, D44, k9);
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) S.Kda[(3 * i + a) * 10 + 3 * m + b] = k9[3 * a + b];
  } else if (lane < 30) {
    const int pm[6] = {0, 0, 0, 1, 1, 2}, pn[6] = {0, 1, 2, 1, 2, 2};
    const int m = pm[lane - 24], n = pn[lane - 24];
    pair_block_iso<11, 8>(S, 8 + m, 8 + n, D11, D12, D44, k9);
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) {
        S.Kaa[(3 * m + a) * 9 + 3 * n + b] = k9[3 * a + b];
        if (m != n) S.Kaa[(3 * n + b) * 9 + 3 * m + a] = k9[3 * a + b];
      }
  }
}

// ---- calInverse(9) (utilities.f90:247-316) with one matrix COLUMN per lane held in registers ---------
// Lane c < 9 owns column c (a[0..8] = rows); values cross lanes only by broadcast / pairwise exchange, which
// the CUDA kernel does with warp shuffles (assembly.cuh: ic_inverse_warp) and the host emulation with plain
// arrays (emul_test.cpp).  K is always a compile-time constant after unrolling, so a[] stays in registers.
// Same operations in the same order as the reference: partial (row) pivoting, IP bookkeeping, the W != 0
// skips and the final column un-permutation "quirks included".
FB_HD void gj_pivot_local(int K, const double (&a)[9], int &LR, double &pivot) {   // meaningful on lane K only
  double wmax = 0.0;
  LR = K;
#pragma unroll
  for (int I = 0; I < 9; I++)
    if (I >= K) { const double w = fabs(a[I]); if (w > wmax) { wmax = w; LR = I; } }
  pivot = 0.0;
#pragma unroll
  for (int I = 0; I < 9; I++) if (I == LR) pivot = a[I];
}
FB_HD void gj_swap_rows(int K, int LR, double (&a)[9], int (&ip)[9]) {
#pragma unroll
  for (int I = 0; I < 9; I++)
    if (I > K && I == LR) {
      const double t = a[K]; a[K] = a[I]; a[I] = t;
      const int q = ip[K]; ip[K] = ip[I]; ip[I] = q;
    }
}
// Step K of the elimination on column c (held by this lane).  w[I] = A(I,K) after the row swap (column K, broadcast
// from lane K); w[K] is the pivot; qn[I] = -w[I] / PIVOT, the new column K, each entry divided by a different lane
// (ic_inverse_warp) so that lane K does not have to run eight FP64 divisions one after the other while 31 lanes wait.
FB_HD void gj_update(int c, int K, double (&a)[9], const double (&w)[9], const double (&qn)[9], double rpivot, double akc) {
  if (c == K) {
#pragma unroll
    for (int I = 0; I < 9; I++)
      if (I != K) { if (w[I] != 0.0) a[I] = qn[I]; }     // -W / PIVOT
    a[K] = rpivot;                                        // 1 / PIVOT
  } else {
    a[K] = akc;                                           // A(K,c) / PIVOT
#pragma unroll
    for (int I = 0; I < 9; I++)
      if (I != K) { if (w[I] != 0.0) a[I] = a[I] - w[I] * akc; }
  }
}
// un-permutation step I (utilities.f90:301-314): returns the column K to exchange with column I, or -1
FB_HD int gj_unpermute_step(int I, int (&ip)[9]) {
  int K = 0;
#pragma unroll
  for (int J = 0; J < 9; J++) if (J == I) K = ip[J];
  if (K == I) return -1;
  int IW = 0;
#pragma unroll
  for (int J = 0; J < 9; J++) if (J == K) IW = ip[J];
#pragma unroll
  for (int J = 0; J < 9; J++) if (J == K) ip[J] = K;   // IP(K) = IP(I) (= K)
#pragma unroll
  for (int J = 0; J < 9; J++) if (J == I) ip[J] = IW;
  return K;
}

#ifdef FB_WARP_FN
// calInverse(9, Kaa) for the IC condensation, executed by all 32 lanes of the warp that owns the element:
// column c of Kaa in the registers of lane c, cross-lane traffic by warp shuffles only.  Reads S.Kaa, writes
// the inverse back to S.Kaa.
FB_WARP_FN void ic_inverse_warp(int lane, Scratch &S) {
  const int c = (lane < 9) ? lane : 8;      // lanes >= 9 shadow column 8 (results discarded)
  double a[9];
  int ip[9];
#pragma unroll
  for (int I = 0; I < 9; I++) { a[I] = S.Kaa[I * 9 + c]; ip[I] = I; }
#pragma unroll
  for (int K = 0; K < 9; K++) {
    int LR;
    double pivot;
    gj_pivot_local(K, a, LR, pivot);
    LR = FB_SHFL(LR, K);
    gj_swap_rows(K, LR, a, ip);
    double w[9], qn[9];
#pragma unroll
    for (int I = 0; I < 9; I++) w[I] = FB_SHFL(a[I], K);
    // three divisions per lane, all lanes in parallel: lane I supplies -W(I) / PIVOT of the new column K
    double wc = 0.0;
#pragma unroll
    for (int I = 0; I < 9; I++) if (I == c) wc = w[I];
    const double q = -wc / w[K], rpivot = 1.0 / w[K], akc = a[K] / w[K];
#pragma unroll
    for (int I = 0; I < 9; I++) qn[I] = FB_SHFL(q, I);
    gj_update(c, K, a, w, qn, rpivot, akc);
  }
  // Un-permutation (utilities.f90:301-314): the reference exchanges columns I and K = IP(I) for I = 1..N in turn.  The
  // net effect is ONE column permutation; it is composed here in registers (ip is identical in every lane, so this is
  // warp-uniform integer work) and applied with a single round of shuffles - none at all when no row was ever
  // swapped, which is the normal case for the positive definite K(alpha,alpha).
  int col[9];            // col[p] = the column (= lane) whose content ends up at position p
  bool moved = false;
#pragma unroll
  for (int p = 0; p < 9; p++) col[p] = p;
#pragma unroll
  for (int I = 0; I < 9; I++) {
    const int K = gj_unpermute_step(I, ip);
    if (K >= 0) {
      moved = true;
#pragma unroll
      for (int J = 0; J < 9; J++)
        if (J == K) { const int t = col[J]; col[J] = col[I]; col[I] = t; }
    }
  }
  if (moved) {           // warp-uniform
    int src = lane;
#pragma unroll
    for (int p = 0; p < 9; p++) if (p == lane) src = col[p];
#pragma unroll
    for (int r = 0; r < 9; r++) a[r] = FB_SHFL(a[r], src);
  }
  FB_SYNCWARP();
  if (lane < 9) {
#pragma unroll
    for (int I = 0; I < 9; I++) S.Kaa[I * 9 + lane] = a[I];
  }
}
#endif

// ---- the same calInverse(9) through shared memory (the first implementation, the one every hardware run of
// round 1 used): the matrix stays in S.Kaa, lane 0 searches the pivot, lanes 0..8 own one column each, warp barriers
// between the phases.  Kept as the arbiter of the register/shuffle version above: the library compares the two on
// the device once per process (assembly.cuh: ic_selftest_kernel) and falls back to this one if they ever disagree.
FB_HD void ic_inv_pivot(int lane, Scratch &S, int K) {
  if (lane != 0) return;
  double wmax = 0.0;
  int LR = K;
  for (int I = K; I < 9; I++) {
    const double w = fabs(S.Kaa[I * 9 + K]);
    if (w > wmax) { wmax = w; LR = I; }
  }
  S.piv[9] = (double)LR;
  S.piv[10] = S.Kaa[LR * 9 + K];
  if (LR != K) { const double t = S.piv[K]; S.piv[K] = S.piv[LR]; S.piv[LR] = t; }
}
FB_HD void ic_inv_swap(int lane, Scratch &S, int K) {
  if (lane >= 9) return;
  const int LR = (int)S.piv[9];
  if (LR != K) { const double t = S.Kaa[K * 9 + lane]; S.Kaa[K * 9 + lane] = S.Kaa[LR * 9 + lane]; S.Kaa[LR * 9 + lane] = t; }
}
// lane c (<9) owns column c: scaling of row K, elimination; colK = A(:,K) after the swap, before this step
FB_HD void ic_inv_step(int lane, Scratch &S, int K, const double *colK) {
  if (lane >= 9) return;
  const int c = lane;
  const double PIVOT = S.piv[10];
  if (c == K) {
    for (int I = 0; I < 9; I++)
      if (I != K) { const double W = colK[I]; if (W != 0.0) S.Kaa[I * 9 + K] = -W / PIVOT; }
    S.Kaa[K * 9 + K] = 1.0 / PIVOT;
  } else {
    const double akc = S.Kaa[K * 9 + c] / PIVOT;
    S.Kaa[K * 9 + c] = akc;
    for (int I = 0; I < 9; I++)
      if (I != K) { const double W = colK[I]; if (W != 0.0) S.Kaa[I * 9 + c] = S.Kaa[I * 9 + c] - W * akc; }
  }
}
FB_HD void ic_inv_unpermute(int lane, Scratch &S) {
  if (lane != 0) return;
  int IP[9];
  for (int I = 0; I < 9; I++) IP[I] = (int)S.piv[I];
  for (int I = 0; I < 9; I++) {   // utilities.f90:301-314, quirks included
    const int K = IP[I];
    if (K != I) {
      const int IW = IP[K]; IP[K] = IP[I]; IP[I] = IW;
      for (int J = 0; J < 9; J++) { const double W = S.Kaa[J * 9 + I]; S.Kaa[J * 9 + I] = S.Kaa[J * 9 + K]; S.Kaa[J * 9 + K] = W; }
    }
  }
}
#ifdef FB_WARP_FN
FB_WARP_FN void ic_inverse_smem_warp(int lane, Scratch &S) {
  if (lane < 9) S.piv[lane] = (double)lane;
  FB_SYNCWARP();
  for (int K = 0; K < 9; K++) {
    ic_inv_pivot(lane, S, K);
    FB_SYNCWARP();
    ic_inv_swap(lane, S, K);
    FB_SYNCWARP();
    double colK[9];
#pragma unroll
    for (int I = 0; I < 9; I++) colK[I] = S.Kaa[I * 9 + K];
    FB_SYNCWARP();
    ic_inv_step(lane, S, K, colK);
    FB_SYNCWARP();
  }
  ic_inv_unpermute(lane, S);
  FB_SYNCWARP();
}
#endif

#ifdef FB_WARP_FN
// ---- third way to W = Kda * Kaa^-1 (ICINV == 2): no inverse at all.  K(alpha,alpha) is symmetric positive definite
// for every valid element, so W^T = Kaa^-1 Kad is obtained from an LDL^T factorisation and 24 triangular solves.
// Every lane factorises the 9x9 matrix redundantly in registers (165 FMA, 9 divisions, NO cross-lane traffic and no
// barrier), then lane r < 24 solves for row r of W.  This is NOT the reference's operation order (calInverse is a
// pivoted Gauss-Jordan): W differs from the reference-order result in the last bits (backward stable, cond(Kaa) is
// O(10)), far inside the 1e-12 gate on K.  A non-positive or non-finite pivot (distorted / inverted element) makes
// the function return false in ALL lanes (they computed the same numbers) and the caller takes the reference-order
// path for that element.  The library uses this kernel only if its device self-test agrees with the shared-memory
// reference-order kernel AND it is the fastest (api_assembly.inc).
FB_WARP_FN bool ic_w_ldlt_warp(int lane, Scratch &S) {
  double L[36], d[9], dinv[9];          // L(i,j), j < i, packed at i (i - 1) / 2 + j
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 9; j++) {
    double v[9];
    double dj = S.Kaa[j * 9 + j];
#pragma unroll
    for (int k = 0; k < 9; k++)
      if (k < j) { v[k] = L[j * (j - 1) / 2 + k] * d[k]; dj -= L[j * (j - 1) / 2 + k] * v[k]; }
    d[j] = dj;
    ok = ok && (dj > 0.0) && (dj < 1.0e300);
    dinv[j] = 1.0 / dj;
#pragma unroll
    for (int i = 0; i < 9; i++)
      if (i > j) {
        double t = S.Kaa[i * 9 + j];
#pragma unroll
        for (int k = 0; k < 9; k++)
          if (k < j) t -= L[i * (i - 1) / 2 + k] * v[k];
        L[i * (i - 1) / 2 + j] = t * dinv[j];
      }
  }
  if (!ok) return false;
  if (lane < 24) {
    double x[9];
#pragma unroll
    for (int i = 0; i < 9; i++) {         // L y = b
      double t = S.Kda[lane * 10 + i];
#pragma unroll
      for (int k = 0; k < 9; k++)
        if (k < i) t -= L[i * (i - 1) / 2 + k] * x[k];
      x[i] = t;
    }
#pragma unroll
    for (int i = 0; i < 9; i++) x[i] *= dinv[i];
#pragma unroll
    for (int i = 8; i >= 0; i--) {        // L^T x = z
      double t = x[i];
#pragma unroll
      for (int k = 0; k < 9; k++)
        if (k > i) t -= L[k * (k - 1) / 2 + i] * x[k];
      x[i] = t;
    }
#pragma unroll
    for (int i = 0; i < 9; i++) S.W[lane * 10 + i] = x[i];
  }
  return true;
}
#endif

// W = Kda * Kaa^-1   (tmpk, static_LIB_3dIC.f90:212): 216 entries over 32 lanes
FB_HD void ic_phase_W(int lane, Scratch &S) {
  for (int t = lane; t < 216; t += 32) {
    const int r = t / 9, c = t % 9;
    double s = 0.0;
    for (int k = 0; k < 9; k++) s += S.Kda[r * 10 + k] * S.Kaa[k * 9 + c];
    S.W[r * 10 + c] = s;
  }
}
// K_ij -= W_i * Kad_j   (static_LIB_3dIC.f90:213; Kad = Kda^T by symmetry of D).  The three Kda rows of
// node j are read once; every W row once.
FB_HD void ic_condense_block(const Scratch &S, int i, int j, double *k9) {
  double kd[3][9];
  for (int b = 0; b < 3; b++)
    for (int k = 0; k < 9; k++) kd[b][k] = S.Kda[(3 * j + b) * 10 + k];
  for (int a = 0; a < 3; a++) {
    double w[9];
    for (int k = 0; k < 9; k++) w[k] = S.W[(3 * i + a) * 10 + k];
    for (int b = 0; b < 3; b++) {
      double s = 0.0;
      for (int k = 0; k < 9; k++) s += w[k] * kd[b][k];
      k9[3 * a + b] -= s;
    }
  }
}

#ifdef FB_WARP_FN
// The whole element: executed by the 32 lanes of one warp after S.xc holds the element coordinates (and the
// warp has synchronised).  On return lane `l` holds in kb[r] the 3x3 block t = 32 r + l (i = t / nn, j = t % nn)
// of the element stiffness matrix.  This is the body of assembly.cuh:stiff_kernel AND of the host emulation.
// ICINV selects how the IC element gets W = Kda Kaa^-1: calInverse in registers + shuffles (1) or through shared
// memory (0), both in the reference's operation order, or LDL^T solves without forming the inverse (2).
template <int FORM, int ICINV = 1>
FB_WARP_FN void element_stiffness_warp(int lane, Scratch &S, double D11, double D12, double D44,
                                       double (&kb)[((FORM == FORM_TET10 ? 100 : 64) + 31) / 32][9],
                                       bool condense = true) {
  constexpr int NN = (FORM == FORM_TET10) ? 10 : 8;
  constexpr int NGP = (FORM == FORM_TET10) ? 4 : 8;
  constexpr int NB = NN * NN;
  constexpr int ROUNDS = (NB + 31) / 32;
  phase_geometry<FORM>(lane, S);
  FB_SYNCWARP();
  // ICINV == 2: the coupling blocks and W first, while the node blocks kb[] are not live yet (the LDL^T factors
  // occupy ~110 registers); the results are the same numbers in either order
  constexpr bool W_FIRST = (FORM == FORM_IC && ICINV == 2);
  bool have_w = false;
  if (W_FIRST && condense) {
    ic_phase_coupling(lane, S, D11, D12, D44);
    FB_SYNCWARP();
    have_w = ic_w_ldlt_warp(lane, S);     // warp-uniform result
  }
#pragma unroll
  for (int r = 0; r < ROUNDS; r++) {
    const int t = r * 32 + lane;
    if (t < NB) {
      if (FORM == FORM_BBAR) pair_block_bbar(S, t / NN, t % NN, D11, D12, D44, kb[r]);
      else pair_block_iso<(FORM == FORM_IC) ? 11 : NN, NGP>(S, t / NN, t % NN, D11, D12, D44, kb[r]);
    }
  }
  if (FORM == FORM_IC && condense) {
    if (!W_FIRST) {
      ic_phase_coupling(lane, S, D11, D12, D44);
      FB_SYNCWARP();
    }
    if (!have_w) {                        // reference-order paths (and the fall-back of the LDL^T kernel)
      if (ICINV == 1) ic_inverse_warp(lane, S);
      else ic_inverse_smem_warp(lane, S);
      FB_SYNCWARP();
      ic_phase_W(lane, S);
    }
    FB_SYNCWARP();
#pragma unroll
    for (int r = 0; r < ROUNDS; r++) {
      const int t = r * 32 + lane;
      if (t < NB) ic_condense_block(S, t / NN, t % NN, kb[r]);
    }
  }
}
#endif

// =============================================================================================
// Geometrically nonlinear path (total Lagrangian): STF_C3 / STF_C3D8Bbar with flag == TOTALLAG
// (static_LIB_3d.f90:47-205, static_LIB_C3D8.f90:23-198) and UPDATE_C3 / Update_C3D8Bbar
// (static_LIB_3d.f90:516-837, static_LIB_C3D8.f90:203-552) for ELASTIC and MOONEYRIVLIN materials.
// One element per warp as above; per-Gauss-point state (Green-Lagrange strain, 2nd Piola-Kirchhoff stress) lives in
// dense device arrays [element][8][6].  SURVEY 8(f) rank 2.
// =============================================================================================
struct alignas(16) ScratchTL {
  double stage[100 * 9];    // gather-mode output staging (must stay first: 16-byte aligned)
  double xc[10 * 3];        // element coordinates
  double uc[10 * 3];        // nodal displacement the deformation gradient is formed from
  double gd[8 * 10 * 3];    // gderiv[g][n][3]
  double bb[8 * 8 * 3];     // B-bar correction (Bbar - gderiv)/3
  double wdet[8];           // weight * det
  double F[8 * 9];          // dudx(i,j) per Gauss point, [g][3*i+j]
  double Dm[8 * 36];        // material tangent per Gauss point, row-major 6x6
  double sg[8 * 6];         // 2nd Piola-Kirchhoff stress per Gauss point (xx yy zz xy yz zx)
  double qf[8 * 30];        // update: internal-force contribution of every Gauss point
  double tc[10], t0c[10];   // linear update with !TEMPERATURE: nodal temperature / initial temperature
};

enum Material { MAT_ELASTIC = 0, MAT_MOONEYRIVLIN = 1 };

// reduced invariants of C = 2E + 1 and their derivatives (cderiv, lib/physics/Hyperelastic.f90:13-131);
// d2 == nullptr skips the second derivatives
FB_HD void mr_cderiv(const double *strain, double &inv3b, double (*dib)[3][3], double (*d2)[3][3][3][3]) {
  const double delta[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  double ctn[3][3], ci[3][3], didc[3][3][3];
  ctn[0][0] = strain[0] * 2.0 + 1.0; ctn[1][1] = strain[1] * 2.0 + 1.0; ctn[2][2] = strain[2] * 2.0 + 1.0;
  ctn[0][1] = ctn[1][0] = strain[3]; ctn[1][2] = ctn[2][1] = strain[4]; ctn[2][0] = ctn[0][2] = strain[5];
  const double inv1 = ctn[0][0] + ctn[1][1] + ctn[2][2];
  const double inv2 = ctn[1][1] * ctn[2][2] + ctn[0][0] * ctn[2][2] + ctn[0][0] * ctn[1][1] - ctn[1][2] * ctn[1][2] -
                      ctn[0][2] * ctn[0][2] - ctn[0][1] * ctn[0][1];
  const double inv3 = ctn[0][0] * ctn[1][1] * ctn[2][2] + ctn[1][0] * ctn[2][1] * ctn[0][2] + ctn[2][0] * ctn[0][1] * ctn[1][2] -
                      ctn[2][0] * ctn[1][1] * ctn[0][2] - ctn[1][0] * ctn[0][1] * ctn[2][2] - ctn[0][0] * ctn[2][1] * ctn[1][2];
  const double inv33 = pow(inv3, -1.0 / 3.0);
  ci[0][0] = (ctn[1][1] * ctn[2][2] - ctn[1][2] * ctn[1][2]) / inv3;
  ci[1][1] = (ctn[0][0] * ctn[2][2] - ctn[0][2] * ctn[0][2]) / inv3;
  ci[2][2] = (ctn[0][0] * ctn[1][1] - ctn[0][1] * ctn[0][1]) / inv3;
  ci[0][1] = ci[1][0] = (ctn[0][2] * ctn[1][2] - ctn[0][1] * ctn[2][2]) / inv3;
  ci[0][2] = ci[2][0] = (ctn[0][1] * ctn[1][2] - ctn[1][1] * ctn[0][2]) / inv3;
  ci[1][2] = ci[2][1] = (ctn[0][1] * ctn[0][2] - ctn[0][0] * ctn[1][2]) / inv3;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      didc[0][i][j] = delta[i][j];
      didc[1][i][j] = inv1 * delta[i][j] - ctn[i][j];
      didc[2][i][j] = inv3 * ci[i][j];
    }
  inv3b = sqrt(inv3);
  const double i2 = inv33 * inv33, i4 = i2 * i2, i5 = i4 * inv33, i7 = i5 * i2, i8 = i4 * i4;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      dib[0][i][j] = -i4 * inv1 * didc[2][i][j] / 3.0 + inv33 * didc[0][i][j];
      dib[1][i][j] = -2.0 * i5 * inv2 * didc[2][i][j] / 3.0 + i2 * didc[1][i][j];
      dib[2][i][j] = didc[2][i][j] / (2.0 * sqrt(inv3));
    }
  if (!d2) return;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      for (int k = 0; k < 3; k++)
        for (int l = 0; l < 3; l++) {
          const double h2 = delta[i][j] * delta[k][l] - (delta[i][k] * delta[j][l] + delta[i][l] * delta[j][k]) / 2.0;
          const double h3 = inv3 * (ci[k][l] * ci[i][j] - (ci[i][k] * ci[l][j] + ci[i][l] * ci[k][j]) / 2.0);
          d2[0][i][j][k][l] = 4.0 / 9.0 * i7 * inv1 * didc[2][i][j] * didc[2][k][l] -
                              i4 / 3.0 * (didc[0][k][l] * didc[2][i][j] + didc[0][i][j] * didc[2][k][l]) - i4 / 3.0 * inv1 * h3;
          d2[1][i][j][k][l] = 10.0 / 9.0 * i8 * inv2 * didc[2][i][j] * didc[2][k][l] -
                              2.0 / 3.0 * i5 * (didc[1][k][l] * didc[2][i][j] + didc[1][i][j] * didc[2][k][l]) -
                              2.0 / 3.0 * i5 * inv2 * h3 + i2 * h2;
          d2[2][i][j][k][l] = -didc[2][i][j] * didc[2][k][l] / (4.0 * pow(inv3, 1.5)) + h3 / (2.0 * sqrt(inv3));
        }
}

// calUpdateElasticMooneyRivlin (Hyperelastic.f90:250-284): S from the Green-Lagrange strain
FB_HD void mr_stress(const double *c, const double *strain, double *stress) {
  double i3, dib[3][3][3];
  mr_cderiv(strain, i3, dib, nullptr);
  const int vi[6] = {0, 1, 2, 0, 1, 0}, vj[6] = {0, 1, 2, 1, 2, 2};
  for (int a = 0; a < 6; a++) {
    const int k = vi[a], l = vj[a];
    stress[a] = 2.0 * (dib[0][k][l] * c[0] + dib[1][k][l] * c[1] + 2.0 * (i3 - 1.0) * dib[2][k][l] / c[2]);
  }
}

// calElasticMooneyRivlin + mat_c2d (Hyperelastic.f90:221-246, calMatMatrix.f90:154-197): D(6,6) row-major
FB_HD void mr_tangent(const double *c, const double *strain, double *D) {
  double i3, dib[3][3][3], d2[3][3][3][3][3];
  mr_cderiv(strain, i3, dib, d2);
  const int vi[6] = {0, 1, 2, 0, 1, 2}, vj[6] = {0, 1, 2, 1, 2, 0};   // 11 22 33 12 23 31
  for (int a = 0; a < 6; a++)
    for (int b = 0; b < 6; b++) {
      const int k = vi[a], l = vj[a], m = vi[b], n = vj[b];
      D[6 * a + b] = 4.0 * (d2[0][k][l][m][n] * c[0] + d2[1][k][l][m][n] * c[1] +
                            2.0 * (dib[2][k][l] * dib[2][m][n] + (i3 - 1.0) * d2[2][k][l][m][n]) / c[2]);
    }
}

// per-section material record handed to the TL kernels
struct MatlTL {
  int mtype;             // MAT_ELASTIC / MAT_MOONEYRIVLIN
  double D11, D12, D44;  // elastic constants (ElasticLinear.f90:48-59)
  double c[3];           // Mooney-Rivlin C10, C01, D
};

FB_HD void tl_tangent(const MatlTL &M, const double *strain, double *D) {   // MatlMatrix, calMatMatrix.f90:29-112
  if (M.mtype == MAT_MOONEYRIVLIN) { mr_tangent(M.c, strain, D); return; }
  for (int a = 0; a < 36; a++) D[a] = 0.0;
  D[0] = D[7] = D[14] = M.D11;
  D[1] = D[2] = D[6] = D[8] = D[12] = D[13] = M.D12;
  D[21] = D[28] = D[35] = M.D44;
}

// BL = BL0 + BL1 of one node (6x3): gderiv g[3], B-bar correction bbn[3] (nullptr: plain), dudx F[9]
FB_HD void tl_build_B(const double *g, const double *bbn, const double *F, double (*B)[3]) {
  for (int c = 0; c < 3; c++) {
    const double d0 = F[3 * c + 0], d1 = F[3 * c + 1], d2 = F[3 * c + 2];   // gdispderiv(c,1..3)
    const double vol = bbn ? bbn[c] : 0.0;
    B[0][c] = vol + (c == 0 ? g[0] : 0.0) + d0 * g[0];
    B[1][c] = vol + (c == 1 ? g[1] : 0.0) + d1 * g[1];
    B[2][c] = vol + (c == 2 ? g[2] : 0.0) + d2 * g[2];
    B[3][c] = (c == 0 ? g[1] : (c == 1 ? g[0] : 0.0)) + d1 * g[0] + d0 * g[1];
    B[4][c] = (c == 1 ? g[2] : (c == 2 ? g[1] : 0.0)) + d1 * g[2] + d2 * g[1];
    B[5][c] = (c == 0 ? g[2] : (c == 2 ? g[0] : 0.0)) + d2 * g[0] + d0 * g[2];
  }
}

// lane g < NGP: dudx at Gauss point g from S.uc; optionally the tangent at the stored strain and the stored stress
template <int FORM>
FB_HD void tl_phase_state(int lane, ScratchTL &S, const MatlTL &M, const double *gp_strain, const double *gp_stress) {
  constexpr int NN = (FORM == FORM_TET10) ? 10 : 8;
  constexpr int NGP = (FORM == FORM_TET10) ? 4 : 8;
  if (lane >= NGP) return;
  const double *gd = S.gd + lane * NN * 3;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0.0;
      for (int n = 0; n < NN; n++) s += S.uc[3 * n + i] * gd[3 * n + j];
      S.F[9 * lane + 3 * i + j] = s;
    }
  if (gp_strain) tl_tangent(M, gp_strain + 6 * lane, S.Dm + 36 * lane);
  if (gp_stress)
    for (int k = 0; k < 6; k++) S.sg[6 * lane + k] = gp_stress[6 * lane + k];
}

// 3x3 block (i,j) of the tangent stiffness: sum_g w (B_i^T D B_j + (g_i^T S g_j) 1)
template <int FORM>
FB_HD void tl_pair_block(const ScratchTL &S, int i, int j, double *k9) {
  constexpr int NN = (FORM == FORM_TET10) ? 10 : 8;
  constexpr int NGP = (FORM == FORM_TET10) ? 4 : 8;
  for (int a = 0; a < 9; a++) k9[a] = 0.0;
  for (int g = 0; g < NGP; g++) {
    const double *gi = S.gd + (g * NN + i) * 3, *gj = S.gd + (g * NN + j) * 3;
    const double *F = S.F + 9 * g, *D = S.Dm + 36 * g, *sg = S.sg + 6 * g;
    const double w = S.wdet[g];
    double Bi[6][3], Bj[6][3];
    tl_build_B(gi, FORM == FORM_BBAR ? S.bb + (g * 8 + i) * 3 : nullptr, F, Bi);
    tl_build_B(gj, FORM == FORM_BBAR ? S.bb + (g * 8 + j) * 3 : nullptr, F, Bj);
    for (int b = 0; b < 3; b++) {
      double DB[6];
      for (int r = 0; r < 6; r++) {
        double s = 0.0;
        for (int k = 0; k < 6; k++) s += D[6 * r + k] * Bj[k][b];
        DB[r] = s;
      }
      for (int a = 0; a < 3; a++) {
        double s = 0.0;
        for (int r = 0; r < 6; r++) s += Bi[r][a] * DB[r];
        k9[3 * a + b] += s * w;
      }
    }
    // initial-stress matrix BN^T S BN (static_LIB_3d.f90:170-201)
    const double t0 = sg[0] * gj[0] + sg[3] * gj[1] + sg[5] * gj[2];
    const double t1 = sg[3] * gj[0] + sg[1] * gj[1] + sg[4] * gj[2];
    const double t2 = sg[5] * gj[0] + sg[4] * gj[1] + sg[2] * gj[2];
    const double geo = (gi[0] * t0 + gi[1] * t1 + gi[2] * t2) * w;
    k9[0] += geo; k9[4] += geo; k9[8] += geo;
  }
}

// lane g < NGP of the update: Green-Lagrange strain, stress, internal-force contribution of Gauss point g
// (S.uc = unode + dunode).  Writes strain/stress of the point to gp_strain/gp_stress (element-local [g][6]).
template <int FORM>
FB_HD void tl_phase_update(int lane, ScratchTL &S, const MatlTL &M, double *gp_strain, double *gp_stress) {
  constexpr int NN = (FORM == FORM_TET10) ? 10 : 8;
  constexpr int NGP = (FORM == FORM_TET10) ? 4 : 8;
  if (lane >= NGP) return;
  const double *gd = S.gd + lane * NN * 3;
  const double *F = S.F + 9 * lane;
  double dvol = 0.0;
  if (FORM == FORM_BBAR) {   // dilatation at the centroid, static_LIB_C3D8.f90:268-272,291
    const double *bb = S.bb + lane * 8 * 3;
    double tr = 0.0;
    for (int n = 0; n < 8; n++)
      for (int c = 0; c < 3; c++) tr += S.uc[3 * n + c] * (gd[3 * n + c] + 3.0 * bb[3 * n + c]);   // Bbar = gderiv + 3 bb
    dvol = tr / 3.0 - (F[0] + F[4] + F[8]) / 3.0;
  }
  double e[6];
  e[0] = F[0] + dvol; e[1] = F[4] + dvol; e[2] = F[8] + dvol;
  e[3] = F[1] + F[3]; e[4] = F[5] + F[7]; e[5] = F[6] + F[2];
  e[0] += 0.5 * (F[0] * F[0] + F[3] * F[3] + F[6] * F[6]);
  e[1] += 0.5 * (F[1] * F[1] + F[4] * F[4] + F[7] * F[7]);
  e[2] += 0.5 * (F[2] * F[2] + F[5] * F[5] + F[8] * F[8]);
  e[3] += F[0] * F[1] + F[3] * F[4] + F[6] * F[7];
  e[4] += F[1] * F[2] + F[4] * F[5] + F[7] * F[8];
  e[5] += F[0] * F[2] + F[3] * F[5] + F[6] * F[8];
  double sg[6];
  if (M.mtype == MAT_MOONEYRIVLIN) mr_stress(M.c, e, sg);
  else {
    double D[36];
    tl_tangent(M, e, D);
    for (int r = 0; r < 6; r++) {
      double s = 0.0;
      for (int k = 0; k < 6; k++) s += D[6 * r + k] * e[k];
      sg[r] = s;
    }
  }
  for (int k = 0; k < 6; k++) { gp_strain[6 * lane + k] = e[k]; gp_stress[6 * lane + k] = sg[k]; }
  const double w = S.wdet[lane];
  for (int n = 0; n < NN; n++) {
    double B[6][3];
    tl_build_B(gd + 3 * n, FORM == FORM_BBAR ? S.bb + (lane * 8 + n) * 3 : nullptr, F, B);
    for (int c = 0; c < 3; c++) {
      double s = 0.0;
      for (int r = 0; r < 6; r++) s += sg[r] * B[r][c];
      S.qf[30 * lane + 3 * n + c] = s * w;
    }
  }
}

#ifdef FB_WARP_FN
// Tangent stiffness of one element (all 32 lanes; S.xc and S.uc = unode + dunode filled and synchronised).
template <int FORM>
FB_WARP_FN void element_stiffness_tl_warp(int lane, ScratchTL &S, const MatlTL &M, const double *gp_strain,
                                          const double *gp_stress, double (&kb)[((FORM == FORM_TET10 ? 100 : 64) + 31) / 32][9]) {
  constexpr int NN = (FORM == FORM_TET10) ? 10 : 8;
  constexpr int NB = NN * NN;
  constexpr int ROUNDS = (NB + 31) / 32;
  phase_geometry<FORM>(lane, S);
  FB_SYNCWARP();
  tl_phase_state<FORM>(lane, S, M, gp_strain, gp_stress);
  FB_SYNCWARP();
#pragma unroll
  for (int r = 0; r < ROUNDS; r++) {
    const int t = r * 32 + lane;
    if (t < NB) tl_pair_block<FORM>(S, t / NN, t % NN, kb[r]);
  }
}

// Strain / stress update + internal force of one element (S.xc, S.uc = unode + dunode filled and synchronised).
// On return lane t < 3 nn holds qf[t] = internal force of local dof t.
template <int FORM>
FB_WARP_FN double element_update_tl_warp(int lane, ScratchTL &S, const MatlTL &M, double *gp_strain, double *gp_stress) {
  constexpr int NN = (FORM == FORM_TET10) ? 10 : 8;
  constexpr int NGP = (FORM == FORM_TET10) ? 4 : 8;
  phase_geometry<FORM>(lane, S);
  FB_SYNCWARP();
  tl_phase_state<FORM>(lane, S, M, nullptr, nullptr);
  FB_SYNCWARP();
  tl_phase_update<FORM>(lane, S, M, gp_strain, gp_stress);
  FB_SYNCWARP();
  double q = 0.0;
  if (lane < 3 * NN)
    for (int g = 0; g < NGP; g++) q += S.qf[30 * g + lane];
  return q;
}
#endif

// =============================================================================================
// Strain / stress recovery of the LINEAR static analysis: fstr_UpdateNewton with nlgeom_flag == INFINITE
// (fstr_Update.f90:25-293): UPDATE_C3 (static_LIB_3d.f90:516-837) for 361 FI / 342, Update_C3D8Bbar
// (static_LIB_C3D8.f90:203-552), UpdateST_C3D8IC (static_LIB_3dIC.f90:220-455).  Isotropic elastic material, thermal
// strain alpha (T - Tref) - alpha (T0 - Tref) when nodal temperatures are given.  SURVEY 8(f) rank 2.
// =============================================================================================
FB_HD void hex_shape_func(double x, double y, double z, double *f) {   // ShapeFunc_hex8n, hex8n.f90:9-21
  const double sx[8] = {-1, 1, 1, -1, -1, 1, 1, -1};
  const double sy[8] = {-1, -1, 1, 1, -1, -1, 1, 1};
  const double sz[8] = {-1, -1, -1, -1, 1, 1, 1, 1};
  for (int n = 0; n < 8; n++) f[n] = 0.125 * (1.0 + sx[n] * x) * (1.0 + sy[n] * y) * (1.0 + sz[n] * z);
}
FB_HD void tet10_shape_func(double xi, double et, double ze, double *s) {   // ShapeFunc_tet10n, tet10n.f90:9-26
  const double a = 1.0 - xi - et - ze;
  s[0] = (2.0 * a - 1.0) * a; s[1] = xi * (2.0 * xi - 1.0); s[2] = et * (2.0 * et - 1.0); s[3] = ze * (2.0 * ze - 1.0);
  s[4] = 4.0 * xi * a; s[5] = 4.0 * xi * et; s[6] = 4.0 * et * a; s[7] = 4.0 * ze * a; s[8] = 4.0 * xi * ze; s[9] = 4.0 * et * ze;
}
// thermal strain at Gauss point g from the nodal temperatures (static_LIB_3d.f90:597-633)
template <int FORM>
FB_HD double gauss_thermal_strain(int g, const double *tc, const double *t0c, double alpha, double ref_temp) {
  constexpr int NN = (FORM == FORM_TET10) ? 10 : 8;
  double H[10];
  if (FORM == FORM_TET10) tet10_shape_func((g == 1) ? kTetB : kTetA, (g == 2) ? kTetB : kTetA, (g == 3) ? kTetB : kTetA, H);
  else hex_shape_func((g & 1) ? kGaussHex : -kGaussHex, (g & 2) ? kGaussHex : -kGaussHex, (g & 4) ? kGaussHex : -kGaussHex, H);
  double ttc = 0.0, tt0 = 0.0;
  for (int n = 0; n < NN; n++) { ttc += tc[n] * H[n]; tt0 += t0c[n] * H[n]; }
  return alpha * (ttc - ref_temp) - alpha * (tt0 - ref_temp);
}

// lane g < NGP of UPDATE_C3 / Update_C3D8Bbar with flag == INFINITE; S.F holds dudx of the point (tl_phase_state)
template <int FORM>
FB_HD void lin_phase_update(int lane, ScratchTL &S, double D11, double D12, double D44, double alpha, double ref_temp,
                            bool has_temp, double *gp_strain, double *gp_stress) {
  constexpr int NN = (FORM == FORM_TET10) ? 10 : 8;
  constexpr int NGP = (FORM == FORM_TET10) ? 4 : 8;
  if (lane >= NGP) return;
  const double *gd = S.gd + lane * NN * 3;
  const double *F = S.F + 9 * lane;
  double dvol = 0.0;
  if (FORM == FORM_BBAR) {   // dilatation at the centroid, static_LIB_C3D8.f90:270-274,295
    const double *bb = S.bb + lane * 8 * 3;
    double tr = 0.0;
    for (int n = 0; n < 8; n++)
      for (int c = 0; c < 3; c++) tr += S.uc[3 * n + c] * (gd[3 * n + c] + 3.0 * bb[3 * n + c]);
    dvol = tr / 3.0 - (F[0] + F[4] + F[8]) / 3.0;
  }
  const double eth = has_temp ? gauss_thermal_strain<FORM>(lane, S.tc, S.t0c, alpha, ref_temp) : 0.0;
  double ds[6];
  ds[0] = (F[0] + dvol) - eth; ds[1] = (F[4] + dvol) - eth; ds[2] = (F[8] + dvol) - eth;
  ds[3] = F[1] + F[3]; ds[4] = F[5] + F[7]; ds[5] = F[6] + F[2];
  double sg[6];
  sg[0] = D11 * ds[0] + D12 * ds[1] + D12 * ds[2];
  sg[1] = D12 * ds[0] + D11 * ds[1] + D12 * ds[2];
  sg[2] = D12 * ds[0] + D12 * ds[1] + D11 * ds[2];
  sg[3] = D44 * ds[3]; sg[4] = D44 * ds[4]; sg[5] = D44 * ds[5];
  for (int k = 0; k < 6; k++) { gp_strain[6 * lane + k] = ds[k] + (k < 3 ? eth : 0.0); gp_stress[6 * lane + k] = sg[k]; }
  const double w = S.wdet[lane];
  const double zeroF[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // BL0 only
  for (int n = 0; n < NN; n++) {
    double B[6][3];
    tl_build_B(gd + 3 * n, FORM == FORM_BBAR ? S.bb + (lane * 8 + n) * 3 : nullptr, zeroF, B);
    for (int c = 0; c < 3; c++) {
      double s = 0.0;
      for (int r = 0; r < 6; r++) s += sg[r] * B[r][c];
      S.qf[30 * lane + 3 * n + c] = s * w;
    }
  }
}

#ifdef FB_WARP_FN
// One element of the linear update (361 FI / B-bar, 342): S.xc, S.uc (= unode + dunode) and, if has_temp, S.tc / S.t0c
// filled and synchronised.  Writes the strain / stress of every Gauss point; lane t < 3 nn returns qf[t].
template <int FORM>
FB_WARP_FN double element_update_lin_warp(int lane, ScratchTL &S, double D11, double D12, double D44, double alpha,
                                          double ref_temp, bool has_temp, double *gp_strain, double *gp_stress) {
  constexpr int NN = (FORM == FORM_TET10) ? 10 : 8;
  constexpr int NGP = (FORM == FORM_TET10) ? 4 : 8;
  MatlTL M;
  M.mtype = MAT_ELASTIC; M.D11 = D11; M.D12 = D12; M.D44 = D44; M.c[0] = M.c[1] = M.c[2] = 0.0;
  phase_geometry<FORM>(lane, S);
  FB_SYNCWARP();
  tl_phase_state<FORM>(lane, S, M, nullptr, nullptr);   // dudx per Gauss point
  FB_SYNCWARP();
  lin_phase_update<FORM>(lane, S, D11, D12, D44, alpha, ref_temp, has_temp, gp_strain, gp_stress);
  FB_SYNCWARP();
  double q = 0.0;
  if (lane < 3 * NN)
    for (int g = 0; g < NGP; g++) q += S.qf[30 * g + lane];
  return q;
}

// UpdateST_C3D8IC: S.xc, S.uc (and S.tc / S.t0c) filled and synchronised.  The enhanced-mode amplitudes are
// recovered from the element's own K(d,alpha), K(alpha,alpha) exactly as the reference does (:329-335), then strain
// = B [u; alpha] at the 8 Gauss points.  Lane t < 24 returns qf[t] = (K [u; alpha])_t (- thermal load vector).
template <int ICINV>
FB_WARP_FN double element_update_ic_warp(int lane, Scratch &S, double D11, double D12, double D44, double alpha,
                                         double ref_temp, bool has_temp, double *gp_strain, double *gp_stress) {
  double kb[2][9];
  element_stiffness_warp<FORM_IC, ICINV>(lane, S, D11, D12, D44, kb, false);   // K(d,d) blocks, not condensed
  ic_phase_coupling(lane, S, D11, D12, D44);
  FB_SYNCWARP();
  if (ICINV == 1) ic_inverse_warp(lane, S);
  else ic_inverse_smem_warp(lane, S);
#pragma unroll
  for (int r = 0; r < 2; r++) {
    const int t = r * 32 + lane;
#pragma unroll
    for (int a = 0; a < 9; a++) S.stage[t * 9 + a] = kb[r][a];
  }
  if (has_temp && lane < 8) S.eth[lane] = gauss_thermal_strain<FORM_IC>(lane, S.tc, S.t0c, alpha, ref_temp);
  FB_SYNCWARP();
  if (lane < 9) {   // tmpforce = Kad u, :332
    double s = 0.0;
    for (int r = 0; r < 24; r++) s += S.Kda[r * 10 + lane] * S.uc[r];
    S.tf[lane] = s;
  }
  FB_SYNCWARP();
  if (lane < 9) {   // cdisp(25:33) = -Kaa^-1 tmpforce, :335
    double s = 0.0;
    for (int k = 0; k < 9; k++) s += S.Kaa[lane * 9 + k] * S.tf[k];
    S.cd[lane] = -s;
  }
  if (has_temp) {   // TLOAD_C3D8IC tmpvect (static_LIB_3dIC.f90:560-590): thermal stress (D11 + 2 D12) eth on the normal rows
    for (int j = lane; j < 33; j += 32) {
      const int n = j / 3, c = j % 3;
      double s = 0.0;
      for (int g = 0; g < 8; g++) s += ((D11 + D12 + D12) * S.eth[g]) * S.gd[(g * 11 + n) * 3 + c] * S.wdet[g];
      S.tv[j] = s;
    }
  }
  FB_SYNCWARP();
  double q = 0.0;
  if (lane < 24) {   // qf = K(1:24, 1:33) cdisp, :338
    const int i = lane / 3, a = lane % 3;
    for (int j = 0; j < 8; j++)
      for (int b = 0; b < 3; b++) q += S.stage[(i * 8 + j) * 9 + 3 * a + b] * S.uc[3 * j + b];
    for (int k = 0; k < 9; k++) q += S.Kda[lane * 10 + k] * S.cd[k];
  }
  if (has_temp) {   // qf -= TLOAD_C3D8IC = tmpvect(1:24) - Kda Kaa^-1 tmpvect(25:33), :340-342
    FB_SYNCWARP();
    if (lane < 9) {
      double s = 0.0;
      for (int k = 0; k < 9; k++) s += S.Kaa[lane * 9 + k] * S.tv[24 + k];
      S.tf[lane] = s;
    }
    FB_SYNCWARP();
    if (lane < 24) {
      double s = 0.0;
      for (int k = 0; k < 9; k++) s += S.Kda[lane * 10 + k] * S.tf[k];
      q -= S.tv[lane] - s;
    }
  }
  if (lane < 8) {   // strain = B cdisp, stress = D (strain - eth), :346-452
    const double *gd = S.gd + lane * 11 * 3;
    double e[6] = {0, 0, 0, 0, 0, 0};
    for (int n = 0; n < 11; n++) {
      const double *u = (n < 8) ? S.uc + 3 * n : S.cd + 3 * (n - 8);
      const double g1 = gd[3 * n], g2 = gd[3 * n + 1], g3 = gd[3 * n + 2];
      e[0] += g1 * u[0]; e[1] += g2 * u[1]; e[2] += g3 * u[2];
      e[3] += g2 * u[0]; e[3] += g1 * u[1];
      e[4] += g3 * u[1]; e[4] += g2 * u[2];
      e[5] += g3 * u[0]; e[5] += g1 * u[2];
    }
    const double eth = has_temp ? S.eth[lane] : 0.0;
    const double d0 = e[0] - eth, d1 = e[1] - eth, d2 = e[2] - eth;
    double *eo = gp_strain + 6 * lane, *so = gp_stress + 6 * lane;
    for (int k = 0; k < 6; k++) eo[k] = e[k];
    so[0] = D11 * d0 + D12 * d1 + D12 * d2;
    so[1] = D12 * d0 + D11 * d1 + D12 * d2;
    so[2] = D12 * d0 + D12 * d1 + D11 * d2;
    so[3] = D44 * e[3]; so[4] = D44 * e[4]; so[5] = D44 * e[5];
  }
  return q;
}
#endif

// ---- fstr_NodalStress3D (fstr_NodalStress.f90:13-272) for 361 / 342 -------------------------------------------
// Extrapolation matrices: the inverse (inverse_func, :714-748: Gauss-Jordan without pivoting) of the hex8n (342: tet4n)
// shape functions evaluated at the Gauss points, :64-81.
struct NodalExtrap {
  double inv8[64];   // 361: [node][gauss point]
  double inv4[16];   // 342: [vertex][gauss point]
};
FB_HD void inverse_func(int n, double *a, double *inv_a) {
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) inv_a[i * n + j] = (i == j) ? 1.0 : 0.0;
  for (int i = 0; i < n; i++) {
    double buf = 1.0 / a[i * n + i];
    for (int j = 0; j < n; j++) { a[i * n + j] = a[i * n + j] * buf; inv_a[i * n + j] = inv_a[i * n + j] * buf; }
    for (int j = 0; j < n; j++)
      if (i != j) {
        buf = a[j * n + i];
        for (int k = 0; k < n; k++) {
          a[j * n + k] = a[j * n + k] - a[i * n + k] * buf;
          inv_a[j * n + k] = inv_a[j * n + k] - inv_a[i * n + k] * buf;
        }
      }
  }
}
inline void nodal_extrap_build(NodalExtrap &X) {   // host, once
  double f8[64], f4[16];
  for (int g = 0; g < 8; g++)
    hex_shape_func((g & 1) ? kGaussHex : -kGaussHex, (g & 2) ? kGaussHex : -kGaussHex, (g & 4) ? kGaussHex : -kGaussHex, f8 + 8 * g);
  inverse_func(8, f8, X.inv8);
  for (int g = 0; g < 4; g++) {   // ShapeFunc_tet4n at the 342 Gauss points, tet4n.f90:11-16
    const double xi = (g == 1) ? kTetB : kTetA, et = (g == 2) ? kTetB : kTetA, ze = (g == 3) ? kTetB : kTetA;
    f4[4 * g] = 1.0 - xi - et - ze; f4[4 * g + 1] = xi; f4[4 * g + 2] = et; f4[4 * g + 3] = ze;
  }
  inverse_func(4, f4, X.inv4);
}
// value of component c (of [gp][6] data `src` of ONE element) extrapolated to local node j (NodalStress_INV3, :337-400)
FB_HD double nodal_extrapolate(const NodalExtrap &X, int nn, int j, const double *src, int c) {
  if (nn == 8) {
    double v = 0.0;
    for (int i = 0; i < 8; i++) v = v + X.inv8[j * 8 + i] * src[6 * i + c];
    return v;
  }
  const int ma[6] = {0, 1, 2, 0, 1, 2}, mb[6] = {1, 2, 0, 3, 3, 3};   // mid-edge nodes 5..10 of tet10n, :361-378
  const int a = (j < 4) ? j : ma[j - 4], b = (j < 4) ? j : mb[j - 4];
  double va = 0.0, vb = 0.0;
  for (int i = 0; i < 4; i++) va = va + X.inv4[a * 4 + i] * src[6 * i + c];
  if (j < 4) return va;
  for (int i = 0; i < 4; i++) vb = vb + X.inv4[b * 4 + i] * src[6 * i + c];
  return (va + vb) / 2.0;
}
FB_HD double mises_of(const double *s) {   // get_mises, fstr_NodalStress.f90:483-500
  const double ps = (s[0] + s[1] + s[2]) / 3.0;
  const double sm = 0.5 * ((s[0] - ps) * (s[0] - ps) + (s[1] - ps) * (s[1] - ps) + (s[2] - ps) * (s[2] - ps)) +
                    s[3] * s[3] + s[4] * s[4] + s[5] * s[5];
  return sqrt(3.0 * sm);
}

// isotropic elastic constants, ElasticLinear.f90:48-59
FB_HD void elastic_constants(double EE, double PP, double &D11, double &D12, double &D44) {
  D11 = EE * (1.0 - PP) / (1.0 - 2.0 * PP) / (1.0 + PP);
  D12 = EE * PP / (1.0 - 2.0 * PP) / (1.0 + PP);
  D44 = EE / (1.0 + PP) * 0.5;
}

}  // namespace elem
}  // namespace fb
