// Multicolour SSOR for hecmw_precond_SSOR_33 (hecmw1/src/solver/precond/33/hecmw_precond_SSOR_33.f90:55-467).
//
// The reference picks its row ordering by OpenMP thread count: natural order with one thread, RCM +
// greedy multicolouring (NCOLOR_IN) with more (SSOR_33.f90:93-115, matrix/hecmw_matrix_ordering_CM.f90,
// hecmw_matrix_ordering_MC.f90).  A GPU is the many-thread case, so this path always uses the
// multicolour ordering; rows of one colour are mutually independent and are swept by one kernel
// launch, one thread per row, with the same per-row accumulation order as the reference.
//
// As in the reference the preconditioner owns a permuted private copy restricted to the internal
// rows/columns (halo columns dropped, hecmw_matrix_reorder.f90:19-65) whose column ids are mapped
// back to the OLD numbering so the sweeps index the work vector directly (:142-157).
#pragma once
#include <algorithm>
#include <numeric>
#include <vector>

#include "common.cuh"

namespace fb {

// ---- host: orderings (sequential graph algorithms, run once per structure) -------------------
inline void ssor_cm_inner(int N, const int *indexL, const int *itemL, const int *indexU, const int *itemU,
                          int nstart, int &nlevel, std::vector<int> &lv_item) {
  // ordering_CM_inner, hecmw_matrix_ordering_CM.f90:64-144 — level-set sweep from nstart
  std::vector<int> iwk((size_t)N + 1, 0), lv_index((size_t)N + 1, 0);
  lv_item.assign(N, 0);
  iwk[nstart] = 1;
  int cntall = 1;
  lv_item[0] = nstart;
  lv_index[1] = 1;
  nlevel = 1;
  for (int level = 2; level <= N; level++) {
    int cnt = 0;
    bool full = false;
    for (int j = lv_index[level - 2]; j < lv_index[level - 1] && !full; j++) {
      const int jnode = lv_item[j];
      for (int k = indexL[jnode - 1]; k < indexL[jnode]; k++) {
        const int kn = itemL[k];
        if (iwk[kn] == 0) { iwk[kn] = level; cnt++; lv_item[cntall++] = kn; if (cntall == N) { full = true; break; } }
      }
      if (full) break;
      for (int k = indexU[jnode - 1]; k < indexU[jnode]; k++) {
        const int kn = itemU[k];
        if (kn > N) continue;
        if (iwk[kn] == 0) { iwk[kn] = level; cnt++; lv_item[cntall++] = kn; if (cntall == N) { full = true; break; } }
      }
    }
    if (cnt == 0) {
      for (int kn = 1; kn <= N; kn++)
        if (iwk[kn] == 0) { iwk[kn] = level; cnt++; lv_item[cntall++] = kn; break; }
    }
    lv_index[level] = cntall;
    if (cntall == N) { nlevel = level; break; }
  }
}

// perm[new-1] = old (1-based), iperm[old-1] = new; colour c owns new rows color_index[c]..color_index[c+1]-1
inline void ssor_multicolor_order(int N, const int *indexL, const int *itemL, const int *indexU,
                                  const int *itemU, int ncolor_in, std::vector<int> &perm,
                                  std::vector<int> &iperm, std::vector<int> &color_index) {
  perm.assign(N, 0); iperm.assign(N, 0);
  std::vector<int> cur(N);
  // find_minimum_degrees (:146-176): up to 5 start nodes of minimum (internal) degree
  int mins[5], nmin = 0, degmin = N;
  for (int i = 1; i <= N; i++) {
    int deg = indexL[i] - indexL[i - 1];
    for (int j = indexU[i - 1]; j < indexU[i]; j++) deg += (itemU[j] <= N);
    if (deg == 0) continue;
    if (deg < degmin) { degmin = deg; nmin = 1; mins[0] = i; }
    else if (deg == degmin) { nmin++; if (nmin <= 5) mins[nmin - 1] = i; }
  }
  if (nmin > 5) nmin = 5;
  if (nmin == 0) std::iota(cur.begin(), cur.end(), 1);
  else {
    int best_level = -1;
    std::vector<int> lv;
    for (int s = 0; s < nmin; s++) {
      int nlevel;
      ssor_cm_inner(N, indexL, itemL, indexU, itemU, mins[s], nlevel, lv);
      if (nlevel > best_level) { best_level = nlevel; cur = lv; }   // first maximum wins (:44-51)
    }
  }
  // reverse_ordering (:178-188) as written in the reference: the node ID is mirrored
  for (int i = 0; i < N; i++) cur[i] = N + 1 - cur[i];
  // hecmw_matrix_ordering_MC (hecmw_matrix_ordering_MC.f90:15-71)
  std::vector<int> iwk((size_t)N + 1, 0);
  const int nn_color = N / std::max(ncolor_in, 1);
  int cntall = 0;
  color_index.assign(1, 0);
  for (int color = 1; color <= N; color++) {
    int cnt = 0;
    for (int i = 0; i < N; i++) {
      const int inode = cur[i];
      if (iwk[inode] > 0 || iwk[inode] == -1) continue;
      iwk[inode] = color;
      perm[cntall++] = inode;
      cnt++;
      if (cnt == nn_color) break;
      if (cntall == N) break;
      for (int j = indexL[inode - 1]; j < indexL[inode]; j++) if (iwk[itemL[j]] == 0) iwk[itemL[j]] = -1;
      for (int j = indexU[inode - 1]; j < indexU[inode]; j++) {
        const int jn = itemU[j];
        if (jn > N) continue;
        if (iwk[jn] == 0) iwk[jn] = -1;
      }
    }
    color_index.push_back(cntall);
    if (cntall == N) break;
    for (int i = 1; i <= N; i++) if (iwk[i] == -1) iwk[i] = 0;
  }
  if (N == 0) color_index.push_back(0);
  for (int i = 0; i < N; i++) iperm[perm[i] - 1] = i + 1;
}

// one triangular part in sliced-ELL form; slices never straddle a colour
struct SsorPart {
  int nslice = 0;
  long long total = 0;                   // block columns over all slices
  std::vector<int> color_slice;          // ncolor+1 slice offsets
  DevBuf<long long> slice_ptr;           // nslice+1 block-column offsets
  DevBuf<int> slice_row0;                // first NEW row (0-based) of the slice
  DevBuf<int> slice_nrow;                // rows in the slice (<=32)
  DevBuf<double> val;                    // [p][9][32]
  DevBuf<int> col;                       // [p][32]  OLD node id, 0-based
  DevBuf<unsigned> src;                  // [p][32]  source slot in the unified value array, or kBad
};

struct Ssor {
  int N = 0, ncolor = 0;
  bool structure_valid = false, values_valid = false;
  std::vector<int> perm, iperm, color_index;
  DevBuf<int> d_perm;                    // new -> old, 0-based
  SsorPart L, U;
  DevBuf<double> alu;                    // SoA [9][ld] by NEW row
  size_t ld = 0;
};

constexpr unsigned kSsorPad = 0xFFFFFFFFu;

// host: build the permuted profile (hecmw_matrix_reorder_profile + _renum_item) as SELL slices
inline void ssor_build_part(const Ssor &S, bool upper, int NP, int NPL, const int *indexL, const int *itemL,
                            const int *indexU, const int *itemU, std::vector<long long> &slice_ptr,
                            std::vector<int> &row0, std::vector<int> &nrow, std::vector<int> &col,
                            std::vector<unsigned> &src, std::vector<int> &color_slice) {
  const int N = S.N;
  // per new row: entries (jnew, old col, source slot), sorted by jnew ascending
  struct Ent { int jnew, jold; unsigned slot; };
  std::vector<std::vector<Ent>> rows(N);
  for (int inew = 1; inew <= N; inew++) {
    const int iold = S.perm[inew - 1];
    auto &r = rows[inew - 1];
    for (int j = indexL[iold - 1]; j < indexL[iold]; j++) {
      const int jold = itemL[j], jnew = S.iperm[jold - 1];
      if ((jnew < inew) != upper) r.push_back({jnew, jold, (unsigned)NP + (unsigned)j});
    }
    for (int j = indexU[iold - 1]; j < indexU[iold]; j++) {
      const int jold = itemU[j];
      if (jold > N) continue;
      const int jnew = S.iperm[jold - 1];
      if ((jnew < inew) != upper) r.push_back({jnew, jold, (unsigned)NP + (unsigned)NPL + (unsigned)j});
    }
    std::sort(r.begin(), r.end(), [](const Ent &a, const Ent &b) { return a.jnew < b.jnew; });
    if (upper) std::reverse(r.begin(), r.end());   // backward sweep runs j = ieU..isU (SSOR_33.f90:366)
  }
  slice_ptr.assign(1, 0); row0.clear(); nrow.clear(); color_slice.assign(1, 0);
  for (int c = 0; c < S.ncolor; c++) {
    for (int r = S.color_index[c]; r < S.color_index[c + 1]; r += 32) {
      const int n = std::min(32, S.color_index[c + 1] - r);
      size_t w = 0;
      for (int l = 0; l < n; l++) w = std::max(w, rows[r + l].size());
      row0.push_back(r); nrow.push_back(n);
      slice_ptr.push_back(slice_ptr.back() + (long long)w);
    }
    color_slice.push_back((int)row0.size());
  }
  const size_t total = (size_t)slice_ptr.back();
  col.assign(total * 32, 0); src.assign(total * 32, kSsorPad);
  for (size_t s = 0; s < row0.size(); s++)
    for (int l = 0; l < nrow[s]; l++) {
      const auto &r = rows[row0[s] + l];
      for (size_t j = 0; j < r.size(); j++) {
        const size_t p = (size_t)slice_ptr[s] + j;
        col[p * 32 + l] = r[j].jold - 1;
        src[p * 32 + l] = r[j].slot;
      }
    }
}

// device: gather values of a part from the unified value array
__global__ void __launch_bounds__(256)
ssor_fill_kernel(size_t total32, const unsigned *__restrict__ src, const double *__restrict__ vals,
                 double *__restrict__ out) {
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total32; t += (size_t)gridDim.x * blockDim.x) {
    const size_t p = t >> 5, lane = t & 31;
    const unsigned s = src[t];
#pragma unroll
    for (int k = 0; k < 9; k++) out[(p * 9 + k) * 32 + lane] = (s == kSsorPad) ? 0.0 : vals[9 * (size_t)s + k];
  }
}

// ALU(new row) = block LU of sigma-scaled D(old row)  (SSOR_33.f90:155-213)
__global__ void __launch_bounds__(256)
ssor_alu_kernel(int N, const int *__restrict__ perm, const double *__restrict__ D, double sigma,
                double *__restrict__ alu, size_t ld) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)N) return;
  const double *d = D + 9 * (size_t)perm[i];
  double T[3][3];
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

__device__ __forceinline__ void ssor_block_solve(const double *__restrict__ alu, size_t ld, size_t i,
                                                 double &X1, double &X2, double &X3) {
  const double a1 = alu[1 * ld + i], a2 = alu[2 * ld + i], a3 = alu[3 * ld + i];
  const double a5 = alu[5 * ld + i], a6 = alu[6 * ld + i], a7 = alu[7 * ld + i];
  const double a0 = alu[0 * ld + i], a4 = alu[4 * ld + i], a8 = alu[8 * ld + i];
  X2 = fma(-a3, X1, X2);
  X3 = fma(-a7, X2, fma(-a6, X1, X3));
  X3 = a8 * X3;
  X2 = a4 * fma(-a5, X3, X2);
  X1 = a0 * fma(-a1, X2, fma(-a2, X3, X1));
}

// one colour of the forward (FWD=true, SSOR_33.f90:300-352) or backward (:355-410) sweep; DOT: also r.z over the rows
// this launch finishes (a row of the backward sweep holds its final z once its own colour has been swept)
template <bool FWD, bool DOT>
__device__ __forceinline__ void ssor_sweep_body(int s0, int s1, const long long *__restrict__ slice_ptr,
                                                const int *__restrict__ slice_row0, const int *__restrict__ slice_nrow,
                                                const double *__restrict__ val, const int *__restrict__ col,
                                                const int *__restrict__ perm, const double *__restrict__ alu, size_t ld,
                                                double *zp, const double *__restrict__ r, double &rz) {
  const int lane = threadIdx.x & 31;
  const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nw = (gridDim.x * blockDim.x) >> 5;
  for (int s = s0 + wg; s < s1; s += nw) {
    const long long p0 = slice_ptr[s], p1 = slice_ptr[s + 1];
    const bool active = lane < slice_nrow[s];
    const int inew = slice_row0[s] + lane;
    const size_t iold = active ? (size_t)perm[inew] : 0;
    double s1v = 0.0, s2v = 0.0, s3v = 0.0;
    if (FWD && active) { s1v = zp[3 * iold]; s2v = zp[3 * iold + 1]; s3v = zp[3 * iold + 2]; }
    const double *v = val + p0 * 288 + lane;
    const int *cp = col + p0 * 32 + lane;
    for (long long p = p0; p < p1; ++p, v += 288, cp += 32) {
      const int c = ld_stream(cp);
      double a[9];
#pragma unroll
      for (int k = 0; k < 9; k++) a[k] = ld_stream(v + k * 32);
      const double x0 = zp[3 * (size_t)c], x1 = zp[3 * (size_t)c + 1], x2 = zp[3 * (size_t)c + 2];
      if (FWD) {
        s1v = fma(-a[2], x2, fma(-a[1], x1, fma(-a[0], x0, s1v)));
        s2v = fma(-a[5], x2, fma(-a[4], x1, fma(-a[3], x0, s2v)));
        s3v = fma(-a[8], x2, fma(-a[7], x1, fma(-a[6], x0, s3v)));
      } else {
        s1v = fma(a[2], x2, fma(a[1], x1, fma(a[0], x0, s1v)));
        s2v = fma(a[5], x2, fma(a[4], x1, fma(a[3], x0, s2v)));
        s3v = fma(a[8], x2, fma(a[7], x1, fma(a[6], x0, s3v)));
      }
    }
    if (active) {
      ssor_block_solve(alu, ld, (size_t)inew, s1v, s2v, s3v);
      if (FWD) { zp[3 * iold] = s1v; zp[3 * iold + 1] = s2v; zp[3 * iold + 2] = s3v; }
      else if (!DOT) { zp[3 * iold] -= s1v; zp[3 * iold + 1] -= s2v; zp[3 * iold + 2] -= s3v; }
      else {
        const double z0 = zp[3 * iold] - s1v, z1 = zp[3 * iold + 1] - s2v, z2 = zp[3 * iold + 2] - s3v;
        zp[3 * iold] = z0; zp[3 * iold + 1] = z1; zp[3 * iold + 2] = z2;
        rz = fma(r[3 * iold], z0, rz); rz = fma(r[3 * iold + 1], z1, rz); rz = fma(r[3 * iold + 2], z2, rz);
      }
    }
  }
}

template <bool FWD>
__global__ void __launch_bounds__(256)
ssor_sweep_kernel(int s0, int s1, const long long *__restrict__ slice_ptr, const int *__restrict__ slice_row0,
                  const int *__restrict__ slice_nrow, const double *__restrict__ val,
                  const int *__restrict__ col, const int *__restrict__ perm,
                  const double *__restrict__ alu, size_t ld, double *zp, const int *phase) {
  if (phase && *phase != 0) return;
  double unused = 0.0;
  ssor_sweep_body<FWD, false>(s0, s1, slice_ptr, slice_row0, slice_nrow, val, col, perm, alu, ld, zp, nullptr, unused);
}

// Backward sweep of one colour fused with the inner product that follows the preconditioner in hecmw_solve_CG
// (RHO = r.z, CG.f90:160-168): the rows of this colour are final after this launch, so their share of r.z is summed here
// (deterministic grid reduction) and added to the shares of the colours swept before (`accumulate`; launches of one
// stream are ordered, so the sum over colours has a fixed order too).  Saves the separate dot-product pass over r and z.
__global__ void __launch_bounds__(256)
ssor_sweep_bwd_dot_kernel(int s0, int s1, const long long *__restrict__ slice_ptr, const int *__restrict__ slice_row0,
                          const int *__restrict__ slice_nrow, const double *__restrict__ val,
                          const int *__restrict__ col, const int *__restrict__ perm,
                          const double *__restrict__ alu, size_t ld, double *zp, const double *__restrict__ r,
                          double *partial, unsigned int *counter, double *rz_out, int accumulate, const int *phase) {
  if (phase && *phase != 0) return;
  double acc[1] = {0.0};
  ssor_sweep_body<false, true>(s0, s1, slice_ptr, slice_row0, slice_nrow, val, col, perm, alu, ld, zp, r, acc[0]);
  grid_reduce<1, 256>(acc, partial, counter, [=](double(&o)[1]) { *rz_out = accumulate ? *rz_out + o[0] : o[0]; });
}

}  // namespace fb
