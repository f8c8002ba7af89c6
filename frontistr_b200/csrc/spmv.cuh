// BCSR33 SpMV for hecmw_matvec_33 (hecmw1/src/solver/las/hecmw_solver_las_33.f90:135-351).
//
// The hecmwST_matrix arrays (D / AL / AU + indexL/itemL/indexU/itemU) stay untouched at the
// interface.  Like the reference's USEJAD option (las/hecmw_jadm_33.f90: a private, machine-friendly
// copy built at solve start) the solver keeps a private device copy in a sliced-ELL layout:
//
//   slice s = rows 32s .. 32s+31, width w_s = max_i (1 + |L_i| + |U_i|) block columns
//   block column p of a slice:  col[p*32 + lane]          0-based column node
//                               val[(p*9 + c)*32 + lane]  c = 3*row_in_block + col_in_block
//   column order inside a row is D, then L ascending, then U ascending — the reference's own
//   accumulation order (las_33.f90:269-295) — so one thread per row reproduces the reference's
//   floating-point result bit for bit.  Rows shorter than the slice are padded with zero blocks
//   pointing at the row's own node.
//
// Every load of the value/index streams is a fully coalesced 256-byte (double) / 512-byte (double2)
// warp access with an L1::no_allocate hint; x is gathered through L1/L2.
// Variant 2 interleaves two block columns so each lane issues 16-byte double2 loads.
#pragma once
#include <algorithm>

#include "common.cuh"
#include "cg_kernels.cuh"

namespace fb {

struct SellMatrix {
  int N = 0, NP = 0, nslice = 0, variant = 0;
  long long total_cols = 0;        // block columns (variant 1) or pair columns (variant 2)
  DevBuf<long long> slice_ptr;     // nslice+1
  DevBuf<double> val;
  DevBuf<int> col;
  DevBuf<unsigned char> halo_slice;   // 1 if any row of the slice has a column in the halo (>= N)
  DevBuf<int> rowperm;                // SELL-C-sigma: matrix row held by lane l of slice s = rowperm[32 s + l] (-1: none)
  bool permuted = false;              // rowperm in use (option "sell_sigma" > 1), else rows are in natural order
  size_t bytes() const { return val.n * 8 + col.n * 4 + slice_ptr.n * 8; }
};

// ---- layout (host, topology only): which row sits in which lane of which slice, and the slice widths -------
// SELL-C-sigma (sigma > 1): rows are sorted by descending block count inside windows of sigma rows before they are
// cut into 32-row slices, which removes most of the zero padding on meshes whose rows differ a lot in length
// (tet10: vertex vs mid-edge nodes, 1.5x -> 1.05x; hex: 1.06x -> 1.04x).  The order INSIDE a row is untouched, so
// each row's result stays bit-identical.  rowperm is left empty for the natural order (sigma <= 1).
inline void sell_layout(int N, const int *indexL, const int *indexU, int sigma, int variant, std::vector<int> &rowperm,
                        std::vector<long long> &slice_ptr) {
  const int nslice = (N + 31) / 32;
  auto width = [&](int r) { return 1 + (indexL[r + 1] - indexL[r]) + (indexU[r + 1] - indexU[r]); };
  rowperm.clear();
  if (sigma > 1 && nslice > 0) {
    rowperm.assign((size_t)nslice * 32, -1);
    for (int r = 0; r < N; r++) rowperm[r] = r;
    for (int s0 = 0; s0 < N; s0 += sigma) {
      const int s1 = std::min(N, s0 + sigma);
      std::stable_sort(rowperm.begin() + s0, rowperm.begin() + s1, [&](int a, int b) { return width(a) > width(b); });
    }
  }
  slice_ptr.assign((size_t)nslice + 1, 0);
  for (int s = 0; s < nslice; s++) {
    int w = 0;
    for (int l = 0; l < 32; l++) {
      const int r = !rowperm.empty() ? rowperm[(size_t)s * 32 + l] : ((s * 32 + l < N) ? s * 32 + l : -1);
      if (r >= 0) w = std::max(w, width(r));
    }
    if (variant == 2) w = (w + 1) / 2;
    slice_ptr[s + 1] = slice_ptr[s] + w;
  }
}

// ---- repack: hecmw layout -> SELL (runs once per numeric change) --------------------------
template <int VARIANT>
__global__ void __launch_bounds__(256)
sell_fill_kernel(int N, const double *__restrict__ D, const double *__restrict__ AL,
                 const double *__restrict__ AU, const int *__restrict__ indexL,
                 const int *__restrict__ itemL, const int *__restrict__ indexU,
                 const int *__restrict__ itemU, const long long *__restrict__ slice_ptr,
                 double *__restrict__ val, int *__restrict__ col, unsigned char *__restrict__ halo_slice,
                 const int *__restrict__ rowperm) {
  const int lane = threadIdx.x & 31;
  const long long s = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (s * 32 >= N) return;
  // 0-based matrix row of this lane; N (= "no row") for the padding lanes of the last slice / of a permuted slice
  long long row = s * 32 + lane;
  if (rowperm) { const int pr = rowperm[row]; row = pr < 0 ? (long long)N : (long long)pr; }
  const long long p0 = slice_ptr[s], p1 = slice_ptr[s + 1];
  const int w = (int)(p1 - p0) * (VARIANT == 2 ? 2 : 1);
  int nL = 0, nU = 0;
  long long l0 = 0, u0 = 0;
  if (row < N) {
    l0 = indexL[row]; nL = indexL[row + 1] - (int)l0;
    u0 = indexU[row]; nU = indexU[row + 1] - (int)u0;
  }
  bool has_halo = false;
  for (int j = 0; j < w; j++) {
    const double *src = nullptr;
    int c = (row < N) ? (int)row : 0;
    if (row < N) {
      if (j == 0) src = D + 9 * row;
      else if (j - 1 < nL) { src = AL + 9 * (l0 + j - 1); c = itemL[l0 + j - 1] - 1; }
      else if (j - 1 - nL < nU) { src = AU + 9 * (u0 + j - 1 - nL); c = itemU[u0 + j - 1 - nL] - 1; }
    }
    has_halo = has_halo || (c >= N);
    if (VARIANT == 1) {
      const long long p = p0 + j;
      col[p * 32 + lane] = c;
#pragma unroll
      for (int k = 0; k < 9; k++) val[(p * 9 + k) * 32 + lane] = src ? src[k] : 0.0;
    } else {
      const long long p = p0 + (j >> 1);
      col[(p * 32 + lane) * 2 + (j & 1)] = c;
#pragma unroll
      for (int k = 0; k < 9; k++) val[((p * 9 + k) * 32 + lane) * 2 + (j & 1)] = src ? src[k] : 0.0;
    }
  }
  const unsigned any = __ballot_sync(0xffffffffu, has_halo);
  if (lane == 0) halo_slice[s] = any ? 1 : 0;
}

#define FB_ROW_FMA(y, a0, a1, a2, x0, x1, x2) y = fma((a2), (x2), fma((a1), (x1), fma((a0), (x0), (y))))

// MODE 0: y = A x,      red = sum_i x_i y_i     (q = A p with p.q folded in, CG.f90:204-211)
// MODE 1: y = b - A x,  red = sum_i y_i y_i     (hecmw_matresid + r.r, las_33.f90:358-380)
template <int MODE, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
spmv_sell1_kernel(int N, int nslice, const long long *__restrict__ slice_ptr,
                  const double *__restrict__ val, const int *__restrict__ col,
                  const double *__restrict__ x, const double *__restrict__ b, double *__restrict__ y,
                  double *partial, unsigned int *counter, double *red_out, const int *phase,
                  const int *__restrict__ rowperm) {
  if (phase && *phase != 0) return;
  const int lane = threadIdx.x & 31;
  const int wg = (blockIdx.x * BLOCK + threadIdx.x) >> 5;
  const int nw = (gridDim.x * BLOCK) >> 5;
  double dot[1] = {0.0};
  for (int s = wg; s < nslice; s += nw) {
    const long long p0 = slice_ptr[s], p1 = slice_ptr[s + 1];
    const double *v = val + p0 * 288 + lane;
    const int *cp = col + p0 * 32 + lane;
    double y0 = 0.0, y1 = 0.0, y2 = 0.0;
    long long p = p0;
    for (; p + 2 <= p1; p += 2, v += 576, cp += 64) {
      const int c0 = ld_stream(cp), c1 = ld_stream(cp + 32);
      double a[18];
#pragma unroll
      for (int k = 0; k < 18; k++) a[k] = ld_stream(v + k * 32);
      const double *xa = x + 3 * (size_t)c0, *xb = x + 3 * (size_t)c1;
      const double xa0 = xa[0], xa1 = xa[1], xa2 = xa[2];
      const double xb0 = xb[0], xb1 = xb[1], xb2 = xb[2];
      FB_ROW_FMA(y0, a[0], a[1], a[2], xa0, xa1, xa2);
      FB_ROW_FMA(y1, a[3], a[4], a[5], xa0, xa1, xa2);
      FB_ROW_FMA(y2, a[6], a[7], a[8], xa0, xa1, xa2);
      FB_ROW_FMA(y0, a[9], a[10], a[11], xb0, xb1, xb2);
      FB_ROW_FMA(y1, a[12], a[13], a[14], xb0, xb1, xb2);
      FB_ROW_FMA(y2, a[15], a[16], a[17], xb0, xb1, xb2);
    }
    if (p < p1) {
      const int c0 = ld_stream(cp);
      double a[9];
#pragma unroll
      for (int k = 0; k < 9; k++) a[k] = ld_stream(v + k * 32);
      const double *xa = x + 3 * (size_t)c0;
      const double xa0 = xa[0], xa1 = xa[1], xa2 = xa[2];
      FB_ROW_FMA(y0, a[0], a[1], a[2], xa0, xa1, xa2);
      FB_ROW_FMA(y1, a[3], a[4], a[5], xa0, xa1, xa2);
      FB_ROW_FMA(y2, a[6], a[7], a[8], xa0, xa1, xa2);
    }
    const int row = rowperm ? rowperm[s * 32 + lane] : s * 32 + lane;   // SELL-C-sigma: permuted row (-1: padding lane)
    if (row >= 0 && row < N) {
      const size_t r3 = 3 * (size_t)row;
      if (MODE == 1) {
        y0 = b[r3] - y0; y1 = b[r3 + 1] - y1; y2 = b[r3 + 2] - y2;
        dot[0] = fma(y0, y0, dot[0]); dot[0] = fma(y1, y1, dot[0]); dot[0] = fma(y2, y2, dot[0]);
      } else {
        dot[0] = fma(x[r3], y0, dot[0]); dot[0] = fma(x[r3 + 1], y1, dot[0]); dot[0] = fma(x[r3 + 2], y2, dot[0]);
      }
      y[r3] = y0; y[r3 + 1] = y1; y[r3 + 2] = y2;
    }
  }
  grid_reduce<1, BLOCK>(dot, partial, counter, [=](double(&o)[1]) { if (red_out) *red_out = o[0]; });
}

// variant 2: pair-interleaved columns, double2 / int2 loads.
// PEER = true is the instantiation launched in peer-memory mode (a.halo_seq != 0): the halo rows of x are then written
// by OTHER GPUs while this kernel runs, so the slices that read them must not take x through the non-coherent L1 /
// read-only path the compiler picks for `const __restrict__` data (an interior warp of the same SM may have pulled the
// sector that straddles rows N-1 | N before the neighbour's store landed).  Those slices gather x with ld.global.cg
// (L2, the point of coherence for peer stores) after their acquire on the neighbours' flags.  PEER = false (one GPU, NCCL)
// contains no peer-memory code at all: the wait, the publication and - what they cost every thread even when never
// taken - the copy of the P2PView / P2PNeighbors tables into a 280-byte stack frame are compiled out (`if constexpr`;
// checked with cuobjdump: STACK:0, no STL / LDL).
template <int MODE, int BLOCK, bool PEER = false>
__global__ void __launch_bounds__(BLOCK)
spmv_sell2_kernel(int N, int nslice, const long long *__restrict__ slice_ptr,
                  const double *__restrict__ val, const int *__restrict__ col,
                  const double *__restrict__ x, const double *__restrict__ b, double *__restrict__ y,
                  double *partial, unsigned int *counter, double *red_out, const int *phase,
                  const __grid_constant__ P2PView pv, const __grid_constant__ P2PNeighbors nb, P2PArgs a,
                  const unsigned char *__restrict__ halo_slice, const int *__restrict__ rowperm) {
  if (phase && *phase != 0) return;
  const int lane = threadIdx.x & 31;
  const int wg = (blockIdx.x * BLOCK + threadIdx.x) >> 5;
  const int nw = (gridDim.x * BLOCK) >> 5;
  double dot[1] = {0.0};
  // (the peer tables are __grid_constant__ parameters: they are indexed, and their address is captured by the publication
  // below, in the constant bank - not through a per-thread copy in local memory)
  // peer-memory mode: the neighbours store their export values of x straight into our halo rows; only a warp
  // about to process a slice that reads halo columns waits for their flags — interior slices overlap the exchange
  // (PEER = false carries none of this: no wait, no publication, no copy of the peer tables into the thread's stack frame)
  bool halo_ready = !PEER || (a.halo_seq == 0);
  for (int s = wg; s < nslice; s += nw) {
    if constexpr (PEER) {
      if (!halo_ready && halo_slice[s]) { p2p_wait_halo_warp(pv, nb, a.halo_seq); halo_ready = true; }
    }
    const bool via_l2 = PEER && halo_slice[s];
    const long long p0 = slice_ptr[s], p1 = slice_ptr[s + 1];
    const double2 *v = reinterpret_cast<const double2 *>(val) + p0 * 288 + lane;
    const int2 *cp = reinterpret_cast<const int2 *>(col) + p0 * 32 + lane;
    double y0 = 0.0, y1 = 0.0, y2 = 0.0;
#pragma unroll 2
    for (long long p = p0; p < p1; ++p, v += 288, cp += 32) {
      const int2 c = ld_stream2(cp);
      double2 a[9];
#pragma unroll
      for (int k = 0; k < 9; k++) a[k] = ld_stream2(v + k * 32);
      const double *xa = x + 3 * (size_t)c.x, *xb = x + 3 * (size_t)c.y;
      double xa0, xa1, xa2, xb0, xb1, xb2;
      if (PEER && via_l2) {
        xa0 = __ldcg(xa); xa1 = __ldcg(xa + 1); xa2 = __ldcg(xa + 2);
        xb0 = __ldcg(xb); xb1 = __ldcg(xb + 1); xb2 = __ldcg(xb + 2);
      } else {
        xa0 = xa[0]; xa1 = xa[1]; xa2 = xa[2];
        xb0 = xb[0]; xb1 = xb[1]; xb2 = xb[2];
      }
      FB_ROW_FMA(y0, a[0].x, a[1].x, a[2].x, xa0, xa1, xa2);
      FB_ROW_FMA(y1, a[3].x, a[4].x, a[5].x, xa0, xa1, xa2);
      FB_ROW_FMA(y2, a[6].x, a[7].x, a[8].x, xa0, xa1, xa2);
      FB_ROW_FMA(y0, a[0].y, a[1].y, a[2].y, xb0, xb1, xb2);
      FB_ROW_FMA(y1, a[3].y, a[4].y, a[5].y, xb0, xb1, xb2);
      FB_ROW_FMA(y2, a[6].y, a[7].y, a[8].y, xb0, xb1, xb2);
    }
    const int row = rowperm ? rowperm[s * 32 + lane] : s * 32 + lane;   // SELL-C-sigma: permuted row (-1: padding lane)
    if (row >= 0 && row < N) {
      const size_t r3 = 3 * (size_t)row;
      if (MODE == 1) {
        y0 = b[r3] - y0; y1 = b[r3 + 1] - y1; y2 = b[r3 + 2] - y2;
        dot[0] = fma(y0, y0, dot[0]); dot[0] = fma(y1, y1, dot[0]); dot[0] = fma(y2, y2, dot[0]);
      } else {
        dot[0] = fma(x[r3], y0, dot[0]); dot[0] = fma(x[r3 + 1], y1, dot[0]); dot[0] = fma(x[r3 + 2], y2, dot[0]);
      }
      y[r3] = y0; y[r3 + 1] = y1; y[r3 + 2] = y2;
    }
  }
  if constexpr (PEER) {
    const P2PView *ppv = &pv;
    grid_reduce<1, BLOCK, true>(dot, partial, counter, [=](double(&o)[1]) {
      if (a.publish_seq) p2p_publish_warp(*ppv, P2P_Q_PQ, a.publish_seq, o[0], 0.0);
      else if (red_out && threadIdx.x == 0) *red_out = o[0];
    });
  } else {
    grid_reduce<1, BLOCK>(dot, partial, counter, [=](double(&o)[1]) { if (red_out) *red_out = o[0]; });
  }
}

// variant 0: no private copy — one thread per scalar row straight on D/AL/AU (same accumulation
// order, so also bit-exact).  Baseline for the roofline comparison and the zero-extra-memory mode.
template <int MODE, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
spmv_native_kernel(int N, const double *__restrict__ D, const double *__restrict__ AL,
                   const double *__restrict__ AU, const int *__restrict__ indexL,
                   const int *__restrict__ itemL, const int *__restrict__ indexU,
                   const int *__restrict__ itemU, const double *__restrict__ x,
                   const double *__restrict__ b, double *__restrict__ y, double *partial,
                   unsigned int *counter, double *red_out, const int *phase) {
  if (phase && *phase != 0) return;
  double dot[1] = {0.0};
  const long long n3 = 3LL * N;
  for (long long t = (long long)blockIdx.x * BLOCK + threadIdx.x; t < n3; t += (long long)gridDim.x * BLOCK) {
    const int i = (int)(t / 3), r = (int)(t - 3LL * i);
    const double *d = D + 9 * (size_t)i + 3 * r;
    const double *xi = x + 3 * (size_t)i;
    double acc = 0.0;
    FB_ROW_FMA(acc, d[0], d[1], d[2], xi[0], xi[1], xi[2]);
    for (int j = indexL[i]; j < indexL[i + 1]; j++) {
      const double *a = AL + 9 * (size_t)j + 3 * r;
      const double *xc = x + 3 * (size_t)(itemL[j] - 1);
      FB_ROW_FMA(acc, a[0], a[1], a[2], xc[0], xc[1], xc[2]);
    }
    for (int j = indexU[i]; j < indexU[i + 1]; j++) {
      const double *a = AU + 9 * (size_t)j + 3 * r;
      const double *xc = x + 3 * (size_t)(itemU[j] - 1);
      FB_ROW_FMA(acc, a[0], a[1], a[2], xc[0], xc[1], xc[2]);
    }
    if (MODE == 1) { acc = b[t] - acc; dot[0] = fma(acc, acc, dot[0]); }
    else dot[0] = fma(x[t], acc, dot[0]);
    y[t] = acc;
  }
  grid_reduce<1, BLOCK>(dot, partial, counter, [=](double(&o)[1]) { if (red_out) *red_out = o[0]; });
}

}  // namespace fb
