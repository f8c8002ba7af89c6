// C-ABI of the fistr_b200 library (include/fistr_b200.h): context, residency, SpMV, preconditioners,
// the CG driver mirroring hecmw_solve -> hecmw_solve_iterative -> hecmw_solve_CG, assembly and BC.
#include <cmath>
#include <cstdarg>

#include <cub/device/device_scan.cuh>

#include "context.hpp"
#include "partition.hpp"

namespace fb {
__global__ void add_const_kernel(size_t n, int *a, int c) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] += c;
}
thread_local std::string g_last_error;
int64_t g_launch_count = 0;
Context &ctx() {
  static Context c;
  return c;
}

static int fail(int code, const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define FB_REQUIRE_INIT()                                                              \
  do {                                                                                 \
    if (!fb::ctx().inited) {                                                           \
      int _rc = fistr_b200_init(-1);                                                   \
      if (_rc) return _rc;                                                             \
    }                                                                                  \
  } while (0)

#define FB_TRY(expr)            \
  do {                          \
    int _rc = (expr);           \
    if (_rc) return _rc;        \
  } while (0)

#define FB_NCCL(expr)                                                                       \
  do {                                                                                      \
    int _r = (expr);                                                                        \
    if (_r != 0) return fail(FISTR_B200_ERR_COMM, "%s: %s", #expr, C.nccl.GetErrorString(_r)); \
  } while (0)

// ------------------------------------------------------------------------------------------
// halo exchange (hecmw_update_3_R) and scalar all-reduce (hecmw_allreduce_R1)
// ------------------------------------------------------------------------------------------
static int exchange(Context &C, double *x, bool packed_p) {
  if (C.petot <= 1 || C.halo.n_neighbor == 0) return 0;
  Halo &H = C.halo;
  double *send = packed_p ? H.send_p.p : H.send_x.p;
  if (!packed_p && H.n_export() > 0) {
    gather3_kernel<<<C.grid_for(3 * (size_t)H.n_export(), 256, 2), 256, 0, C.stream>>>(
        H.n_export(), H.d_export_item.p, x, send);
    FB_LAUNCH_CHECK();
  }
  double *recv = H.import_contiguous ? x + 3 * (size_t)C.N : H.recv.p;
  FB_NCCL(C.nccl.GroupStart());
  for (int nb = 0; nb < H.n_neighbor; nb++) {
    const int peer = H.neighbor_pe[nb];
    const size_t ns = 3 * (size_t)(H.export_index[nb + 1] - H.export_index[nb]);
    const size_t nr = 3 * (size_t)(H.import_index[nb + 1] - H.import_index[nb]);
    if (ns) FB_NCCL(C.nccl.Send(send + 3 * (size_t)H.export_index[nb], ns, ncclFloat64_, peer, C.comm, C.stream));
    if (nr) FB_NCCL(C.nccl.Recv(recv + 3 * (size_t)H.import_index[nb], nr, ncclFloat64_, peer, C.comm, C.stream));
  }
  FB_NCCL(C.nccl.GroupEnd());
  if (!H.import_contiguous && H.n_import() > 0) {
    scatter3_kernel<<<C.grid_for(3 * (size_t)H.n_import(), 256, 2), 256, 0, C.stream>>>(
        H.n_import(), H.d_import_item.p, recv, x);
    FB_LAUNCH_CHECK();
  }
  return 0;
}

static int allreduce(Context &C, double *buf, int count) {
  if (C.petot <= 1) return 0;
  FB_NCCL(C.nccl.AllReduce(buf, buf, (size_t)count, ncclFloat64_, ncclSum_, C.comm, C.stream));
  return 0;
}

// ------------------------------------------------------------------------------------------
// private copies: SELL repack, preconditioner factors
// ------------------------------------------------------------------------------------------
static int build_sell(Context &C) {
  const int variant = (int)C.get("spmv_variant", 2);
  if (variant == 0) { C.sell_valid = true; C.sell.variant = 0; return 0; }
  SellMatrix &S = C.sell;
  const int N = C.N;
  const int nslice = (N + 31) / 32;
  // row placement and slice widths (spmv.cuh: sell_layout; SELL-C-sigma when option "sell_sigma" > 1, off by default)
  std::vector<int> rowperm;
  std::vector<long long> sp;
  sell_layout(N, C.h_indexL.data(), C.h_indexU.data(), (int)C.get("sell_sigma", 0), variant, rowperm, sp);
  S.permuted = !rowperm.empty();
  if (S.permuted) {
    if (S.rowperm.alloc(rowperm.size())) return FISTR_B200_ERR_CUDA;
    FB_CUDA(cudaMemcpyAsync(S.rowperm.p, rowperm.data(), sizeof(int) * rowperm.size(), cudaMemcpyHostToDevice, C.stream));
  }
  S.N = N; S.NP = C.NP; S.nslice = nslice; S.variant = variant; S.total_cols = sp[nslice];
  const size_t cols = (size_t)S.total_cols * (variant == 2 ? 2 : 1);
  if (S.slice_ptr.alloc((size_t)nslice + 1) || S.val.alloc(cols * 288) || S.col.alloc(cols * 32) || S.halo_slice.alloc((size_t)nslice + 1)) return FISTR_B200_ERR_CUDA;
  FB_CUDA(cudaMemcpyAsync(S.slice_ptr.p, sp.data(), sizeof(long long) * sp.size(), cudaMemcpyHostToDevice, C.stream));
  FB_CUDA(cudaStreamSynchronize(C.stream));   // sp, rowperm are locals
  const int *rp = S.permuted ? S.rowperm.p : nullptr;
  if (nslice > 0) {
    const int blocks = (nslice * 32 + 255) / 256;
    if (variant == 2)
      sell_fill_kernel<2><<<blocks, 256, 0, C.stream>>>(N, C.dD(), C.dAL(), C.dAU(), C.indexL.p, C.itemL.p, C.indexU.p, C.itemU.p, S.slice_ptr.p, S.val.p, S.col.p, S.halo_slice.p, rp);
    else
      sell_fill_kernel<1><<<blocks, 256, 0, C.stream>>>(N, C.dD(), C.dAL(), C.dAU(), C.indexL.p, C.itemL.p, C.indexU.p, C.itemU.p, S.slice_ptr.p, S.val.p, S.col.p, S.halo_slice.p, rp);
    FB_LAUNCH_CHECK();
  }
  C.sell_valid = true;
  return 0;
}

// mode 0: y = A x (red_out <- x.y), mode 1: y = b - A x (red_out <- y.y).  x must hold its halo.
static int launch_spmv(Context &C, int mode, const double *x, const double *b, double *y, double *red_out,
                       const int *phase, const P2PArgs &pa = P2PArgs()) {
  if (!C.sell_valid) FB_TRY(build_sell(C));
  constexpr int BLOCK = 256;
  const int per_sm = (int)C.get("spmv_blocks_per_sm", 8);
  if (C.N == 0) return 0;
  if (C.sell.variant == 0) {
    const int grid = C.grid_for(3 * (size_t)C.N, BLOCK, per_sm);
    if (mode == 0) spmv_native_kernel<0, BLOCK><<<grid, BLOCK, 0, C.stream>>>(C.N, C.dD(), C.dAL(), C.dAU(), C.indexL.p, C.itemL.p, C.indexU.p, C.itemU.p, x, b, y, C.partial.p, C.counter.p, red_out, phase);
    else spmv_native_kernel<1, BLOCK><<<grid, BLOCK, 0, C.stream>>>(C.N, C.dD(), C.dAL(), C.dAU(), C.indexL.p, C.itemL.p, C.indexU.p, C.itemU.p, x, b, y, C.partial.p, C.counter.p, red_out, phase);
  } else {
    const SellMatrix &S = C.sell;
    const int *rp = S.permuted ? S.rowperm.p : nullptr;
    const int grid = C.grid_for((size_t)S.nslice * 32, BLOCK, per_sm);
    if (S.variant == 1) {
      if (mode == 0) spmv_sell1_kernel<0, BLOCK><<<grid, BLOCK, 0, C.stream>>>(S.N, S.nslice, S.slice_ptr.p, S.val.p, S.col.p, x, b, y, C.partial.p, C.counter.p, red_out, phase, rp);
      else spmv_sell1_kernel<1, BLOCK><<<grid, BLOCK, 0, C.stream>>>(S.N, S.nslice, S.slice_ptr.p, S.val.p, S.col.p, x, b, y, C.partial.p, C.counter.p, red_out, phase, rp);
    } else {
      if (mode == 0) spmv_sell2_kernel<0, BLOCK><<<grid, BLOCK, 0, C.stream>>>(S.N, S.nslice, S.slice_ptr.p, S.val.p, S.col.p, x, b, y, C.partial.p, C.counter.p, red_out, phase, C.p2p.view, C.p2p.nbrs, pa, S.halo_slice.p, rp);
      else spmv_sell2_kernel<1, BLOCK><<<grid, BLOCK, 0, C.stream>>>(S.N, S.nslice, S.slice_ptr.p, S.val.p, S.col.p, x, b, y, C.partial.p, C.counter.p, red_out, phase, C.p2p.view, C.p2p.nbrs, pa, S.halo_slice.p, rp);
    }
  }
  FB_LAUNCH_CHECK();
  return 0;
}

static int ssor_setup(Context &C, double sigma, int ncolor_in, bool structure_changed) {
  Ssor &S = C.ssor;
  const int N = C.N;
  if (structure_changed || !S.structure_valid || S.N != N) {
    std::vector<int> itemL((size_t)C.NPL), itemU((size_t)C.NPU);
    FB_CUDA(cudaMemcpy(itemL.data(), C.itemL.p, sizeof(int) * itemL.size(), cudaMemcpyDeviceToHost));
    FB_CUDA(cudaMemcpy(itemU.data(), C.itemU.p, sizeof(int) * itemU.size(), cudaMemcpyDeviceToHost));
    S.N = N;
    ssor_multicolor_order(N, C.h_indexL.data(), itemL.data(), C.h_indexU.data(), itemU.data(), ncolor_in,
                          S.perm, S.iperm, S.color_index);
    S.ncolor = (int)S.color_index.size() - 1;
    std::vector<int> perm0(N);
    for (int i = 0; i < N; i++) perm0[i] = S.perm[i] - 1;
    if (S.d_perm.alloc(N)) return FISTR_B200_ERR_CUDA;
    FB_CUDA(cudaMemcpy(S.d_perm.p, perm0.data(), sizeof(int) * N, cudaMemcpyHostToDevice));
    for (int up = 0; up < 2; up++) {
      SsorPart &P = up ? S.U : S.L;
      std::vector<long long> sp; std::vector<int> row0, nrow, col; std::vector<unsigned> src;
      ssor_build_part(S, up != 0, C.NP, C.NPL, C.h_indexL.data(), itemL.data(), C.h_indexU.data(), itemU.data(),
                      sp, row0, nrow, col, src, P.color_slice);
      P.nslice = (int)row0.size();
      P.total = sp.back();
      if (P.slice_ptr.alloc(sp.size()) || P.slice_row0.alloc(row0.size()) || P.slice_nrow.alloc(nrow.size()) ||
          P.col.alloc(col.size()) || P.src.alloc(src.size()) || P.val.alloc((size_t)sp.back() * 288)) return FISTR_B200_ERR_CUDA;
      FB_CUDA(cudaMemcpy(P.slice_ptr.p, sp.data(), sizeof(long long) * sp.size(), cudaMemcpyHostToDevice));
      if (!row0.empty()) {
        FB_CUDA(cudaMemcpy(P.slice_row0.p, row0.data(), sizeof(int) * row0.size(), cudaMemcpyHostToDevice));
        FB_CUDA(cudaMemcpy(P.slice_nrow.p, nrow.data(), sizeof(int) * nrow.size(), cudaMemcpyHostToDevice));
      }
      if (!col.empty()) {
        FB_CUDA(cudaMemcpy(P.col.p, col.data(), sizeof(int) * col.size(), cudaMemcpyHostToDevice));
        FB_CUDA(cudaMemcpy(P.src.p, src.data(), sizeof(unsigned) * src.size(), cudaMemcpyHostToDevice));
      }
    }
    S.ld = ((size_t)N + 31) / 32 * 32;
    if (S.alu.alloc(9 * std::max<size_t>(S.ld, 32))) return FISTR_B200_ERR_CUDA;
    S.structure_valid = true;
  }
  // numeric part
  for (int up = 0; up < 2; up++) {
    SsorPart &P = up ? S.U : S.L;
    const size_t total32 = (size_t)P.total * 32;
    if (total32 > 0) {
      ssor_fill_kernel<<<C.grid_for(total32, 256, 8), 256, 0, C.stream>>>(total32, P.src.p, C.vals.p, P.val.p);
      FB_LAUNCH_CHECK();
    }
  }
  if (N > 0) {
    ssor_alu_kernel<<<(N + 255) / 256, 256, 0, C.stream>>>(N, S.d_perm.p, C.dD(), sigma, S.alu.p, S.ld);
    FB_LAUNCH_CHECK();
  }
  S.values_valid = true;
  return 0;
}

static int precond_setup(Context &C, int precond, double sigma, int ncolor_in, bool structure_changed) {
  if (precond == 3) {
    C.alu_ld = ((size_t)C.NP + 31) / 32 * 32;
    if (C.alu.alloc(9 * std::max<size_t>(C.alu_ld, 32))) return FISTR_B200_ERR_CUDA;
    if (C.N > 0) {
      diag_setup_kernel<<<(C.N + 255) / 256, 256, 0, C.stream>>>(C.N, C.dD(), sigma, C.alu.p, C.alu_ld);
      FB_LAUNCH_CHECK();
    }
  } else if (precond == 1 || precond == 2) {
    FB_TRY(ssor_setup(C, sigma, ncolor_in, structure_changed));
  } else {
    return fail(HECMW_SOLVER_ERROR_INCONS_PC_, "PRECOND=%d is outside the hot-path scope (1,2 SSOR; 3 DIAG)", precond);
  }
  C.pc_kind = (precond == 3) ? 3 : 1;
  C.pc_sigma = sigma;
  return 0;
}

// hecmw_precond_{DIAG,SSOR}_33_apply on a device vector (in place)
static int precond_apply_dev(Context &C, double *zp, const int *phase) {
  if (C.pc_kind == 3) {
    if (C.N > 0) {
      diag_apply_kernel<<<C.grid_for((size_t)C.N, 256, 8), 256, 0, C.stream>>>(C.N, C.alu.p, C.alu_ld, zp, phase);
      FB_LAUNCH_CHECK();
    }
    return 0;
  }
  Ssor &S = C.ssor;
  for (int c = 0; c < S.ncolor; c++) {
    const int s0 = S.L.color_slice[c], s1 = S.L.color_slice[c + 1];
    if (s1 > s0) {
      ssor_sweep_kernel<true><<<C.grid_for((size_t)(s1 - s0) * 32, 256, 8), 256, 0, C.stream>>>(
          s0, s1, S.L.slice_ptr.p, S.L.slice_row0.p, S.L.slice_nrow.p, S.L.val.p, S.L.col.p, S.d_perm.p, S.alu.p, S.ld, zp, phase);
      FB_LAUNCH_CHECK();
    }
  }
  for (int c = S.ncolor - 1; c >= 0; c--) {
    const int s0 = S.U.color_slice[c], s1 = S.U.color_slice[c + 1];
    if (s1 > s0) {
      ssor_sweep_kernel<false><<<C.grid_for((size_t)(s1 - s0) * 32, 256, 8), 256, 0, C.stream>>>(
          s0, s1, S.U.slice_ptr.p, S.U.slice_row0.p, S.U.slice_nrow.p, S.U.val.p, S.U.col.p, S.d_perm.p, S.alu.p, S.ld, zp, phase);
      FB_LAUNCH_CHECK();
    }
  }
  return 0;
}

// hecmw_precond_apply + hecmw_precond_33_apply (precond.f90:75-123, precond_33.f90:74-115):
// Z = Minv R with iterPREmax additive-Schwarz sweeps; WK is the ZP work vector.
static int precond_apply_generic(Context &C, int iterPREmax, const int *phase) {
  const size_t n3 = 3 * (size_t)C.N, np3 = 3 * (size_t)C.NP;
  const int g = C.grid_for(np3, 256, 8);
  if (iterPREmax <= 0) {
    vec_copy_zero_tail_kernel<<<g, 256, 0, C.stream>>>(n3, n3, C.R.p, C.Z.p, phase);
    FB_LAUNCH_CHECK();
    return 0;
  }
  vec_copy_zero_tail_kernel<<<g, 256, 0, C.stream>>>(n3, np3, C.R.p, C.WK.p, phase);
  FB_LAUNCH_CHECK();
  if (iterPREmax > 1) {   // Z(1:NP)=0 matters only when Z's halo is read by the inner matresid
    vec_copy_zero_tail_kernel<<<g, 256, 0, C.stream>>>(0, np3, C.R.p, C.Z.p, phase);
    FB_LAUNCH_CHECK();
  }
  for (int it = 1; it <= iterPREmax; it++) {
    FB_TRY(precond_apply_dev(C, C.WK.p, phase));
    vec_add_kernel<<<C.grid_for(n3, 256, 8), 256, 0, C.stream>>>(n3, C.Z.p, C.WK.p, it == 1, phase);
    FB_LAUNCH_CHECK();
    if (it == iterPREmax) break;
    FB_TRY(exchange(C, C.Z.p, false));
    FB_TRY(launch_spmv(C, 1, C.Z.p, C.R.p, C.WK.p, C.red.p + RED_TMP, phase));   // ZP = R - A Z
  }
  return 0;
}

// ------------------------------------------------------------------------------------------
// peer-memory arena (p2p.cuh): allocate, exchange IPC handles over the NCCL communicator, map the peers
// ------------------------------------------------------------------------------------------
static void p2p_teardown(Context &C) {
  Context::P2P &T = C.p2p;
  if (!T.arena) return;
  cudaStreamSynchronize(C.stream);
  for (int r = 0; r < kP2PMaxRanks; r++) {
    if (T.peer_base[r] && r != C.rank) cudaIpcCloseMemHandle(T.peer_base[r]);
    T.peer_base[r] = nullptr;
  }
  if (C.comm) {   // nobody may free its arena while a peer could still store into it
    if (C.red.p && C.nccl.AllReduce(C.red.p + RED_TMP, C.red.p + RED_TMP, 1, ncclFloat64_, ncclSum_, C.comm, C.stream) == 0)
      cudaStreamSynchronize(C.stream);
  }
  if (!C.P.owned) C.P.release();
  cudaFree(T.arena);
  T.arena = nullptr; T.active = false; T.bytes = 0;
  T.view = P2PView();
}

static int p2p_setup(Context &C) {
  Context::P2P &T = C.p2p;
  p2p_teardown(C);
  T.active = false;
  if (C.petot <= 1 || C.petot > kP2PMaxRanks || C.get("p2p", 1) == 0 || !C.comm) return 0;
  const Halo &H = C.halo;
  constexpr size_t OFF_FLAGS = sizeof(P2PEntry) * P2P_NQ * 2 * kP2PMaxRanks, OFF_P = 4096;
  static_assert(OFF_FLAGS + 8 * kP2PMaxRanks <= OFF_P, "arena header");
  const size_t np3 = 3 * (size_t)std::max(C.NP, 1) + 8;
  size_t bytes = OFF_P + np3 * sizeof(double);
  bytes = (bytes + (2u << 20) - 1) / (2u << 20) * (2u << 20);   // own 2 MiB-granular allocation => own IPC handle
  FB_CUDA(cudaMalloc(&T.arena, bytes));
  T.bytes = bytes;
  FB_CUDA(cudaMemsetAsync(T.arena, 0, bytes, C.stream));
  struct Card { cudaIpcMemHandle_t h; long long np; long long ok; };
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
  Card mine;
  memset(&mine, 0, sizeof mine);
  mine.ok = (cudaIpcGetMemHandle(&mine.h, T.arena) == cudaSuccess);
  mine.np = C.NP;
  DevBuf<char> d_send, d_recv;
  if (d_send.alloc(sizeof(Card)) || d_recv.alloc(sizeof(Card) * C.petot)) return FISTR_B200_ERR_CUDA;
  FB_CUDA(cudaMemcpyAsync(d_send.p, &mine, sizeof mine, cudaMemcpyHostToDevice, C.stream));
  FB_NCCL(C.nccl.AllGather(d_send.p, d_recv.p, sizeof(Card), ncclChar_, C.comm, C.stream));
  std::vector<Card> cards(C.petot);
  FB_CUDA(cudaMemcpyAsync(cards.data(), d_recv.p, sizeof(Card) * C.petot, cudaMemcpyDeviceToHost, C.stream));
  FB_CUDA(cudaStreamSynchronize(C.stream));
  bool ok = true;
  for (int r = 0; r < C.petot; r++) ok = ok && cards[r].ok;
  for (int r = 0; r < C.petot && ok; r++) {
    if (r == C.rank) { T.peer_base[r] = T.arena; continue; }
    void *base = nullptr;
    if (cudaIpcOpenMemHandle(&base, cards[r].h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); ok = false; break; }
    T.peer_base[r] = base;
  }
  // every rank must take the same decision
  double flag = ok ? 0.0 : 1.0;
  FB_CUDA(cudaMemcpyAsync(C.red.p + RED_TMP, &flag, sizeof flag, cudaMemcpyHostToDevice, C.stream));
  FB_NCCL(C.nccl.AllReduce(C.red.p + RED_TMP, C.red.p + RED_TMP, 1, ncclFloat64_, ncclSum_, C.comm, C.stream));
  FB_CUDA(cudaMemcpyAsync(&flag, C.red.p + RED_TMP, sizeof flag, cudaMemcpyDeviceToHost, C.stream));
  FB_CUDA(cudaStreamSynchronize(C.stream));
  d_send.release(); d_recv.release();
  if (flag != 0.0) {   // no peer access on this box: stay on the NCCL path
    for (int r = 0; r < kP2PMaxRanks; r++) {
      if (T.peer_base[r] && r != C.rank) cudaIpcCloseMemHandle(T.peer_base[r]);
      T.peer_base[r] = nullptr;
    }
    cudaFree(T.arena); T.arena = nullptr; T.bytes = 0;
    if (C.get("p2p", 1) >= 2) return fail(FISTR_B200_ERR_COMM, "p2p: peer memory could not be mapped");
    return 0;
  }
  T.view = P2PView();
  T.view.nranks = C.petot; T.view.me = C.rank;
  for (int r = 0; r < C.petot; r++) {
    char *b = (char *)T.peer_base[r];
    T.view.ar[r] = (P2PEntry *)b;
    T.view.hflag[r] = (unsigned long long *)(b + OFF_FLAGS);
    T.view.vecP[r] = (double *)(b + OFF_P);
  }
  if (T.err.alloc(4)) return FISTR_B200_ERR_CUDA;
  FB_CUDA(cudaMemsetAsync(T.err.p, 0, sizeof(int) * 4, C.stream));
  T.view.err = T.err.p;
  T.nbrs = P2PNeighbors();
  if (H.n_neighbor > kP2PMaxRanks) return fail(FISTR_B200_ERR_COMM, "p2p: too many neighbours");
  T.nbrs.n = H.n_neighbor;
  for (int a = 0; a < H.n_neighbor; a++) T.nbrs.rank[a] = H.neighbor_pe[a];
  // destination of every export entry: the neighbour's own local id of that node = its import_item list for us
  const size_t nexp = (size_t)H.n_export();
  if (T.exp_peer.alloc(nexp) || T.exp_remote.alloc(nexp)) return FISTR_B200_ERR_CUDA;
  std::vector<int> peer(nexp);
  for (int a = 0; a < H.n_neighbor; a++)
    for (int e = H.export_index[a]; e < H.export_index[a + 1]; e++) peer[e] = H.neighbor_pe[a];
  if (nexp) FB_CUDA(cudaMemcpyAsync(T.exp_peer.p, peer.data(), sizeof(int) * nexp, cudaMemcpyHostToDevice, C.stream));
  FB_NCCL(C.nccl.GroupStart());
  for (int a = 0; a < H.n_neighbor; a++) {
    const size_t ns = (size_t)(H.import_index[a + 1] - H.import_index[a]);
    const size_t nr = (size_t)(H.export_index[a + 1] - H.export_index[a]);
    if (ns) FB_NCCL(C.nccl.Send(H.d_import_item.p + H.import_index[a], ns, ncclInt32_, H.neighbor_pe[a], C.comm, C.stream));
    if (nr) FB_NCCL(C.nccl.Recv(T.exp_remote.p + H.export_index[a], nr, ncclInt32_, H.neighbor_pe[a], C.comm, C.stream));
  }
  FB_NCCL(C.nccl.GroupEnd());
  if (nexp) {
    add_const_kernel<<<(int)((nexp + 255) / 256), 256, 0, C.stream>>>(nexp, T.exp_remote.p, -1);   // 1-based -> 0-based
    FB_LAUNCH_CHECK();
  }
  FB_CUDA(cudaStreamSynchronize(C.stream));
  C.P.view(T.view.vecP[C.rank], np3);
  T.seq[0] = T.seq[1] = 0; T.halo_seq = 0;
  T.active = true;
  return 0;
}

static int ensure_work(Context &C) {
  const size_t np3 = 3 * (size_t)std::max(C.NP, 1) + 8;
  if (C.B.alloc(np3) || C.X.alloc(np3) || C.R.alloc(np3) || C.Z.alloc(np3) || C.P.alloc(np3) ||
      C.Q.alloc(np3) || C.WK.alloc(np3)) return FISTR_B200_ERR_CUDA;
  if (C.red.alloc(RED_COUNT) || C.partial.alloc(65536) || C.counter.alloc(4) || C.ctl.alloc(1) || C.flag.alloc(4)) return FISTR_B200_ERR_CUDA;
  return 0;
}

}  // namespace fb

using namespace fb;

// =============================================================================================
extern "C" {

int fistr_b200_init(int device_ordinal) {
  Context &C = ctx();
  if (C.inited) return 0;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(FISTR_B200_ERR_CUDA, "no CUDA device (%s); fistr_b200 has no CPU fallback", cudaGetErrorString(e));
  if (device_ordinal < 0) {
    const char *lr = getenv("LOCAL_RANK");
    device_ordinal = lr ? atoi(lr) % ndev : 0;
  }
  FB_CUDA(cudaSetDevice(device_ordinal));
  C.device = device_ordinal;
  cudaDeviceProp prop;
  FB_CUDA(cudaGetDeviceProperties(&prop, device_ordinal));
  C.sm_count = prop.multiProcessorCount;
  FB_CUDA(cudaStreamCreateWithFlags(&C.stream, cudaStreamNonBlocking));
  if (C.counter.alloc(4)) return FISTR_B200_ERR_CUDA;
  FB_CUDA(cudaMemset(C.counter.p, 0, 4 * sizeof(unsigned)));
  C.inited = true;
  return 0;
}

int fistr_b200_finalize(void) {
  Context &C = ctx();
  if (!C.inited) return 0;
  cudaStreamSynchronize(C.stream);
  p2p_teardown(C);
  if (C.comm) { C.nccl.CommDestroy(C.comm); C.comm = nullptr; }
  cudaStreamDestroy(C.stream);
  C = Context();
  cudaDeviceReset();
  return 0;
}

const char *fistr_b200_last_error(void) { return g_last_error.c_str(); }

int fistr_b200_set_option(const char *key, int64_t value) {
  Context &C = ctx();
  C.opt[key] = value;
  if (!strcmp(key, "spmv_variant") || !strcmp(key, "sell_sigma")) C.sell_valid = false;
  if (!strcmp(key, "launch_count_reset")) g_launch_count = 0;
  return 0;
}
int64_t fistr_b200_get_option(const char *key) {
  if (!strcmp(key, "p2p_active")) return ctx().p2p.active ? 1 : 0;
  return ctx().get(key, -1);
}

int64_t fistr_b200_launch_count(void) { return g_launch_count; }

int fistr_b200_device_info(char *name, int name_len, int *sm_count, int64_t *mem_bytes) {
  FB_REQUIRE_INIT();
  cudaDeviceProp prop;
  FB_CUDA(cudaGetDeviceProperties(&prop, ctx().device));
  if (name && name_len > 0) { strncpy(name, prop.name, name_len - 1); name[name_len - 1] = 0; }
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (mem_bytes) *mem_bytes = (int64_t)prop.totalGlobalMem;
  return 0;
}

// ---- communication ---------------------------------------------------------------------------
int fistr_b200_comm_get_unique_id(void *id128) {
  Context &C = ctx();
  std::string err;
  if (!C.nccl.load(err)) return fail(FISTR_B200_ERR_COMM, "%s", err.c_str());
  ncclUniqueId_ id;
  FB_NCCL(C.nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof id);
  return 0;
}

int fistr_b200_comm_init(const void *id128, int my_rank, int petot) {
  FB_REQUIRE_INIT();
  Context &C = ctx();
  if (petot <= 1) { C.rank = 0; C.petot = 1; return 0; }
  std::string err;
  if (!C.nccl.load(err)) return fail(FISTR_B200_ERR_COMM, "%s", err.c_str());
  ncclUniqueId_ id;
  memcpy(&id, id128, sizeof id);
  FB_NCCL(C.nccl.CommInitRank(&C.comm, petot, id, my_rank));
  C.rank = my_rank; C.petot = petot;
  return 0;
}

int fistr_b200_comm_finalize(void) {
  Context &C = ctx();
  p2p_teardown(C);
  if (C.comm) { cudaStreamSynchronize(C.stream); C.nccl.CommDestroy(C.comm); C.comm = nullptr; }
  C.petot = 1; C.rank = 0;
  return 0;
}

// ---- hecmw_mat_con ---------------------------------------------------------------------------
int fistr_b200_mat_con_count(int n_node, int nn_internal, int n_elem_type, const int *elem_type_index,
                             const int *elem_type_item, const int *elem_node_index,
                             const int *elem_node_item, int *N, int *NP, int *NPL, int *NPU) {
  Context &C = ctx();
  int rc = C.con.build(n_node, nn_internal, n_elem_type, elem_type_index, elem_type_item, elem_node_index, elem_node_item);
  if (rc == -1) return fail(FISTR_B200_ERR_ARG, "mat_con: unsupported element type");
  if (rc == -2) return fail(FISTR_B200_ERR_ARG, "mat_con: node id out of range");
  if (rc == -3) return fail(FISTR_B200_ERR_ARG, "mat_con: 9*NPL exceeds the reference's 32-bit index range");
  *N = C.con.N; *NP = C.con.NP; *NPL = (int)C.con.itemL.size(); *NPU = (int)C.con.itemU.size();
  return 0;
}

int fistr_b200_mat_con_fill(int *indexL, int *itemL, int *indexU, int *itemU) {
  Context &C = ctx();
  if (C.con.indexL.empty()) return fail(FISTR_B200_ERR_STATE, "mat_con_fill before mat_con_count");
  memcpy(indexL, C.con.indexL.data(), sizeof(int) * C.con.indexL.size());
  memcpy(indexU, C.con.indexU.data(), sizeof(int) * C.con.indexU.size());
  if (!C.con.itemL.empty()) memcpy(itemL, C.con.itemL.data(), sizeof(int) * C.con.itemL.size());
  if (!C.con.itemU.empty()) memcpy(itemU, C.con.itemU.data(), sizeof(int) * C.con.itemU.size());
  C.con = MatCon();
  return 0;
}

// ---- residency -------------------------------------------------------------------------------
int fistr_b200_mat_structure(int N, int NP, int NPL, int NPU, const int *indexL, const int *itemL,
                             const int *indexU, const int *itemU, int n_neighbor_pe,
                             const int *neighbor_pe, const int *import_index, const int *import_item,
                             const int *export_index, const int *export_item) {
  FB_REQUIRE_INIT();
  Context &C = ctx();
  if (N < 0 || NP < N || NPL < 0 || NPU < 0 || !indexL || !indexU) return fail(FISTR_B200_ERR_ARG, "mat_structure: bad sizes");
  if ((long long)NP + NPL + NPU >= 0xFFFFFFFELL) return fail(FISTR_B200_ERR_ARG, "mat_structure: too many blocks for 32-bit slots");
  C.N = N; C.NP = NP; C.NPL = NPL; C.NPU = NPU;
  C.h_indexL.assign(indexL, indexL + NP + 1);
  C.h_indexU.assign(indexU, indexU + NP + 1);
  if (C.indexL.alloc((size_t)NP + 1) || C.indexU.alloc((size_t)NP + 1) || C.itemL.alloc((size_t)NPL) || C.itemU.alloc((size_t)NPU)) return FISTR_B200_ERR_CUDA;
  FB_CUDA(cudaMemcpy(C.indexL.p, indexL, sizeof(int) * ((size_t)NP + 1), cudaMemcpyHostToDevice));
  FB_CUDA(cudaMemcpy(C.indexU.p, indexU, sizeof(int) * ((size_t)NP + 1), cudaMemcpyHostToDevice));
  if (NPL) FB_CUDA(cudaMemcpy(C.itemL.p, itemL, sizeof(int) * (size_t)NPL, cudaMemcpyHostToDevice));
  if (NPU) FB_CUDA(cudaMemcpy(C.itemU.p, itemU, sizeof(int) * (size_t)NPU, cudaMemcpyHostToDevice));
  const size_t nblk = (size_t)NP + NPL + NPU;
  if (C.vals.alloc(9 * nblk)) return FISTR_B200_ERR_CUDA;
  p2p_teardown(C);   // P may alias the old peer arena
  FB_TRY(ensure_work(C));
  FB_CUDA(cudaMemsetAsync(C.X.p, 0, sizeof(double) * 3 * (size_t)NP, C.stream));
  FB_CUDA(cudaMemsetAsync(C.B.p, 0, sizeof(double) * 3 * (size_t)NP, C.stream));
  // halo tables
  Halo &H = C.halo;
  H.n_neighbor = n_neighbor_pe;
  H.neighbor_pe.assign(neighbor_pe, neighbor_pe + n_neighbor_pe);
  if (n_neighbor_pe > 0) {
    H.import_index.assign(import_index, import_index + n_neighbor_pe + 1);
    H.export_index.assign(export_index, export_index + n_neighbor_pe + 1);
    H.import_item.assign(import_item, import_item + H.import_index.back());
    H.export_item.assign(export_item, export_item + H.export_index.back());
  } else {
    H.import_index.assign(1, 0); H.export_index.assign(1, 0); H.import_item.clear(); H.export_item.clear();
  }
  H.import_contiguous = (H.n_import() == NP - N);
  for (int k = 0; k < H.n_import() && H.import_contiguous; k++) H.import_contiguous = (H.import_item[k] == N + 1 + k);
  if (H.d_export_item.alloc(H.export_item.size()) || H.d_import_item.alloc(H.import_item.size()) ||
      H.send_p.alloc(3 * H.export_item.size()) || H.send_x.alloc(3 * H.export_item.size()) ||
      H.recv.alloc(3 * H.import_item.size())) return FISTR_B200_ERR_CUDA;
  if (!H.export_item.empty()) FB_CUDA(cudaMemcpy(H.d_export_item.p, H.export_item.data(), sizeof(int) * H.export_item.size(), cudaMemcpyHostToDevice));
  if (!H.import_item.empty()) FB_CUDA(cudaMemcpy(H.d_import_item.p, H.import_item.data(), sizeof(int) * H.import_item.size(), cudaMemcpyHostToDevice));
  FB_TRY(p2p_setup(C));
  C.key_indexL = indexL; C.key_itemL = itemL;
  C.have_structure = true;
  C.values_on_device = false;
  C.sell_valid = false;
  C.slot_valid = false;
  C.pc_kind = 0;
  C.ssor.structure_valid = false;
  return 0;
}

int fistr_b200_mat_upload(const double *D, const double *AL, const double *AU, const double *B) {
  FB_REQUIRE_INIT();
  Context &C = ctx();
  if (!C.have_structure) return fail(FISTR_B200_ERR_STATE, "mat_upload before mat_structure");
  if (D) FB_CUDA(cudaMemcpyAsync(C.dD(), D, sizeof(double) * 9 * (size_t)C.NP, cudaMemcpyHostToDevice, C.stream));
  if (AL && C.NPL) FB_CUDA(cudaMemcpyAsync(C.dAL(), AL, sizeof(double) * 9 * (size_t)C.NPL, cudaMemcpyHostToDevice, C.stream));
  if (AU && C.NPU) FB_CUDA(cudaMemcpyAsync(C.dAU(), AU, sizeof(double) * 9 * (size_t)C.NPU, cudaMemcpyHostToDevice, C.stream));
  if (B) FB_CUDA(cudaMemcpyAsync(C.B.p, B, sizeof(double) * 3 * (size_t)C.NP, cudaMemcpyHostToDevice, C.stream));
  FB_CUDA(cudaStreamSynchronize(C.stream));
  if (D || AL || AU) { C.values_on_device = true; C.sell_valid = false; }
  return 0;
}

int fistr_b200_mat_download(double *D, double *AL, double *AU, double *B, double *X) {
  FB_REQUIRE_INIT();
  Context &C = ctx();
  if (!C.have_structure) return fail(FISTR_B200_ERR_STATE, "mat_download before mat_structure");
  FB_CUDA(cudaStreamSynchronize(C.stream));
  if (D) FB_CUDA(cudaMemcpy(D, C.dD(), sizeof(double) * 9 * (size_t)C.NP, cudaMemcpyDeviceToHost));
  if (AL && C.NPL) FB_CUDA(cudaMemcpy(AL, C.dAL(), sizeof(double) * 9 * (size_t)C.NPL, cudaMemcpyDeviceToHost));
  if (AU && C.NPU) FB_CUDA(cudaMemcpy(AU, C.dAU(), sizeof(double) * 9 * (size_t)C.NPU, cudaMemcpyDeviceToHost));
  if (B) FB_CUDA(cudaMemcpy(B, C.B.p, sizeof(double) * 3 * (size_t)C.NP, cudaMemcpyDeviceToHost));
  if (X) FB_CUDA(cudaMemcpy(X, C.X.p, sizeof(double) * 3 * (size_t)C.NP, cudaMemcpyDeviceToHost));
  return 0;
}

int fistr_b200_host_register(void *ptr, int64_t bytes) {
  FB_REQUIRE_INIT();
  if (!ptr || bytes <= 0) return 0;
  FB_CUDA(cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable));
  return 0;
}
int fistr_b200_host_unregister(void *ptr) {
  if (!ptr || !ctx().inited) return 0;
  FB_CUDA(cudaHostUnregister(ptr));
  return 0;
}

// ---- fstr_StiffMatrix --------------------------------------------------------------------------
// Shared front half of the assembly entry points: mesh / section upload (connectivity only when the host array
// changes identity), slot table and its inverse (gather map), element-matrix scratch.
static int assembly_prepare(const char *who, int n_node, const double *node, int n_elem, int n_elem_type,
                            const int *elem_type_index, const int *elem_type_item, const int *elem_node_index,
                            const int *elem_node_item, const int *section_ID, int n_sect,
                            const int *sect_elemopt361, const double *sect_E, const double *sect_nu,
                            std::vector<size_t> &slot_base, double *&ke_out, bool &use_gather) {
  FB_REQUIRE_INIT();
  Context &C = ctx();
  if (!C.have_structure) return fail(FISTR_B200_ERR_STATE, "%s before mat_structure", who);
  if (n_node != C.NP) return fail(FISTR_B200_ERR_ARG, "%s: n_node (%d) != NP (%d)", who, n_node, C.NP);
  for (int it = 0; it < n_elem_type; it++)
    if (elem_type_item[it] != 361 && elem_type_item[it] != 342)
      return fail(FISTR_B200_ERR_ARG, "%s: element type %d is outside the hot-path scope (361, 342)", who, elem_type_item[it]);
  const size_t nconn = (size_t)elem_node_index[n_elem];
  // mesh upload (connectivity is re-uploaded only when the host array changes identity)
  const bool new_mesh = (C.key_elem_node_item != elem_node_item) || !C.slot_valid;
  if (new_mesh) {
    if (C.m_eni.alloc((size_t)n_elem + 1) || C.m_enn.alloc(nconn) || C.m_sec.alloc((size_t)n_elem)) return FISTR_B200_ERR_CUDA;
    FB_CUDA(cudaMemcpy(C.m_eni.p, elem_node_index, sizeof(int) * ((size_t)n_elem + 1), cudaMemcpyHostToDevice));
    FB_CUDA(cudaMemcpy(C.m_enn.p, elem_node_item, sizeof(int) * nconn, cudaMemcpyHostToDevice));
  }
  if (C.m_node.alloc(3 * (size_t)n_node) || C.m_sectD.alloc(3 * (size_t)n_sect) || C.m_form.alloc((size_t)n_sect) || C.m_sec.alloc((size_t)n_elem)) return FISTR_B200_ERR_CUDA;
  FB_CUDA(cudaMemcpy(C.m_node.p, node, sizeof(double) * 3 * (size_t)n_node, cudaMemcpyHostToDevice));
  {
    std::vector<int> sec((size_t)n_elem, 1);
    if (section_ID) sec.assign(section_ID, section_ID + n_elem);
    for (int e = 0; e < n_elem; e++)
      if (sec[e] < 1 || sec[e] > n_sect) return fail(FISTR_B200_ERR_ARG, "%s: section_ID out of range", who);
    FB_CUDA(cudaMemcpy(C.m_sec.p, sec.data(), sizeof(int) * (size_t)n_elem, cudaMemcpyHostToDevice));
    std::vector<double> sd(3 * (size_t)n_sect);
    for (int s = 0; s < n_sect; s++) elem::elastic_constants(sect_E[s], sect_nu[s], sd[3 * s], sd[3 * s + 1], sd[3 * s + 2]);
    FB_CUDA(cudaMemcpy(C.m_sectD.p, sd.data(), sizeof(double) * sd.size(), cudaMemcpyHostToDevice));
    FB_CUDA(cudaMemcpy(C.m_form.p, sect_elemopt361, sizeof(int) * (size_t)n_sect, cudaMemcpyHostToDevice));
  }
  // slot table (topology only)
  slot_base.assign((size_t)n_elem_type + 1, 0);
  for (int it = 0; it < n_elem_type; it++) {
    const int nn = etype_max_node(elem_type_item[it]);
    slot_base[it + 1] = slot_base[it] + (size_t)(elem_type_index[it + 1] - elem_type_index[it]) * nn * nn;
  }
  if (new_mesh) {
    if (C.slot.alloc(slot_base.back())) return FISTR_B200_ERR_CUDA;
    FB_CUDA(cudaMemset(C.flag.p, 0, sizeof(int)));
    for (int it = 0; it < n_elem_type; it++) {
      const int nn = etype_max_node(elem_type_item[it]);
      const int e0 = elem_type_index[it], e1 = elem_type_index[it + 1];
      if (e1 <= e0) continue;
      slot_build_kernel<<<C.grid_for((size_t)(e1 - e0) * nn * nn, 256, 16), 256, 0, C.stream>>>(
          e0, e1, nn, C.m_eni.p, C.m_enn.p, C.NP, C.NPL, C.indexL.p, C.itemL.p, C.indexU.p, C.itemU.p,
          C.slot.p, slot_base[it], C.flag.p);
      FB_LAUNCH_CHECK();
    }
    int err = 0;
    FB_CUDA(cudaMemcpyAsync(&err, C.flag.p, sizeof(int), cudaMemcpyDeviceToHost, C.stream));
    FB_CUDA(cudaStreamSynchronize(C.stream));
    if (err) return fail(FISTR_B200_ERR_CONNECTIVITY, "###ERROR### : cannot find connectivity");
    C.slot_valid = true;
    C.key_elem_node_item = elem_node_item;
    // inverse map of the slot table for the gather assembly (topology only)
    C.gather_valid = false;
    const size_t ncontrib = slot_base.back(), nblk = (size_t)C.NP + C.NPL + C.NPU;
    if (C.get("asm_gather", 1) != 0 && ncontrib < 0xFFFFFFF0ull && ncontrib > 0) {
      DevBuf<unsigned> cursor;
      DevBuf<unsigned char> tmp;
      size_t tmp_bytes = 0;
      cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, (unsigned *)nullptr, (unsigned *)nullptr, (int)1);
      if (!(C.gptr.alloc(nblk + 1) || C.gsrc.alloc(ncontrib) || cursor.alloc(nblk + 1))) {
        FB_CUDA(cudaMemsetAsync(cursor.p, 0, sizeof(unsigned) * (nblk + 1), C.stream));
        gather_count_kernel<<<C.grid_for(ncontrib, 256, 16), 256, 0, C.stream>>>(ncontrib, C.slot.p, cursor.p);
        FB_LAUNCH_CHECK();
        // exclusive scan in chunks of < 2^31 items (CUB takes an int count), carrying the running offset
        FB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cursor.p, C.gptr.p, (int)std::min<size_t>(nblk + 1, 1u << 30), C.stream));
        if (tmp.alloc(tmp_bytes + 16)) return FISTR_B200_ERR_CUDA;
        if (nblk + 1 > (1u << 30)) return fail(FISTR_B200_ERR_ARG, "%s: too many blocks for the gather map", who);
        FB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tmp_bytes, cursor.p, C.gptr.p, (int)(nblk + 1), C.stream));
        FB_CUDA(cudaMemsetAsync(cursor.p, 0, sizeof(unsigned) * (nblk + 1), C.stream));
        gather_fill_kernel<<<C.grid_for(ncontrib, 256, 16), 256, 0, C.stream>>>(ncontrib, C.slot.p, C.gptr.p, cursor.p, C.gsrc.p);
        FB_LAUNCH_CHECK();
        gather_sort_kernel<<<C.grid_for(nblk, 256, 16), 256, 0, C.stream>>>(nblk, C.gptr.p, C.gsrc.p);
        FB_LAUNCH_CHECK();
        FB_CUDA(cudaStreamSynchronize(C.stream));
        cursor.release(); tmp.release();
        C.gather_valid = true;
      } else {
        C.gptr.release(); C.gsrc.release(); cursor.release();   // not enough memory: atomic scatter path
        cudaGetLastError();
      }
    }
  }
  // element-matrix scratch of the gather assembly; falls back to the atomic scatter if it does not fit
  use_gather = C.gather_valid && C.get("asm_gather", 1) != 0;
  if (use_gather && C.ke.alloc(9 * slot_base.back())) { cudaGetLastError(); use_gather = false; }
  ke_out = use_gather ? C.ke.p : nullptr;
  return 0;
}

int fistr_b200_stiffmatrix_3d(int n_node, const double *node, int n_elem, int n_elem_type,
                              const int *elem_type_index, const int *elem_type_item,
                              const int *elem_node_index, const int *elem_node_item,
                              const int *section_ID, int n_sect, const int *sect_elemopt361,
                              const double *sect_E, const double *sect_nu, double *D, double *AL,
                              double *AU, double *time_ms) {
  const double t_begin = now_s();
  std::vector<size_t> slot_base;
  double *ke_out = nullptr;
  bool use_gather = false;
  FB_TRY(assembly_prepare("stiffmatrix_3d", n_node, node, n_elem, n_elem_type, elem_type_index, elem_type_item,
                          elem_node_index, elem_node_item, section_ID, n_sect, sect_elemopt361, sect_E, sect_nu,
                          slot_base, ke_out, use_gather));
  Context &C = ctx();
  // hecmw_mat_clear + element loop
  cudaEvent_t ev0, ev1;
  FB_CUDA(cudaEventCreate(&ev0)); FB_CUDA(cudaEventCreate(&ev1));
  FB_CUDA(cudaEventRecord(ev0, C.stream));
  if (!use_gather) FB_CUDA(cudaMemsetAsync(C.vals.p, 0, sizeof(double) * 9 * ((size_t)C.NP + C.NPL + C.NPU), C.stream));   // hecmw_mat_clear
  constexpr int WARPS = 4;
  const size_t smem = sizeof(elem::Scratch) * WARPS;
  static bool attr_set = false;
  if (!attr_set) {
    FB_CUDA(cudaFuncSetAttribute(stiff_kernel<elem::FORM_IC, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FB_CUDA(cudaFuncSetAttribute(stiff_kernel<elem::FORM_FI, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FB_CUDA(cudaFuncSetAttribute(stiff_kernel<elem::FORM_BBAR, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    FB_CUDA(cudaFuncSetAttribute(stiff_kernel<elem::FORM_TET10, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int probe = (int)C.get("asm_probe", 0);
  for (int it = 0; it < n_elem_type; it++) {
    const int e0 = elem_type_index[it], e1 = elem_type_index[it + 1];
    if (e1 <= e0) continue;
    const int grid = C.grid_for((size_t)(e1 - e0) * 32, WARPS * 32, 8);
    if (elem_type_item[it] == 342) {
      stiff_kernel<elem::FORM_TET10, WARPS><<<grid, WARPS * 32, smem, C.stream>>>(e0, e1, C.m_eni.p, C.m_enn.p, C.m_node.p, C.m_sec.p, C.m_sectD.p, C.m_form.p, C.slot.p, slot_base[it], C.vals.p, ke_out, probe);
      FB_LAUNCH_CHECK();
    } else {
      bool has[5] = {false, false, false, false, false};
      for (int s = 0; s < n_sect; s++) {
        if (sect_elemopt361[s] < 1 || sect_elemopt361[s] > 3)
          return fail(FISTR_B200_ERR_ARG, "stiffmatrix_3d: elemopt361=%d not supported (1 FI, 2 BBAR, 3 IC)", sect_elemopt361[s]);
        has[sect_elemopt361[s]] = true;
      }
      if (has[3]) { stiff_kernel<elem::FORM_IC, WARPS><<<grid, WARPS * 32, smem, C.stream>>>(e0, e1, C.m_eni.p, C.m_enn.p, C.m_node.p, C.m_sec.p, C.m_sectD.p, C.m_form.p, C.slot.p, slot_base[it], C.vals.p, ke_out, probe); FB_LAUNCH_CHECK(); }
      if (has[1]) { stiff_kernel<elem::FORM_FI, WARPS><<<grid, WARPS * 32, smem, C.stream>>>(e0, e1, C.m_eni.p, C.m_enn.p, C.m_node.p, C.m_sec.p, C.m_sectD.p, C.m_form.p, C.slot.p, slot_base[it], C.vals.p, ke_out, probe); FB_LAUNCH_CHECK(); }
      if (has[2]) { stiff_kernel<elem::FORM_BBAR, WARPS><<<grid, WARPS * 32, smem, C.stream>>>(e0, e1, C.m_eni.p, C.m_enn.p, C.m_node.p, C.m_sec.p, C.m_sectD.p, C.m_form.p, C.slot.p, slot_base[it], C.vals.p, ke_out, probe); FB_LAUNCH_CHECK(); }
    }
  }
  if (use_gather) {
    const size_t nblk = (size_t)C.NP + C.NPL + C.NPU;
    gather_sum_kernel<<<C.grid_for(9 * nblk, 256, 8), 256, 0, C.stream>>>(nblk, C.gptr.p, C.gsrc.p, C.ke.p, C.vals.p);
    FB_LAUNCH_CHECK();
  }
  FB_CUDA(cudaEventRecord(ev1, C.stream));
  FB_CUDA(cudaStreamSynchronize(C.stream));
  float ms = 0.f;
  FB_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
  cudaEventDestroy(ev0); cudaEventDestroy(ev1);
  C.values_on_device = true;
  C.sell_valid = false;
  FB_TRY(fistr_b200_mat_download(D, AL, AU, nullptr, nullptr));
  if (time_ms) { time_ms[0] = ms; time_ms[1] = (now_s() - t_begin) * 1e3; }
  return 0;
}

// ---- nonlinear (total Lagrangian) assembly / update: fstr_StiffMatrix + fstr_UpdateNewton under NLGEOM ------
static int nl_upload_materials(Context &C, int n_sect, const int *sect_elemopt361, const double *sect_E,
                               const double *sect_nu, const int *sect_mtype, const double *sect_hyper, const int *ett,
                               int n_elem_type) {
  std::vector<elem::MatlTL> ml((size_t)n_sect);
  bool has361 = false;
  for (int it = 0; it < n_elem_type; it++) has361 = has361 || (ett[it] == 361);
  for (int s = 0; s < n_sect; s++) {
    elem::MatlTL &M = ml[s];
    M.mtype = sect_mtype ? sect_mtype[s] : 0;
    if (M.mtype != elem::MAT_ELASTIC && M.mtype != elem::MAT_MOONEYRIVLIN)
      return fail(FISTR_B200_ERR_ARG, "nl: material type %d not supported (0 ELASTIC, 1 MOONEYRIVLIN)", M.mtype);
    elem::elastic_constants(sect_E[s], sect_nu[s], M.D11, M.D12, M.D44);
    for (int k = 0; k < 3; k++) M.c[k] = sect_hyper ? sect_hyper[3 * s + k] : 0.0;
    if (M.mtype == elem::MAT_MOONEYRIVLIN && !(M.c[2] > 0.0)) return fail(FISTR_B200_ERR_ARG, "nl: Mooney-Rivlin D must be > 0");
    // the 361 IC element aborts under NLGEOM in the reference (fstr_StiffMatrix.f90:118, fstr_Update.f90:163)
    if (has361 && sect_elemopt361[s] != 1 && sect_elemopt361[s] != 2)
      return fail(FISTR_B200_ERR_ARG, "###ERROR### : nonlinear analysis not supported with 361 IC element (elemopt361=%d)", sect_elemopt361[s]);
  }
  if (C.m_matl.alloc((size_t)n_sect)) return FISTR_B200_ERR_CUDA;
  FB_CUDA(cudaMemcpy(C.m_matl.p, ml.data(), sizeof(elem::MatlTL) * ml.size(), cudaMemcpyHostToDevice));
  return 0;
}

static int nl_ensure_state(Context &C, int n_elem) {
  if (C.nl_n_elem == n_elem && C.gp_strain.p) return 0;
  const size_t n = 48 * (size_t)std::max(n_elem, 1);
  if (C.gp_strain.alloc(n) || C.gp_stress.alloc(n)) return FISTR_B200_ERR_CUDA;
  FB_CUDA(cudaMemsetAsync(C.gp_strain.p, 0, sizeof(double) * n, C.stream));
  FB_CUDA(cudaMemsetAsync(C.gp_stress.p, 0, sizeof(double) * n, C.stream));
  C.nl_n_elem = n_elem;
  return 0;
}

constexpr int kNlWarps = 2;   // warps (= elements in flight) per block of the TL kernels: ScratchTL is 16 KB per warp
static int nl_set_attrs() {
  constexpr int WARPS = kNlWarps;
  static bool done = false;
  if (done) return 0;
  const int smem = (int)(sizeof(elem::ScratchTL) * WARPS);
  FB_CUDA(cudaFuncSetAttribute(stiff_tl_kernel<elem::FORM_FI, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  FB_CUDA(cudaFuncSetAttribute(stiff_tl_kernel<elem::FORM_BBAR, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  FB_CUDA(cudaFuncSetAttribute(stiff_tl_kernel<elem::FORM_TET10, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  FB_CUDA(cudaFuncSetAttribute(update_tl_kernel<elem::FORM_FI, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  FB_CUDA(cudaFuncSetAttribute(update_tl_kernel<elem::FORM_BBAR, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  FB_CUDA(cudaFuncSetAttribute(update_tl_kernel<elem::FORM_TET10, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  done = true;
  return 0;
}

int fistr_b200_nl_state_reset(int n_elem) {
  FB_REQUIRE_INIT();
  Context &C = ctx();
  C.nl_n_elem = -1;
  FB_TRY(nl_ensure_state(C, n_elem));
  FB_CUDA(cudaStreamSynchronize(C.stream));
  return 0;
}

int fistr_b200_nl_state_download(double *strain, double *stress) {
  FB_REQUIRE_INIT();
  Context &C = ctx();
  if (C.nl_n_elem <= 0) return fail(FISTR_B200_ERR_STATE, "nl_state_download: no Gauss-point state");
  FB_CUDA(cudaStreamSynchronize(C.stream));
  if (strain) FB_CUDA(cudaMemcpy(strain, C.gp_strain.p, sizeof(double) * 48 * (size_t)C.nl_n_elem, cudaMemcpyDeviceToHost));
  if (stress) FB_CUDA(cudaMemcpy(stress, C.gp_stress.p, sizeof(double) * 48 * (size_t)C.nl_n_elem, cudaMemcpyDeviceToHost));
  return 0;
}

int fistr_b200_nl_stiffmatrix_3d(int n_node, const double *node, int n_elem, int n_elem_type,
                                 const int *elem_type_index, const int *elem_type_item,
                                 const int *elem_node_index, const int *elem_node_item, const int *section_ID,
                                 int n_sect, const int *sect_elemopt361, const double *sect_E, const double *sect_nu,
                                 const int *sect_mtype, const double *sect_hyper, const double *u, double *D,
                                 double *AL, double *AU, double *time_ms) {
  const double t_begin = now_s();
  std::vector<size_t> slot_base;
  double *ke_out = nullptr;
  bool use_gather = false;
  FB_TRY(assembly_prepare("nl_stiffmatrix_3d", n_node, node, n_elem, n_elem_type, elem_type_index, elem_type_item,
                          elem_node_index, elem_node_item, section_ID, n_sect, sect_elemopt361, sect_E, sect_nu,
                          slot_base, ke_out, use_gather));
  Context &C = ctx();
  FB_TRY(nl_upload_materials(C, n_sect, sect_elemopt361, sect_E, sect_nu, sect_mtype, sect_hyper, elem_type_item, n_elem_type));
  FB_TRY(nl_ensure_state(C, n_elem));
  if (C.m_u.alloc(3 * (size_t)n_node)) return FISTR_B200_ERR_CUDA;
  if (u) FB_CUDA(cudaMemcpyAsync(C.m_u.p, u, sizeof(double) * 3 * (size_t)n_node, cudaMemcpyHostToDevice, C.stream));
  else FB_CUDA(cudaMemsetAsync(C.m_u.p, 0, sizeof(double) * 3 * (size_t)n_node, C.stream));
  constexpr int WARPS = kNlWarps;
  FB_TRY(nl_set_attrs());
  const size_t smem = sizeof(elem::ScratchTL) * WARPS;
  cudaEvent_t ev0, ev1;
  FB_CUDA(cudaEventCreate(&ev0)); FB_CUDA(cudaEventCreate(&ev1));
  FB_CUDA(cudaEventRecord(ev0, C.stream));
  if (!use_gather) FB_CUDA(cudaMemsetAsync(C.vals.p, 0, sizeof(double) * 9 * ((size_t)C.NP + C.NPL + C.NPU), C.stream));
  for (int it = 0; it < n_elem_type; it++) {
    const int e0 = elem_type_index[it], e1 = elem_type_index[it + 1];
    if (e1 <= e0) continue;
    const int grid = C.grid_for((size_t)(e1 - e0) * 32, WARPS * 32, 8);
#define FB_TL_STIFF(FORM_)                                                                                              \
  stiff_tl_kernel<FORM_, WARPS><<<grid, WARPS * 32, smem, C.stream>>>(e0, e1, C.m_eni.p, C.m_enn.p, C.m_node.p, C.m_u.p, \
      C.m_sec.p, C.m_matl.p, C.m_form.p, C.gp_strain.p, C.gp_stress.p, C.slot.p, slot_base[it], C.vals.p, ke_out);       \
  FB_LAUNCH_CHECK();
    if (elem_type_item[it] == 342) { FB_TL_STIFF(elem::FORM_TET10) }
    else {
      bool has[3] = {false, false, false};
      for (int s = 0; s < n_sect; s++) has[sect_elemopt361[s]] = true;   // 1 or 2, validated above
      if (has[1]) { FB_TL_STIFF(elem::FORM_FI) }
      if (has[2]) { FB_TL_STIFF(elem::FORM_BBAR) }
    }
#undef FB_TL_STIFF
  }
  if (use_gather) {
    const size_t nblk = (size_t)C.NP + C.NPL + C.NPU;
    gather_sum_kernel<<<C.grid_for(9 * nblk, 256, 8), 256, 0, C.stream>>>(nblk, C.gptr.p, C.gsrc.p, C.ke.p, C.vals.p);
    FB_LAUNCH_CHECK();
  }
  FB_CUDA(cudaEventRecord(ev1, C.stream));
  FB_CUDA(cudaStreamSynchronize(C.stream));
  float ms = 0.f;
  FB_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
  cudaEventDestroy(ev0); cudaEventDestroy(ev1);
  C.values_on_device = true;
  C.sell_valid = false;
  FB_TRY(fistr_b200_mat_download(D, AL, AU, nullptr, nullptr));
  if (time_ms) { time_ms[0] = ms; time_ms[1] = (now_s() - t_begin) * 1e3; }
  return 0;
}

int fistr_b200_nl_update_3d(int n_node, const double *node, int n_elem, int n_elem_type,
                            const int *elem_type_index, const int *elem_type_item, const int *elem_node_index,
                            const int *elem_node_item, const int *section_ID, int n_sect,
                            const int *sect_elemopt361, const double *sect_E, const double *sect_nu,
                            const int *sect_mtype, const double *sect_hyper, const double *unode,
                            const double *dunode, double *QFORCE) {
  std::vector<size_t> slot_base;
  double *ke_out = nullptr;
  bool use_gather = false;
  FB_TRY(assembly_prepare("nl_update_3d", n_node, node, n_elem, n_elem_type, elem_type_index, elem_type_item,
                          elem_node_index, elem_node_item, section_ID, n_sect, sect_elemopt361, sect_E, sect_nu,
                          slot_base, ke_out, use_gather));
  Context &C = ctx();
  FB_TRY(nl_upload_materials(C, n_sect, sect_elemopt361, sect_E, sect_nu, sect_mtype, sect_hyper, elem_type_item, n_elem_type));
  FB_TRY(nl_ensure_state(C, n_elem));
  const size_t n3 = 3 * (size_t)n_node;
  if (C.m_u.alloc(n3) || C.m_qf.alloc(n3)) return FISTR_B200_ERR_CUDA;
  {   // totaldisp = unode + dunode (UPDATE_C3: static_LIB_3d.f90:561)
    std::vector<double> tot(n3);
    for (size_t i = 0; i < n3; i++) tot[i] = (unode ? unode[i] : 0.0) + (dunode ? dunode[i] : 0.0);
    FB_CUDA(cudaMemcpy(C.m_u.p, tot.data(), sizeof(double) * n3, cudaMemcpyHostToDevice));
  }
  FB_CUDA(cudaMemsetAsync(C.m_qf.p, 0, sizeof(double) * n3, C.stream));
  constexpr int WARPS = kNlWarps;
  FB_TRY(nl_set_attrs());
  const size_t smem = sizeof(elem::ScratchTL) * WARPS;
  for (int it = 0; it < n_elem_type; it++) {
    const int e0 = elem_type_index[it], e1 = elem_type_index[it + 1];
    if (e1 <= e0) continue;
    const int grid = C.grid_for((size_t)(e1 - e0) * 32, WARPS * 32, 8);
#define FB_TL_UPDATE(FORM_)                                                                                              \
  update_tl_kernel<FORM_, WARPS><<<grid, WARPS * 32, smem, C.stream>>>(e0, e1, C.m_eni.p, C.m_enn.p, C.m_node.p, C.m_u.p, \
      C.m_sec.p, C.m_matl.p, C.m_form.p, C.gp_strain.p, C.gp_stress.p, C.m_qf.p);                                         \
  FB_LAUNCH_CHECK();
    if (elem_type_item[it] == 342) { FB_TL_UPDATE(elem::FORM_TET10) }
    else {
      bool has[3] = {false, false, false};
      for (int s = 0; s < n_sect; s++) has[sect_elemopt361[s]] = true;
      if (has[1]) { FB_TL_UPDATE(elem::FORM_FI) }
      if (has[2]) { FB_TL_UPDATE(elem::FORM_BBAR) }
    }
#undef FB_TL_UPDATE
  }
  FB_CUDA(cudaMemcpyAsync(QFORCE, C.m_qf.p, sizeof(double) * n3, cudaMemcpyDeviceToHost, C.stream));
  FB_CUDA(cudaStreamSynchronize(C.stream));
  return 0;
}

// ---- hecmw_mat_ass_bc --------------------------------------------------------------------------
int fistr_b200_mat_ass_bc(int nbc, const int *inode, const int *idof, const double *rhs) {
  FB_REQUIRE_INIT();
  Context &C = ctx();
  if (!C.have_structure || !C.values_on_device) return fail(FISTR_B200_ERR_STATE, "mat_ass_bc needs a device-resident matrix");
  if (nbc <= 0) return 0;
  std::vector<int> nodes, dofs, ptr(1, 0);
  std::vector<double> vals;
  for (int c = 0; c < nbc; c++) {
    if (idof[c] > 3) continue;   // 'if( NDOF < idof ) return', hecmw_mat_ass.f90:301
    if (inode[c] < 1 || inode[c] > C.NP || idof[c] < 1) return fail(FISTR_B200_ERR_ARG, "mat_ass_bc: bad node/dof");
    const int n = inode[c];
    nodes.push_back(n); dofs.push_back(idof[c]); vals.push_back(rhs[c]);
    ptr.push_back(ptr.back() + 1 + (C.h_indexL[n] - C.h_indexL[n - 1]) + (C.h_indexU[n] - C.h_indexU[n - 1]));
  }
  const int m = (int)nodes.size();
  if (m == 0) return 0;
  DevBuf<int> dn, dd, dp;
  DevBuf<double> dv;
  if (dn.alloc(m) || dd.alloc(m) || dp.alloc(m + 1) || dv.alloc(m)) return FISTR_B200_ERR_CUDA;
  FB_CUDA(cudaMemcpy(dn.p, nodes.data(), sizeof(int) * m, cudaMemcpyHostToDevice));
  FB_CUDA(cudaMemcpy(dd.p, dofs.data(), sizeof(int) * m, cudaMemcpyHostToDevice));
  FB_CUDA(cudaMemcpy(dp.p, ptr.data(), sizeof(int) * (m + 1), cudaMemcpyHostToDevice));
  FB_CUDA(cudaMemcpy(dv.p, vals.data(), sizeof(double) * m, cudaMemcpyHostToDevice));
  const int g = C.grid_for((size_t)ptr.back(), 256, 8);
  bc_kernel<0><<<g, 256, 0, C.stream>>>(m, dn.p, dd.p, dv.p, dp.p, C.NP, C.NPL, C.indexL.p, C.itemL.p, C.indexU.p, C.itemU.p, C.vals.p, C.B.p);
  FB_LAUNCH_CHECK();
  bc_kernel<1><<<g, 256, 0, C.stream>>>(m, dn.p, dd.p, dv.p, dp.p, C.NP, C.NPL, C.indexL.p, C.itemL.p, C.indexU.p, C.itemU.p, C.vals.p, C.B.p);
  FB_LAUNCH_CHECK();
  bc_set_rhs_kernel<<<(m + 255) / 256, 256, 0, C.stream>>>(m, dn.p, dd.p, dv.p, C.B.p);
  FB_LAUNCH_CHECK();
  FB_CUDA(cudaStreamSynchronize(C.stream));
  dn.release(); dd.release(); dp.release(); dv.release();
  C.sell_valid = false;
  return 0;
}

// ---- hecmw_matvec_33 / update_3_R / precond entry points ----------------------------------------
int fistr_b200_update_3_R(double *X) {
  FB_REQUIRE_INIT();
  Context &C = ctx();
  if (!C.have_structure) return fail(FISTR_B200_ERR_STATE, "update_3_R before mat_structure");
  FB_CUDA(cudaMemcpyAsync(C.WK.p, X, sizeof(double) * 3 * (size_t)C.NP, cudaMemcpyHostToDevice, C.stream));
  FB_TRY(exchange(C, C.WK.p, false));
  FB_CUDA(cudaMemcpyAsync(X, C.WK.p, sizeof(double) * 3 * (size_t)C.NP, cudaMemcpyDeviceToHost, C.stream));
  FB_CUDA(cudaStreamSynchronize(C.stream));
  return 0;
}

int fistr_b200_matvec_33(const double *X, double *Y) {
  FB_REQUIRE_INIT();
  Context &C = ctx();
  if (!C.have_structure || !C.values_on_device) return fail(FISTR_B200_ERR_STATE, "matvec_33 needs a device-resident matrix");
  FB_CUDA(cudaMemcpyAsync(C.P.p, X, sizeof(double) * 3 * (size_t)C.NP, cudaMemcpyHostToDevice, C.stream));
  FB_TRY(exchange(C, C.P.p, false));
  FB_TRY(launch_spmv(C, 0, C.P.p, nullptr, C.Q.p, C.red.p + RED_TMP, nullptr));
  FB_CUDA(cudaMemcpyAsync(Y, C.Q.p, sizeof(double) * 3 * (size_t)C.N, cudaMemcpyDeviceToHost, C.stream));
  FB_CUDA(cudaStreamSynchronize(C.stream));
  return 0;
}

__global__ void sin_fill_kernel(size_t n, double *x) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] = sin((double)(i + 1));
}
__global__ void flush_kernel(size_t n, double *x) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] = x[i] * 0.5 + 1.0;
}

int fistr_b200_matvec_bench(int nrep, int flush_l2, double *ms_mean, double *ms_min) {
  FB_REQUIRE_INIT();
  Context &C = ctx();
  if (!C.have_structure || !C.values_on_device) return fail(FISTR_B200_ERR_STATE, "matvec_bench needs a device-resident matrix");
  sin_fill_kernel<<<C.grid_for(3 * (size_t)C.NP, 256, 8), 256, 0, C.stream>>>(3 * (size_t)C.NP, C.P.p);
  FB_LAUNCH_CHECK();
  DevBuf<double> fl;
  const size_t nfl = 48u << 20;   // 384 MB > L2
  if (flush_l2 && fl.alloc(nfl)) return FISTR_B200_ERR_CUDA;
  for (int w = 0; w < 3; w++) FB_TRY(launch_spmv(C, 0, C.P.p, nullptr, C.Q.p, C.red.p + RED_TMP, nullptr));
  cudaEvent_t e0, e1;
  FB_CUDA(cudaEventCreate(&e0)); FB_CUDA(cudaEventCreate(&e1));
  double sum = 0.0, mn = 1e30;
  for (int r = 0; r < nrep; r++) {
    if (flush_l2) { flush_kernel<<<C.sm_count * 8, 256, 0, C.stream>>>(nfl, fl.p); FB_LAUNCH_CHECK(); }
    FB_CUDA(cudaEventRecord(e0, C.stream));
    FB_TRY(launch_spmv(C, 0, C.P.p, nullptr, C.Q.p, C.red.p + RED_TMP, nullptr));
    FB_CUDA(cudaEventRecord(e1, C.stream));
    FB_CUDA(cudaEventSynchronize(e1));
    float ms;
    FB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    sum += ms; mn = std::min(mn, (double)ms);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  fl.release();
  if (ms_mean) *ms_mean = sum / std::max(nrep, 1);
  if (ms_min) *ms_min = mn;
  return 0;
}

int fistr_b200_precond_33_setup(int precond, double sigma_diag, int ncolor_in) {
  FB_REQUIRE_INIT();
  Context &C = ctx();
  if (!C.have_structure || !C.values_on_device) return fail(FISTR_B200_ERR_STATE, "precond_33_setup needs a device-resident matrix");
  FB_TRY(precond_setup(C, precond, sigma_diag, ncolor_in, true));
  FB_CUDA(cudaStreamSynchronize(C.stream));
  return 0;
}
int fistr_b200_precond_33_apply(double *ZP) {
  FB_REQUIRE_INIT();
  Context &C = ctx();
  if (C.pc_kind == 0) return fail(FISTR_B200_ERR_STATE, "precond_33_apply before precond_33_setup");
  FB_CUDA(cudaMemcpyAsync(C.WK.p, ZP, sizeof(double) * 3 * (size_t)C.NP, cudaMemcpyHostToDevice, C.stream));
  FB_TRY(precond_apply_dev(C, C.WK.p, nullptr));
  FB_CUDA(cudaMemcpyAsync(ZP, C.WK.p, sizeof(double) * 3 * (size_t)C.NP, cudaMemcpyDeviceToHost, C.stream));
  FB_CUDA(cudaStreamSynchronize(C.stream));
  return 0;
}
int fistr_b200_precond_33_clear(void) {
  Context &C = ctx();
  C.pc_kind = 0;
  C.ssor.structure_valid = false;
  C.ssor.values_valid = false;
  return 0;
}
int fistr_b200_precond_33_info(int *ncolor, int *perm) {
  Context &C = ctx();
  if (C.pc_kind != 1) { if (ncolor) *ncolor = 0; return 0; }
  if (ncolor) *ncolor = C.ssor.ncolor;
  if (perm) memcpy(perm, C.ssor.perm.data(), sizeof(int) * C.ssor.perm.size());
  return 0;
}

// ---- hecmw_solve ---------------------------------------------------------------------------------
#define IA(slot) Iarray[(slot)-1]
#define RA(slot) Rarray[(slot)-1]

int fistr_b200_solve_resid2(double *resid2) {
  *resid2 = ctx().last_resid2;
  return 0;
}

int fistr_b200_solve_profile(double *seconds4) {
  memcpy(seconds4, ctx().last_profile, sizeof(double) * 4);
  return 0;
}

int fistr_b200_solve_history(double *hist, int max_len) {
  Context &C = ctx();
  const int n = (int)std::min<size_t>(C.last_hist.size(), (size_t)std::max(max_len, 0));
  if (n > 0) memcpy(hist, C.last_hist.data(), sizeof(double) * n);
  return n;
}

__global__ void zero_diag_check_kernel(int N, const double *__restrict__ D, int *flag) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < (size_t)N) {
    const double *d = D + 9 * i;
    if (fabs(d[0]) == 0.0 || fabs(d[4]) == 0.0 || fabs(d[8]) == 0.0) atomicExch(flag, 1);
  }
}

int fistr_b200_solve_33(int N, int NP, int NPL, int NPU, const double *D, const double *AL,
                        const double *AU, const int *indexL, const int *itemL, const int *indexU,
                        const int *itemU, const double *B, double *X, int *Iarray, double *Rarray,
                        int n_neighbor_pe, const int *neighbor_pe, const int *import_index,
                        const int *import_item, const int *export_index, const int *export_item,
                        int my_rank, int petot, int *iter_out, double *resid_out, double *time_out) {
  FB_REQUIRE_INIT();
  Context &C = ctx();
  const double t_entry = now_s();
  cudaEvent_t ev_in = nullptr, ev_out = nullptr;
  FB_CUDA(cudaEventCreate(&ev_in)); FB_CUDA(cudaEventCreate(&ev_out));
  FB_CUDA(cudaEventRecord(ev_in, C.stream));
  if (petot != C.petot) return fail(FISTR_B200_ERR_COMM, "solve_33: petot=%d but the communicator has %d ranks", petot, C.petot);
  (void)my_rank;
  if (IA(99) != 1) return fail(FISTR_B200_ERR_ARG, "solve_33: Iarray(99)=%d (direct solvers are not on this path)", IA(99));
  // --- parameters, Iterative.f90:55-77
  int MAXIT = IA(1);
  const int METHOD = IA(2), PRECOND = IA(3), iterPREmax = IA(5), ITERlog = IA(21), TIMElog = IA(22);
  const double TOL = RA(1);
  double SIGMA_DIAG = RA(2);
  if (SIGMA_DIAG < 0.0) SIGMA_DIAG = 1.0;   // auto mode only matters for the ILU family
  // --- structure / values residency
  bool structure_changed = false;
  const bool same_sizes = C.have_structure && C.N == N && C.NP == NP && C.NPL == NPL && C.NPU == NPU;
  // Iarray(98)=1 with host values given: the caller declares a (possibly) new structure -> re-upload it
  if (!same_sizes || (indexL && (C.key_indexL != indexL || C.key_itemL != itemL)) || (IA(98) >= 1 && indexL && D)) {
    if (!indexL) return fail(FISTR_B200_ERR_STATE, "solve_33: no structure on the device and none given");
    const bool keep_vals = same_sizes && C.values_on_device && !D;   // same shape re-declared, values device-only
    FB_TRY(fistr_b200_mat_structure(N, NP, NPL, NPU, indexL, itemL, indexU, itemU, n_neighbor_pe, neighbor_pe,
                                    import_index, import_item, export_index, export_item));
    if (keep_vals) C.values_on_device = true;
    structure_changed = true;
  }
  if (IA(98) >= 1) structure_changed = true;   // caller asks for a symbolic re-setup (fstr_solve_NonLinear.f90:81-86)
  if (D || AL || AU) {
    if (!(D && (AL || NPL == 0) && (AU || NPU == 0))) return fail(FISTR_B200_ERR_ARG, "solve_33: D, AL, AU must be given together");
    FB_TRY(fistr_b200_mat_upload(D, AL, AU, nullptr));
  }
  if (!C.values_on_device) return fail(FISTR_B200_ERR_STATE, "solve_33: no matrix values on the device");
  if (B) FB_CUDA(cudaMemcpyAsync(C.B.p, B, sizeof(double) * 3 * (size_t)NP, cudaMemcpyHostToDevice, C.stream));
  if (X) FB_CUDA(cudaMemcpyAsync(C.X.p, X, sizeof(double) * 3 * (size_t)NP, cudaMemcpyHostToDevice, C.stream));
  else FB_CUDA(cudaMemsetAsync(C.X.p, 0, sizeof(double) * 3 * (size_t)NP, C.stream));

  const int BLOCK = 256;
  const size_t n3 = 3 * (size_t)N, np3 = 3 * (size_t)NP;
  const int gvec = C.grid_for(n3, BLOCK, 8), gnode = C.grid_for((size_t)N, BLOCK, 8);
  double *red = C.red.p;
  int ret_warn = 0;

  // --- hecmw_solve_check_zerorhs / zerodiag, Iterative.f90:212-278
  FB_CUDA(cudaMemsetAsync(red, 0, sizeof(double) * RED_COUNT, C.stream));
  FB_CUDA(cudaMemsetAsync(C.flag.p, 0, sizeof(int) * 4, C.stream));
  dot_kernel<256><<<gvec, BLOCK, 0, C.stream>>>(n3, C.B.p, C.B.p, C.partial.p, C.counter.p, red + RED_BNRM2, nullptr);
  FB_LAUNCH_CHECK();
  FB_TRY(allreduce(C, red + RED_BNRM2, 1));
  if (N > 0) { zero_diag_check_kernel<<<(N + 255) / 256, 256, 0, C.stream>>>(N, C.dD(), C.flag.p); FB_LAUNCH_CHECK(); }
  double h_red[RED_COUNT];
  int h_flag[4];
  FB_CUDA(cudaMemcpyAsync(h_red, red, sizeof h_red, cudaMemcpyDeviceToHost, C.stream));
  FB_CUDA(cudaMemcpyAsync(h_flag, C.flag.p, sizeof h_flag, cudaMemcpyDeviceToHost, C.stream));
  FB_CUDA(cudaStreamSynchronize(C.stream));
  const double BNRM2 = h_red[RED_BNRM2];
  if (C.petot > 1) {   // hecmw_allreduce_I1(error, hecmw_max)
    double f = h_flag[0];
    FB_CUDA(cudaMemcpyAsync(red + RED_TMP, &f, sizeof f, cudaMemcpyHostToDevice, C.stream));
    FB_TRY(allreduce(C, red + RED_TMP, 1));
    FB_CUDA(cudaMemcpyAsync(&f, red + RED_TMP, sizeof f, cudaMemcpyDeviceToHost, C.stream));
    FB_CUDA(cudaStreamSynchronize(C.stream));
    h_flag[0] = f > 0.0;
  }
  if (BNRM2 == 0.0) ret_warn = HECMW_SOLVER_ERROR_ZERO_RHS_;
  if (h_flag[0] && PRECOND < 10 && iterPREmax > 0) {
    g_last_error = "#### HEC-MW-SOLVER-E-2001: ZERO component in diagonal block";
    return HECMW_SOLVER_ERROR_ZERO_DIAG_;
  }
  // --- hecmw_mat_recycle_precond_setting, matrix_misc.f90:678-697
  if (IA(98) >= 1) { IA(97) = 1; IA(96) = 0; }
  else if (IA(97) > 1) { IA(96) = 0; IA(97) = 1; }
  else if (IA(97) == 1) {
    if (IA(96) < IA(35)) { IA(97) = 0; IA(96) = IA(96) + 1; }
    else IA(96) = 0;
  }
  IA(81) = 0; IA(82) = 0;
  if (METHOD != 1) {
    g_last_error = "#### HEC-MW-SOLVER-E-1001: only METHOD=1 (CG) is on this path";
    return HECMW_SOLVER_ERROR_INCONS_PC_;
  }
  // --- hecmw_precond_setup honouring the recycle flags (DIAG_33.f90:38-41, SSOR_33.f90:72-80)
  if (iterPREmax > 0) {
    const int want = (PRECOND == 3) ? 3 : ((PRECOND == 1 || PRECOND == 2) ? 1 : -1);
    if (want < 0) {
      g_last_error = "#### HEC-MW-SOLVER-E-1001: PRECOND outside 1,2,3";
      return HECMW_SOLVER_ERROR_INCONS_PC_;
    }
    const bool reuse = (C.pc_kind == want) && !structure_changed && IA(98) == 0 && IA(97) == 0;
    if (!reuse) FB_TRY(precond_setup(C, PRECOND, SIGMA_DIAG, IA(34), structure_changed || IA(98) >= 1));
    IA(98) = 0; IA(97) = 0;
  }
  if (!C.sell_valid) FB_TRY(build_sell(C));
  const bool fused = (PRECOND == 3 && iterPREmax == 1 && C.get("cg_fused", 1) != 0);
  // peer-memory mode has no host-visible communication inside the loop: look at the control block less often
  const bool p2p_mode = C.p2p.active && fused && C.petot > 1;
  const int check_interval = (int)std::max<int64_t>(1, C.get("cg_check_interval", p2p_mode ? 16 : 8));

  // --- control block, history
  const int hist_len = std::max(MAXIT, 1);
  if (C.hist.alloc((size_t)hist_len)) return FISTR_B200_ERR_CUDA;
  CgCtl h_ctl;
  memset(&h_ctl, 0, sizeof h_ctl);
  h_ctl.tol = TOL;
  FB_CUDA(cudaMemcpyAsync(C.ctl.p, &h_ctl, sizeof h_ctl, cudaMemcpyHostToDevice, C.stream));
  int *phase = &C.ctl.p->phase;
  // WW = 0 (CG.f90:94-95)
  FB_CUDA(cudaMemsetAsync(C.R.p, 0, sizeof(double) * np3, C.stream));
  FB_CUDA(cudaMemsetAsync(C.Z.p, 0, sizeof(double) * np3, C.stream));
  FB_CUDA(cudaMemsetAsync(C.P.p, 0, sizeof(double) * np3, C.stream));
  FB_CUDA(cudaMemsetAsync(C.Q.p, 0, sizeof(double) * np3, C.stream));
  FB_CUDA(cudaMemsetAsync(C.WK.p, 0, sizeof(double) * np3, C.stream));

  // --- r0 = b - A x0 (CG.f90:120), z1 = Minv r0, rho1
  auto z_and_rho = [&](int par) -> int {   // z = Minv r ; red[2*par+1] = r.z (+allreduce)
    if (fused) {
      cg_precond_dot_kernel<256><<<gnode, BLOCK, 0, C.stream>>>(N, C.alu.p, C.alu_ld, C.R.p, C.Z.p, C.partial.p, C.counter.p, red + 2 * par + 1, phase);
      FB_LAUNCH_CHECK();
    } else {
      FB_TRY(precond_apply_generic(C, iterPREmax, phase));
      dot_kernel<256><<<gvec, BLOCK, 0, C.stream>>>(n3, C.R.p, C.Z.p, C.partial.p, C.counter.p, red + 2 * par + 1, phase);
      FB_LAUNCH_CHECK();
    }
    return allreduce(C, red + 2 * par + 1, 1);
  };
  int ITER = 0, error = 0;
  double RESID = TOL;
  C.last_hist.clear();
  double Tcomm = 0.0, time_Ax = 0.0;
  cudaEvent_t evA0 = nullptr, evA1 = nullptr, evP0 = nullptr, evF1 = nullptr;
  double prof_p = 0.0, prof_f = 0.0, prof_it = 0.0;
  const bool time_ax = TIMElog >= 1 || C.get("time_ax", 0) != 0;
  if (time_ax) { FB_CUDA(cudaEventCreate(&evA0)); FB_CUDA(cudaEventCreate(&evA1)); FB_CUDA(cudaEventCreate(&evP0)); FB_CUDA(cudaEventCreate(&evF1)); }
  int ax_samples = 0;

  if (BNRM2 == 0.0) {
    // CG.f90:124-129: iter=0, MAXIT=0, X=0; the DO loop then leaves ITER=1
    FB_CUDA(cudaMemsetAsync(C.X.p, 0, sizeof(double) * np3, C.stream));
    MAXIT = 0; ITER = 1; RESID = 0.0;
  }
  FB_TRY(exchange(C, C.X.p, false));
  FB_TRY(launch_spmv(C, 1, C.X.p, C.B.p, C.R.p, red + RED_TMP, nullptr));
  FB_CUDA(cudaStreamSynchronize(C.stream));
  const double t_setup_end = now_s();

  if (MAXIT >= 1) FB_TRY(z_and_rho(1));
  int k = 1;
  bool done = (MAXIT < 1);
  // peer-memory mode (p2p.cuh): inside the fused DIAG iteration the halo exchange and the two all-reduces are
  // issued by the compute kernels themselves; the 1-in-50 true-residual iterations stay on NCCL.
  Context::P2P &T = C.p2p;
  const bool p2p = T.active && fused && C.sell.variant == 2 && C.petot > 1;
  unsigned long long pending_rrrz = 0;   // sequence number of the (r.r, r.z) publication the next prologue must consume
  while (!done) {
    // ---- launch iterations k .. kend speculatively
    const int kend = std::min(MAXIT, (k + check_interval - 1) / check_interval * check_interval);
    for (; k <= kend; k++) {
      const int parn = (k + 1) & 1;
      const bool recompute = (k % 50 == 0);   // N_ITER_RECOMPUTE_R, CG.f90:232-233
      P2PArgs ap, as, af;
      ap.consume_seq = pending_rrrz; pending_rrrz = 0;
      if (p2p) { ap.halo_seq = ++T.halo_seq; as.halo_seq = ap.halo_seq; }
      if (p2p && !recompute) { as.publish_seq = ++T.seq[P2P_Q_PQ]; af.consume_seq = as.publish_seq; af.publish_seq = ++T.seq[P2P_Q_RRRZ]; }
      const bool sample = time_ax && (k % check_interval == 1 || check_interval == 1);
      if (sample) FB_CUDA(cudaEventRecord(evP0, C.stream));
      cg_update_p_kernel<256><<<gvec, BLOCK, 0, C.stream>>>(k, n3, C.Z.p, C.P.p, red, C.ctl.p, C.hist.p, hist_len,
                                                            C.halo.n_export(), C.halo.d_export_item.p, C.halo.send_p.p,
                                                            T.view, T.nbrs, ap, T.exp_peer.p, T.exp_remote.p, C.counter.p + 1);
      FB_LAUNCH_CHECK();
      if (!p2p) FB_TRY(exchange(C, C.P.p, true));
      if (sample) FB_CUDA(cudaEventRecord(evA0, C.stream));
      FB_TRY(launch_spmv(C, 0, C.P.p, nullptr, C.Q.p, red + RED_PQ, phase, as));
      if (sample) FB_CUDA(cudaEventRecord(evA1, C.stream));
      if (!as.publish_seq) FB_TRY(allreduce(C, red + RED_PQ, 1));
      if (recompute) {
        cg_update_xr_kernel<256><<<gvec, BLOCK, 0, C.stream>>>(k, n3, C.P.p, C.Q.p, C.X.p, nullptr, red, C.ctl.p, C.partial.p, C.counter.p);
        FB_LAUNCH_CHECK();
        FB_TRY(exchange(C, C.X.p, false));
        FB_TRY(launch_spmv(C, 1, C.X.p, C.B.p, C.R.p, red + 2 * parn, phase));
        FB_TRY(allreduce(C, red + 2 * parn, 1));
        FB_TRY(z_and_rho(parn));
      } else if (fused) {
        cg_fused_update_kernel<256><<<gnode, BLOCK, 0, C.stream>>>(k, N, C.P.p, C.Q.p, C.X.p, C.R.p, C.Z.p, C.alu.p, C.alu_ld, red, C.ctl.p, C.partial.p, C.counter.p, T.view, af);
        FB_LAUNCH_CHECK();
        if (af.publish_seq) pending_rrrz = af.publish_seq;
        else FB_TRY(allreduce(C, red + 2 * parn, 2));
      } else {
        cg_update_xr_kernel<256><<<gvec, BLOCK, 0, C.stream>>>(k, n3, C.P.p, C.Q.p, C.X.p, C.R.p, red, C.ctl.p, C.partial.p, C.counter.p);
        FB_LAUNCH_CHECK();
        FB_TRY(allreduce(C, red + 2 * parn, 1));
        FB_TRY(z_and_rho(parn));
      }
      if (sample) FB_CUDA(cudaEventRecord(evF1, C.stream));
    }
    // ---- evaluate the last launched iteration's RESID, then look at the control block
    {
      P2PArgs ac;
      ac.consume_seq = pending_rrrz; pending_rrrz = 0;   // the check kernel leaves the sums in red[] for the next batch
      cg_check_kernel<<<1, 32, 0, C.stream>>>(k, red, C.ctl.p, C.hist.p, hist_len, T.view, ac);
      FB_LAUNCH_CHECK();
    }
    FB_CUDA(cudaMemcpyAsync(&h_ctl, C.ctl.p, sizeof h_ctl, cudaMemcpyDeviceToHost, C.stream));
    FB_CUDA(cudaStreamSynchronize(C.stream));
    if (time_ax) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, evA0, evA1) == cudaSuccess) { time_Ax += ms * 1e-3; ax_samples++; }
      if (cudaEventElapsedTime(&ms, evP0, evA0) == cudaSuccess) prof_p += ms * 1e-3;
      if (cudaEventElapsedTime(&ms, evA1, evF1) == cudaSuccess) prof_f += ms * 1e-3;
      if (cudaEventElapsedTime(&ms, evP0, evF1) == cudaSuccess) prof_it += ms * 1e-3;
    }
    if (p2p) {
      int perr = 0;
      FB_CUDA(cudaMemcpy(&perr, T.err.p, sizeof perr, cudaMemcpyDeviceToHost));
      if (perr) return fail(FISTR_B200_ERR_COMM, "solve_33: peer-memory wait timed out (code %d) at iteration <= %d", perr, k);
    }
    if (h_ctl.phase == 0) {
      if (k > MAXIT) { error = HECMW_SOLVER_ERROR_NOCONV_MAXIT_; ITER = MAXIT + 1; done = true; }
      continue;
    }
    const int j = h_ctl.event_iter;
    if (h_ctl.phase == 1) {
      RESID = h_ctl.event_resid;
      if (j % 50 == 0) { ITER = j; done = true; break; }
      // recompute R to make sure it is really converged, CG.f90:261-265
      FB_TRY(exchange(C, C.X.p, false));
      FB_TRY(launch_spmv(C, 1, C.X.p, C.B.p, C.R.p, red + RED_TMP, nullptr));
      FB_TRY(allreduce(C, red + RED_TMP, 1));
      double rr;
      FB_CUDA(cudaMemcpyAsync(&rr, red + RED_TMP, sizeof rr, cudaMemcpyDeviceToHost, C.stream));
      FB_CUDA(cudaStreamSynchronize(C.stream));
      RESID = std::sqrt(rr / BNRM2);
      if (RESID <= TOL) { ITER = j; done = true; break; }
      if (j == MAXIT) { error = HECMW_SOLVER_ERROR_NOCONV_MAXIT_; ITER = MAXIT + 1; done = true; break; }
      // not converged after all: resume at iteration j+1 with the recomputed residual
      const int parn = (j + 1) & 1;
      FB_CUDA(cudaMemcpyAsync(red + 2 * parn, red + RED_TMP, sizeof(double), cudaMemcpyDeviceToDevice, C.stream));
      h_ctl.phase = 0;
      FB_CUDA(cudaMemcpyAsync(&C.ctl.p->phase, &h_ctl.phase, sizeof(int), cudaMemcpyHostToDevice, C.stream));
      FB_TRY(z_and_rho(parn));
      k = j + 1;
      // RESID history of iteration j keeps the recurrence value, as the reference prints it
      continue;
    }
    if (h_ctl.phase == 2) { ITER = j; done = true; break; }   // RHO == 0, CG.f90:170-172
    if (h_ctl.phase == 3002) { error = HECMW_SOLVER_ERROR_DIVERGE_MAT_; ITER = j; done = true; break; }
    if (h_ctl.phase == 3003) { error = HECMW_SOLVER_ERROR_DIVERGE_PC_; ITER = j; done = true; break; }
    return fail(FISTR_B200_ERR_STATE, "solve_33: unknown phase %d", h_ctl.phase);
  }
  // history (ITERLOG)
  {
    int nvalid = (ITER > MAXIT) ? MAXIT : ITER;
    if (h_ctl.phase == 2 || h_ctl.phase == 3002 || h_ctl.phase == 3003) nvalid = ITER - 1;   // left before RESID of ITER
    const int nh = std::max(0, std::min(nvalid, hist_len));
    C.last_hist.resize(nh);
    if (nh > 0) FB_CUDA(cudaMemcpy(C.last_hist.data(), C.hist.p, sizeof(double) * nh, cudaMemcpyDeviceToHost));
    if (error == 0 && nh > 0 && h_ctl.phase != 1) RESID = C.last_hist[nh - 1];
    if (error == HECMW_SOLVER_ERROR_NOCONV_MAXIT_ && nh > 0) RESID = C.last_hist[nh - 1];
    if (ITERlog == 1 && C.rank == 0)
      for (int i = 0; i < nh; i++) printf("%7d%16.6E\n", i + 1, C.last_hist[i]);
  }
  // --- hecmw_update_m_R (CG.f90:280)
  FB_TRY(exchange(C, C.X.p, false));
  FB_CUDA(cudaStreamSynchronize(C.stream));
  const double t_sol_end = now_s();
  // --- back in hecmw_solve_iterative: flags, final true residual (Iterative.f90:145-170)
  if (error == HECMW_SOLVER_ERROR_DIVERGE_PC_ || error == HECMW_SOLVER_ERROR_DIVERGE_MAT_) IA(82) = 1;
  FB_TRY(launch_spmv(C, 1, C.X.p, C.B.p, C.WK.p, red + RED_TMP, nullptr));
  FB_TRY(allreduce(C, red + RED_TMP, 1));
  double rnorm2;
  FB_CUDA(cudaMemcpyAsync(&rnorm2, red + RED_TMP, sizeof rnorm2, cudaMemcpyDeviceToHost, C.stream));
  if (X) FB_CUDA(cudaMemcpyAsync(X, C.X.p, sizeof(double) * np3, cudaMemcpyDeviceToHost, C.stream));
  FB_CUDA(cudaEventRecord(ev_out, C.stream));
  FB_CUDA(cudaStreamSynchronize(C.stream));
  float ev_ms = 0.f;
  cudaEventElapsedTime(&ev_ms, ev_in, ev_out);
  cudaEventDestroy(ev_in); cudaEventDestroy(ev_out);
  const double resid2 = std::sqrt(rnorm2 / (BNRM2 == 0.0 ? 1.0 : BNRM2));
  if (resid2 < TOL) IA(81) = 1;
  C.last_resid2 = resid2;
  if (time_ax) { cudaEventDestroy(evA0); cudaEventDestroy(evA1); cudaEventDestroy(evP0); cudaEventDestroy(evF1); }
  if (ax_samples > 0) {   // per-iteration profile of the sampled iterations: p-update(+halo), SpMV, tail(+all-reduces), whole
    C.last_profile[0] = prof_p / ax_samples; C.last_profile[1] = time_Ax / ax_samples;
    C.last_profile[2] = prof_f / ax_samples; C.last_profile[3] = prof_it / ax_samples;
  }
  const int niter_done = std::max(1, std::min(ITER, MAXIT));
  if (ax_samples > 0) time_Ax = time_Ax / ax_samples * niter_done;
  if (C.rank == 0 && (ITERlog == 1 || TIMElog >= 1)) printf("### Relative residual =%12.5E\n", resid2);
  const double Tset = t_setup_end - t_entry, Tsol = t_sol_end - t_setup_end;
  if (C.rank == 0 && TIMElog >= 1) {
    printf("\n### summary of linear solver\n%10d iterations  %16.6E\n", ITER, RESID);
    printf("    set-up time      : %16.6E\n    solver time      : %16.6E\n    solver/comm time : %16.6E\n", Tset, Tsol, Tcomm);
    printf("    solver/matvec    : %16.6E\n    solver/precond   : %16.6E\n", time_Ax, 0.0);
    if (ITER > 0) printf("    solver/1 iter    : %16.6E\n", Tsol / ITER);
    printf("    work ratio (%%)   : %16.6E\n\n", (Tsol - Tcomm) / (Tsol + 1e-24) * 100.0);
  }
  if (iter_out) *iter_out = ITER;
  if (resid_out) *resid_out = RESID;
  if (time_out) { time_out[0] = Tset; time_out[1] = Tsol; time_out[2] = Tcomm; time_out[3] = time_Ax; time_out[4] = ev_ms * 1e-3; }
  if (error) {
    static const char *msg[] = {"not converged within ceratin iterations", "diverged due to indefinite or negative definite matrix", "diverged due to indefinite preconditioner"};
    if (C.rank == 0) printf("\n  #### HEC-MW-SOLVER-W-%d: \n    %s\n\n", error, msg[error - 3001]);
    return error;
  }
  if (ret_warn) {
    if (C.rank == 0) printf("\n  #### HEC-MW-SOLVER-W-2002: \n    ZERO RHS norm\n\n");
    return ret_warn;
  }
  return 0;
}

// hecmw_part1's RCB (host-only, no GPU): domain of every node, hecmw_partition.c:1570-1660
int fistr_b200_part_rcb(int n_node, const double *node, int n_div, const int *axes, int *domain) {
  const int rc = rcb_partition(n_node, node, n_div, axes, domain);
  if (rc) return fail(FISTR_B200_ERR_ARG, "part_rcb failed (%d)", rc);
  return 0;
}

// hecmw_part1 in-process, second half: the distributed mesh of one rank from the node->domain map (host-only)
int fistr_b200_part_local_count(int n_node, int n_elem_type, const int *elem_type_index, const int *elem_type_item,
                                const int *elem_node_index, const int *elem_node_item, const int *domain, int rank,
                                int nparts, int *sizes /* [8] */) {
  LocalMesh &M = ctx().part;
  const int rc = M.build(n_node, n_elem_type, elem_type_index, elem_type_item, elem_node_index, elem_node_item, domain,
                         rank, nparts, etype_max_node);
  if (rc) return fail(FISTR_B200_ERR_ARG, "part_local: invalid input (%d)", rc);
  sizes[0] = M.n_node; sizes[1] = M.nn_internal; sizes[2] = M.n_elem; sizes[3] = (int)M.elem_type_item.size();
  sizes[4] = (int)M.elem_node_item.size(); sizes[5] = (int)M.neighbor_pe.size();
  sizes[6] = (int)M.import_item.size(); sizes[7] = (int)M.export_item.size();
  return 0;
}

int fistr_b200_part_local_fill(int *global_node, int *node_ID, int *elem_type_index, int *elem_type_item,
                               int *elem_node_index, int *elem_node_item, int *global_elem, int *neighbor_pe,
                               int *import_index, int *import_item, int *export_index, int *export_item) {
  LocalMesh &M = ctx().part;
  if (M.elem_type_index.empty()) return fail(FISTR_B200_ERR_STATE, "part_local_fill before part_local_count");
  auto put = [](int *dst, const std::vector<int> &v) { if (dst && !v.empty()) memcpy(dst, v.data(), sizeof(int) * v.size()); };
  put(global_node, M.global_node); put(node_ID, M.node_ID); put(elem_type_index, M.elem_type_index);
  put(elem_type_item, M.elem_type_item); put(elem_node_index, M.elem_node_index); put(elem_node_item, M.elem_node_item);
  put(global_elem, M.global_elem); put(neighbor_pe, M.neighbor_pe); put(import_index, M.import_index);
  put(import_item, M.import_item); put(export_index, M.export_index); put(export_item, M.export_item);
  M = LocalMesh();
  return 0;
}

// host-only: the SELL(-C-sigma) layout the SpMV copy would use (rowperm: 32*nslice entries, -1 = empty lane; left
// untouched for the natural order; slice_ptr: nslice+1 block-column offsets).  Returns 1 if rows are permuted.
int fistr_b200_debug_sell_layout(int N, const int *indexL, const int *indexU, int sigma, int variant, int *rowperm,
                                 long long *slice_ptr) {
  std::vector<int> rp;
  std::vector<long long> sp;
  sell_layout(N, indexL, indexU, sigma, variant, rp, sp);
  if (!rp.empty() && rowperm) memcpy(rowperm, rp.data(), sizeof(int) * rp.size());
  if (slice_ptr) memcpy(slice_ptr, sp.data(), sizeof(long long) * sp.size());
  return rp.empty() ? 0 : 1;
}

// host-only view of the multicolour ordering (no GPU needed): lets CPU tests compare the ordering
// with the oracle's restatement of hecmw_matrix_ordering_{CM,MC}
int fistr_b200_debug_ssor_order(int N, const int *indexL, const int *itemL, const int *indexU, const int *itemU,
                                int ncolor_in, int *perm, int *ncolor, int *color_index /* N+1 */) {
  std::vector<int> p, ip, ci;
  ssor_multicolor_order(N, indexL, itemL, indexU, itemU, ncolor_in, p, ip, ci);
  memcpy(perm, p.data(), sizeof(int) * (size_t)N);
  *ncolor = (int)ci.size() - 1;
  memcpy(color_index, ci.data(), sizeof(int) * ci.size());
  return 0;
}

// ---- micro-benchmarks for the roofline denominators ------------------------------------------------
__global__ void __launch_bounds__(256) fp64_fma_kernel(double *out, int iters) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-7;
  for (int i = 0; i < iters; i++) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  if (a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 0.123) out[0] = a0;
}
__global__ void __launch_bounds__(256) copy_kernel(size_t n, const double2 *__restrict__ a, double2 *__restrict__ b) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}

int fistr_b200_fp64_peak(double *tflops) {
  FB_REQUIRE_INIT();
  Context &C = ctx();
  DevBuf<double> o;
  if (o.alloc(8)) return FISTR_B200_ERR_CUDA;
  const int iters = 1 << 14, grid = C.sm_count * 8;
  cudaEvent_t e0, e1;
  FB_CUDA(cudaEventCreate(&e0)); FB_CUDA(cudaEventCreate(&e1));
  double best = 0.0;
  for (int r = 0; r < 5; r++) {
    FB_CUDA(cudaEventRecord(e0, C.stream));
    fp64_fma_kernel<<<grid, 256, 0, C.stream>>>(o.p, iters);
    FB_LAUNCH_CHECK();
    FB_CUDA(cudaEventRecord(e1, C.stream));
    FB_CUDA(cudaEventSynchronize(e1));
    float ms;
    FB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    const double fl = 2.0 * 8 * (double)iters * grid * 256;
    best = std::max(best, fl / (ms * 1e-3) * 1e-12);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  o.release();
  *tflops = best;
  return 0;
}

int fistr_b200_hbm_copy_peak(double *gbs) {
  FB_REQUIRE_INIT();
  Context &C = ctx();
  const size_t n = 64u << 20;   // 64 Mi double2 = 1 GiB per buffer
  DevBuf<double2> a, b;
  if (a.alloc(n) || b.alloc(n)) return FISTR_B200_ERR_CUDA;
  FB_CUDA(cudaMemsetAsync(a.p, 0, n * sizeof(double2), C.stream));
  cudaEvent_t e0, e1;
  FB_CUDA(cudaEventCreate(&e0)); FB_CUDA(cudaEventCreate(&e1));
  double best = 0.0;
  for (int r = 0; r < 6; r++) {
    FB_CUDA(cudaEventRecord(e0, C.stream));
    copy_kernel<<<C.sm_count * 16, 256, 0, C.stream>>>(n, a.p, b.p);
    FB_LAUNCH_CHECK();
    FB_CUDA(cudaEventRecord(e1, C.stream));
    FB_CUDA(cudaEventSynchronize(e1));
    float ms;
    FB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    if (r > 0) best = std::max(best, 2.0 * n * sizeof(double2) / (ms * 1e-3) * 1e-9);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  a.release(); b.release();
  *gbs = best;
  return 0;
}

}  // extern "C"
