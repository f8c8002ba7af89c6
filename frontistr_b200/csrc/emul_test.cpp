// TEST-ONLY host emulation of the warp-cooperative element kernel (assembly.cuh: stiff_kernel) and of
// the host-side orderings.  The per-lane phase functions of element_math.cuh are run in a lane loop,
// with the warp barriers of the CUDA kernel becoming loop boundaries, so the CPU unit tests can check
// the element algebra against the oracle without a GPU.  Built as libfistr_b200_emul.so; the product
// library never links it and no product entry point can reach it.
#include <pthread.h>

#include <cstring>
#include <thread>
#include <vector>

#define FB_EMUL_WARP 1
#define FB_EMUL_KERNELS 1
#include "assembly.cuh"   // element_math.cuh + the grid-stride kernels, run on the host through kernel_emul.hpp

using namespace fb::elem;

// ---- lock-step warp emulation: 32 host threads, every collective is a barrier-separated exchange -----
namespace fb { namespace emulwarp {
static pthread_barrier_t g_bar;
static thread_local int t_lane = 0;
static double g_slot_d[32];
static int g_slot_i[32];
double shfl(double v, int src) {
  g_slot_d[t_lane] = v;
  pthread_barrier_wait(&g_bar);
  const double r = g_slot_d[src & 31];
  pthread_barrier_wait(&g_bar);
  return r;
}
int shfl(int v, int src) {
  g_slot_i[t_lane] = v;
  pthread_barrier_wait(&g_bar);
  const int r = g_slot_i[src & 31];
  pthread_barrier_wait(&g_bar);
  return r;
}
void syncwarp() { pthread_barrier_wait(&g_bar); }
template <class F>
static void run_warp(F f) {   // f(lane) on 32 threads in lock step
  pthread_barrier_init(&g_bar, nullptr, 32);
  std::vector<std::thread> th;
  for (int lane = 0; lane < 32; lane++) th.emplace_back([lane, &f]() { t_lane = lane; f(lane); });
  for (auto &t : th) t.join();
  pthread_barrier_destroy(&g_bar);
}
} }

// calInverse(9) exactly as the CUDA warp executes it (same source: element_math.cuh ic_inverse_warp)
static void emul_inverse_warp(Scratch &S) {
  fb::emulwarp::run_warp([&S](int lane) { ic_inverse_warp(lane, S); });
}

template <int FORM, int ICINV = 1>
static void run_element(const double *ecoord /* (nn,3) */, double E, double nu, double *K /* (3nn,3nn) row-major */) {
  constexpr int NN = (FORM == FORM_TET10) ? 10 : 8;
  constexpr int NB = NN * NN;
  constexpr int ROUNDS = (NB + 31) / 32;
  static Scratch S;
  std::memset(&S, 0, sizeof S);
  double D11, D12, D44;
  elastic_constants(E, nu, D11, D12, D44);
  // what stiff_kernel does before the body: lanes < nn fetch the coordinates, then the warp synchronises
  for (int lane = 0; lane < NN; lane++)
    for (int d = 0; d < 3; d++) S.xc[3 * lane + d] = ecoord[3 * lane + d];
  fb::emulwarp::run_warp([&](int lane) {
    double kb[ROUNDS][9];
    element_stiffness_warp<FORM, ICINV>(lane, S, D11, D12, D44, kb);
    for (int r = 0; r < ROUNDS; r++) {
      const int t = r * 32 + lane;
      if (t >= NB) continue;
      const int i = t / NN, j = t % NN;
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) K[(3 * i + a) * (3 * NN) + 3 * j + b] = kb[r][3 * a + b];
    }
  });
}

static MatlTL make_matl(int mtype, double E, double nu, const double *hyper) {
  MatlTL M;
  M.mtype = mtype;
  elastic_constants(E, nu, M.D11, M.D12, M.D44);
  for (int k = 0; k < 3; k++) M.c[k] = hyper ? hyper[k] : 0.0;
  return M;
}

template <int FORM>
static void run_element_tl(const double *ecoord, const double *u, const double *gp_strain, const double *gp_stress,
                           const MatlTL &M, double *K) {
  constexpr int NN = (FORM == FORM_TET10) ? 10 : 8;
  constexpr int NB = NN * NN;
  constexpr int ROUNDS = (NB + 31) / 32;
  static ScratchTL S;
  std::memset(&S, 0, sizeof S);
  for (int k = 0; k < 3 * NN; k++) { S.xc[k] = ecoord[k]; S.uc[k] = u[k]; }
  fb::emulwarp::run_warp([&](int lane) {
    double kb[ROUNDS][9];
    element_stiffness_tl_warp<FORM>(lane, S, M, gp_strain, gp_stress, kb);
    for (int r = 0; r < ROUNDS; r++) {
      const int t = r * 32 + lane;
      if (t >= NB) continue;
      const int i = t / NN, j = t % NN;
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) K[(3 * i + a) * (3 * NN) + 3 * j + b] = kb[r][3 * a + b];
    }
  });
}

template <int FORM>
static void run_update_tl(const double *ecoord, const double *utot, const MatlTL &M, double *qf, double *strain,
                          double *stress) {
  constexpr int NN = (FORM == FORM_TET10) ? 10 : 8;
  static ScratchTL S;
  std::memset(&S, 0, sizeof S);
  for (int k = 0; k < 3 * NN; k++) { S.xc[k] = ecoord[k]; S.uc[k] = utot[k]; }
  fb::emulwarp::run_warp([&](int lane) {
    const double q = element_update_tl_warp<FORM>(lane, S, M, strain, stress);
    if (lane < 3 * NN) qf[lane] = q;
  });
}

template <int FORM>
static void run_update_lin(const double *ecoord, const double *utot, double E, double nu, const double *TT, const double *T0,
                           double alpha, double ref_temp, double *qf, double *strain, double *stress) {
  constexpr int NN = (FORM == FORM_TET10) ? 10 : 8;
  static ScratchTL S;
  std::memset(&S, 0, sizeof S);
  double D11, D12, D44;
  elastic_constants(E, nu, D11, D12, D44);
  for (int k = 0; k < 3 * NN; k++) { S.xc[k] = ecoord[k]; S.uc[k] = utot[k]; }
  if (TT) for (int k = 0; k < NN; k++) { S.tc[k] = TT[k]; S.t0c[k] = T0[k]; }
  fb::emulwarp::run_warp([&](int lane) {
    const double q = element_update_lin_warp<FORM>(lane, S, D11, D12, D44, alpha, ref_temp, TT != nullptr, strain, stress);
    if (lane < 3 * NN) qf[lane] = q;
  });
}
template <int ICINV>
static void run_update_ic(const double *ecoord, const double *utot, double E, double nu, const double *TT, const double *T0,
                          double alpha, double ref_temp, double *qf, double *strain, double *stress) {
  static Scratch S;
  std::memset(&S, 0, sizeof S);
  double D11, D12, D44;
  elastic_constants(E, nu, D11, D12, D44);
  for (int k = 0; k < 24; k++) { S.xc[k] = ecoord[k]; S.uc[k] = utot[k]; }
  if (TT) for (int k = 0; k < 8; k++) { S.tc[k] = TT[k]; S.t0c[k] = T0[k]; }
  fb::emulwarp::run_warp([&](int lane) {
    const double q = element_update_ic_warp<ICINV>(lane, S, D11, D12, D44, alpha, ref_temp, TT != nullptr, strain, stress);
    if (lane < 24) qf[lane] = q;
  });
}

extern "C" {

// linear strain / stress update of one element exactly as the CUDA warp executes it (form 3 = IC with the
// register/shuffle inverse, 30 = IC with the shared-memory inverse); TT / T0 may be NULL
int fbemul_element_update_lin(int form, const double *ecoord, const double *utot, double E, double nu, const double *TT,
                              const double *T0, double alpha, double ref_temp, double *qf, double *strain, double *stress) {
  switch (form) {
    case FORM_FI: run_update_lin<FORM_FI>(ecoord, utot, E, nu, TT, T0, alpha, ref_temp, qf, strain, stress); return 0;
    case FORM_BBAR: run_update_lin<FORM_BBAR>(ecoord, utot, E, nu, TT, T0, alpha, ref_temp, qf, strain, stress); return 0;
    case FORM_TET10: run_update_lin<FORM_TET10>(ecoord, utot, E, nu, TT, T0, alpha, ref_temp, qf, strain, stress); return 0;
    case FORM_IC: run_update_ic<1>(ecoord, utot, E, nu, TT, T0, alpha, ref_temp, qf, strain, stress); return 0;
    case 30: run_update_ic<0>(ecoord, utot, E, nu, TT, T0, alpha, ref_temp, qf, strain, stress); return 0;
  }
  return -1;
}
// fstr_NodalStress3D building blocks of the CUDA kernels: extrapolation matrices and one extrapolated value
int fbemul_nodal_extrap(double *inv8, double *inv4) {
  NodalExtrap X;
  nodal_extrap_build(X);
  std::memcpy(inv8, X.inv8, sizeof X.inv8); std::memcpy(inv4, X.inv4, sizeof X.inv4);
  return 0;
}
double fbemul_nodal_extrapolate(int nn, int j, const double *src, int c) {
  NodalExtrap X;
  nodal_extrap_build(X);
  return nodal_extrapolate(X, nn, j, src, c);
}
double fbemul_mises(const double *s) { return mises_of(s); }

// The post-processing kernels of update_3d / nodal_stress_3d (assembly.cuh) run UNCHANGED on the host, threads in a
// scrambled order: node incidence (node_slot / pos_elem / gather_count / scan / gather_fill / gather_sort), QFORCE
// gather, nodal extrapolation + averaging, elemental averages, von Mises.  Same launch sequence as api_assembly.inc.
int fbemul_nodal_pipeline(int n_node, int n_elem, const int *elem_node_index, const int *elem_node_item,
                          const double *gp_strain, const double *gp_stress, const double *eqf, unsigned seed, double *QFORCE,
                          double *STRAIN, double *STRESS, double *MISES, double *ESTRAIN, double *ESTRESS, double *EMISES) {
  using namespace fb;
  const size_t nconn = (size_t)elem_node_index[n_elem];
  std::vector<unsigned> nslot(nconn), cursor((size_t)n_node + 1, 0u), nptr((size_t)n_node + 1, 0u), npos(nconn);
  std::vector<int> pos_elem(nconn);
  emul_launch(3, 64, seed, node_slot_kernel, nconn, elem_node_item, nslot.data());
  emul_launch(2, 64, seed + 1, pos_elem_kernel, n_elem, elem_node_index, pos_elem.data());
  emul_launch(3, 64, seed + 2, gather_count_kernel, nconn, (const unsigned *)nslot.data(), cursor.data());
  for (int i = 0; i < n_node; i++) nptr[i + 1] = nptr[i] + cursor[i];   // cub::DeviceScan::ExclusiveSum
  std::fill(cursor.begin(), cursor.end(), 0u);
  emul_launch(3, 64, seed + 3, gather_fill_kernel, nconn, (const unsigned *)nslot.data(), (const unsigned *)nptr.data(), cursor.data(), npos.data());
  emul_launch(2, 64, seed + 4, gather_sort_kernel, (size_t)n_node, (const unsigned *)nptr.data(), npos.data());
  emul_launch(3, 64, seed + 5, node_gather3_kernel, n_node, (const unsigned *)nptr.data(), (const unsigned *)npos.data(), eqf, QFORCE);
  NodalExtrap X;
  nodal_extrap_build(X);
  emul_launch(5, 64, seed + 6, nodal_stress_kernel, n_node, (const unsigned *)nptr.data(), (const unsigned *)npos.data(),
              (const int *)pos_elem.data(), elem_node_index, gp_strain, gp_stress, X, STRAIN, STRESS);
  emul_launch(5, 64, seed + 7, elem_stress_kernel, n_elem, elem_node_index, gp_strain, gp_stress, ESTRAIN, ESTRESS);
  emul_launch(2, 64, seed + 8, mises_kernel, (size_t)n_node, (const double *)STRESS, MISES);
  emul_launch(2, 64, seed + 9, mises_kernel, (size_t)n_elem, (const double *)ESTRESS, EMISES);
  return 0;
}

// total-Lagrangian tangent stiffness / update of one element exactly as the CUDA warp executes them
int fbemul_element_stiffness_tl(int form, const double *ecoord, const double *u, const double *gp_strain,
                                const double *gp_stress, int mtype, double E, double nu, const double *hyper, double *K) {
  const MatlTL M = make_matl(mtype, E, nu, hyper);
  switch (form) {
    case FORM_FI: run_element_tl<FORM_FI>(ecoord, u, gp_strain, gp_stress, M, K); return 0;
    case FORM_BBAR: run_element_tl<FORM_BBAR>(ecoord, u, gp_strain, gp_stress, M, K); return 0;
    case FORM_TET10: run_element_tl<FORM_TET10>(ecoord, u, gp_strain, gp_stress, M, K); return 0;
  }
  return -1;
}
int fbemul_element_update_tl(int form, const double *ecoord, const double *utot, int mtype, double E, double nu,
                             const double *hyper, double *qf, double *strain, double *stress) {
  const MatlTL M = make_matl(mtype, E, nu, hyper);
  switch (form) {
    case FORM_FI: run_update_tl<FORM_FI>(ecoord, utot, M, qf, strain, stress); return 0;
    case FORM_BBAR: run_update_tl<FORM_BBAR>(ecoord, utot, M, qf, strain, stress); return 0;
    case FORM_TET10: run_update_tl<FORM_TET10>(ecoord, utot, M, qf, strain, stress); return 0;
  }
  return -1;
}

int fbemul_element_stiffness(int form, const double *ecoord, double E, double nu, double *K) {
  switch (form) {
    case FORM_FI: run_element<FORM_FI>(ecoord, E, nu, K); return 0;
    case FORM_BBAR: run_element<FORM_BBAR>(ecoord, E, nu, K); return 0;
    case FORM_IC: run_element<FORM_IC>(ecoord, E, nu, K); return 0;
    case FORM_TET10: run_element<FORM_TET10>(ecoord, E, nu, K); return 0;
  }
  return -1;
}

// IC element through the shared-memory calInverse (the arbiter of the device self-test)
int fbemul_element_stiffness_ic_smem(const double *ecoord, double E, double nu, double *K) {
  run_element<FORM_IC, 0>(ecoord, E, nu, K);
  return 0;
}
int fbemul_inverse9_smem(double *A /* row-major 9x9, in place */) {
  Scratch S;
  std::memset(&S, 0, sizeof S);
  std::memcpy(S.Kaa, A, sizeof(double) * 81);
  fb::emulwarp::run_warp([&S](int lane) { ic_inverse_smem_warp(lane, S); });
  std::memcpy(A, S.Kaa, sizeof(double) * 81);
  return 0;
}

// calInverse(9) through the column-parallel phases alone
int fbemul_inverse9(double *A /* row-major 9x9, in place */) {
  Scratch S;
  std::memset(&S, 0, sizeof S);
  std::memcpy(S.Kaa, A, sizeof(double) * 81);
  emul_inverse_warp(S);
  std::memcpy(A, S.Kaa, sizeof(double) * 81);
  return 0;
}

}  // extern "C"

// =============================================================================================
// TEST-ONLY model of the peer-memory protocol of the fused CG iteration (p2p.cuh + the sequence numbering of
// api_solver.inc): R ranks = R host threads, each rank's "arena" = a block of atomics that every other rank may write.
// One iteration per rank is   update_p: consume (rr,rz)[k-1] -> push halo values to the neighbours -> signal
//                             spmv    : wait for the neighbours' halo flags -> read halo -> publish p.q
//                             tail    : consume p.q -> publish (rr,rz)
// exactly as the kernels do (tagged partial sums in parity-double-buffered slots, one flag word per source rank,
// monotone sequence numbers).  Threads are de-synchronised with random spinning.  Checked invariants: every consume
// sees, for every rank, the payload that belongs to ITS sequence number (no slot is overwritten before all readers
// are done), and every halo read sees the value of the current iteration (no neighbour runs ahead into our halo).
// Lets the N = 8 logic (incl. ranks that are not neighbours of each other) be exercised without 8 GPUs.
// =============================================================================================
#include <atomic>
#include <random>

namespace p2pmodel {
constexpr int MAXR = 8;
struct Entry { std::atomic<double> v0{0}, v1{0}; std::atomic<unsigned long long> tag{0}; };
struct Arena {
  Entry ar[2][2][MAXR];                       // [quantity][parity][source rank]
  std::atomic<unsigned long long> hflag[MAXR];
  std::atomic<double> halo[MAXR];             // one halo value per source rank (stands for the halo rows of p)
};

static int run(int R, const int *nbr /* R x R adjacency */, int iters, unsigned seed, int recompute_every,
               int break_halo_wait /* negative control: do not wait for the neighbours' flags */) {
  std::vector<Arena> arena(R);
  for (auto &a : arena) for (int r = 0; r < MAXR; r++) { a.hflag[r] = 0; a.halo[r] = -1.0; }
  std::atomic<int> errors{0};
  auto payload = [](int q, int rank, unsigned long long seq) { return 1000.0 * q + 10.0 * rank + 1e-3 * (double)seq; };
  std::vector<std::thread> th;
  // the 1-in-`recompute_every` iterations go through the "NCCL" path: modelled as a full barrier that publishes nothing
  pthread_barrier_t bar;
  pthread_barrier_init(&bar, nullptr, R);
  for (int me = 0; me < R; me++)
    th.emplace_back([&, me]() {
      std::mt19937 rng(seed * 131 + me);
      auto jitter = [&]() { volatile int x = 0; for (int i = 0, n = rng() % 400; i < n; i++) x += i; if (rng() % 7 == 0) std::this_thread::yield(); };
      unsigned long long seq_pq = 0, seq_rrrz = 0, halo_seq = 0, pending = 0;
      auto publish = [&](int q, unsigned long long seq) {
        for (int r = 0; r < R; r++) {
          Entry &e = arena[r].ar[q][seq & 1][me];
          e.v0.store(payload(q, me, seq), std::memory_order_relaxed);
          e.v1.store(-payload(q, me, seq), std::memory_order_relaxed);
          e.tag.store(seq, std::memory_order_release);
          jitter();
        }
      };
      auto consume = [&](int q, unsigned long long seq) {
        for (int r = 0; r < R; r++) {
          Entry &e = arena[me].ar[q][seq & 1][r];
          while (e.tag.load(std::memory_order_acquire) != seq) std::this_thread::yield();
          const double v0 = e.v0.load(std::memory_order_relaxed), v1 = e.v1.load(std::memory_order_relaxed);
          jitter();
          if (v0 != payload(q, r, seq) || v1 != -v0 || e.tag.load(std::memory_order_acquire) != seq) errors++;
        }
      };
      for (int k = 1; k <= iters; k++) {
        const bool recompute = recompute_every > 0 && k % recompute_every == 0;
        // ---- update_p(k)
        if (pending) consume(1, pending);
        pending = 0;
        jitter();
        ++halo_seq;
        for (int q = 0; q < R; q++) if (nbr[me * R + q]) arena[q].halo[me].store((double)k, std::memory_order_relaxed);
        for (int q = 0; q < R; q++) if (nbr[me * R + q]) arena[q].hflag[me].store(halo_seq, std::memory_order_release);
        // ---- spmv(k)
        jitter();
        if (!break_halo_wait)
          for (int q = 0; q < R; q++) if (nbr[me * R + q]) while (arena[me].hflag[q].load(std::memory_order_acquire) < halo_seq) std::this_thread::yield();
        for (int q = 0; q < R; q++) if (nbr[me * R + q]) { jitter(); if (arena[me].halo[q].load(std::memory_order_relaxed) != (double)k) errors++; }
        if (recompute) {   // NCCL iteration: collectives = barriers, nothing published
          pthread_barrier_wait(&bar);
          jitter();
          pthread_barrier_wait(&bar);
          continue;
        }
        const unsigned long long spq = ++seq_pq;
        publish(0, spq);
        // ---- tail(k)
        jitter();
        consume(0, spq);
        const unsigned long long srz = ++seq_rrrz;
        publish(1, srz);
        pending = srz;
      }
    });
  for (auto &t : th) t.join();
  pthread_barrier_destroy(&bar);
  return errors.load();
}
}  // namespace p2pmodel

extern "C" int fbemul_p2p_protocol_model(int R, const int *nbr, int iters, unsigned seed, int recompute_every,
                                         int break_halo_wait) {
  if (R < 1 || R > p2pmodel::MAXR) return -1;
  return p2pmodel::run(R, nbr, iters, seed, recompute_every, break_halo_wait);
}
