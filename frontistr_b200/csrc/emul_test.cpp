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
#include "element_math.cuh"

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

template <int FORM>
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
    element_stiffness_warp<FORM>(lane, S, D11, D12, D44, kb);
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

extern "C" {

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
