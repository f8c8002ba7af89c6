// TEST-ONLY host emulation of the warp-cooperative element kernel (assembly.cuh: stiff_kernel) and of
// the host-side orderings.  The per-lane phase functions of element_math.cuh are run in a lane loop,
// with the warp barriers of the CUDA kernel becoming loop boundaries, so the CPU unit tests can check
// the element algebra against the oracle without a GPU.  Built as libfistr_b200_emul.so; the product
// library never links it and no product entry point can reach it.
#include <pthread.h>

#include <cstring>
#include <map>
#include <thread>
#include <vector>

#define FB_EMUL_WARP 1
#define FB_EMUL_KERNELS 1
#include "assembly.cuh"   // element_math.cuh + the assembly / update kernels, run on the host through kernel_emul.hpp
#include "spmv.cuh"       // + SpMV, CG and peer-memory kernels (cg_kernels.cuh, p2p.cuh)

using namespace fb::elem;

namespace fb {
thread_local std::string g_last_error;
int64_t g_launch_count = 0;
}
#include "kernel_emul_impl.hpp"   // lock-step warps (host threads) and blocks of fibers; fb::EmulCfg

// kernel<<<grid, warps*32, smem>>>(args...) for the one-element-per-warp kernels: every warp of every block by 32 host
// threads in lock step, warps one after the other (the warps of these kernels never talk to each other)
template <class K, class... A>
static void emul_launch_warps(unsigned grid, unsigned warps, size_t smem_bytes, K kernel, A... args) {
  fb::EmulSmem sm(smem_bytes);   // ends at a guard page, 0xFF-filled
  for (unsigned b = 0; b < grid; b++) {
    fb::emul_smem() = sm.fresh();
    for (unsigned w = 0; w < warps; w++) {
      fb::emulwarp::FiberBlock B;
      B.grid = grid; B.block = 32; B.blockdim = warps * 32; B.tid0 = w * 32; B.bidx = b;
      B.body = [&]() { kernel(args...); };
      fb::emulwarp::run_fiber_block(B);
    }
  }
  fb::emul_smem() = nullptr;
}

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

static int nn_of(int etype) { return etype == 342 ? 10 : 8; }

// fistr_b200_stiffmatrix_3d on the host: the SAME kernels in the SAME launch sequence as api_assembly.inc
// (slot_build_kernel; stiff_kernel<FORM> per element-type block and formulation; gather map + gather_sum_kernel, or the
// atomic scatter) through kernel_emul.hpp.  vals = [ D | AL | AU ].  Returns the slot-table error flag.
int fbemul_stiffmatrix(int n_node, const double *node, int n_elem, int n_elem_type, const int *elem_type_index,
                       const int *elem_type_item, const int *elem_node_index, const int *elem_node_item,
                       const int *section_ID, int n_sect, const int *sect_form, const double *sect_E, const double *sect_nu,
                       int NP, int NPL, int NPU, const int *indexL, const int *itemL, const int *indexU, const int *itemU,
                       int gather, int icinv, double *vals) {
  using namespace fb;
  (void)n_node;
  std::vector<size_t> slot_base((size_t)n_elem_type + 1, 0);
  for (int it = 0; it < n_elem_type; it++) {
    const int nn = nn_of(elem_type_item[it]);
    slot_base[it + 1] = slot_base[it] + (size_t)(elem_type_index[it + 1] - elem_type_index[it]) * nn * nn;
  }
  const size_t ncontrib = slot_base.back(), nblk = (size_t)NP + NPL + NPU;
  std::vector<unsigned> slot(ncontrib);
  std::vector<double> sectD(3 * (size_t)n_sect), ke(gather ? 9 * ncontrib : 1);
  for (int s = 0; s < n_sect; s++) elastic_constants(sect_E[s], sect_nu[s], sectD[3 * s], sectD[3 * s + 1], sectD[3 * s + 2]);
  int err = 0;
  for (int it = 0; it < n_elem_type; it++) {
    const int nn = nn_of(elem_type_item[it]), e0 = elem_type_index[it], e1 = elem_type_index[it + 1];
    if (e1 > e0)
      emul_launch(3, 64, 7u + it, slot_build_kernel, e0, e1, nn, elem_node_index, elem_node_item, NP, NPL, indexL, itemL, indexU,
                  itemU, slot.data(), slot_base[it], &err);
  }
  std::fill(vals, vals + 9 * nblk, 0.0);
  constexpr int WARPS = 4;
  const size_t smem = sizeof(Scratch) * WARPS;
  double *ke_out = gather ? ke.data() : nullptr;
  for (int it = 0; it < n_elem_type; it++) {
    const int e0 = elem_type_index[it], e1 = elem_type_index[it + 1];
    if (e1 <= e0) continue;
#define FB_ARGS e0, e1, elem_node_index, elem_node_item, node, section_ID, (const double *)sectD.data(), sect_form, \
                (const unsigned *)slot.data(), slot_base[it], vals, ke_out, 0
    if (elem_type_item[it] == 342) emul_launch_warps(2, WARPS, smem, stiff_kernel<FORM_TET10, WARPS, 1>, FB_ARGS);
    else {
      bool has[4] = {false, false, false, false};
      for (int s = 0; s < n_sect; s++) has[sect_form[s]] = true;
      if (has[3]) {
        if (icinv == 1) emul_launch_warps(2, WARPS, smem, stiff_kernel<FORM_IC, WARPS, 1>, FB_ARGS);
        else if (icinv == 2) emul_launch_warps(2, WARPS, smem, stiff_kernel<FORM_IC, WARPS, 2>, FB_ARGS);
        else emul_launch_warps(2, WARPS, smem, stiff_kernel<FORM_IC, WARPS, 0>, FB_ARGS);
      }
      if (has[1]) emul_launch_warps(2, WARPS, smem, stiff_kernel<FORM_FI, WARPS, 1>, FB_ARGS);
      if (has[2]) emul_launch_warps(2, WARPS, smem, stiff_kernel<FORM_BBAR, WARPS, 1>, FB_ARGS);
    }
#undef FB_ARGS
  }
  if (gather) {
    std::vector<unsigned> cursor(nblk + 1, 0u), gptr(nblk + 1, 0u), gsrc(ncontrib);
    emul_launch(3, 64, 21, gather_count_kernel, ncontrib, (const unsigned *)slot.data(), cursor.data());
    for (size_t i = 0; i < nblk; i++) gptr[i + 1] = gptr[i] + cursor[i];
    std::fill(cursor.begin(), cursor.end(), 0u);
    emul_launch(3, 64, 22, gather_fill_kernel, ncontrib, (const unsigned *)slot.data(), (const unsigned *)gptr.data(), cursor.data(), gsrc.data());
    emul_launch(3, 64, 23, gather_sort_kernel, nblk, (const unsigned *)gptr.data(), gsrc.data());
    emul_launch(4, 64, 24, gather_sum_kernel, nblk, (const unsigned *)gptr.data(), (const unsigned *)gsrc.data(), (const double *)ke.data(), vals);
  }
  return err;
}

// fistr_b200_update_3d + fistr_b200_nodal_stress_3d on the host: update_lin_kernel / update_ic_kernel per block and
// formulation, node incidence, QFORCE gather, nodal / elemental strain, stress, von Mises — the launch sequence of
// api_assembly.inc.  sect_alpha / temp / temp0 may be NULL.
int fbemul_update_pipeline(int n_node, const double *node, int n_elem, int n_elem_type, const int *elem_type_index,
                           const int *elem_type_item, const int *elem_node_index, const int *elem_node_item,
                           const int *section_ID, int n_sect, const int *sect_form, const double *sect_E, const double *sect_nu,
                           const double *sect_alpha, const double *utot, const double *temp, const double *temp0,
                           double ref_temp, int icinv, double *gp_strain, double *gp_stress, double *QFORCE, double *STRAIN,
                           double *STRESS, double *MISES, double *ESTRAIN, double *ESTRESS, double *EMISES) {
  using namespace fb;
  const size_t nconn = (size_t)elem_node_index[n_elem];
  std::vector<double> sectD(3 * (size_t)n_sect), eqf(3 * nconn, 0.0);
  for (int s = 0; s < n_sect; s++) elastic_constants(sect_E[s], sect_nu[s], sectD[3 * s], sectD[3 * s + 1], sectD[3 * s + 2]);
  constexpr int WL = 2, WI = 4;
  for (int it = 0; it < n_elem_type; it++) {
    const int e0 = elem_type_index[it], e1 = elem_type_index[it + 1];
    if (e1 <= e0) continue;
#define FB_ARGS e0, e1, elem_node_index, elem_node_item, node, utot, section_ID, (const double *)sectD.data(), sect_alpha, \
                sect_form, temp, temp0, ref_temp, gp_strain, gp_stress, eqf.data()
    if (elem_type_item[it] == 342) emul_launch_warps(2, WL, sizeof(ScratchTL) * WL, update_lin_kernel<FORM_TET10, WL>, FB_ARGS);
    else {
      bool has[4] = {false, false, false, false};
      for (int s = 0; s < n_sect; s++) has[sect_form[s]] = true;
      if (has[3]) {
        if (icinv) emul_launch_warps(2, WI, sizeof(Scratch) * WI, update_ic_kernel<WI, 1>, FB_ARGS);
        else emul_launch_warps(2, WI, sizeof(Scratch) * WI, update_ic_kernel<WI, 0>, FB_ARGS);
      }
      if (has[1]) emul_launch_warps(2, WL, sizeof(ScratchTL) * WL, update_lin_kernel<FORM_FI, WL>, FB_ARGS);
      if (has[2]) emul_launch_warps(2, WL, sizeof(ScratchTL) * WL, update_lin_kernel<FORM_BBAR, WL>, FB_ARGS);
    }
#undef FB_ARGS
  }
  return fbemul_nodal_pipeline(n_node, n_elem, elem_node_index, elem_node_item, gp_strain, gp_stress, eqf.data(), 55u, QFORCE,
                               STRAIN, STRESS, MISES, ESTRAIN, ESTRESS, EMISES);
}

// fistr_b200_nl_stiffmatrix_3d / fistr_b200_nl_update_3d on the host (stiff_tl_kernel / update_tl_kernel, launch
// sequence of api_assembly.inc).  sect_mtype / sect_hyper may be NULL.  gp_strain / gp_stress: [n_elem][8][6] state.
static std::vector<MatlTL> make_matls(int n_sect, const double *sect_E, const double *sect_nu, const int *sect_mtype,
                                      const double *sect_hyper) {
  std::vector<MatlTL> ml((size_t)n_sect);
  for (int s = 0; s < n_sect; s++) ml[s] = make_matl(sect_mtype ? sect_mtype[s] : 0, sect_E[s], sect_nu[s], sect_hyper ? sect_hyper + 3 * s : nullptr);
  return ml;
}
int fbemul_nl_stiffmatrix(int n_node, const double *node, int n_elem, int n_elem_type, const int *elem_type_index,
                          const int *elem_type_item, const int *elem_node_index, const int *elem_node_item,
                          const int *section_ID, int n_sect, const int *sect_form, const double *sect_E, const double *sect_nu,
                          const int *sect_mtype, const double *sect_hyper, const double *u, const double *gp_strain,
                          const double *gp_stress, int NP, int NPL, int NPU, const int *indexL, const int *itemL,
                          const int *indexU, const int *itemU, int gather, double *vals) {
  using namespace fb;
  (void)n_node; (void)n_elem;
  std::vector<size_t> slot_base((size_t)n_elem_type + 1, 0);
  for (int it = 0; it < n_elem_type; it++) {
    const int nn = nn_of(elem_type_item[it]);
    slot_base[it + 1] = slot_base[it] + (size_t)(elem_type_index[it + 1] - elem_type_index[it]) * nn * nn;
  }
  const size_t ncontrib = slot_base.back(), nblk = (size_t)NP + NPL + NPU;
  std::vector<unsigned> slot(ncontrib);
  std::vector<double> ke(gather ? 9 * ncontrib : 1);
  const std::vector<MatlTL> ml = make_matls(n_sect, sect_E, sect_nu, sect_mtype, sect_hyper);
  int err = 0;
  for (int it = 0; it < n_elem_type; it++) {
    const int nn = nn_of(elem_type_item[it]), e0 = elem_type_index[it], e1 = elem_type_index[it + 1];
    if (e1 > e0)
      emul_launch(3, 64, 7u + it, slot_build_kernel, e0, e1, nn, elem_node_index, elem_node_item, NP, NPL, indexL, itemL, indexU,
                  itemU, slot.data(), slot_base[it], &err);
  }
  std::fill(vals, vals + 9 * nblk, 0.0);
  constexpr int WARPS = 2;
  const size_t smem = sizeof(ScratchTL) * WARPS;
  double *ke_out = gather ? ke.data() : nullptr;
  for (int it = 0; it < n_elem_type; it++) {
    const int e0 = elem_type_index[it], e1 = elem_type_index[it + 1];
    if (e1 <= e0) continue;
#define FB_ARGS e0, e1, elem_node_index, elem_node_item, node, u, section_ID, (const MatlTL *)ml.data(), sect_form, gp_strain, \
                gp_stress, (const unsigned *)slot.data(), slot_base[it], vals, ke_out
    if (elem_type_item[it] == 342) emul_launch_warps(2, WARPS, smem, stiff_tl_kernel<FORM_TET10, WARPS>, FB_ARGS);
    else {
      bool has[4] = {false, false, false, false};
      for (int s = 0; s < n_sect; s++) has[sect_form[s]] = true;
      if (has[1]) emul_launch_warps(2, WARPS, smem, stiff_tl_kernel<FORM_FI, WARPS>, FB_ARGS);
      if (has[2]) emul_launch_warps(2, WARPS, smem, stiff_tl_kernel<FORM_BBAR, WARPS>, FB_ARGS);
    }
#undef FB_ARGS
  }
  if (gather) {
    std::vector<unsigned> cursor(nblk + 1, 0u), gptr(nblk + 1, 0u), gsrc(ncontrib);
    emul_launch(3, 64, 21, gather_count_kernel, ncontrib, (const unsigned *)slot.data(), cursor.data());
    for (size_t i = 0; i < nblk; i++) gptr[i + 1] = gptr[i] + cursor[i];
    std::fill(cursor.begin(), cursor.end(), 0u);
    emul_launch(3, 64, 22, gather_fill_kernel, ncontrib, (const unsigned *)slot.data(), (const unsigned *)gptr.data(), cursor.data(), gsrc.data());
    emul_launch(3, 64, 23, gather_sort_kernel, nblk, (const unsigned *)gptr.data(), gsrc.data());
    emul_launch(4, 64, 24, gather_sum_kernel, nblk, (const unsigned *)gptr.data(), (const unsigned *)gsrc.data(), (const double *)ke.data(), vals);
  }
  return err;
}
int fbemul_nl_update(int n_node, const double *node, int n_elem, int n_elem_type, const int *elem_type_index,
                     const int *elem_type_item, const int *elem_node_index, const int *elem_node_item, const int *section_ID,
                     int n_sect, const int *sect_form, const double *sect_E, const double *sect_nu, const int *sect_mtype,
                     const double *sect_hyper, const double *utot, double *gp_strain, double *gp_stress, double *QFORCE) {
  using namespace fb;
  (void)n_elem;
  const std::vector<MatlTL> ml = make_matls(n_sect, sect_E, sect_nu, sect_mtype, sect_hyper);
  std::fill(QFORCE, QFORCE + 3 * (size_t)n_node, 0.0);
  constexpr int WARPS = 2;
  const size_t smem = sizeof(ScratchTL) * WARPS;
  for (int it = 0; it < n_elem_type; it++) {
    const int e0 = elem_type_index[it], e1 = elem_type_index[it + 1];
    if (e1 <= e0) continue;
#define FB_ARGS e0, e1, elem_node_index, elem_node_item, node, utot, section_ID, (const MatlTL *)ml.data(), sect_form, gp_strain, \
                gp_stress, QFORCE
    if (elem_type_item[it] == 342) emul_launch_warps(2, WARPS, smem, update_tl_kernel<FORM_TET10, WARPS>, FB_ARGS);
    else {
      bool has[4] = {false, false, false, false};
      for (int s = 0; s < n_sect; s++) has[sect_form[s]] = true;
      if (has[1]) emul_launch_warps(2, WARPS, smem, update_tl_kernel<FORM_FI, WARPS>, FB_ARGS);
      if (has[2]) emul_launch_warps(2, WARPS, smem, update_tl_kernel<FORM_BBAR, WARPS>, FB_ARGS);
    }
#undef FB_ARGS
  }
  return 0;
}

// fistr_b200_mat_ass_bc on the host: bc_kernel<0>, bc_kernel<1>, bc_set_rhs_kernel as api_assembly.inc launches them
int fbemul_mat_ass_bc(int nbc, const int *inode, const int *idof, const double *rhs, int NP, int NPL, const int *indexL,
                      const int *itemL, const int *indexU, const int *itemU, double *vals, double *B) {
  using namespace fb;
  std::vector<int> nodes, dofs, ptr(1, 0);
  std::vector<double> v, vlast;      // a DOF named twice: eliminate with its first value, set B(c) to its last (api_assembly.inc)
  std::map<long long, int> seen;
  for (int c = 0; c < nbc; c++) {
    if (idof[c] > 3) continue;
    const int n = inode[c];
    const long long key = 3LL * n + idof[c];
    auto it = seen.find(key);
    if (it != seen.end()) { vlast[it->second] = rhs[c]; continue; }
    seen[key] = (int)nodes.size();
    nodes.push_back(n); dofs.push_back(idof[c]); v.push_back(rhs[c]); vlast.push_back(rhs[c]);
    ptr.push_back(ptr.back() + 1 + (indexL[n] - indexL[n - 1]) + (indexU[n] - indexU[n - 1]));
  }
  const int m = (int)nodes.size();
  if (m == 0) return 0;
  emul_launch(3, 64, 31, bc_kernel<0>, m, (const int *)nodes.data(), (const int *)dofs.data(), (const double *)v.data(),
              (const int *)ptr.data(), NP, NPL, indexL, itemL, indexU, itemU, vals, B);
  emul_launch(3, 64, 32, bc_kernel<1>, m, (const int *)nodes.data(), (const int *)dofs.data(), (const double *)v.data(),
              (const int *)ptr.data(), NP, NPL, indexL, itemL, indexU, itemU, vals, B);
  emul_launch((unsigned)(m + 63) / 64, 64, 33, bc_set_rhs_kernel, m, (const int *)nodes.data(), (const int *)dofs.data(),
              (const double *)vlast.data(), B);
  return 0;
}

// =============================================================================================
// The fused CG iteration of fistr_b200_solve_33 in PEER-MEMORY mode on R emulated GPUs: R host threads, each running
// the product's kernels (sell_fill_kernel, diag_setup_kernel, cg_precond_dot_kernel, cg_update_p_kernel,
// spmv_sell2_kernel<0, BLOCK, PEER>, cg_fused_update_kernel, cg_check_kernel; UNCHANGED, blocks of fibers) in the launch
// order and with the sequence numbering of api_solver.inc, the arenas (tagged partial sums, halo flags, p vectors)
// wired as p2p_setup does.  No NCCL: the three scalars of the set-up are summed on the host; `iters` < 50 so that no
// true-residual iteration (which the product runs over NCCL) falls inside.  Ranks run de-synchronised (random spins
// between launches; break_halo_wait = negative control).  Output per rank: X (3*NP) after `iters` iterations, the residual history, the error word.
// =============================================================================================
struct P2PRank {
  int N, NP, nslice;
  std::vector<long long> sp;
  std::vector<int> rowperm, col, exp_peer, exp_remote;
  std::vector<double> val, alu, X, R, Z, Q, sendbuf, red, partial, hist;
  std::vector<unsigned char> halo_slice, arena;
  std::vector<unsigned> counter;
  fb::CgCtl ctl;
  int err[4];
};

int fbemul_p2p_cg(int R, const int *N, const int *NP, const double *const *D, const double *const *AL, const double *const *AU,
                  const int *const *indexL, const int *const *itemL, const int *const *indexU, const int *const *itemU,
                  const double *const *Bv, const int *n_nbr, const int *const *nbr_pe, const int *const *imp_index,
                  const int *const *imp_item, const int *const *exp_index, const int *const *exp_item, int iters, double tol,
                  int sigma, int peer_kernel, unsigned seed, int break_halo_wait, double *const *X_out, double *const *hist_out, int *err_out) {
  using namespace fb;
  constexpr int BLOCK = 256;
  if (R < 1 || R > kP2PMaxRanks || iters >= 50) return -1;
  constexpr size_t OFF_FLAGS = sizeof(P2PEntry) * P2P_NQ * 2 * kP2PMaxRanks, OFF_P = 4096;
  std::vector<P2PRank> rk(R);
  for (int r = 0; r < R; r++) {
    rk[r].N = N[r]; rk[r].NP = NP[r];
    rk[r].arena.assign(OFF_P + sizeof(double) * (3 * (size_t)NP[r] + 8) + 64, 0);
  }
  auto arena_base = [&](int r) { unsigned char *b = rk[r].arena.data(); return b + (64 - reinterpret_cast<uintptr_t>(b) % 64) % 64; };
  pthread_barrier_t bar;
  pthread_barrier_init(&bar, nullptr, R);
  std::vector<double> gsum(R, 0.0);
  auto allreduce = [&](int me, double v) {   // host stand-in of the set-up all-reduces (rank order, like the device path)
    gsum[me] = v;
    pthread_barrier_wait(&bar);
    double s = 0.0;
    for (int r = 0; r < R; r++) s += gsum[r];
    pthread_barrier_wait(&bar);
    return s;
  };
  std::vector<std::thread> th;
  for (int me = 0; me < R; me++)
    th.emplace_back([&, me]() {
      P2PRank &K = rk[me];
      std::mt19937 rng(seed * 977u + me);
      auto jitter = [&]() { volatile unsigned x = 0; for (unsigned i = 0, n = rng() % 20000; i < n; i++) x += i; if (rng() % 5 == 0) std::this_thread::yield(); };
      const int n = K.N, np = K.NP;
      const size_t n3 = 3 * (size_t)n, np3 = 3 * (size_t)np + 8;
      // ---- private SELL copy (build_sell)
      sell_layout(n, indexL[me], indexU[me], sigma, 2, K.rowperm, K.sp);
      K.nslice = (n + 31) / 32;
      const size_t cols = (size_t)K.sp[K.nslice] * 2;
      K.val.assign(cols * 288 + 2, 0.0); K.col.assign(cols * 32 + 2, 0); K.halo_slice.assign((size_t)K.nslice + 1, 0);
      double *val = K.val.data() + ((reinterpret_cast<uintptr_t>(K.val.data()) % 16) ? 1 : 0);   // 16-byte aligned, like cudaMalloc
      int *col = K.col.data() + ((reinterpret_cast<uintptr_t>(K.col.data()) % 8) ? 1 : 0);
      const int *rp = K.rowperm.empty() ? nullptr : K.rowperm.data();
      if (K.nslice > 0)
        emul_launch_blocks((unsigned)((K.nslice * 32 + 255) / 256), 256, sell_fill_kernel<2>, n, D[me], AL[me], AU[me], indexL[me], itemL[me],
                           indexU[me], itemU[me], (const long long *)K.sp.data(), val, col, K.halo_slice.data(), rp);
      // ---- DIAG factors
      K.alu.assign(9 * (size_t)std::max(n, 1), 0.0);
      if (n > 0) emul_launch((unsigned)((n + 63) / 64), 64, 3u + me, diag_setup_kernel, n, D[me], 1.0, K.alu.data(), (size_t)n);
      // ---- work vectors, control block, peer view (p2p_setup)
      K.X.assign(np3, 0.0); K.R.assign(np3, 0.0); K.Z.assign(np3, 0.0); K.Q.assign(np3, 0.0);
      K.red.assign(RED_COUNT, 0.0); K.partial.assign(65536, 0.0); K.counter.assign(4, 0u); K.hist.assign((size_t)iters + 2, 0.0);
      memset(&K.ctl, 0, sizeof K.ctl); K.ctl.tol = tol;
      memset(K.err, 0, sizeof K.err);
      P2PView pv;
      pv.nranks = R; pv.me = me; pv.err = K.err;
      for (int r = 0; r < R; r++) {
        unsigned char *b = arena_base(r);
        pv.ar[r] = (P2PEntry *)b; pv.hflag[r] = (unsigned long long *)(b + OFF_FLAGS); pv.vecP[r] = (double *)(b + OFF_P);
      }
      double *P = pv.vecP[me];
      P2PNeighbors nb;
      nb.n = n_nbr[me];
      for (int a = 0; a < nb.n; a++) nb.rank[a] = nbr_pe[me][a];
      const int nexp = nb.n ? exp_index[me][nb.n] : 0;
      K.exp_peer.assign((size_t)std::max(nexp, 1), 0); K.exp_remote.assign((size_t)std::max(nexp, 1), 0);
      K.sendbuf.assign(3 * (size_t)std::max(nexp, 1), 0.0);
      for (int a = 0; a < nb.n; a++) {
        const int q = nbr_pe[me][a];
        int b = -1;
        for (int t = 0; t < n_nbr[q]; t++) if (nbr_pe[q][t] == me) b = t;
        for (int e = exp_index[me][a]; e < exp_index[me][a + 1]; e++) {
          K.exp_peer[e] = q;
          K.exp_remote[e] = imp_item[q][imp_index[q][b] + (e - exp_index[me][a])] - 1;
        }
      }
      pthread_barrier_wait(&bar);   // every arena is wired before anybody stores into a peer
      double *red = K.red.data();
      const int gvec = 2, gnode = 2, gsp = 3;
      // ---- b.b, r0 = b - A x0 = b, z1 = Minv r0, rho1 (solve_33 set-up)
      {
        double s = 0.0;
        for (size_t i = 0; i < n3; i++) { K.R[i] = Bv[me][i]; s = std::fma(Bv[me][i], Bv[me][i], s); }
        red[RED_BNRM2] = allreduce(me, s);
      }
      emul_launch_blocks(gnode, BLOCK, cg_precond_dot_kernel<BLOCK>, n, (const double *)K.alu.data(), (size_t)n, (const double *)K.R.data(),
                         K.Z.data(), K.partial.data(), K.counter.data(), red + 3, (const int *)nullptr);
      red[3] = allreduce(me, red[3]);
      // ---- iterations (api_solver.inc: the p2p branch of the launch loop)
      unsigned long long seq_pq = 0, seq_rrrz = 0, halo_seq = 0, pending = 0;
      int *phase = &K.ctl.phase;
      for (int k = 1; k <= iters; k++) {
        P2PArgs ap, as, af;
        ap.consume_seq = pending; pending = 0;
        ap.halo_seq = ++halo_seq; as.halo_seq = break_halo_wait ? 0 : ap.halo_seq;   // negative control: SpMV does not wait for the neighbours
        as.publish_seq = ++seq_pq; af.consume_seq = as.publish_seq; af.publish_seq = ++seq_rrrz;
        jitter();
        emul_launch_blocks(gvec, BLOCK, cg_update_p_kernel<BLOCK>, k, n3, (const double *)K.Z.data(), P, red, &K.ctl, K.hist.data(), iters + 1,
                           nexp, exp_item[me], K.sendbuf.data(), pv, nb, ap, (const int *)K.exp_peer.data(), (const int *)K.exp_remote.data(),
                           K.counter.data() + 1);
        jitter();
        (void)peer_kernel;   // (the PEER = false instantiations carry no peer-memory code since round 2: they cannot run this loop)
        emul_launch_blocks(gsp, BLOCK, spmv_sell2_kernel<0, BLOCK, true>, n, K.nslice, (const long long *)K.sp.data(), (const double *)val,
                           (const int *)col, (const double *)P, (const double *)nullptr, K.Q.data(), K.partial.data(), K.counter.data(),
                           red + RED_PQ, (const int *)phase, pv, nb, as, (const unsigned char *)K.halo_slice.data(), rp);
        jitter();
        emul_launch_blocks(gnode, BLOCK, cg_fused_update_kernel<BLOCK, true>, k, n, (const double *)P, (const double *)K.Q.data(), K.X.data(),
                           K.R.data(), K.Z.data(), (const double *)K.alu.data(), (size_t)n, red, &K.ctl, K.partial.data(), K.counter.data(), pv, af);
        pending = af.publish_seq;
      }
      {
        P2PArgs ac;
        ac.consume_seq = pending;
        emul_launch_blocks(1, 32, cg_check_kernel, iters + 1, red, &K.ctl, K.hist.data(), iters + 1, pv, ac);
      }
      memcpy(X_out[me], K.X.data(), sizeof(double) * 3 * (size_t)np);
      memcpy(hist_out[me], K.hist.data(), sizeof(double) * (size_t)iters);
      err_out[me] = K.err[0] | (K.ctl.phase << 8);
    });
  for (auto &t : th) t.join();
  pthread_barrier_destroy(&bar);
  return 0;
}

// fistr_b200_matvec_33 on one emulated GPU: the private SELL copy (sell_layout + sell_fill_kernel) and the SpMV kernel of
// the chosen variant (0 native, 1 SELL, 2 pair-interleaved SELL; sigma > 1: SELL-C-sigma), mode 0 y = A x / 1 y = b - A x,
// launched as api_internal.inc: build_sell / launch_spmv do.  red_out receives the fused dot product.
int fbemul_spmv(int N, int NP, const double *D, const double *AL, const double *AU, const int *indexL, const int *itemL,
                const int *indexU, const int *itemU, int variant, int sigma, int mode, const double *x, const double *b,
                double *y, double *red_out) {
  using namespace fb;
  (void)NP;
  constexpr int BLOCK = 256;
  std::vector<double> partial(65536, 0.0);
  std::vector<unsigned> counter(4, 0u);
  double red = 0.0;
  if (variant == 0) {
    if (mode == 0) emul_launch_blocks(3, BLOCK, spmv_native_kernel<0, BLOCK>, N, D, AL, AU, indexL, itemL, indexU, itemU, x, b, y, partial.data(), counter.data(), &red, (const int *)nullptr);
    else emul_launch_blocks(3, BLOCK, spmv_native_kernel<1, BLOCK>, N, D, AL, AU, indexL, itemL, indexU, itemU, x, b, y, partial.data(), counter.data(), &red, (const int *)nullptr);
  } else {
    std::vector<int> rowperm;
    std::vector<long long> sp;
    sell_layout(N, indexL, indexU, sigma, variant, rowperm, sp);
    const int nslice = (N + 31) / 32;
    const size_t cols = (size_t)sp[nslice] * (variant == 2 ? 2 : 1);
    std::vector<double> valv(cols * 288 + 2, 0.0);
    std::vector<int> colv(cols * 32 + 2, 0);
    std::vector<unsigned char> halo((size_t)nslice + 1, 0);
    double *val = valv.data() + ((reinterpret_cast<uintptr_t>(valv.data()) % 16) ? 1 : 0);
    int *col = colv.data() + ((reinterpret_cast<uintptr_t>(colv.data()) % 8) ? 1 : 0);
    const int *rp = rowperm.empty() ? nullptr : rowperm.data();
    const unsigned fb_blocks = (unsigned)((nslice * 32 + 255) / 256);
    if (nslice > 0) {
      if (variant == 2) emul_launch_blocks(fb_blocks, 256, sell_fill_kernel<2>, N, D, AL, AU, indexL, itemL, indexU, itemU, (const long long *)sp.data(), val, col, halo.data(), rp);
      else emul_launch_blocks(fb_blocks, 256, sell_fill_kernel<1>, N, D, AL, AU, indexL, itemL, indexU, itemU, (const long long *)sp.data(), val, col, halo.data(), rp);
    }
    P2PView pv; P2PNeighbors nb; P2PArgs pa;
#define FB_S1 N, nslice, (const long long *)sp.data(), (const double *)val, (const int *)col, x, b, y, partial.data(), counter.data(), &red, (const int *)nullptr, rp
#define FB_S2 N, nslice, (const long long *)sp.data(), (const double *)val, (const int *)col, x, b, y, partial.data(), counter.data(), &red, (const int *)nullptr, pv, nb, pa, (const unsigned char *)halo.data(), rp
    if (variant == 1) {
      if (mode == 0) emul_launch_blocks(3, BLOCK, spmv_sell1_kernel<0, BLOCK>, FB_S1);
      else emul_launch_blocks(3, BLOCK, spmv_sell1_kernel<1, BLOCK>, FB_S1);
    } else {
      if (mode == 0) emul_launch_blocks(3, BLOCK, spmv_sell2_kernel<0, BLOCK, false>, FB_S2);
      else emul_launch_blocks(3, BLOCK, spmv_sell2_kernel<1, BLOCK, false>, FB_S2);
    }
#undef FB_S1
#undef FB_S2
  }
  if (red_out) *red_out = red;
  return 0;
}

// number of misaligned double2 accesses seen so far by the emulated kernels (must stay 0), and a negative control
long fbemul_misaligned_count(void) { return emul_misaligned_count(); }
long fbemul_misaligned_probe(void) {
  alignas(16) double buf[6] = {1, 2, 3, 4, 5, 6};
  const long before = emul_misaligned_count();
  double2 *good = reinterpret_cast<double2 *>(buf), *bad = reinterpret_cast<double2 *>(buf + 1);
  good[1] = good[0];           // aligned: not counted
  volatile double sink = good[1].x;
  (void)sink;
  bad[1] = bad[0];             // 8-byte aligned only: counted (source and destination)
  const long seen = emul_misaligned_count() - before;
  emul_misaligned_count() = before;
  return seen;
}

// Negative / positive control of the fiber-order race detector (kernel_emul_impl.hpp: fbemul_set_fiber_order): a warp
// whose lanes hand a value to their neighbour through shared storage, with (sync = 1) and without (sync = 0) the
// __syncwarp that the exchange needs.  out[lane] = what lane read from lane + 1.  With the barrier the result cannot
// depend on the order in which the emulator runs the lanes; without it, it does.
}
namespace fb {
static void race_probe_kernel(int sync, double *box, double *out) {
  const int lane = threadIdx.x & 31;
  box[lane] = 100.0 + lane;
  if (sync) __syncwarp();
  out[lane] = box[(lane + 1) & 31];
}
}
extern "C" {
int fbemul_race_probe(int sync, double *out) {
  double box[32];
  for (int i = 0; i < 32; i++) box[i] = -1.0;
  emul_launch_blocks(1, 32, fb::race_probe_kernel, sync, (double *)box, out);
  return 0;
}

// Control of the divergent-collective detector (kernel_emul_impl.hpp): mode 0 - every lane shuffles at the same call site
// (fine); mode 1 - the lower and the upper half of the warp execute DIFFERENT __shfl_sync instructions with a full mask;
// mode 2 - half of the lanes wait at a __syncwarp while the others shuffle.  Modes 1 and 2 are undefined behaviour on
// hardware; an emulator that only counts arrivals would run them through.  The detector aborts the process.
}
namespace fb {
static void divergent_probe_kernel(int mode, double *out) {
  const int lane = threadIdx.x & 31;
  double v = lane;
  if (mode == 0) v = __shfl_sync(0xffffffffu, v, (lane + 1) & 31);
  else if (lane < 16) v = __shfl_sync(0xffffffffu, v, (lane + 1) & 31);
  else if (mode == 1) v = __shfl_sync(0xffffffffu, v, (lane + 2) & 31);
  else __syncwarp();
  out[lane] = v;
}
}
extern "C" {
int fbemul_divergent_probe(int mode, double *out) {
  emul_launch_blocks(1, 32, fb::divergent_probe_kernel, mode, out);
  return 0;
}

// the device self-test of the three condensation kernels (assembly.cuh: ic_selftest_kernel):
// out[0] max |K_reg - K_smem|, out[1] max entry, out[2] max |K_ldlt - K_smem|
int fbemul_ic_selftest(double *out) {
  emul_launch_warps(1, 1, sizeof(Scratch), fb::ic_selftest_kernel, out);
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
// IC element condensed through LDL^T solves (ICINV == 2; falls back to the shared-memory calInverse by itself when
// K(alpha,alpha) is not positive definite)
int fbemul_element_stiffness_ic_ldlt(const double *ecoord, double E, double nu, double *K) {
  run_element<FORM_IC, 2>(ecoord, E, nu, K);
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
