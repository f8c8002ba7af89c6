// Process-wide state of the fistr_b200 library: one GPU, one live matrix (as the reference's
// preconditioner modules keep module-level SAVE state, precond/33/hecmw_precond_DIAG_33.f90:20-23).
#pragma once
#include <chrono>
#include <map>
#include <vector>

#include "../../include/fistr_b200.h"
#include "assembly.cuh"
#include "bilu.cuh"
#include "cg_kernels.cuh"
#include "cg_zhalo.cuh"
#include "comm.hpp"
#include "common.cuh"
#include "matcon.hpp"
#include "matcon_dev.cuh"
#include "matdump.hpp"
#include "p2p.cuh"
#include "scaling.cuh"
#include "partition.hpp"
#include "partition_metis.hpp"
#include "spmv.cuh"
#include "ssor.cuh"

namespace fb {

enum {
  HECMW_SOLVER_ERROR_INCONS_PC_ = 1001, HECMW_SOLVER_ERROR_ZERO_DIAG_ = 2001, HECMW_SOLVER_ERROR_ZERO_RHS_ = 2002,
  HECMW_SOLVER_ERROR_NOCONV_MAXIT_ = 3001, HECMW_SOLVER_ERROR_DIVERGE_MAT_ = 3002, HECMW_SOLVER_ERROR_DIVERGE_PC_ = 3003
};   // hecmw1/src/solver/init/hecmw_solve_error.f90:9-15

struct Halo {
  int n_neighbor = 0;
  std::vector<int> neighbor_pe, import_index, import_item, export_index, export_item;
  DevBuf<int> d_export_item, d_import_item;
  DevBuf<double> send_p, send_x, recv;
  bool import_contiguous = true;   // import_item == N+1..NP in order -> receive straight into x
  int n_export() const { return export_index.empty() ? 0 : export_index.back(); }
  int n_import() const { return import_index.empty() ? 0 : import_index.back(); }
};

struct Context {
  bool inited = false;
  int device = 0, sm_count = kNumSM;
  cudaStream_t stream = nullptr;
  std::map<std::string, int64_t> opt;

  // ---- matrix structure (hecmwST_matrix index part)
  int N = 0, NP = 0, NPL = 0, NPU = 0;
  bool have_structure = false;
  const int *key_indexL = nullptr, *key_itemL = nullptr;   // identity of the host arrays last declared
  std::vector<int> h_indexL, h_indexU;
  DevBuf<int> indexL, itemL, indexU, itemU;
  Halo halo;

  // ---- values: [D | AL | AU], RHS, solution, CG work vectors
  DevBuf<double> vals, B, X, R, Z, P, Q, WK;
  DevBuf<double> scale;    // SCALING=YES: s = 1 / sqrt(|diag|), 3*NP (scaling.cuh)
  bool values_on_device = false;
  double *dD() { return vals.p; }
  double *dAL() { return vals.p + 9 * (size_t)NP; }
  double *dAU() { return vals.p + 9 * ((size_t)NP + NPL); }

  // ---- solver-private copies
  SellMatrix sell;
  bool sell_valid = false;
  DevBuf<double> alu;      // DIAG: SoA [9][alu_ld]
  size_t alu_ld = 0;
  int pc_kind = 0;         // 0 none, 3 DIAG, 1 SSOR, 10 BILU(0)
  double pc_sigma = 0.0;
  Ssor ssor;
  Bilu bilu;

  // ---- reductions / control
  DevBuf<double> red, partial, hist;
  DevBuf<double> rtab;     // z-halo schedule (cg_zhalo.cuh): every rank's (r.r, r.z) partials, [parity][rank][2]
  DevBuf<unsigned> counter;
  DevBuf<CgCtl> ctl;
  DevBuf<int> flag;
  std::vector<double> last_hist;
  double last_resid2 = 0.0;
  double last_profile[4] = {0, 0, 0, 0};

  // ---- assembly
  DevBuf<unsigned> gptr, gsrc;     // inverse of the slot table: block -> its contributions (ascending)
  DevBuf<double> ke;               // element-matrix scratch of the gather assembly
  DevBuf<double> gp_strain, gp_stress;   // nonlinear path: Gauss-point state [element][8][6]
  DevBuf<elem::MatlTL> m_matl;           // nonlinear path: per-section material records
  DevBuf<double> m_u, m_qf;              // nonlinear path: nodal displacement in, internal force out
  int nl_n_elem = -1;                    // elements the Gauss-point state is sized for (-1: none yet; 0 is a valid, empty mesh)
  // linear strain / stress recovery (update_3d / nodal_stress_3d): node incidence of the resident mesh, per-position
  // element internal force, section expansion coefficients, nodal temperatures
  DevBuf<unsigned> nptr, npos;
  DevBuf<int> pos_elem;
  DevBuf<double> eqf, m_sectA, m_temp, m_temp0;
  const void *key_node_inc = nullptr;    // elem_node_item the incidence was built from
  int upd_n_node = 0, upd_n_elem = 0;    // mesh of the last update call (what nodal_stress_3d works on)
  bool gather_valid = false;
  DevBuf<unsigned> slot;
  DevBuf<int> m_eni, m_enn, m_sec, m_form;
  DevBuf<double> m_node, m_sectD;
  const int *key_elem_node_item = nullptr;
  bool slot_valid = false;

  // ---- hecmw_mat_con result held between _count and _fill
  MatCon con;
  struct ConDev {   // the same result built on the device (option matcon_device), held until _fill
    bool ready = false;
    int N = 0, NP = 0, NPL = 0, NPU = 0;
    DevBuf<int> indexL, indexU, itemL, itemU;
    void clear() { ready = false; indexL.release(); indexU.release(); itemL.release(); itemU.release(); }
  } con_dev;
  LocalMesh part;   // hecmw_part1 result of one rank held between _count and _fill

  // ---- peer-memory communication of the fused CG loop (p2p.cuh)
  struct P2P {
    bool active = false;
    void *arena = nullptr;
    size_t bytes = 0;
    void *peer_base[kP2PMaxRanks] = {nullptr};
    P2PView view;
    P2PNeighbors nbrs;
    DevBuf<int> exp_peer, exp_remote;     // per export entry: destination rank, 0-based node index there
    unsigned long long seq[P2P_NQ] = {0, 0}, halo_seq = 0;
    DevBuf<int> err;
  } p2p;

  // ---- communication
  Nccl nccl;
  ncclComm_t_ comm = nullptr;
  int rank = 0, petot = 1;
  int64_t comm_steps = 0;   // NCCL communication steps issued so far (grouped send/recv or all-reduce): read-out "comm_steps"

  int64_t get(const char *k, int64_t dflt) const {
    auto it = opt.find(k);
    return it == opt.end() ? dflt : it->second;
  }
  int grid_for(size_t items, int block, int per_sm = 8) const {
    size_t g = (items + block - 1) / block;
    size_t cap = (size_t)sm_count * per_sm;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
  }
};

Context &ctx();

inline double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace fb
