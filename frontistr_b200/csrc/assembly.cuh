// fstr_StiffMatrix + hecmw_mat_ass_elem on the device
// (fistr1/src/analysis/static/fstr_StiffMatrix.f90:18-212, hecmw1/src/solver/matrix/hecmw_mat_ass.f90:31-166)
// and hecmw_mat_ass_bc (hecmw_mat_ass.f90:292-429).
//
// The matrix values live in ONE device allocation laid out [ D (9*NP) | AL (9*NPL) | AU (9*NPU) ],
// each part exactly in the hecmwST_matrix layout, so a block is addressed by a single 32-bit "slot"
// (D: node-1, AL: NP+k, AU: NP+NPL+k).  The binary search that hecmw_mat_add_node repeats for every
// element and every call (hecmw_mat_ass.f90:81-121) depends on topology only: it is done once per
// structure into slot[e*nn*nn + i*nn + j].
#pragma once
#include "common.cuh"   // (under FB_EMUL_KERNELS: the TEST-ONLY host stand-in kernel_emul.hpp, libfistr_b200_emul.so)
#include "element_math.cuh"

namespace fb {

constexpr unsigned kBadSlot = 0xFFFFFFFFu;

__device__ __forceinline__ int bsearch_row(const int *__restrict__ item, int lo, int hi, int key) {
  // hecmw_array_search_i, hecmw_mat_ass.f90:136-166 (0-based positions here)
  while (lo <= hi) {
    const int c = (lo + hi) >> 1;
    const int v = item[c];
    if (key < v) hi = c - 1;
    else if (v < key) lo = c + 1;
    else return c;
  }
  return -1;
}

__global__ void __launch_bounds__(256)
slot_build_kernel(int e0, int e1, int nn, const int *__restrict__ elem_node_index,
                  const int *__restrict__ elem_node_item, int NP, int NPL,
                  const int *__restrict__ indexL, const int *__restrict__ itemL,
                  const int *__restrict__ indexU, const int *__restrict__ itemU,
                  unsigned *__restrict__ slot, size_t slot_base, int *err) {
  const size_t nb = (size_t)nn * nn;
  const size_t total = (size_t)(e1 - e0) * nb;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const int e = e0 + (int)(t / nb), ij = (int)(t % nb), i = ij / nn, j = ij % nn;
    const int *c = elem_node_item + elem_node_index[e];
    const int inod = c[i], jnod = c[j];
    unsigned s;
    if (inod == jnod) s = (unsigned)(inod - 1);
    else if (inod < jnod) {
      const int k = bsearch_row(itemU, indexU[inod - 1], indexU[inod] - 1, jnod);
      s = (k < 0) ? kBadSlot : (unsigned)NP + (unsigned)NPL + (unsigned)k;
    } else {
      const int k = bsearch_row(itemL, indexL[inod - 1], indexL[inod] - 1, jnod);
      s = (k < 0) ? kBadSlot : (unsigned)NP + (unsigned)k;
    }
    if (s == kBadSlot) atomicExch(err, 1);   // "cannot find connectivity", hecmw_mat_ass.f90:86-90
    slot[slot_base + t] = s;
  }
}

// ---- deterministic assembly: element matrices are stored, every matrix block gathers its own -------
// hecmw_mat_add_node adds nn*nn blocks per element with 9 scalar atomics each (hecmw_mat_ass.f90:96,117,128).
// On the GPU that is ~2 G atomic adds at H150.  Instead the element kernel writes its nn*nn blocks to a
// scratch buffer ke[g*9 .. g*9+8], g = contribution index (the same index the slot table uses), and a second
// kernel lets every matrix block sum the contributions that point at it — in ascending g, i.e. in the
// reference's own 1-thread element order (mat_ass.f90:42-50) — and write the block ONCE.  No atomics, no
// zero-fill pass, bit-reproducible.  The inverse map (gptr/gsrc) depends on topology only.
__global__ void __launch_bounds__(256)
gather_count_kernel(size_t ncontrib, const unsigned *__restrict__ slot, unsigned *__restrict__ cnt) {
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < ncontrib; g += (size_t)gridDim.x * blockDim.x) {
    const unsigned s = slot[g];
    if (s != kBadSlot) atomicAdd(cnt + s, 1u);
  }
}
__global__ void __launch_bounds__(256)
gather_fill_kernel(size_t ncontrib, const unsigned *__restrict__ slot, const unsigned *__restrict__ gptr,
                   unsigned *__restrict__ cursor, unsigned *__restrict__ gsrc) {
  for (size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x; g < ncontrib; g += (size_t)gridDim.x * blockDim.x) {
    const unsigned s = slot[g];
    if (s != kBadSlot) gsrc[gptr[s] + atomicAdd(cursor + s, 1u)] = (unsigned)g;
  }
}
// ascending contribution index inside every block's list (lists are short: <= 8 for a hex mesh)
__global__ void __launch_bounds__(256)
gather_sort_kernel(size_t nblk, const unsigned *__restrict__ gptr, unsigned *__restrict__ gsrc) {
  for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < nblk; b += (size_t)gridDim.x * blockDim.x) {
    const unsigned k0 = gptr[b], k1 = gptr[b + 1];
    for (unsigned i = k0 + 1; i < k1; i++) {
      const unsigned v = gsrc[i];
      unsigned j = i;
      while (j > k0 && gsrc[j - 1] > v) { gsrc[j] = gsrc[j - 1]; j--; }
      gsrc[j] = v;
    }
  }
}
// one thread per scalar entry of the matrix: vals[9*b + c] = sum_k ke[9*gsrc[k] + c]
__global__ void __launch_bounds__(256)
gather_sum_kernel(size_t nblk, const unsigned *__restrict__ gptr, const unsigned *__restrict__ gsrc,
                  const double *__restrict__ ke, double *__restrict__ vals) {
  const size_t total = 9 * nblk;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const size_t b = t / 9;
    const unsigned c = (unsigned)(t - 9 * b);
    const unsigned k0 = gptr[b], k1 = gptr[b + 1];
    double acc = 0.0;
    for (unsigned k = k0; k < k1; k++) acc += ld_stream(ke + 9 * (size_t)gsrc[k] + c);
    vals[t] = acc;
  }
}

// One element per warp.  sect[3*s..] = D11, D12, D44 of section s; sect_form[s] = elemopt361.
// ICINV (IC element only): calInverse in registers + shuffles (1) or through shared memory (0), or LDL^T solves (2);
// see element_math.cuh.
template <int FORM, int WARPS, int ICINV = 1>
__global__ void __launch_bounds__(WARPS * 32)
stiff_kernel(int e0, int e1, const int *__restrict__ elem_node_index, const int *__restrict__ elem_node_item,
             const double *__restrict__ node, const int *__restrict__ section_ID,
             const double *__restrict__ sectD, const int *__restrict__ sect_form,
             const unsigned *__restrict__ slot, size_t slot_base, double *__restrict__ vals,
             double *__restrict__ ke_out, int probe = 0) {
  // probe (development timing switches, 0 in production): 1 = no scatter, 2 = no IC condensation
  using namespace elem;
  constexpr int NN = (FORM == FORM_TET10) ? 10 : 8;
  constexpr int NB = NN * NN;
  constexpr int ROUNDS = (NB + 31) / 32;
  FB_DYNAMIC_SMEM(smem_raw);
  Scratch &S = reinterpret_cast<Scratch *>(smem_raw)[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const int wg = blockIdx.x * WARPS + (threadIdx.x >> 5);
  const int nw = gridDim.x * WARPS;
  for (int e = e0 + wg; e < e1; e += nw) {
    const int sec = section_ID[e] - 1;
    if (FORM != FORM_TET10 && sect_form[sec] != FORM) continue;   // warp-uniform
    const double D11 = sectD[3 * sec], D12 = sectD[3 * sec + 1], D44 = sectD[3 * sec + 2];
    const int *conn = elem_node_item + elem_node_index[e];
    if (lane < NN) {
      const size_t n3 = 3 * (size_t)(conn[lane] - 1);
      S.xc[3 * lane] = node[n3]; S.xc[3 * lane + 1] = node[n3 + 1]; S.xc[3 * lane + 2] = node[n3 + 2];
    }
    __syncwarp();
    double kb[ROUNDS][9];
    element_stiffness_warp<FORM, ICINV>(lane, S, D11, D12, D44, kb, !(probe & 2));
    if (probe & 1) {
      double sum = 0.0;
#pragma unroll
      for (int r = 0; r < ROUNDS; r++)
#pragma unroll
        for (int a = 0; a < 9; a++) sum += kb[r][a];
      if (sum == 1.2345e300) vals[0] = sum;
      __syncwarp();
      continue;
    }
    if (ke_out) {
      // gather mode: stage the nn*nn blocks in shared memory, then stream them out as coalesced 16-byte stores
      __syncwarp();
#pragma unroll
      for (int r = 0; r < ROUNDS; r++) {
        const int t = r * 32 + lane;
        if (t < NB) {
#pragma unroll
          for (int a = 0; a < 9; a++) S.stage[t * 9 + a] = kb[r][a];
        }
      }
      __syncwarp();
      double2 *dst = reinterpret_cast<double2 *>(ke_out + 9 * (slot_base + (size_t)(e - e0) * NB));
      const double2 *src = reinterpret_cast<const double2 *>(S.stage);
      for (int q = lane; q < NB * 9 / 2; q += 32) dst[q] = src[q];
      __syncwarp();
      continue;
    }
    // hecmw_mat_add_node: 9 scalar atomic adds per block ('!$omp atomic', hecmw_mat_ass.f90:96,117,128)
#pragma unroll
    for (int r = 0; r < ROUNDS; r++) {
      const int t = r * 32 + lane;
      if (t < NB) {
        const unsigned s = slot[slot_base + (size_t)(e - e0) * NB + t];
        if (s != kBadSlot) {
          double *dst = vals + 9 * (size_t)s;
#pragma unroll
          for (int a = 0; a < 9; a++) atomicAdd(dst + a, kb[r][a]);
        }
      }
    }
    __syncwarp();
  }
}

// Device self-test of the IC element: the condensed stiffness of a few distorted hexahedra through ALL THREE ways of
// condensing (one warp, one block).  out[0] = max |K_shuffle - K_smem|, out[1] = max |K_smem|, out[2] = max |K_ldlt -
// K_smem|.  The host (api_assembly.inc: ic_inverse_mode) uses a kernel other than the shared-memory one only if it
// agrees with it to rounding.
__global__ void __launch_bounds__(32) ic_selftest_kernel(double *__restrict__ out) {
  using namespace elem;
  FB_DYNAMIC_SMEM(smem_raw);
  Scratch &S = *reinterpret_cast<Scratch *>(smem_raw);
  const int lane = threadIdx.x & 31;
  double D11, D12, D44;
  elastic_constants(210000.0, 0.3, D11, D12, D44);
  double maxd1 = 0.0, maxd2 = 0.0, maxabs = 0.0;
  for (int trial = 0; trial < 4; trial++) {
    if (lane < 8) {
      const double corner[3] = {(double)((lane ^ (lane >> 1)) & 1), (double)((lane >> 1) & 1), (double)((lane >> 2) & 1)};
      for (int d = 0; d < 3; d++)   // unit cube (hex8n node order) with a deterministic distortion that grows with trial
        S.xc[3 * lane + d] = corner[d] * (1.0 + 0.5 * d) + 0.07 * trial * sin(1.7 * (3 * lane + d) + 0.9 * trial);
    }
    __syncwarp();
    double ka[2][9], kb[2][9];
    element_stiffness_warp<FORM_IC, 0>(lane, S, D11, D12, D44, ka);
    __syncwarp();
#pragma unroll
    for (int variant = 1; variant <= 2; variant++) {
      if (variant == 1) element_stiffness_warp<FORM_IC, 1>(lane, S, D11, D12, D44, kb);
      else element_stiffness_warp<FORM_IC, 2>(lane, S, D11, D12, D44, kb);
      __syncwarp();
#pragma unroll
      for (int r = 0; r < 2; r++)
#pragma unroll
        for (int a = 0; a < 9; a++) {
          const double d = fabs(ka[r][a] - kb[r][a]);
          double &m = (variant == 1) ? maxd1 : maxd2;
          m = (d != d || m != m) ? d + m : fmax(m, d);   // NaN-propagating max
          maxabs = fmax(maxabs, fabs(ka[r][a]));
        }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double d1 = __shfl_xor_sync(0xffffffffu, maxd1, o), d2 = __shfl_xor_sync(0xffffffffu, maxd2, o);
    const double m = __shfl_xor_sync(0xffffffffu, maxabs, o);
    maxd1 = (d1 != d1 || maxd1 != maxd1) ? d1 + maxd1 : fmax(maxd1, d1);
    maxd2 = (d2 != d2 || maxd2 != maxd2) ? d2 + maxd2 : fmax(maxd2, d2);
    maxabs = fmax(maxabs, m);
  }
  if (lane == 0) { out[0] = maxd1; out[1] = maxabs; out[2] = maxd2; }
}

// ---- geometrically nonlinear (total Lagrangian) assembly and update, SURVEY 8(f) rank 2 -----------
// fstr_StiffMatrix with material%nlgeom_flag == TOTALLAG (fstr_StiffMatrix.f90:60-150 -> STF_C3 / STF_C3D8Bbar) and
// fstr_UpdateNewton (fstr_Update.f90:25-293 -> UPDATE_C3 / Update_C3D8Bbar).  One element per warp; the
// Gauss-point state lives in gp_strain / gp_stress [element][8][6].  u = unode + dunode (3 per node).
template <int FORM, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
stiff_tl_kernel(int e0, int e1, const int *__restrict__ elem_node_index, const int *__restrict__ elem_node_item,
                const double *__restrict__ node, const double *__restrict__ u, const int *__restrict__ section_ID,
                const elem::MatlTL *__restrict__ matl, const int *__restrict__ sect_form,
                const double *__restrict__ gp_strain, const double *__restrict__ gp_stress,
                const unsigned *__restrict__ slot, size_t slot_base, double *__restrict__ vals,
                double *__restrict__ ke_out) {
  using namespace elem;
  constexpr int NN = (FORM == FORM_TET10) ? 10 : 8;
  constexpr int NB = NN * NN;
  constexpr int ROUNDS = (NB + 31) / 32;
  FB_DYNAMIC_SMEM(smem_raw);
  ScratchTL &S = reinterpret_cast<ScratchTL *>(smem_raw)[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const int wg = blockIdx.x * WARPS + (threadIdx.x >> 5);
  const int nw = gridDim.x * WARPS;
  for (int e = e0 + wg; e < e1; e += nw) {
    const int sec = section_ID[e] - 1;
    if (FORM != FORM_TET10 && sect_form[sec] != FORM) continue;   // warp-uniform
    const MatlTL M = matl[sec];
    const int *conn = elem_node_item + elem_node_index[e];
    if (lane < NN) {
      const size_t n3 = 3 * (size_t)(conn[lane] - 1);
#pragma unroll
      for (int c = 0; c < 3; c++) { S.xc[3 * lane + c] = node[n3 + c]; S.uc[3 * lane + c] = u[n3 + c]; }
    }
    __syncwarp();
    double kb[ROUNDS][9];
    element_stiffness_tl_warp<FORM>(lane, S, M, gp_strain + 48 * (size_t)e, gp_stress + 48 * (size_t)e, kb);
    if (ke_out) {
      __syncwarp();
#pragma unroll
      for (int r = 0; r < ROUNDS; r++) {
        const int t = r * 32 + lane;
        if (t < NB) {
#pragma unroll
          for (int a = 0; a < 9; a++) S.stage[t * 9 + a] = kb[r][a];
        }
      }
      __syncwarp();
      double2 *dst = reinterpret_cast<double2 *>(ke_out + 9 * (slot_base + (size_t)(e - e0) * NB));
      const double2 *src = reinterpret_cast<const double2 *>(S.stage);
      for (int q = lane; q < NB * 9 / 2; q += 32) dst[q] = src[q];
      __syncwarp();
      continue;
    }
#pragma unroll
    for (int r = 0; r < ROUNDS; r++) {
      const int t = r * 32 + lane;
      if (t < NB) {
        const unsigned s = slot[slot_base + (size_t)(e - e0) * NB + t];
        if (s != kBadSlot) {
          double *dst = vals + 9 * (size_t)s;
#pragma unroll
          for (int a = 0; a < 9; a++) atomicAdd(dst + a, kb[r][a]);
        }
      }
    }
    __syncwarp();
  }
}

// utot = unode + dunode; writes the new strain / stress of every Gauss point and adds the element's internal
// force into qforce (zeroed by the caller)
template <int FORM, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
update_tl_kernel(int e0, int e1, const int *__restrict__ elem_node_index, const int *__restrict__ elem_node_item,
                 const double *__restrict__ node, const double *__restrict__ utot, const int *__restrict__ section_ID,
                 const elem::MatlTL *__restrict__ matl, const int *__restrict__ sect_form,
                 double *__restrict__ gp_strain, double *__restrict__ gp_stress, double *__restrict__ qforce) {
  using namespace elem;
  constexpr int NN = (FORM == FORM_TET10) ? 10 : 8;
  FB_DYNAMIC_SMEM(smem_raw);
  ScratchTL &S = reinterpret_cast<ScratchTL *>(smem_raw)[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const int wg = blockIdx.x * WARPS + (threadIdx.x >> 5);
  const int nw = gridDim.x * WARPS;
  for (int e = e0 + wg; e < e1; e += nw) {
    const int sec = section_ID[e] - 1;
    if (FORM != FORM_TET10 && sect_form[sec] != FORM) continue;
    const MatlTL M = matl[sec];
    const int *conn = elem_node_item + elem_node_index[e];
    if (lane < NN) {
      const size_t n3 = 3 * (size_t)(conn[lane] - 1);
#pragma unroll
      for (int c = 0; c < 3; c++) { S.xc[3 * lane + c] = node[n3 + c]; S.uc[3 * lane + c] = utot[n3 + c]; }
    }
    __syncwarp();
    const double q = element_update_tl_warp<FORM>(lane, S, M, gp_strain + 48 * (size_t)e, gp_stress + 48 * (size_t)e);
    if (lane < 3 * NN) atomicAdd(qforce + 3 * (size_t)(conn[lane / 3] - 1) + lane % 3, q);
    __syncwarp();
  }
}

// ---- linear strain / stress recovery: fstr_UpdateNewton (nlgeom_flag == INFINITE) + fstr_NodalStress3D ------------
// (fstr_Update.f90:25-293, fstr_NodalStress.f90:13-272), SURVEY 8(f) rank 2.  One element per warp.  The element's
// internal force goes to eqf[3*p + c] with p the element's position in elem_node_item (NOT added to the nodes here):
// the nodes gather their contributions afterwards in ascending element order through the node incidence (below),
// which is the reference's own 1-thread summation order and needs no atomics.
// sectA[s] = thermal expansion coefficient of section s; temp / temp0 == nullptr: no !TEMPERATURE.
template <int FORM, int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
update_lin_kernel(int e0, int e1, const int *__restrict__ elem_node_index, const int *__restrict__ elem_node_item,
                  const double *__restrict__ node, const double *__restrict__ utot, const int *__restrict__ section_ID,
                  const double *__restrict__ sectD, const double *__restrict__ sectA, const int *__restrict__ sect_form,
                  const double *__restrict__ temp, const double *__restrict__ temp0, double ref_temp,
                  double *__restrict__ gp_strain, double *__restrict__ gp_stress, double *__restrict__ eqf) {
  using namespace elem;
  constexpr int NN = (FORM == FORM_TET10) ? 10 : 8;
  FB_DYNAMIC_SMEM(smem_raw);
  ScratchTL &S = reinterpret_cast<ScratchTL *>(smem_raw)[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const int wg = blockIdx.x * WARPS + (threadIdx.x >> 5);
  const int nw = gridDim.x * WARPS;
  const bool has_temp = temp != nullptr;
  for (int e = e0 + wg; e < e1; e += nw) {
    const int sec = section_ID[e] - 1;
    if (FORM != FORM_TET10 && sect_form[sec] != FORM) continue;   // warp-uniform
    const double D11 = sectD[3 * sec], D12 = sectD[3 * sec + 1], D44 = sectD[3 * sec + 2];
    const double alpha = has_temp ? sectA[sec] : 0.0;
    const int p0 = elem_node_index[e];
    const int *conn = elem_node_item + p0;
    if (lane < NN) {
      const size_t nd = (size_t)(conn[lane] - 1);
#pragma unroll
      for (int c = 0; c < 3; c++) { S.xc[3 * lane + c] = node[3 * nd + c]; S.uc[3 * lane + c] = utot[3 * nd + c]; }
      if (has_temp) { S.tc[lane] = temp[nd]; S.t0c[lane] = temp0[nd]; }
    }
    __syncwarp();
    const double q = element_update_lin_warp<FORM>(lane, S, D11, D12, D44, alpha, ref_temp, has_temp,
                                                   gp_strain + 48 * (size_t)e, gp_stress + 48 * (size_t)e);
    if (lane < 3 * NN) eqf[3 * (size_t)p0 + lane] = q;
    __syncwarp();
  }
}

template <int WARPS, int ICINV>
__global__ void __launch_bounds__(WARPS * 32)
update_ic_kernel(int e0, int e1, const int *__restrict__ elem_node_index, const int *__restrict__ elem_node_item,
                 const double *__restrict__ node, const double *__restrict__ utot, const int *__restrict__ section_ID,
                 const double *__restrict__ sectD, const double *__restrict__ sectA, const int *__restrict__ sect_form,
                 const double *__restrict__ temp, const double *__restrict__ temp0, double ref_temp,
                 double *__restrict__ gp_strain, double *__restrict__ gp_stress, double *__restrict__ eqf) {
  using namespace elem;
  FB_DYNAMIC_SMEM(smem_raw);
  Scratch &S = reinterpret_cast<Scratch *>(smem_raw)[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const int wg = blockIdx.x * WARPS + (threadIdx.x >> 5);
  const int nw = gridDim.x * WARPS;
  const bool has_temp = temp != nullptr;
  for (int e = e0 + wg; e < e1; e += nw) {
    const int sec = section_ID[e] - 1;
    if (sect_form[sec] != FORM_IC) continue;   // warp-uniform
    const double D11 = sectD[3 * sec], D12 = sectD[3 * sec + 1], D44 = sectD[3 * sec + 2];
    const double alpha = has_temp ? sectA[sec] : 0.0;
    const int p0 = elem_node_index[e];
    const int *conn = elem_node_item + p0;
    if (lane < 8) {
      const size_t nd = (size_t)(conn[lane] - 1);
#pragma unroll
      for (int c = 0; c < 3; c++) { S.xc[3 * lane + c] = node[3 * nd + c]; S.uc[3 * lane + c] = utot[3 * nd + c]; }
      if (has_temp) { S.tc[lane] = temp[nd]; S.t0c[lane] = temp0[nd]; }
    }
    __syncwarp();
    const double q = element_update_ic_warp<ICINV>(lane, S, D11, D12, D44, alpha, ref_temp, has_temp,
                                                   gp_strain + 48 * (size_t)e, gp_stress + 48 * (size_t)e);
    if (lane < 24) eqf[3 * (size_t)p0 + lane] = q;
    __syncwarp();
  }
}

// ---- node incidence (topology only): node -> the positions p in elem_node_item that name it, ascending --------
// Built with the gather_count / gather_fill / gather_sort kernels above from nslot[p] = elem_node_item[p] - 1.
__global__ void __launch_bounds__(256)
node_slot_kernel(size_t nconn, const int *__restrict__ elem_node_item, unsigned *__restrict__ nslot) {
  for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < nconn; p += (size_t)gridDim.x * blockDim.x)
    nslot[p] = (unsigned)(elem_node_item[p] - 1);
}
__global__ void __launch_bounds__(256)
pos_elem_kernel(int n_elem, const int *__restrict__ elem_node_index, int *__restrict__ pos_elem) {
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_elem; e += gridDim.x * blockDim.x)
    for (int p = elem_node_index[e]; p < elem_node_index[e + 1]; p++) pos_elem[p] = e;
}
// QFORCE(3*node + c) = sum over the incident (element, local node) pairs, ascending element order
__global__ void __launch_bounds__(256)
node_gather3_kernel(int n_node, const unsigned *__restrict__ nptr, const unsigned *__restrict__ npos,
                    const double *__restrict__ eqf, double *__restrict__ out) {
  const size_t total = 3 * (size_t)n_node;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const size_t n = t / 3;
    const unsigned c = (unsigned)(t - 3 * n);
    double acc = 0.0;
    for (unsigned k = nptr[n]; k < nptr[n + 1]; k++) acc += eqf[3 * (size_t)npos[k] + c];
    out[t] = acc;
  }
}
// fstr_NodalStress3D, nodal part: thread (node, c) with c < 6 strain, c >= 6 stress.  Extrapolates the Gauss-point
// values of every incident element to this node (NodalStress_INV3), sums them in ascending element order and divides
// by their number (:204-228).
__global__ void __launch_bounds__(256)
nodal_stress_kernel(int n_node, const unsigned *__restrict__ nptr, const unsigned *__restrict__ npos,
                    const int *__restrict__ pos_elem, const int *__restrict__ elem_node_index,
                    const double *__restrict__ gp_strain, const double *__restrict__ gp_stress, elem::NodalExtrap X,
                    double *__restrict__ STRAIN, double *__restrict__ STRESS) {
  const size_t total = 12 * (size_t)n_node;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const size_t n = t / 12;
    const int c12 = (int)(t - 12 * n), c = c12 % 6;
    const double *gp = (c12 < 6) ? gp_strain : gp_stress;
    const unsigned k0 = nptr[n], k1 = nptr[n + 1];
    double acc = 0.0;
    for (unsigned k = k0; k < k1; k++) {
      const int p = (int)npos[k], e = pos_elem[p], pe = elem_node_index[e];
      acc = acc + elem::nodal_extrapolate(X, elem_node_index[e + 1] - pe, p - pe, gp + 48 * (size_t)e, c);
    }
    if (k1 > k0) acc = acc / (double)(k1 - k0);
    ((c12 < 6) ? STRAIN : STRESS)[6 * n + c] = acc;
  }
}
// elemental part (ElementStress_C3, static_LIB_3d.f90:881-913): average over the Gauss points
__global__ void __launch_bounds__(256)
elem_stress_kernel(int n_elem, const int *__restrict__ elem_node_index, const double *__restrict__ gp_strain,
                   const double *__restrict__ gp_stress, double *__restrict__ ESTRAIN, double *__restrict__ ESTRESS) {
  const size_t total = 12 * (size_t)n_elem;
  for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
    const size_t e = t / 12;
    const int c12 = (int)(t - 12 * e), c = c12 % 6;
    const int ni = (elem_node_index[e + 1] - elem_node_index[e] == 8) ? 8 : 4;
    const double *gp = ((c12 < 6) ? gp_strain : gp_stress) + 48 * e;
    double acc = 0.0;
    for (int i = 0; i < ni; i++) acc = acc + gp[6 * i + c];
    ((c12 < 6) ? ESTRAIN : ESTRESS)[6 * e + c] = acc / ni;
  }
}
__global__ void __launch_bounds__(256)
mises_kernel(size_t n, const double *__restrict__ stress, double *__restrict__ mises) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    mises[i] = elem::mises_of(stress + 6 * i);
}

// ---- hecmw_mat_ass_bc, batched ---------------------------------------------------------------
// One thread per (constraint c, entry t of node n's row): t = 0 is the diagonal block,
// 1..nL the L entries, then the U entries.  STAGE 0 subtracts K(:,c)*rhs from B (reading the
// still-unmodified matrix); STAGE 1 zeroes row/column and puts 1 on the diagonal; the caller then
// writes B(c) = rhs.  Equals the reference's one-constraint-at-a-time elimination because a
// constraint only ever zeroes its own row and column — provided every DOF appears ONCE in the batch: the host side
// (api_assembly.inc) folds repeated DOFs into one entry that eliminates with the first value and sets the last.
// Reproducibility: a constraint with rhs == 0 (the usual fixed support) contributes exactly nothing to B and issues no
// atomic at all, so B is bit-reproducible; with a NON-ZERO prescribed value several constraints may add into the same
// B entry through atomicAdd in an order that differs from run to run (and from the reference's constraint order): B then
// agrees with the reference to the last bit or two, not bitwise.
template <int STAGE>
__global__ void __launch_bounds__(256)
bc_kernel(int nbc, const int *__restrict__ bc_node, const int *__restrict__ bc_dof,
          const double *__restrict__ bc_val, const int *__restrict__ bc_ptr /* nbc+1 work offsets */,
          int NP, int NPL, const int *__restrict__ indexL, const int *__restrict__ itemL,
          const int *__restrict__ indexU, const int *__restrict__ itemU, double *__restrict__ vals,
          double *__restrict__ B) {
  const int total = bc_ptr[nbc];
  for (int w = blockIdx.x * blockDim.x + threadIdx.x; w < total; w += gridDim.x * blockDim.x) {
    int lo = 0, hi = nbc - 1;   // constraint owning work item w
    while (lo < hi) { const int m = (lo + hi + 1) >> 1; if (bc_ptr[m] <= w) lo = m; else hi = m - 1; }
    const int c = lo, t = w - bc_ptr[c];
    const int n = bc_node[c], d = bc_dof[c] - 1;
    const double rhs = bc_val[c];
    if (STAGE == 0 && rhs == 0.0) continue;   // B - K * 0 = B
    double *Dn = vals + 9 * (size_t)(n - 1);
    const int nL = indexL[n] - indexL[n - 1];
    if (t == 0) {
      if (STAGE == 0) {
        for (int i = 0; i < 3; i++)
          if (i != d) atomicAdd(&B[3 * (size_t)(n - 1) + i], -(Dn[3 * i + d] * rhs));
      } else {
        for (int i = 0; i < 3; i++) { Dn[3 * d + i] = 0.0; Dn[3 * i + d] = 0.0; }
        Dn[3 * d + d] = 1.0;
      }
      continue;
    }
    double *own, *tr;   // block (n, in) in n's row; transposed-position block (in, n)
    int in;
    if (t - 1 < nL) {
      const int k = indexL[n - 1] + t - 1;
      in = itemL[k];
      own = vals + 9 * ((size_t)NP + k);
      const int ik = bsearch_row(itemU, indexU[in - 1], indexU[in] - 1, n);
      tr = (ik < 0) ? nullptr : vals + 9 * ((size_t)NP + NPL + ik);
    } else {
      const int k = indexU[n - 1] + t - 1 - nL;
      in = itemU[k];
      own = vals + 9 * ((size_t)NP + NPL + k);
      const int ik = bsearch_row(itemL, indexL[in - 1], indexL[in] - 1, n);
      tr = (ik < 0) ? nullptr : vals + 9 * ((size_t)NP + ik);
    }
    if (STAGE == 0) {
      if (tr)
        for (int i = 0; i < 3; i++) atomicAdd(&B[3 * (size_t)(in - 1) + i], -(tr[3 * i + d] * rhs));
    } else {
      for (int i = 0; i < 3; i++) own[3 * d + i] = 0.0;
      if (tr)
        for (int i = 0; i < 3; i++) tr[3 * i + d] = 0.0;
    }
  }
}

__global__ void bc_set_rhs_kernel(int nbc, const int *__restrict__ bc_node, const int *__restrict__ bc_dof,
                                  const double *__restrict__ bc_val, double *__restrict__ B) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < nbc) B[3 * (size_t)(bc_node[c] - 1) + bc_dof[c] - 1] = bc_val[c];
}

}  // namespace fb
