// NVLink peer-memory communication for the CG hot loop: replaces, inside the fused DIAG-CG iteration, the
// MPI traffic of hecmw_update_3_R (communication/hecmw_comm_f.F90:669-694, hecmw_solver_SR_33.F90:42-124)
// and hecmw_allreduce_R1 (hecmw_comm_f.F90:346-393) by stores into the neighbours' memory issued from the
// compute kernels themselves:
//
//   * halo:  the kernel that forms p = z + beta p writes the export values straight into the halo part of
//            each neighbour's p vector (peer stores over NVLink); when the whole grid has done so the last
//            block publishes a sequence number in every neighbour's flag word.  The neighbour's SpMV waits
//            for the flags of all its neighbours before it reads x.
//   * dots:  the last block of the deterministic grid reduction writes its rank's partial sum (with a
//            sequence tag) into a slot of EVERY rank's table; the prologue of the next kernel waits for all
//            tags and sums the partials in rank order, so every rank obtains the bit-identical value — no
//            separate all-reduce kernel, no host involvement.
//
// Every rank owns one "arena" (plain cudaMalloc memory exported with cudaIpcGetMemHandle); layout:
//   [ P2PEntry table [P2P_NQ][2][8] | halo flags u64[8] | p vector (3*NP+8 doubles) ]
// One process drives one GPU, so a kernel spinning on a word written by another GPU cannot starve its
// producer.  Every spin has a time-out that raises an error flag instead of hanging the device; after the first
// time-out all later waits on that GPU return immediately, so a broken exchange costs seconds, not minutes.
#pragma once
#include "common.cuh"

namespace fb {

constexpr int kP2PMaxRanks = 8;
enum { P2P_Q_PQ = 0, P2P_Q_RRRZ = 1, P2P_NQ = 2 };

struct P2PEntry {
  double v[2];
  unsigned long long tag;
  unsigned long long pad;
};

struct P2PView {
  int nranks = 0, me = 0;
  P2PEntry *ar[kP2PMaxRanks];              // ar[r]: entry table of rank r (peer-mapped for r != me)
  unsigned long long *hflag[kP2PMaxRanks]; // hflag[r]: halo flags of rank r, indexed by source rank
  double *vecP[kP2PMaxRanks];              // p vector of rank r
  int *err;                                // local error word (spin time-out)
};

struct P2PNeighbors {
  int n = 0;
  int rank[kP2PMaxRanks];
};

#ifdef FB_EMUL_KERNELS   // TEST-ONLY host stand-ins: the "GPUs" are host threads, system scope = process memory
inline unsigned long long ld_acquire_sys(const unsigned long long *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
inline void st_release_sys(unsigned long long *p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
inline void st_relaxed_sys(double *p, double v) { uint64_t b; __builtin_memcpy(&b, &v, 8); __atomic_store_n(reinterpret_cast<uint64_t *>(p), b, __ATOMIC_RELAXED); }
inline double ld_relaxed_sys(const double *p) { const uint64_t b = __atomic_load_n(reinterpret_cast<const uint64_t *>(p), __ATOMIC_RELAXED); double v; __builtin_memcpy(&v, &b, 8); return v; }
inline void st_relaxed_sys(unsigned long long *p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELAXED); }
unsigned long long p2p_now_ns();   // emul_test.cpp (steady clock)
#else
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(double *p, double v) {
  asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys(const double *p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long p2p_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

#endif
constexpr unsigned long long kP2PTimeoutNs = 2000000000ull;   // 2 s; once any wait on this GPU has timed out, later waits give up at once

// ONE WARP (all 32 lanes call): lane r hands this rank's partial sums to rank r (itself included):
// values, one system-scope fence, then the tag (fence + relaxed store = release pattern).
__device__ __forceinline__ void p2p_publish_warp(const P2PView &pv, int q, unsigned long long seq, double v0, double v1) {
  const int lane = threadIdx.x & 31;
  if (lane < pv.nranks) {
    P2PEntry *e = pv.ar[lane] + (q * 2 + (int)(seq & 1ull)) * kP2PMaxRanks + pv.me;
    st_relaxed_sys(&e->v[0], v0);
    st_relaxed_sys(&e->v[1], v1);
    __threadfence_system();
    st_relaxed_sys(&e->tag, seq);
  }
}

// FIRST WARP of a block (all 32 lanes call): lane r waits for rank r's partials of (q, seq); the partials are
// then summed in rank order by every lane, so all lanes (and all ranks) obtain the bit-identical sums.
__device__ __forceinline__ void p2p_consume_warp(const P2PView &pv, int q, unsigned long long seq, double &s0, double &s1) {
  const int lane = threadIdx.x & 31;
  const P2PEntry *base = pv.ar[pv.me] + (q * 2 + (int)(seq & 1ull)) * kP2PMaxRanks;
  double v0 = 0.0, v1 = 0.0;
  if (lane < pv.nranks) {
    const unsigned long long t0 = p2p_now_ns();
    while (ld_acquire_sys(&base[lane].tag) != seq) {
      if (*(volatile int *)pv.err != 0) break;
      if (p2p_now_ns() - t0 > kP2PTimeoutNs) { atomicExch(pv.err, 1); break; }
    }
    v0 = ld_relaxed_sys(&base[lane].v[0]);
    v1 = ld_relaxed_sys(&base[lane].v[1]);
  }
  s0 = 0.0; s1 = 0.0;
  for (int r = 0; r < pv.nranks; r++) {
    s0 += __shfl_sync(0xffffffffu, v0, r);
    s1 += __shfl_sync(0xffffffffu, v1, r);
  }
}

// ONE WARP (all 32 lanes call): lane a waits until neighbour a has delivered halo exchange number `seq`
__device__ __forceinline__ void p2p_wait_halo_warp(const P2PView &pv, const P2PNeighbors &nb, unsigned long long seq) {
  const int lane = threadIdx.x & 31;
  if (lane < nb.n) {
    const unsigned long long *f = pv.hflag[pv.me] + nb.rank[lane];
    const unsigned long long t0 = p2p_now_ns();
    while (ld_acquire_sys(f) < seq) {
      if (*(volatile int *)pv.err != 0) break;
      if (p2p_now_ns() - t0 > kP2PTimeoutNs) { atomicExch(pv.err, 2); break; }
    }
  }
  __syncwarp();
}

// ONE thread (of the last block to finish its peer stores): tell every neighbour that exchange `seq` landed
__device__ __forceinline__ void p2p_signal_halo(const P2PView &pv, const P2PNeighbors &nb, unsigned long long seq) {
  __threadfence_system();
  for (int a = 0; a < nb.n; a++) st_relaxed_sys(pv.hflag[nb.rank[a]] + pv.me, seq);
}

}  // namespace fb
