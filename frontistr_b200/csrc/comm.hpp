// NVLink communication backend that replaces the MPI layer under hecmw_update_3_R and
// hecmw_allreduce_R1 (hecmw1/src/solver/communication/hecmw_comm_f.F90:346-393,669-694).
// NCCL is resolved at run time with dlopen (the process usually has torch's libnccl.so.2
// loaded already); a single-rank run never touches it.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <string>
#ifdef FB_EMUL_KERNELS
#include "nccl_shim.hpp"   // TEST-ONLY (host-compiled library): ranks = processes meeting in shared memory
#endif

namespace fb {

typedef struct ncclComm *ncclComm_t_;
struct ncclUniqueId_ { char internal[128]; };
enum { ncclSum_ = 0, ncclChar_ = 0, ncclInt32_ = 2, ncclFloat64_ = 8 };

struct Nccl {
  void *h = nullptr;
  int (*GetUniqueId)(ncclUniqueId_ *) = nullptr;
  int (*CommInitRank)(ncclComm_t_ *, int, ncclUniqueId_, int) = nullptr;
  int (*CommDestroy)(ncclComm_t_) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t_, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, ncclComm_t_, cudaStream_t) = nullptr;
  int (*Send)(const void *, size_t, int, int, ncclComm_t_, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, ncclComm_t_, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;

#ifdef FB_EMUL_KERNELS
  bool load(std::string &) {   // TEST-ONLY binding to the shared-memory stand-in
    using namespace ncclshim;
    if (GetUniqueId) return true;
    GetUniqueId = [](ncclUniqueId_ *id) { return ncclshim::GetUniqueId(id); };
    CommInitRank = [](ncclComm_t_ *c, int n, ncclUniqueId_ id, int r) { return ncclshim::CommInitRank((Comm **)c, n, id.internal, r); };
    CommDestroy = [](ncclComm_t_ c) { return ncclshim::CommDestroy((Comm *)c); };
    AllReduce = [](const void *s, void *r, size_t n, int dt, int op, ncclComm_t_ c, cudaStream_t) { return ncclshim::AllReduce(s, r, n, dt, op, (Comm *)c, nullptr); };
    AllGather = [](const void *s, void *r, size_t n, int dt, ncclComm_t_ c, cudaStream_t) { return ncclshim::AllGather(s, r, n, dt, (Comm *)c, nullptr); };
    Send = [](const void *b, size_t n, int dt, int peer, ncclComm_t_ c, cudaStream_t) { return ncclshim::Send(b, n, dt, peer, (Comm *)c, nullptr); };
    Recv = [](void *b, size_t n, int dt, int peer, ncclComm_t_ c, cudaStream_t) { return ncclshim::Recv(b, n, dt, peer, (Comm *)c, nullptr); };
    GroupStart = []() { return ncclshim::GroupStart(); };
    GroupEnd = []() { return ncclshim::GroupEnd(); };
    GetErrorString = [](int) -> const char * { return "ncclshim error"; };
    return true;
  }
#else
  bool load(std::string &err) {
    if (h) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
      h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (h) break;
    }
    if (!h) { err = std::string("dlopen(libnccl.so.2): ") + dlerror(); return false; }
#define FB_SYM(field, name)                                   \
  *(void **)(&field) = dlsym(h, name);                        \
  if (!field) { err = std::string("dlsym ") + name; return false; }
    FB_SYM(GetUniqueId, "ncclGetUniqueId");
    FB_SYM(CommInitRank, "ncclCommInitRank");
    FB_SYM(CommDestroy, "ncclCommDestroy");
    FB_SYM(AllReduce, "ncclAllReduce");
    FB_SYM(AllGather, "ncclAllGather");
    FB_SYM(Send, "ncclSend");
    FB_SYM(Recv, "ncclRecv");
    FB_SYM(GroupStart, "ncclGroupStart");
    FB_SYM(GroupEnd, "ncclGroupEnd");
    FB_SYM(GetErrorString, "ncclGetErrorString");
#undef FB_SYM
    return true;
  }
#endif
};

}  // namespace fb
