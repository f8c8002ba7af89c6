// NVLink communication backend that replaces the MPI layer under hecmw_update_3_R and
// hecmw_allreduce_R1 (hecmw1/src/solver/communication/hecmw_comm_f.F90:346-393,669-694).
// NCCL is resolved at run time with dlopen (the process usually has torch's libnccl.so.2
// loaded already); a single-rank run never touches it.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <string>

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
};

}  // namespace fb
