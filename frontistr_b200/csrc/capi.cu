// C-ABI of the fistr_b200 library (include/fistr_b200.h): context, residency, SpMV, preconditioners,
// the CG driver mirroring hecmw_solve -> hecmw_solve_iterative -> hecmw_solve_CG, assembly and BC.
#include <cmath>
#include <cstdarg>

#include <cub/device/device_radix_sort.cuh>
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
  // never destroyed: at process exit the CUDA runtime may already be gone, and the owning buffers of the context must not
  // call into it then (fistr_b200_finalize is the orderly way to release everything)
  static Context *c = new Context;
  return *c;
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

#include "api_internal.inc"
}  // namespace fb

using namespace fb;

// =============================================================================================
extern "C" {
#include "api_core.inc"
#include "api_assembly.inc"
#include "api_solver.inc"
#include "api_misc.inc"
}  // extern "C"

