// TEST-ONLY: implementation of the execution models behind kernel_emul.hpp.  Included by exactly one translation unit
// per shared library (emul_test.cpp for libfistr_b200_emul.so, the rewritten capi.cu for libfistr_b200_hostemu.so).
#pragma once
#include <map>
#include <mutex>
#include <string>
#include <pthread.h>
#include <sys/mman.h>
#include <unistd.h>
#include <ucontext.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

// ---- execution models behind the warp / block collectives -------------------------------------------------------
// (a) emul_launch_warps: the 32 lanes of a warp as fibers (FB_EMUL_PTHREAD_WARPS: 32 host threads and pthread barriers instead);
// (b) emul_launch_blocks: the threads of a block are cooperative fibers (ucontext) of the calling host thread; a fiber
//     that reaches __syncthreads / __syncwarp / a shuffle parks itself and the scheduler releases the block (warp) once
//     every live fiber of it has arrived.  Everything is per host thread, so several host threads = several "GPUs".

namespace fb {
unsigned long long p2p_now_ns() {
  // FB_EMUL_P2P_TIMEOUT=1 (test of the time-out handling): every look at the clock jumps 10 s ahead, so the first wait that
  // has to spin at all gives up
  static const bool jump = getenv("FB_EMUL_P2P_TIMEOUT") != nullptr;
  static thread_local unsigned long long skew = 0;
  if (jump) skew += 10000000000ull;
  return skew + (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
namespace emulwarp {
static pthread_barrier_t g_bar;
static thread_local int t_lane = 0;
static double g_slot_d[32];
static int g_slot_i[32];

// Context switch.  glibc's swapcontext saves / restores the signal mask with two system calls per switch, and a block
// of 256 fibers switches ~10^4 times per kernel: on x86-64 a 12-instruction switch of the callee-saved registers is
// used instead (about 50x faster); ucontext stays as the portable fall-back.
#if defined(__x86_64__) && !defined(FB_EMUL_UCONTEXT)
#define FB_EMUL_ASM_SWITCH 1
extern "C" void fb_emul_ctx_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.hidden fb_emul_ctx_switch
.globl fb_emul_ctx_switch
.type fb_emul_ctx_switch,@function
fb_emul_ctx_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size fb_emul_ctx_switch,.-fb_emul_ctx_switch
)");
#endif

struct Fiber {
#ifdef FB_EMUL_ASM_SWITCH
  void *sp = nullptr;
#else
  ucontext_t ctx;
#endif
  std::unique_ptr<char[]> stack;
  int state = 0;   // 0 runnable, 1 parked at the block barrier, 2 parked at the warp barrier, 3 finished
};
struct FiberBlock {
  std::vector<Fiber> f;
#ifdef FB_EMUL_ASM_SWITCH
  void *main_sp = nullptr;
#else
  ucontext_t main;
#endif
  int cur = -1;
  unsigned grid = 1, block = 1, bidx = 0;   // block = number of fibers run together
  unsigned blockdim = 0, tid0 = 0;          // reported blockDim.x (0: = block) and threadIdx.x of fiber 0 (one warp of a larger block)
  std::function<void()> body;
  std::vector<double> slot_d;
  std::vector<int> slot_i;
  std::vector<unsigned long long> site;     // per fiber: call site (<< 2) | kind of the warp collective it is parked at
};
enum { kKindSync = 0, kKindShflPut = 1, kKindShflGet = 2 };
static thread_local FiberBlock *t_fb = nullptr;   // non-null while this host thread runs a fiber block

static void fiber_to_main(FiberBlock &B) {
#ifdef FB_EMUL_ASM_SWITCH
  fb_emul_ctx_switch(&B.f[B.cur].sp, B.main_sp);
#else
  swapcontext(&B.f[B.cur].ctx, &B.main);
#endif
}
static void fiber_yield(int state) {
  FiberBlock &B = *t_fb;
  B.f[B.cur].state = state;
  fiber_to_main(B);
}
static void fiber_entry() {
  FiberBlock &B = *t_fb;
  B.body();
  B.f[B.cur].state = 3;
  fiber_to_main(B);
  abort();   // a finished fiber is never resumed
}
static thread_local std::vector<Fiber> t_fiber_pool;   // stacks are reused from launch to launch
static int g_order_mode = -1;                 // -1: take FB_EMUL_FIBER_ORDER at the first launch; 0 ascending, 1 descending, 2 random
static unsigned long long g_order_seed = 0;
static void run_fiber_block(FiberBlock &B) {
  constexpr size_t STACK = 96 * 1024;
  const int n = (int)B.block;
  B.f.swap(t_fiber_pool);
  if ((int)B.f.size() < n) B.f.resize(n);
  B.slot_d.assign(n, 0.0); B.slot_i.assign(n, 0); B.site.assign(n, 0ull);
  t_fb = &B;
  for (int i = 0; i < n; i++) {
    Fiber &F = B.f[i];
    if (!F.stack) F.stack.reset(new char[STACK]);
#ifdef FB_EMUL_ASM_SWITCH
    {   // initial frame: six callee-saved registers (zero) below the "return address" fiber_entry; at its first
        // instruction rsp == 8 (mod 16), as after a call
      uintptr_t top = (reinterpret_cast<uintptr_t>(F.stack.get()) + STACK) & ~(uintptr_t)15;
      void **a = reinterpret_cast<void **>(top - 16);
      a[0] = reinterpret_cast<void *>(&fiber_entry);
      a[1] = nullptr;
      for (int r = 1; r <= 6; r++) a[-r] = nullptr;
      F.sp = a - 6;
    }
#else
    getcontext(&F.ctx);
    F.ctx.uc_stack.ss_sp = F.stack.get(); F.ctx.uc_stack.ss_size = STACK; F.ctx.uc_link = &B.main;
    makecontext(&F.ctx, fiber_entry, 0);
#endif
    F.state = 0;
  }
  // Order in which the runnable fibers of a scheduling round get the processor: ascending (default), descending
  // (FB_EMUL_FIBER_ORDER=reverse) or a fresh random permutation per round (FB_EMUL_FIBER_ORDER=random[:seed]).  Between two
  // barriers / shuffles the lanes of a real warp are NOT ordered, so a kernel whose result depends on this order has an
  // intra-warp (or intra-block) race - a missing __syncwarp / __syncthreads - that the default order would hide.  The
  // test-suite runs the emulated kernels under all three (tests/test_emul_fiber_order.py).
  if (g_order_mode < 0) {
    const char *e = getenv("FB_EMUL_FIBER_ORDER");
    const char *c = e ? strchr(e, ':') : nullptr;
    g_order_mode = !e ? 0 : (!strncmp(e, "reverse", 7) ? 1 : (!strncmp(e, "random", 6) ? 2 : 0));
    g_order_seed = c ? strtoull(c + 1, nullptr, 10) : 0ull;
  }
  const int order_mode = g_order_mode;
  static thread_local unsigned long long rng = 0, rng_seed = ~0ull;
  if (rng_seed != g_order_seed) { rng_seed = g_order_seed; rng = 0x9E3779B97F4A7C15ull ^ (g_order_seed * 0xD1B54A32D192ED03ull); }
  std::vector<int> order(n);
  for (int i = 0; i < n; i++) order[i] = i;
  for (;;) {
    bool ran = false, alive = false;
    if (order_mode == 1) { for (int i = 0; i < n; i++) order[i] = n - 1 - i; }
    else if (order_mode == 2) {
      for (int i = n - 1; i > 0; i--) {
        rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17;
        std::swap(order[i], order[(int)(rng % (unsigned long long)(i + 1))]);
      }
    }
    for (int oi = 0; oi < n; oi++) {
      const int i = order[oi];
      if (B.f[i].state != 0) continue;
      B.cur = i;
      fb::gridDim.x = B.grid; fb::blockDim.x = B.blockdim ? B.blockdim : B.block; fb::blockIdx.x = B.bidx; fb::threadIdx.x = B.tid0 + (unsigned)i;
#ifdef FB_EMUL_ASM_SWITCH
      fb_emul_ctx_switch(&B.main_sp, B.f[i].sp);
#else
      swapcontext(&B.main, &B.f[i].ctx);
#endif
      ran = true;
    }
    // release barriers whose every live participant has arrived
    bool all_block = true, any_block = false;
    for (int i = 0; i < n; i++) {
      if (B.f[i].state != 3) alive = true;
      if (B.f[i].state == 1) any_block = true;
      else if (B.f[i].state != 3) all_block = false;
    }
    if (!alive) break;
    bool released = false;
    if (any_block && all_block) { for (int i = 0; i < n; i++) if (B.f[i].state == 1) B.f[i].state = 0; released = true; }
    for (int w = 0; w * 32 < n; w++) {
      bool all = true, any = false;
      for (int i = w * 32; i < std::min(n, w * 32 + 32); i++) {
        if (B.f[i].state == 2) any = true;
        else if (B.f[i].state != 3) all = false;
      }
      if (any && all) {
        // all live lanes have arrived: they must be at the same collective.  Lanes of one shuffle at different call sites, or
        // some lanes at a shuffle and others at a __syncwarp, is a divergent collective (undefined on hardware).
        unsigned long long first = ~0ull;
        for (int i = w * 32; i < std::min(n, w * 32 + 32); i++) {
          if (B.f[i].state != 2) continue;
          const unsigned long long t = B.site[i];
          if (first == ~0ull) first = t;
          const bool both_sync = (t & 3ull) == kKindSync && (first & 3ull) == kKindSync;
          if (t != first && !both_sync && (t >> 2) != 0 && (first >> 2) != 0) {
            fprintf(stderr, "emul: divergent warp collective in block %u warp %d: lane %d is at site %llx kind %llu, lane %d at site %llx kind %llu\n",
                    B.bidx, w, i & 31, t >> 2, t & 3ull, 0, first >> 2, first & 3ull);
            abort();
          }
        }
        for (int i = w * 32; i < std::min(n, w * 32 + 32); i++) if (B.f[i].state == 2) B.f[i].state = 0;
        released = true;
      }
    }
    if (!ran && !released) { fprintf(stderr, "emul: fiber block dead-locked (divergent barrier)\n"); abort(); }
  }
  t_fb = nullptr;
  B.f.swap(t_fiber_pool);
}

double shfl(double v, int src, unsigned long long site) {
  if (t_fb) {
    FiberBlock &B = *t_fb;
    const int me = B.cur, base = me & ~31;
    B.slot_d[me] = v;
    B.site[me] = (site << 2) | kKindShflPut;
    fiber_yield(2);
    const double r = B.slot_d[base + (src & 31)];
    B.site[me] = (site << 2) | kKindShflGet;
    fiber_yield(2);
    return r;
  }
  g_slot_d[t_lane] = v;
  pthread_barrier_wait(&g_bar);
  const double r = g_slot_d[src & 31];
  pthread_barrier_wait(&g_bar);
  return r;
}
int shfl(int v, int src, unsigned long long site) {
  if (t_fb) {
    FiberBlock &B = *t_fb;
    const int me = B.cur, base = me & ~31;
    B.slot_i[me] = v;
    B.site[me] = (site << 2) | kKindShflPut;
    fiber_yield(2);
    const int r = B.slot_i[base + (src & 31)];
    B.site[me] = (site << 2) | kKindShflGet;
    fiber_yield(2);
    return r;
  }
  g_slot_i[t_lane] = v;
  pthread_barrier_wait(&g_bar);
  const int r = g_slot_i[src & 31];
  pthread_barrier_wait(&g_bar);
  return r;
}
void syncwarp(unsigned long long site) {
  if (t_fb) { t_fb->site[t_fb->cur] = (site << 2) | kKindSync; fiber_yield(2); }
  else pthread_barrier_wait(&g_bar);
}
void syncthreads() {
  if (!t_fb) { fprintf(stderr, "emul: __syncthreads outside emul_launch_blocks\n"); abort(); }
  fiber_yield(1);
}
template <class F>
static void run_warp(F f) {   // f(lane) on the 32 lanes of one warp in lock step (fibers; FB_EMUL_PTHREAD_WARPS: host threads)
#ifdef FB_EMUL_PTHREAD_WARPS
  pthread_barrier_init(&g_bar, nullptr, 32);
  std::vector<std::thread> th;
  for (int lane = 0; lane < 32; lane++) th.emplace_back([lane, &f]() { t_lane = lane; f(lane); });
  for (auto &t : th) t.join();
  pthread_barrier_destroy(&g_bar);
#else
  FiberBlock B;
  B.block = 32;
  B.body = [&f]() { f((int)(fb::threadIdx.x & 31)); };
  run_fiber_block(B);
#endif
}
} }

// kernel<<<grid, block>>>(args...) for block-cooperative kernels: blocks one after the other, the threads of a block as
// fibers of the calling host thread (see above)
template <class K, class... A>
static void emul_launch_blocks(unsigned grid, unsigned block, K kernel, A... args) {
  for (unsigned b = 0; b < grid; b++) {
    fb::emulwarp::FiberBlock B;
    B.grid = grid; B.block = block; B.bidx = b;
    B.body = [&]() { kernel(args...); };
    fb::emulwarp::run_fiber_block(B);
  }
}


// `kernel<<<grid, block, smem, stream>>>(args...)` after the rewrite of tools/build_hostemu.py:
//   fb::EmulCfg(grid, block, smem, stream).launch([&]() { kernel(args...); });
namespace fb {
// Dynamic shared memory of an emulated block, electric-fence style like the "device" allocations of cuda_shim.hpp: the
// buffer ENDS at an inaccessible page, so an index past the launch's shared-memory size dies on the spot (on hardware: an
// out-of-range shared address, Xid 13), and it is refilled with 0xFF bytes for every block (NaN as a double, -1 as an int):
// shared memory is not zeroed on hardware either.
struct EmulSmem {
  unsigned char *base = nullptr, *user = nullptr;
  size_t len = 0, bytes = 0;
  explicit EmulSmem(size_t n) {
    const size_t page = (size_t)sysconf(_SC_PAGESIZE);
    bytes = (std::max<size_t>(n, 16) + 15) & ~(size_t)15;
    const size_t body = (bytes + page - 1) / page * page;
    len = body + 2 * page;
    base = (unsigned char *)mmap(nullptr, len, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (base == (unsigned char *)MAP_FAILED) { fprintf(stderr, "emul: mmap of %zu bytes of shared memory failed\n", len); abort(); }
    mprotect(base, page, PROT_NONE);
    mprotect(base + page + body, page, PROT_NONE);
    user = base + page + body - bytes;
  }
  ~EmulSmem() { if (base) munmap(base, len); }
  EmulSmem(const EmulSmem &) = delete;
  EmulSmem &operator=(const EmulSmem &) = delete;
  unsigned char *fresh() { memset(user, 0xFF, bytes); return user; }
};

// every launch configuration the library has used in this process, per kernel as named at the launch site: the largest
// block, dynamic shared memory and grid (tests/test_gpu_suite_on_host_api.py holds them against the resource usage of the
// kernels nvcc compiled for sm_100a - registers x threads, shared-memory opt-in - which no CPU run can exercise)
struct EmulLaunchStat { unsigned block = 0, grid = 0; size_t smem = 0; long long count = 0; };
inline std::map<std::string, EmulLaunchStat> &emul_launch_log() { static std::map<std::string, EmulLaunchStat> m; return m; }

struct EmulCfg {
  unsigned grid, block;
  size_t smem;
  EmulCfg(long long g, long long b, size_t s = 0, cudaStream_t = nullptr) : grid((unsigned)g), block((unsigned)b), smem(s) {}
  template <class F>
  void launch(const char *name, F body) const {
    {
      static std::mutex mu;
      std::lock_guard<std::mutex> lk(mu);
      EmulLaunchStat &st = emul_launch_log()[name];
      st.block = std::max(st.block, block); st.grid = std::max(st.grid, grid); st.smem = std::max(st.smem, smem); st.count++;
    }
    launch(body);
  }
  template <class F>
  void launch(F body) const {
    // what the CUDA runtime would refuse at launch ("invalid configuration argument" / "too much shared data")
    if (grid == 0 || block == 0 || block > 1024 || smem > 227 * 1024) {
      fprintf(stderr, "emul: invalid launch configuration <<<%u, %u, %zu>>>\n", grid, block, smem);
      abort();
    }
    EmulSmem sm(smem);
    for (unsigned b = 0; b < grid; b++) {
      emulwarp::FiberBlock B;
      B.grid = grid; B.block = block; B.bidx = b;
      B.body = body;
      emul_smem() = sm.fresh();
      emulwarp::run_fiber_block(B);
    }
    emul_smem() = nullptr;
  }
};
}  // namespace fb

// "name\tblock\tsmem\tgrid\tcount\n" per kernel launched so far; returns the length needed (call again with a larger buffer)
extern "C" int fbemul_launch_log(char *buf, int len) {
  std::string out;
  for (const auto &kv : fb::emul_launch_log())
    out += kv.first + "\t" + std::to_string(kv.second.block) + "\t" + std::to_string(kv.second.smem) + "\t" +
           std::to_string(kv.second.grid) + "\t" + std::to_string(kv.second.count) + "\n";
  if (buf && len > 0) { strncpy(buf, out.c_str(), (size_t)len - 1); buf[len - 1] = 0; }
  return (int)out.size() + 1;
}

// order in which the runnable fibers of a block are scheduled (see run_fiber_block): 0 ascending, 1 descending, 2 random
extern "C" void fbemul_set_fiber_order(int mode, unsigned long long seed) {
  fb::emulwarp::g_order_mode = mode;
  fb::emulwarp::g_order_seed = seed;
}
