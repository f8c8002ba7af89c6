// TEST-ONLY stand-in for the NCCL entry points the library resolves with dlopen (comm.hpp), for the host-compiled
// library (tools/build_hostemu.py): ranks are PROCESSES on one machine that meet in a POSIX shared-memory segment named
// after the unique id.  Every collective is "write my part into my outbox, barrier, read the others' parts, barrier";
// grouped send/recv pairs are matched in call order per (source, destination), which is how the library uses them.
// Streams execute at once in this build, so every call completes before it returns.  Never part of the product build.
#pragma once
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <unistd.h>

#include <cstdint>
#include <ctime>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace fb {
namespace ncclshim {

constexpr int kMaxRanks = 8;
constexpr size_t kOutbox = (size_t)16 << 20;   // bytes of payload per rank and collective

struct Msg { int dst; size_t bytes, off; };
struct Shared {
  volatile int arrived, sense;             // sense-reversing barrier
  volatile int joined;
  struct PerRank {
    volatile int nmsg;
    Msg msg[64];
    volatile size_t used;
  } r[kMaxRanks];
  // outboxes follow
};
struct Comm {
  int nranks = 0, rank = 0, local_sense = 0;
  Shared *sh = nullptr;
  size_t len = 0;
  char name[64];
  // pending grouped operations
  struct Op { bool send; void *buf; size_t bytes; int peer; };
  std::vector<Op> ops;
  int group_depth = 0;
  char *outbox(int r) const { return reinterpret_cast<char *>(sh) + ((sizeof(Shared) + 4095) & ~(size_t)4095) + (size_t)r * kOutbox; }
};

inline Comm *&current() { static Comm *c = nullptr; return c; }   // NCCL's group calls carry no communicator: one per process

inline void barrier(Comm *c) {
  Shared *s = c->sh;
  c->local_sense ^= 1;
  if (__atomic_add_fetch(&s->arrived, 1, __ATOMIC_ACQ_REL) == c->nranks) {
    __atomic_store_n(&s->arrived, 0, __ATOMIC_RELAXED);
    __atomic_store_n(&s->sense, c->local_sense, __ATOMIC_RELEASE);
  } else {
    long spins = 0;
    while (__atomic_load_n(&s->sense, __ATOMIC_ACQUIRE) != c->local_sense)
      if (++spins > 2000) sched_yield();
  }
}

inline size_t dtype_bytes(int dt) { return dt == 8 ? 8 : (dt == 2 ? 4 : 1); }   // ncclFloat64_, ncclInt32_, ncclChar_

inline int GetUniqueId(void *id128) {
  memset(id128, 0, 128);
  unsigned long long a = (unsigned long long)getpid() * 2654435761ull ^ (unsigned long long)time(nullptr);
  int fd = open("/dev/urandom", O_RDONLY);
  if (fd >= 0) { unsigned long long r = 0; if (read(fd, &r, sizeof r) == (ssize_t)sizeof r) a ^= r; close(fd); }
  snprintf((char *)id128, 64, "/fb_ncclshim_%016llx", a);
  return 0;
}

inline int CommInitRank(Comm **out, int nranks, const char *id, int rank) {
  if (nranks > kMaxRanks) return 1;
  Comm *c = new Comm;
  c->nranks = nranks; c->rank = rank;
  strncpy(c->name, id, sizeof c->name - 1); c->name[sizeof c->name - 1] = 0;
  c->len = ((sizeof(Shared) + 4095) & ~(size_t)4095) + (size_t)nranks * kOutbox;
  int fd = shm_open(c->name, O_CREAT | O_RDWR, 0600);
  if (fd < 0) { delete c; return 2; }
  if (ftruncate(fd, (off_t)c->len) != 0) { close(fd); delete c; return 3; }   // every rank sets the same size; new pages read as zero
  void *p = mmap(nullptr, c->len, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) { delete c; return 4; }
  c->sh = (Shared *)p;
  __atomic_add_fetch(&c->sh->joined, 1, __ATOMIC_ACQ_REL);
  long spins = 0;
  while (__atomic_load_n(&c->sh->joined, __ATOMIC_ACQUIRE) < nranks)
    if (++spins > 2000) sched_yield();
  barrier(c);
  if (rank == 0) shm_unlink(c->name);   // everybody has it mapped
  *out = c;
  current() = c;
  return 0;
}

inline int CommDestroy(Comm *c) {
  if (!c) return 0;
  if (current() == c) current() = nullptr;
  munmap(c->sh, c->len);
  delete c;
  return 0;
}

// all ranks: contribute `bytes` from `send`, afterwards every outbox holds every rank's contribution
inline void publish_all(Comm *c, const void *send, size_t bytes) {
  if (bytes > kOutbox) { fprintf(stderr, "ncclshim: message of %zu bytes exceeds the outbox\n", bytes); abort(); }
  memcpy(c->outbox(c->rank), send, bytes);
  barrier(c);
}

inline int AllReduce(const void *send, void *recv, size_t count, int dtype, int /*op: sum*/, Comm *c, void * /*stream*/) {
  if (dtype != 8) return 5;   // the library reduces doubles only
  std::vector<double> mine(count);
  memcpy(mine.data(), send, count * 8);
  publish_all(c, mine.data(), count * 8);
  double *out = (double *)recv;
  for (size_t i = 0; i < count; i++) {
    double s = 0.0;
    for (int r = 0; r < c->nranks; r++) s += reinterpret_cast<const double *>(c->outbox(r))[i];
    out[i] = s;
  }
  barrier(c);
  return 0;
}

inline int AllGather(const void *send, void *recv, size_t count, int dtype, Comm *c, void *) {
  const size_t b = count * dtype_bytes(dtype);
  publish_all(c, send, b);
  for (int r = 0; r < c->nranks; r++) memcpy((char *)recv + (size_t)r * b, c->outbox(r), b);
  barrier(c);
  return 0;
}

inline int GroupStart() { current()->group_depth++; return 0; }
inline int flush_group(Comm *c) {
  Shared::PerRank &me = c->sh->r[c->rank];
  me.nmsg = 0; me.used = 0;
  for (const Comm::Op &o : c->ops)
    if (o.send) {
      if (me.nmsg >= 64 || me.used + o.bytes > kOutbox) { fprintf(stderr, "ncclshim: group too large\n"); abort(); }
      memcpy(c->outbox(c->rank) + me.used, o.buf, o.bytes);
      me.msg[me.nmsg].dst = o.peer; me.msg[me.nmsg].bytes = o.bytes; me.msg[me.nmsg].off = me.used;
      me.used = me.used + ((o.bytes + 15) & ~(size_t)15);
      me.nmsg = me.nmsg + 1;
    }
  barrier(c);
  int taken[kMaxRanks] = {0, 0, 0, 0, 0, 0, 0, 0};   // per source: how many of its messages to me were consumed
  for (const Comm::Op &o : c->ops)
    if (!o.send) {
      const Shared::PerRank &src = c->sh->r[o.peer];
      int seen = 0, found = -1;
      for (int m = 0; m < src.nmsg; m++)
        if (src.msg[m].dst == c->rank) { if (seen == taken[o.peer]) { found = m; break; } seen++; }
      if (found < 0 || src.msg[found].bytes != o.bytes) { fprintf(stderr, "ncclshim: unmatched recv (rank %d from %d)\n", c->rank, o.peer); abort(); }
      memcpy(o.buf, c->outbox(o.peer) + src.msg[found].off, o.bytes);
      taken[o.peer]++;
    }
  barrier(c);
  c->ops.clear();
  return 0;
}
inline int GroupEnd() {
  Comm *c = current();
  if (--c->group_depth > 0) return 0;
  return flush_group(c);
}
inline int Send(const void *buf, size_t count, int dtype, int peer, Comm *c, void *) {
  c->ops.push_back({true, const_cast<void *>(buf), count * dtype_bytes(dtype), peer});
  return c->group_depth ? 0 : flush_group(c);
}
inline int Recv(void *buf, size_t count, int dtype, int peer, Comm *c, void *) {
  c->ops.push_back({false, buf, count * dtype_bytes(dtype), peer});
  return c->group_depth ? 0 : flush_group(c);
}

}  // namespace ncclshim
}  // namespace fb
