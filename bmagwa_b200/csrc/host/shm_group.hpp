// shm_group.hpp -- the POSIX shared-memory segment the ranks of a shard group (group.cu) meet in, and its barrier.
// Plain host code (no CUDA), so that the multi-process logic is unit-tested on the CPU (tests/test_cpu_shm_group.py:
// two real processes).  The ranks of a group live on one box: one process per GPU.
#pragma once
#include <fcntl.h>
#include <immintrin.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>
#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

namespace bmg {

constexpr int kShmMaxRanks = 16;

struct GroupShm {
  std::atomic<uint32_t> magic;
  std::atomic<uint32_t> attached;
  std::atomic<uint32_t> bar_count, bar_gen;
  std::atomic<uint32_t> failed;
  uint32_t pad[11];
  unsigned char handle[kShmMaxRanks][64];   // one CUDA IPC handle per rank (the exchange buffer)
  int64_t lo[kShmMaxRanks], hi[kShmMaxRanks];
  int64_t scratch[kShmMaxRanks];            // free for the ranks' use between barriers (tests)
};

inline double shm_now_seconds()
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

inline double shm_barrier_timeout()
{
  static const double t = getenv("BMG_GROUP_TIMEOUT") ? atof(getenv("BMG_GROUP_TIMEOUT")) : 120.0;   // seconds: a missing peer must not hang the box
  return t;
}

// Maps the segment `name` ("/..."): rank 0 creates it (exclusive: the name is unique per job) and marks it initialised,
// the others wait for it.  Throws std::runtime_error.
inline GroupShm* shm_group_open(const char* name, int rank)
{
  constexpr uint32_t kMagic = 0x424D4731u;   // "BMG1"
  int fd = -1;
  const double t0 = shm_now_seconds();
  if (rank == 0) {
    fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
    if (fd < 0) throw std::runtime_error(std::string("shard group: shm_open(create) failed for ") + name);
    if (ftruncate(fd, sizeof(GroupShm)) != 0) { close(fd); throw std::runtime_error("shard group: ftruncate failed"); }
  } else {
    while ((fd = shm_open(name, O_RDWR, 0600)) < 0) {
      if (shm_now_seconds() - t0 > shm_barrier_timeout()) throw std::runtime_error(std::string("shard group: rank 0 never created ") + name);
      usleep(1000);
    }
    struct stat sb;
    while (fstat(fd, &sb) == 0 && (size_t)sb.st_size < sizeof(GroupShm)) {
      if (shm_now_seconds() - t0 > shm_barrier_timeout()) { close(fd); throw std::runtime_error("shard group: shared segment never sized"); }
      usleep(1000);
    }
  }
  void* p = mmap(nullptr, sizeof(GroupShm), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) throw std::runtime_error("shard group: mmap failed");
  GroupShm* shm = reinterpret_cast<GroupShm*>(p);
  if (rank == 0) shm->magic.store(kMagic, std::memory_order_release);   // a fresh segment is zero-filled
  else
    while (shm->magic.load(std::memory_order_acquire) != kMagic) {
      if (shm_now_seconds() - t0 > shm_barrier_timeout()) throw std::runtime_error("shard group: shared segment never initialised");
      usleep(100);
    }
  return shm;
}

inline void shm_group_close(GroupShm* shm) { if (shm) munmap(shm, sizeof(GroupShm)); }

// Sense-reversing barrier over `world` processes.  Leaves with an exception when a peer has failed (shm->failed) or does
// not arrive within the time-out (and then marks the group failed, so that the others leave as well).
inline void shm_group_barrier(GroupShm* s, int world, double* waited_seconds = nullptr)
{
  if (world <= 1) return;
  const double t0 = shm_now_seconds();
  const uint32_t gen = s->bar_gen.load(std::memory_order_acquire);
  if (s->bar_count.fetch_add(1u, std::memory_order_acq_rel) + 1u == (uint32_t)world) {
    s->bar_count.store(0u, std::memory_order_relaxed);
    s->bar_gen.fetch_add(1u, std::memory_order_release);
  } else {
    unsigned long spins = 0;
    while (s->bar_gen.load(std::memory_order_acquire) == gen) {
      if (s->failed.load(std::memory_order_acquire)) throw std::runtime_error("shard group: a peer rank failed");
      _mm_pause();
      if ((++spins & 0x3FF) == 0) {
        if (spins > 200000) usleep(20);   // a rank without a chain waits for a whole Rao-Blackwell period: leave the core
        if (shm_now_seconds() - t0 > shm_barrier_timeout()) {
          s->failed.store(1u, std::memory_order_release);
          throw std::runtime_error("shard group: barrier timed out (a peer rank is missing)");
        }
      }
    }
  }
  if (waited_seconds) *waited_seconds += shm_now_seconds() - t0;
}

}  // namespace bmg
