// shm_group_harness.cpp -- C window onto csrc/host/shm_group.hpp for tests/test_cpu_shm_group.py (two real processes).
#include "../../bmagwa_b200/csrc/host/shm_group.hpp"

using namespace bmg;

// Every rank: `rounds` times { scratch[rank] = round; barrier; check that every rank's scratch holds `round`; barrier }.
// Returns 0, or 1 + the round at which a peer's value was stale, or -1 on an exception (message in err).
extern "C" long harness_shm_rounds(const char* name, int rank, int world, long rounds, int unlink_after_attach, char* err, int err_len)
{
  try {
    GroupShm* shm = shm_group_open(name, rank);
    shm->attached.fetch_add(1u);
    shm_group_barrier(shm, world);
    if (rank == 0 && unlink_after_attach) shm_unlink(name);
    long bad = 0;
    for (long r = 1; r <= rounds && !bad; ++r) {
      shm->scratch[rank] = r;
      shm_group_barrier(shm, world);
      for (int q = 0; q < world; ++q)
        if (shm->scratch[q] != r) bad = 1 + r;
      shm_group_barrier(shm, world);
    }
    const unsigned attached = shm->attached.load();
    shm_group_close(shm);
    return bad ? bad : (attached == (unsigned)world ? 0 : -2);
  } catch (const std::exception& e) {
    snprintf(err, (size_t)err_len, "%s", e.what());
    return -1;
  }
}

// a rank that never shows up: the others must leave the barrier with the time-out error, not hang
extern "C" long harness_shm_missing_peer(const char* name, int world, char* err, int err_len)
{
  try {
    GroupShm* shm = shm_group_open(name, 0);
    shm_unlink(name);
    shm_group_barrier(shm, world);
    shm_group_close(shm);
    return 0;
  } catch (const std::exception& e) {
    snprintf(err, (size_t)err_len, "%s", e.what());
    return -1;
  }
}
