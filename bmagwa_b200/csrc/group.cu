// group.cu -- several chains over ONE SNP-sharded store (BASELINE.json configs[4]: "sharded over 8xB200, 4 parallel
// chains").  The reference's chains are threads of one process sharing one Data (src/main.cpp:54-85);
// here a "shard group" is one process per GPU: rank r holds the packed SNPs [r stride, (r+1) stride) and -- for
// r < n_chains -- the host sampler of chain r.  Nothing of a chain is replicated on other ranks, and chains never wait
// for each other:
//
//   per iteration   chain c asks ITS GPU for the column statistics of the SNPs it proposes; columns of other shards are
//                   read over NVLink through CUDA-IPC peer mappings (store.cu).  No collective, no other rank involved.
//   per scan        (chain c, every n_rao iterations of ITS OWN clock)
//     1. chain c quantises its residual into the tensor-core scan's limb format (scan_imma.cu) inside its exchange
//        buffer, which every peer has mapped, and posts a request number in a POSIX shared-memory segment;
//     2. every rank runs a SCAN SERVICE thread with its own CUDA stream: it picks the request up, pulls the chain's limbs
//        over NVLink (n x 8 bytes), runs the scan kernel over ITS shard, and a small kernel adds the per-chunk partial sums
//        and stores the shard's dot products STRAIGHT INTO THE OWNING CHAIN'S GPU (peer stores, 8 bytes per SNP) --
//        compute and exchange in one pass, no NCCL call, no barrier and no Python on the data path -- then acknowledges;
//     3. with every rank's acknowledgement chain c finishes the scan on its own GPU (per-SNP algebra over all m_g SNPs
//        with its own tau draws) and carries on.  Integer accumulation makes the dot products independent of the
//        sharding, so every chain writes the bytes its single-GPU run writes (tests/test_gpu_sharded.py).
//
// A chain's scan therefore costs it one pass over 1/world of the store per rank, all ranks in parallel, whatever the other
// chains are doing; a slow chain (large model) delays nobody.  Ranks without a chain (n_chains < world, e.g. 4 chains over
// 8 GPUs) only run the service.  Collective operations (creation, the start-up exchanges, destruction) use a
// shared-memory barrier.
#include <fcntl.h>
#include <immintrin.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <unistd.h>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include "common.cuh"
#include "store.cuh"
#include "group.cuh"

namespace bmg {

namespace {
constexpr uint32_t kShmMagic = 0x424D4731u;   // "BMG1"
constexpr double kBarrierTimeout = 300.0;     // seconds; a missing peer must not hang the box

struct GroupShm {
  std::atomic<uint32_t> magic;
  std::atomic<uint32_t> attached;
  std::atomic<uint32_t> bar_count, bar_gen;
  std::atomic<uint32_t> failed;
  uint32_t pad[11];
  unsigned char handle[kGroupMaxRanks][64];
  int64_t lo[kGroupMaxRanks], hi[kGroupMaxRanks];
  // scan service: request[c] = number of the latest scan chain c asked for; done[c][r] = the latest rank r has served
  struct alignas(64) Flag { std::atomic<uint64_t> v; };
  Flag request[kGroupMaxRanks];
  Flag done[kGroupMaxRanks][kGroupMaxRanks];
  double request_time[kGroupMaxRanks];   // CLOCK_MONOTONIC of the latest request (diagnostics only)
};

double now_seconds()
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

__global__ void k_group_combine(const double* __restrict__ partial, int n_chunks, int64_t m, double* __restrict__ out)
{
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= m) return;
  double d = 0.0;   // the order k_scan_finalize adds the chunks in: same bits as the single-GPU scan
  for (int c = 0; c < n_chunks; ++c) d += partial[(int64_t)c * m + j];
  out[j] = d;       // `out` is the owning chain's buffer: a peer store over NVLink unless the chain is local
}
}  // namespace

struct Group {
  int world = 1, rank = 0, n_chains = 1;
  int64_t stride = 0;
  Store* store = nullptr;
  Chain* scan_chain = nullptr;   // geometry, partial sums and stream of this rank's share of every scan
  GroupShm* shm = nullptr;
  std::string shm_name;
  DevBuf<unsigned char> xbuf;    // | limbs of this rank's chain | exponent | dots of this rank's chain over all SNPs |
  size_t q_bytes = 0, off_exp = 0, off_dots = 0, xbuf_bytes = 0;
  unsigned char* peer[kGroupMaxRanks] = {nullptr};
  bool peer_opened[kGroupMaxRanks] = {false};
  DevBuf<unsigned char> q_stage[2];   // a peer chain's limbs + exponent, double-buffered
  DevBuf<int32_t> n1_all, n2_all;
  double barrier_seconds = 0.0;
  // scan service
  GroupShm local_shm;                  // world == 1: the flags live here
  std::thread service;
  std::atomic<bool> stop{false};
  std::atomic<int64_t> served{0};      // requests this rank's service has completed
  uint64_t served_seq[kGroupMaxRanks] = {0};
  std::string service_error;
  double pickup_seconds = 0.0, serve_seconds = 0.0, pickup_max = 0.0, serve_max = 0.0;   // request -> pick-up, pick-up -> acknowledged
  // this rank's chain
  uint64_t my_seq = 0;
  int64_t my_scans = 0;
  double scan_wait_seconds = 0.0;
};

static void group_fail(Group* g)
{
  if (g && g->shm) g->shm->failed.store(1u, std::memory_order_release);
}

void group_barrier(Group* g)
{
  if (g->world <= 1) return;
  GroupShm* s = g->shm;
  const double t0 = now_seconds();
  const uint32_t gen = s->bar_gen.load(std::memory_order_acquire);
  if (s->bar_count.fetch_add(1u, std::memory_order_acq_rel) + 1u == (uint32_t)g->world) {
    s->bar_count.store(0u, std::memory_order_relaxed);
    s->bar_gen.fetch_add(1u, std::memory_order_release);
  } else {
    unsigned long spins = 0;
    while (s->bar_gen.load(std::memory_order_acquire) == gen) {
      if (s->failed.load(std::memory_order_acquire)) throw Error("shard group: a peer rank failed");
      _mm_pause();
      if ((++spins & 0x3FF) == 0) {
        if (spins > 200000) usleep(20);   // a rank without a chain waits for a whole Rao-Blackwell period: leave the core
        if (now_seconds() - t0 > kBarrierTimeout) {
          group_fail(g);
          throw Error("shard group: barrier timed out (a peer rank is missing)");
        }
      }
    }
  }
  g->barrier_seconds += now_seconds() - t0;
}

// In-place all-gather over the group: every rank owns elements [rank per, (rank+1) per) of dev_buffer (world x per
// elements).  Same contract as the host-supplied bmg_allgather_fn, served natively through the peer mappings.
int group_allgather(void* ctx, void* dev_buffer, int64_t elems_per_rank, int elem_bytes, void* cuda_stream)
{
  Group* g = reinterpret_cast<Group*>(ctx);
  try {
    if (g->world <= 1) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
    const size_t per = (size_t)elems_per_rank * (size_t)elem_bytes;
    BMG_REQUIRE(per <= g->xbuf_bytes - g->off_dots, "shard group: all-gather block larger than the exchange buffer");
    unsigned char* buf = reinterpret_cast<unsigned char*>(dev_buffer);
    BMG_CUDA(cudaMemcpyAsync(g->xbuf.p + g->off_dots, buf + (size_t)g->rank * per, per, cudaMemcpyDeviceToDevice, st));
    BMG_CUDA(cudaStreamSynchronize(st));
    group_barrier(g);
    for (int i = 1; i < g->world; ++i) {
      const int r = (g->rank + i) % g->world;
      BMG_CUDA(cudaMemcpyAsync(buf + (size_t)r * per, g->peer[r] + g->off_dots, per, cudaMemcpyDefault, st));
    }
    BMG_CUDA(cudaStreamSynchronize(st));
    group_barrier(g);   // nobody overwrites its staging area before everyone has read it
    return 0;
  } catch (const std::exception& e) {
    group_fail(g);
    set_last_error(e.what());
    return 1;
  }
}

static void group_service_loop(Group* g);

Group* group_create(Store* s, int world, int rank, int n_chains, int64_t stride, const char* shm_name)
{
  BMG_REQUIRE(world >= 1 && world <= kGroupMaxRanks && rank >= 0 && rank < world, "shard group: invalid world / rank");
  BMG_REQUIRE(n_chains >= 1 && n_chains <= world, "shard group: 1 <= n_chains <= world (one chain per rank at most)");
  BMG_REQUIRE(stride > 0 && (int64_t)world * stride >= s->m_g, "shard group: world x stride does not cover m_g");
  BMG_REQUIRE(s->lo == std::min(s->m_g, (int64_t)rank * stride) && s->hi == std::min(s->m_g, (int64_t)(rank + 1) * stride),
              "shard group: the store's SNP range is not this rank's shard");
  BMG_REQUIRE(s->m_e >= 1, "shard group: call bmg_store_set_phenotype on the shard first");
  BMG_REQUIRE(world == 1 || (shm_name != nullptr && shm_name[0] == '/'), "shard group: a POSIX shared-memory name (\"/...\") is required");
  BMG_CUDA(cudaSetDevice(s->device));
  std::unique_ptr<Group> g(new Group());
  g->world = world; g->rank = rank; g->n_chains = n_chains; g->stride = stride; g->store = s;
  g->scan_chain = chain_create(s);
  imma_prepare(g->scan_chain);
  const int64_t n_pad = 16 * (int64_t)g->scan_chain->imma_chunks * g->scan_chain->imma_chunk_words;
  g->q_bytes = (size_t)n_pad * 8;
  g->off_exp = g->q_bytes;
  g->off_dots = g->q_bytes + 256;
  g->xbuf_bytes = g->off_dots + (size_t)world * (size_t)stride * sizeof(double);
  g->xbuf.alloc(g->xbuf_bytes);
  BMG_CUDA(cudaMemset(g->xbuf.p, 0, g->xbuf_bytes));
  g->q_stage[0].alloc(g->q_bytes + 256);
  g->q_stage[1].alloc(g->q_bytes + 256);
  BMG_CUDA(cudaDeviceSynchronize());   // the memset ran on the null stream; everything below uses non-blocking streams
  g->peer[rank] = g->xbuf.p;
  if (world > 1) {
    g->shm_name = shm_name;
    int fd = -1;
    if (rank == 0) {
      fd = shm_open(shm_name, O_CREAT | O_EXCL | O_RDWR, 0600);
      BMG_REQUIRE(fd >= 0, std::string("shard group: shm_open(create) failed for ") + shm_name);
      BMG_REQUIRE(ftruncate(fd, sizeof(GroupShm)) == 0, "shard group: ftruncate failed");
    } else {
      const double t0 = now_seconds();
      while ((fd = shm_open(shm_name, O_RDWR, 0600)) < 0) {
        BMG_REQUIRE(now_seconds() - t0 < 120.0, std::string("shard group: rank 0 never created ") + shm_name);
        usleep(1000);
      }
      struct stat sb;
      while (fstat(fd, &sb) == 0 && (size_t)sb.st_size < sizeof(GroupShm)) {
        BMG_REQUIRE(now_seconds() - t0 < 120.0, "shard group: shared segment never sized");
        usleep(1000);
      }
    }
    void* p = mmap(nullptr, sizeof(GroupShm), PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    BMG_REQUIRE(p != MAP_FAILED, "shard group: mmap failed");
    g->shm = reinterpret_cast<GroupShm*>(p);
    if (rank == 0) g->shm->magic.store(kShmMagic, std::memory_order_release);   // a fresh segment is zero-filled
    else {
      const double t0 = now_seconds();
      while (g->shm->magic.load(std::memory_order_acquire) != kShmMagic) {
        BMG_REQUIRE(now_seconds() - t0 < 120.0, "shard group: shared segment never initialised");
        usleep(100);
      }
    }
    cudaIpcMemHandle_t h;
    BMG_CUDA(cudaIpcGetMemHandle(&h, g->xbuf.p));
    std::memcpy(g->shm->handle[rank], &h, 64);
    g->shm->lo[rank] = s->lo; g->shm->hi[rank] = s->hi;
    g->shm->attached.fetch_add(1u, std::memory_order_acq_rel);
    group_barrier(g.get());
    if (rank == 0) shm_unlink(shm_name);   // everyone has it mapped: nothing is left behind whatever happens next
    for (int r = 0; r < world; ++r) {
      if (r == rank) continue;
      cudaIpcMemHandle_t hr;
      std::memcpy(&hr, g->shm->handle[r], 64);
      void* ptr = nullptr;
      BMG_CUDA(cudaIpcOpenMemHandle(&ptr, hr, cudaIpcMemLazyEnablePeerAccess));
      g->peer[r] = reinterpret_cast<unsigned char*>(ptr);
      g->peer_opened[r] = true;
    }
    group_barrier(g.get());
  }
  // genotype counts of every SNP on every rank: the per-SNP algebra of a chain's scan runs on the chain's own GPU
  const int64_t total = (int64_t)world * stride;
  g->n1_all.alloc(total); g->n2_all.alloc(total);
  cudaStream_t st = g->scan_chain->stream;
  BMG_CUDA(cudaMemsetAsync(g->n1_all.p, 0, total * sizeof(int32_t), st));
  BMG_CUDA(cudaMemsetAsync(g->n2_all.p, 0, total * sizeof(int32_t), st));
  BMG_CUDA(cudaMemcpyAsync(g->n1_all.p + (int64_t)rank * stride, s->n1.p, s->m * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  BMG_CUDA(cudaMemcpyAsync(g->n2_all.p + (int64_t)rank * stride, s->n2.p, s->m * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  BMG_REQUIRE(group_allgather(g.get(), g->n1_all.p, stride, (int)sizeof(int32_t), (void*)st) == 0, "shard group: exchange of the genotype counts failed");
  BMG_REQUIRE(group_allgather(g.get(), g->n2_all.p, stride, (int)sizeof(int32_t), (void*)st) == 0, "shard group: exchange of the genotype counts failed");
  BMG_CUDA(cudaStreamSynchronize(st));
  if (g->shm == nullptr) {   // one rank: no segment to share
    std::memset(static_cast<void*>(&g->local_shm), 0, sizeof(GroupShm));
    g->shm = &g->local_shm;
  }
  Group* raw = g.get();
  g->service = std::thread([raw] { group_service_loop(raw); });
  return g.release();
}

// collective: every rank's service must stay up until every chain of the group has ended
void group_destroy(Group* g)
{
  if (!g) return;
  cudaSetDevice(g->store->device);
  try { group_barrier(g); } catch (...) {}
  g->stop.store(true, std::memory_order_release);
  if (g->service.joinable()) g->service.join();
  if (getenv("BMG_TIMING") && g->served.load() > 0)
    fprintf(stderr, "[bmg timing] shard group rank %d: scan service served %lld requests; request -> pick-up mean %.0f us (max %.0f), pick-up -> acknowledged "
                    "mean %.0f us (max %.0f); own chain waited %.3f s for %lld scans\n", g->rank, (long long)g->served.load(),
            1e6 * g->pickup_seconds / (double)g->served.load(), 1e6 * g->pickup_max, 1e6 * g->serve_seconds / (double)g->served.load(),
            1e6 * g->serve_max, g->scan_wait_seconds, (long long)g->my_scans);
  if (g->scan_chain) { cudaStreamSynchronize(g->scan_chain->stream); }
  for (int r = 0; r < g->world; ++r)
    if (g->peer_opened[r]) cudaIpcCloseMemHandle(g->peer[r]);
  chain_destroy(g->scan_chain);
  if (g->shm && g->shm != &g->local_shm) munmap(g->shm, sizeof(GroupShm));
  delete g;
}

int group_world(const Group* g) { return g->world; }
int group_rank(const Group* g) { return g->rank; }
int group_chains(const Group* g) { return g->n_chains; }
int64_t group_stride(const Group* g) { return g->stride; }
const int32_t* group_n1(const Group* g) { return g->n1_all.p; }
const int32_t* group_n2(const Group* g) { return g->n2_all.p; }
Chain* group_scan_chain(Group* g) { return g->scan_chain; }
void group_stats(const Group* g, double* out4)
{
  out4[0] = (double)g->served.load(std::memory_order_acquire); out4[1] = g->scan_wait_seconds; out4[2] = (double)g->my_scans;
  out4[3] = g->barrier_seconds;
}

// ---- the scan service of this rank: serves every chain's requests over this rank's shard ----------------------------
static void group_serve_one(Group* g, int c)
{
  Store* s = g->store;
  Chain* sc = g->scan_chain;
  cudaStream_t st = sc->stream;
  const unsigned char* q = g->xbuf.p;
  if (c != g->rank) {
    BMG_CUDA(cudaMemcpyAsync(g->q_stage[0].p, g->peer[c], g->q_bytes + 256, cudaMemcpyDefault, st));
    q = g->q_stage[0].p;
  }
  imma_launch_on(sc, reinterpret_cast<const uint4*>(q), reinterpret_cast<const int*>(q + g->q_bytes), sc->imma_partial.p, false, st, sc);
  double* out = reinterpret_cast<double*>(g->peer[c] + g->off_dots) + s->lo;
  k_group_combine<<<(unsigned)((s->m + 255) / 256), 256, 0, st>>>(sc->imma_partial.p, sc->imma_chunks, s->m, out);
  count_launch();
  BMG_CUDA(cudaGetLastError());
  BMG_CUDA(cudaStreamSynchronize(st));   // the peer stores have landed before the acknowledgement is visible
}

static void group_service_loop(Group* g)
{
  try {
    BMG_CUDA(cudaSetDevice(g->store->device));
    GroupShm* shm = g->shm;
    unsigned idle = 0;
    while (!g->stop.load(std::memory_order_acquire)) {
      bool found = false;
      for (int i = 0; i < g->n_chains; ++i) {
        const int c = (g->rank + i) % g->n_chains;   // the local chain first
        const uint64_t req = shm->request[c].v.load(std::memory_order_acquire);
        if (req == g->served_seq[c]) continue;
        const double t_pick = now_seconds();
        const double waited = t_pick - shm->request_time[c];
        group_serve_one(g, c);
        g->served_seq[c] = req;
        shm->done[c][g->rank].v.store(req, std::memory_order_release);
        const double took = now_seconds() - t_pick;
        g->pickup_seconds += waited; g->serve_seconds += took;
        if (waited > g->pickup_max) g->pickup_max = waited;
        if (took > g->serve_max) g->serve_max = took;
        g->served.fetch_add(1, std::memory_order_acq_rel);
        found = true;
      }
      if (found) { idle = 0; continue; }
      if (shm->failed.load(std::memory_order_acquire)) break;
      if (++idle < 2000) _mm_pause();
      else usleep(20);   // a request comes once per Rao-Blackwell period and chain: do not burn a core on the wait
    }
  } catch (const std::exception& e) {
    g->service_error = e.what();
    group_fail(g);
  }
}

// Chain `mine` (residual ready) asks every rank for its scan and waits for the dot products over all m_g SNPs
// (device pointer on this GPU, complete when the call returns).
const double* group_scan_round(Group* g, Chain* mine)
{
  Store* s = g->store;
  try {
    BMG_CUDA(cudaSetDevice(s->device));
    BMG_REQUIRE(mine != nullptr && g->rank < g->n_chains, "shard group: only ranks below n_chains hold a chain");
    BMG_REQUIRE(mine->residual_valid, "scan: call bmg_chain_residual first");
    cudaStream_t st = mine->stream;
    imma_quantize(mine);
    BMG_REQUIRE(mine->imma_q.n == g->q_bytes, "shard group: limb layout of the chain differs from the group's");
    BMG_CUDA(cudaMemcpyAsync(g->xbuf.p, mine->imma_q.p, g->q_bytes, cudaMemcpyDeviceToDevice, st));
    BMG_CUDA(cudaMemcpyAsync(g->xbuf.p + g->off_exp, mine->imma_exp.p, sizeof(int), cudaMemcpyDeviceToDevice, st));
    BMG_CUDA(cudaStreamSynchronize(st));   // also: the previous scan's per-SNP algebra has finished reading the dot products
    GroupShm* shm = g->shm;
    const uint64_t seq = ++g->my_seq;
    const double t0 = now_seconds();
    shm->request_time[g->rank] = t0;
    shm->request[g->rank].v.store(seq, std::memory_order_release);
    unsigned long spins = 0;
    for (int r = 0; r < g->world; ++r) {
      while (shm->done[g->rank][r].v.load(std::memory_order_acquire) != seq) {
        _mm_pause();
        if ((++spins & 0xFFF) == 0) {
          if (shm->failed.load(std::memory_order_acquire))
            throw Error("shard group: a rank's scan service failed" + (g->service_error.empty() ? std::string() : ": " + g->service_error));
          if (now_seconds() - t0 > kBarrierTimeout) throw Error("shard group: a rank's scan service does not answer");
        }
      }
    }
    g->scan_wait_seconds += now_seconds() - t0;
    ++g->my_scans;
    return reinterpret_cast<const double*>(g->xbuf.p + g->off_dots);
  } catch (...) {
    group_fail(g);
    throw;
  }
}

// a rank without a chain: returns once its service has completed n_rounds more requests of every chain
void group_serve(Group* g, int64_t n_rounds)
{
  const int64_t target = g->served.load(std::memory_order_acquire) + n_rounds * g->n_chains;
  const double t0 = now_seconds();
  while (g->served.load(std::memory_order_acquire) < target) {
    if (g->shm->failed.load(std::memory_order_acquire)) throw Error("shard group: a peer rank failed" + (g->service_error.empty() ? std::string() : ": " + g->service_error));
    if (now_seconds() - t0 > 20 * kBarrierTimeout) throw Error("shard group: no scan request arrives");
    usleep(200);
  }
}

}  // namespace bmg
