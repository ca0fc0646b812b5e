// group.cu -- several chains over ONE SNP-sharded store (BASELINE.json configs[4]: "sharded over 8xB200, 4 parallel
// chains").  The reference's chains are threads of one process sharing one Data (src/main.cpp:54-85);
// here a "shard group" is one process per GPU: rank r holds the packed SNPs [r stride, (r+1) stride) and -- for
// r < n_chains -- the host sampler of chain r.  Nothing of a chain is replicated on other ranks:
//
//   per iteration   chain c asks ITS GPU for the column statistics of the SNPs it proposes; columns of other shards are
//                   read over NVLink through CUDA-IPC peer mappings (store.cu).  No collective, no other rank involved.
//   per scan        (every n_rao iterations; all chains of a group run the same schedule)
//     1. chain c quantises its residual into the tensor-core scan's limb format (scan_imma.cu) inside its exchange
//        buffer, which every peer has mapped;
//     2. barrier (host, POSIX shared memory: the ranks of a group live on one box);
//     3. every rank scans ITS shard once per chain: the chain's limbs are pulled over NVLink (n x 8 bytes, peer loads), the
//        scan kernel runs, and a small kernel adds the per-chunk partial sums into this rank's results area;
//     4. barrier; chain c pulls its results from every rank's results area (one kernel, peer loads, 8 bytes per SNP) and
//        finishes the scan on its own GPU (per-SNP algebra over all m_g SNPs with its own tau draws).
//   No NCCL call and no Python on the data path.  Integer accumulation makes the dot products independent of the
//   sharding, so every chain writes the bytes its single-GPU run writes (tests/test_gpu_sharded.py).
//
// Ranks without a chain (n_chains < world, e.g. 4 chains over 8 GPUs) only take part in steps 2-4 (bmg_group_serve).
//
// Why the scans are collective.  A variant in which a per-rank service thread scanned for whichever chain asked, while the
// other chains kept running (no barrier), was built and measured on 2 x B200: chains then no longer wait for each other,
// but about one scan in ten delivered wrong dot products for a handful of SNPs (rows 0-7 of a few 16-SNP tiles) -- only
// across two real GPUs, only for scans running while other chains' kernels were reading this GPU's memory over NVLink,
// with bit-identical inputs and a scan kernel that is bit-reproducible under every single-GPU concurrency test
// (tools/scan_under_server.py, tools/scan_beside_chain.py, compute-sanitizer racecheck).  Unexplained, so not shipped:
// with the barriers every scan runs while all chains are paused, and chains with the same seed stay byte-identical on any
// number of GPUs (tools/group_same_seed_check.py).  profiles/round2_notes.md has the measurements.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "common.cuh"
#include "store.cuh"
#include "group.cuh"
#include "host/shm_group.hpp"

namespace bmg {

namespace {
inline double now_seconds() { return shm_now_seconds(); }
static_assert(kShmMaxRanks == kGroupMaxRanks, "rank limits of the segment and of the group");

__global__ void k_group_combine(const double* __restrict__ partial, int n_chunks, int64_t m, double* __restrict__ out)
{
  const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (j >= m) return;
  double d = 0.0;   // the order k_scan_finalize adds the chunks in: same bits as the single-GPU scan
  for (int c = 0; c < n_chunks; ++c) d += partial[(int64_t)c * m + j];
  out[j] = d;
}

// Data crosses NVLink by peer LOADS only: the consumer pulls what a finished kernel of the producer left in the producer's
// own memory.  (Peer STORES followed by a stream synchronisation and a host-side acknowledgement were tried first: on two
// B200s a handful of the stored values were seen by the consumer one scan late, even with a system-scope fence at the end
// of the storing kernel.)
__global__ void k_group_pull_bytes(const uint4* __restrict__ src, uint4* __restrict__ dst, int64_t n16)
{
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) dst[i] = __ldcv(src + i);
}
struct GatherArgs {
  const double* src[kGroupMaxRanks];   // rank r's results for this chain over its shard (peer memory)
  int64_t count[kGroupMaxRanks];
  int64_t stride;
  int world;
  double* dst;                         // the chain's dot products over all SNPs (local)
};
__global__ void k_group_gather(const __grid_constant__ GatherArgs a)
{
  const int r = blockIdx.y;
  const double* src = a.src[r];
  double* dst = a.dst + (int64_t)r * a.stride;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.count[r]; i += (int64_t)gridDim.x * blockDim.x) dst[i] = __ldcv(src + i);
}
}  // namespace

struct Group {
  int world = 1, rank = 0, n_chains = 1;
  int64_t stride = 0;
  Store* store = nullptr;
  Chain* scan_chain = nullptr;   // geometry, partial sums and stream of this rank's share of every scan
  GroupShm* shm = nullptr;
  std::string shm_name;
  DevBuf<unsigned char> xbuf;    // | limbs of this rank's chain | exponent | this rank's results per chain: n_chains x stride doubles |
  size_t q_bytes = 0, off_exp = 0, off_dots = 0, xbuf_bytes = 0;   // (the results area doubles as the all-gather staging)
  DevBuf<double> dots;           // this rank's chain: dot products over all SNPs, gathered from every rank's results area
  unsigned char* peer[kGroupMaxRanks] = {nullptr};
  bool peer_opened[kGroupMaxRanks] = {false};
  DevBuf<unsigned char> q_stage[2];   // peer chains' limbs + exponent: the two residuals of one pass
  DevBuf<double> partial2;            // per-chunk partial sums of the second residual of a pass
  int64_t pair_launches = 0;          // passes over the shard that served two chains
  std::unique_ptr<GlobalMissing> gm;   // genotype counts and missing-call index of all SNPs (every rank holds them)
  double barrier_seconds = 0.0;
  GroupShm local_shm;                  // world == 1: the flags live here
  int64_t rounds = 0;                  // scan rounds this rank took part in
  int64_t my_scans = 0;
  double scan_wait_seconds = 0.0;      // this rank's chain: seconds between asking for a scan and having its results
};

static void group_fail(Group* g)
{
  if (g && g->shm) g->shm->failed.store(1u, std::memory_order_release);
}

void group_barrier(Group* g)
{
  if (g->world <= 1) return;
  try {
    shm_group_barrier(g->shm, g->world, &g->barrier_seconds);
  } catch (const std::exception& e) {
    throw Error(e.what());
  }
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
    const size_t cap = (g->xbuf_bytes - g->off_dots) & ~(size_t)15;   // the staging area: larger blocks travel in pieces
    unsigned char* buf = reinterpret_cast<unsigned char*>(dev_buffer);
    for (size_t o = 0; o < per; o += cap) {
      const size_t len = std::min(cap, per - o);
      BMG_CUDA(cudaMemcpyAsync(g->xbuf.p + g->off_dots, buf + (size_t)g->rank * per + o, len, cudaMemcpyDeviceToDevice, st));
      BMG_CUDA(cudaStreamSynchronize(st));
      group_barrier(g);
      for (int i = 1; i < g->world; ++i) {
        const int r = (g->rank + i) % g->world;
        BMG_CUDA(cudaMemcpyAsync(buf + (size_t)r * per + o, g->peer[r] + g->off_dots, len, cudaMemcpyDefault, st));
      }
      BMG_CUDA(cudaStreamSynchronize(st));
      group_barrier(g);   // nobody overwrites its staging area before everyone has read it
    }
    return 0;
  } catch (const std::exception& e) {
    group_fail(g);
    set_last_error(e.what());
    return 1;
  }
}

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
  g->xbuf_bytes = g->off_dots + (size_t)std::max(world, n_chains) * (size_t)stride * sizeof(double);
  g->dots.alloc((size_t)world * (size_t)stride);
  g->xbuf.alloc(g->xbuf_bytes);
  BMG_CUDA(cudaMemset(g->xbuf.p, 0, g->xbuf_bytes));
  g->q_stage[0].alloc(g->q_bytes + 256);
  g->q_stage[1].alloc(g->q_bytes + 256);
  if (g->scan_chain->imma_slices2 > 0 && n_chains > 1) g->partial2.alloc(std::max<size_t>(1, (size_t)g->scan_chain->imma_chunks * (size_t)s->m));
  BMG_CUDA(cudaDeviceSynchronize());   // the memset ran on the null stream; everything below uses non-blocking streams
  g->peer[rank] = g->xbuf.p;
  if (world > 1) {
    g->shm_name = shm_name;
    try {
      g->shm = shm_group_open(shm_name, rank);
    } catch (const std::exception& e) {
      throw Error(e.what());
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
  // genotype counts (and the missing-call index) of every SNP on every rank: the per-SNP algebra of a chain's scan runs on
  // the chain's own GPU, and the chain's imputed values cover all SNPs
  if (g->shm == nullptr) {   // one rank: no segment to share
    std::memset(static_cast<void*>(&g->local_shm), 0, sizeof(GroupShm));
    g->shm = &g->local_shm;
  }
  g->gm.reset(build_global_missing(s, world, rank, stride, group_allgather, g.get()));
  return g.release();
}

// collective: the peer mappings of this rank's exchange buffer must outlive every other rank's last scan
void group_destroy(Group* g)
{
  if (!g) return;
  cudaSetDevice(g->store->device);
  try { group_barrier(g); } catch (...) {}
  if (g->scan_chain) { cudaStreamSynchronize(g->scan_chain->stream); }
  for (int r = 0; r < g->world; ++r)
    if (g->peer_opened[r]) cudaIpcCloseMemHandle(g->peer[r]);
  chain_destroy(g->scan_chain);
  if (g->shm && g->shm != &g->local_shm) shm_group_close(g->shm);
  delete g;
}

int group_world(const Group* g) { return g->world; }
int group_rank(const Group* g) { return g->rank; }
int group_chains(const Group* g) { return g->n_chains; }
int64_t group_stride(const Group* g) { return g->stride; }
const int32_t* group_n1(const Group* g) { return g->gm->n1.p; }
const int32_t* group_n2(const Group* g) { return g->gm->n2.p; }
const GlobalMissing* group_missing(const Group* g) { return g->gm.get(); }
Chain* group_scan_chain(Group* g) { return g->scan_chain; }
void group_stats(const Group* g, double* out4)
{
  out4[0] = (double)g->rounds; out4[1] = g->scan_wait_seconds; out4[2] = (double)g->my_scans; out4[3] = g->barrier_seconds;
}

// every rank's results for this rank's chain -> the chain's full-length array (one launch, peer loads)
static void group_gather_dots(Group* g, cudaStream_t st)
{
  GatherArgs a;
  int64_t most = 0;
  for (int r = 0; r < g->world; ++r) {
    a.src[r] = reinterpret_cast<const double*>(g->peer[r] + g->off_dots) + (int64_t)g->rank * g->stride;
    a.count[r] = std::max<int64_t>(0, std::min<int64_t>(g->stride, g->store->m_g - (int64_t)r * g->stride));
    most = std::max(most, a.count[r]);
  }
  a.stride = g->stride; a.world = g->world; a.dst = g->dots.p;
  const unsigned bx = (unsigned)std::max<int64_t>(1, std::min<int64_t>(64, (most + 1023) / 1024));
  k_group_gather<<<dim3(bx, (unsigned)g->world), 256, 0, st>>>(a);
  count_launch();
  BMG_CUDA(cudaGetLastError());
}

// One scan round (see the header of this file); collective.  mine: this rank's chain with its residual ready, or nullptr
// on a rank without a chain.  Returns the chain's dot products over all m_g SNPs (device pointer on this GPU; the gather
// that fills it is queued on the chain's stream, so work queued behind it sees them).
const double* group_scan_round(Group* g, Chain* mine)
{
  Store* s = g->store;
  try {
    BMG_CUDA(cudaSetDevice(s->device));
    BMG_REQUIRE((mine != nullptr) == (g->rank < g->n_chains), "shard group: ranks below n_chains scan through their chain, the others through bmg_group_serve");
    Chain* sc = g->scan_chain;
    cudaStream_t st = mine ? mine->stream : sc->stream;
    const double t0 = now_seconds();
    if (mine) {
      BMG_REQUIRE(mine->residual_valid, "scan: call bmg_chain_residual first");
      imma_quantize(mine);
      BMG_REQUIRE(mine->imma_q.n == g->q_bytes, "shard group: limb layout of the chain differs from the group's");
      BMG_CUDA(cudaMemcpyAsync(g->xbuf.p, mine->imma_q.p, g->q_bytes, cudaMemcpyDeviceToDevice, st));
      BMG_CUDA(cudaMemcpyAsync(g->xbuf.p + g->off_exp, mine->imma_exp.p, sizeof(int), cudaMemcpyDeviceToDevice, st));
      BMG_CUDA(cudaStreamSynchronize(st));   // also: the previous round's gather has finished reading the peers' results
    }
    group_barrier(g);   // every chain's limbs are in place, every chain is paused
    // the chains two at a time: one pass over the shard serves both residuals (k_scan_dots_imma2); a last odd chain, or a
    // geometry without room for the two-residual kernel, takes the one-residual kernel
    const bool pairs = sc->imma_slices2 > 0;
    // development (profiling the two-residual kernel on one GPU): a one-chain group scans its residual as both halves of a pair
    static const bool self_pair = getenv("BMG_GROUP_SELF_PAIR") != nullptr;
    if (self_pair && pairs && g->n_chains == 1 && g->partial2.n == 0) {
      BMG_CUDA(cudaStreamSynchronize(st));
      g->partial2.alloc(std::max<size_t>(1, (size_t)sc->imma_chunks * (size_t)s->m));
    }
    for (int i = 0; i < g->n_chains; i += pairs ? 2 : 1) {
      int n_here = pairs && i + 1 < g->n_chains ? 2 : 1;
      if (self_pair && pairs && g->n_chains == 1) {
        const unsigned char* q0 = g->xbuf.p;
        imma_launch2_on(sc, reinterpret_cast<const uint4*>(q0), reinterpret_cast<const int*>(q0 + g->q_bytes), sc->imma_partial.p,
                        reinterpret_cast<const uint4*>(q0), reinterpret_cast<const int*>(q0 + g->q_bytes), g->partial2.p, st, sc);
        k_group_combine<<<(unsigned)((s->m + 255) / 256), 256, 0, st>>>(sc->imma_partial.p, sc->imma_chunks, s->m,
                                                                        reinterpret_cast<double*>(g->xbuf.p + g->off_dots) + (int64_t)g->rank * g->stride);
        count_launch();
        break;
      }
      const unsigned char* q[2] = {nullptr, nullptr};
      double* out[2] = {nullptr, nullptr};
      for (int k = 0; k < n_here; ++k) {
        const int c = (g->rank + i + k) % g->n_chains;   // start with the nearest chain: the pulls spread over the peers
        q[k] = g->xbuf.p;
        if (c != g->rank) {   // the chain's limbs + exponent, pulled over NVLink
          const int64_t n16 = (int64_t)((g->q_bytes + 256) / 16);
          k_group_pull_bytes<<<(unsigned)std::min<int64_t>(296, (n16 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const uint4*>(g->peer[c]),
                                                                                               reinterpret_cast<uint4*>(g->q_stage[k].p), n16);
          count_launch();
          q[k] = g->q_stage[k].p;
        }
        out[k] = reinterpret_cast<double*>(g->xbuf.p + g->off_dots) + (int64_t)c * g->stride;   // this rank's results for chain c
      }
      if (n_here == 2) {
        const bool ok = imma_launch2_on(sc, reinterpret_cast<const uint4*>(q[0]), reinterpret_cast<const int*>(q[0] + g->q_bytes), sc->imma_partial.p,
                                        reinterpret_cast<const uint4*>(q[1]), reinterpret_cast<const int*>(q[1] + g->q_bytes), g->partial2.p, st, sc);
        BMG_REQUIRE(ok, "shard group: the two-residual scan kernel is not available");
        ++g->pair_launches;
      } else {
        imma_launch_on(sc, reinterpret_cast<const uint4*>(q[0]), reinterpret_cast<const int*>(q[0] + g->q_bytes), sc->imma_partial.p, false, st, sc);
      }
      for (int k = 0; k < n_here; ++k) {
        k_group_combine<<<(unsigned)((s->m + 255) / 256), 256, 0, st>>>(k == 0 ? sc->imma_partial.p : g->partial2.p, sc->imma_chunks, s->m, out[k]);
        count_launch();
      }
    }
    BMG_CUDA(cudaGetLastError());
    BMG_CUDA(cudaStreamSynchronize(st));
    group_barrier(g);   // every rank's results for every chain are in that rank's memory
    ++g->rounds;
    if (mine == nullptr) return nullptr;
    group_gather_dots(g, st);
    g->scan_wait_seconds += now_seconds() - t0;
    ++g->my_scans;
    return g->dots.p;
  } catch (...) {
    group_fail(g);
    throw;
  }
}

// a rank without a chain: takes part in the next n_rounds scans of the group
void group_serve(Group* g, int64_t n_rounds)
{
  for (int64_t i = 0; i < n_rounds; ++i) group_scan_round(g, nullptr);
}

}  // namespace bmg
