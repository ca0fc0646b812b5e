// sampler_capi.cpp -- extern "C" entry points of the host sampler (include/bmagwa_b200.h, bmg_sampler_*).
//
// bmg_sampler_create does what main() does for one chain (src/main.cpp:47-76): parse the INI file,
// load .fam/.y/.e/.bed, build the device store (recode, counts, moment cache), construct the sampler.
#include <time.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <cmath>
#include <memory>
#include "../common.cuh"
#include "../store.cuh"
#include "../group.cuh"
#include "../../../include/bmagwa_b200.h"
#include "dataset.hpp"
#include "options.hpp"
#include "sampler.hpp"

using namespace bmg;

#define BMG_API extern "C" __attribute__((visibility("default")))

namespace {
// a shard group with the data-set summaries every chain of it needs (computed collectively at creation, because ranks
// without a chain never construct a sampler)
struct GroupHandle {
  Group* g = nullptr;
  double summaries[6] = {0, 0, 0, 0, 0, 0};
  int64_t missing_cells = 0;
  ~GroupHandle() { group_destroy(g); }
};

struct SamplerHandle {
  std::unique_ptr<Options> opt;
  std::unique_ptr<Dataset> data;
  Store* store = nullptr;
  bool owns_store = false;
  std::unique_ptr<Sampler> sampler;
  int chain_index = 0;
  bool flat_done = false;
  ~SamplerHandle()
  {
    sampler.reset();
    if (owns_store && store) {
      cudaSetDevice(store->device);
      delete store;
    }
  }
};

SamplerHandle* H(bmg_sampler* sp)
{
  BMG_REQUIRE(sp != nullptr, "null sampler handle");
  return reinterpret_cast<SamplerHandle*>(sp);
}

// var_x / mean_x of a SNP-sharded data set exactly as the unsharded store computes them (store.cu, as
// Data::compute_g_var_and_mean): per-SNP means, variances and missing counts are all-gathered and summed on the
// host in SNP order.
static int64_t global_summaries(Store* st, const Sampler::ShardComm& cm, double* out4)
{
  BMG_CUDA(cudaSetDevice(st->device));
  const int64_t total = (int64_t)cm.world * cm.stride, off = (int64_t)cm.rank * cm.stride;
  DevBuf<double> buf; buf.alloc(total);
  DevBuf<int32_t> ibuf; ibuf.alloc(total);
  cudaStream_t stm;
  BMG_CUDA(cudaStreamCreateWithFlags(&stm, cudaStreamNonBlocking));
  std::vector<double> hm(total), hv(total);
  std::vector<int32_t> hn(total);
  auto gather_d = [&](const double* local, std::vector<double>& host) {
    BMG_CUDA(cudaMemsetAsync(buf.p, 0, total * sizeof(double), stm));
    BMG_CUDA(cudaMemcpyAsync(buf.p + off, local, st->m * sizeof(double), cudaMemcpyDeviceToDevice, stm));
    if (cm.allgather(cm.ctx, buf.p, cm.stride, (int)sizeof(double), (void*)stm) != 0) throw Error("sharded sampler: all-gather callback failed");
    BMG_CUDA(cudaMemcpyAsync(host.data(), buf.p, total * sizeof(double), cudaMemcpyDeviceToHost, stm));
    BMG_CUDA(cudaStreamSynchronize(stm));
  };
  gather_d(st->snp_mean.p, hm);
  gather_d(st->snp_var.p, hv);
  BMG_CUDA(cudaMemsetAsync(ibuf.p, 0, total * sizeof(int32_t), stm));
  BMG_CUDA(cudaMemcpyAsync(ibuf.p + off, st->nmiss.p, st->m * sizeof(int32_t), cudaMemcpyDeviceToDevice, stm));
  if (cm.allgather(cm.ctx, ibuf.p, cm.stride, (int)sizeof(int32_t), (void*)stm) != 0) throw Error("sharded sampler: all-gather callback failed");
  BMG_CUDA(cudaMemcpyAsync(hn.data(), ibuf.p, total * sizeof(int32_t), cudaMemcpyDeviceToHost, stm));
  BMG_CUDA(cudaStreamSynchronize(stm));
  cudaStreamDestroy(stm);
  double tm = 0, tv = 0, nm = 0, nv = 0;
  for (int64_t j = 0; j < st->m_g; ++j) {
    const int ng = (int)st->n - hn[j];
    if (ng > 1) { tm += hm[j]; tv += hv[j]; nm += 1; nv += 1; }
    else if (ng == 1) { tm += hm[j]; nm += 1; }
  }
  out4[0] = tm; out4[1] = nm; out4[2] = tv; out4[3] = nv;
  int64_t missing = 0;
  for (int64_t j = 0; j < st->m_g; ++j) missing += hn[j];
  return missing;   // over all shards: the same number on every rank
}

SamplerHandle* make(const char* ini, int chain_index, int device, Store* existing, const Sampler::ShardComm* comm = nullptr,
                    const GroupHandle* group = nullptr)
{
  BMG_REQUIRE(ini != nullptr, "bmg_sampler_create: null ini path");
  const bool timing = getenv("BMG_TIMING") != nullptr;
  auto now = [] { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; };
  const double t_a = now();
  std::unique_ptr<SamplerHandle> h(new SamplerHandle());
  h->chain_index = chain_index;
  h->opt.reset(new Options(ini, /*quiet=*/chain_index != 0));
  const Options& o = *h->opt;
  BMG_REQUIRE(chain_index >= 0 && (size_t)chain_index < o.n_threads, "bmg_sampler_create: chain_index must be < thread.n_threads");
  h->data.reset(new Dataset(o.n, o.m_g, o.m_e, o.file_fam, o.file_g, o.file_e, o.file_y, /*load_bed=*/false));
  const double t_b = now();
  if (existing) {
    BMG_REQUIRE(existing->n == (int64_t)o.n && existing->m_g == (int64_t)o.m_g, "bmg_sampler_create_on_store: store dimensions differ from the INI file");
    // the host side sizes its Gram blocks from the INI's covariates and reads y / e from the INI's files: the store must carry
    // the same phenotype and covariates (the device kernels use the store's)
    BMG_REQUIRE(existing->m_e == (int)o.m_e + 1, "bmg_sampler_create_on_store: the store's covariates differ from the INI file (sizes.m_e)");
    {
      double yy = 0.0;
      for (double v : h->data->y) yy += v * v;
      const double syy = existing->summaries[5];
      BMG_REQUIRE(std::fabs(yy - syy) <= 1e-9 * std::max(1.0, std::fabs(yy)),
                  "bmg_sampler_create_on_store: the store's phenotype differs from the INI file's (file_y / file_fam)");
    }
    h->store = existing;
  } else {
    // the packed genotypes are streamed from the file to the device and live there only
    h->store = store_create_from_bed(o.file_g.c_str(), (int64_t)o.n, (int64_t)o.m_g, 0, (int64_t)o.m_g,
                                     o.recode_g_to_minor_allele_count, device >= 0 ? device : o.device);
    h->owns_store = true;
    store_set_phenotype(h->store, h->data->y.data(), h->data->e.data(), (int)h->data->m_e);
  }
  const double t_c = now();
  double sm[6];
  for (int i = 0; i < 6; ++i) sm[i] = h->store->summaries[i];
  if (comm != nullptr) {
    BMG_REQUIRE(existing != nullptr, "bmg_sampler_create_sharded: a shard store is required");
    BMG_REQUIRE(existing->m_e >= 1, "bmg_sampler_create_sharded: call bmg_store_set_phenotype on the shard first");
    if (group != nullptr) {
      for (int i = 0; i < 4; ++i) sm[i] = group->summaries[i];
    } else {
      (void)global_summaries(h->store, *comm, sm);
    }
  }
  const double mean_x = sm[0] / sm[1], var_x = sm[2] / sm[3];
  h->sampler.reset(new Sampler(o, chain_index, h->store, h->data->y, h->data->e, sm[4], sm[5], var_x, mean_x, comm));
  if (timing)
    fprintf(stderr, "[bmg timing] create: options + fam/y/e files %.1f ms, store (bed -> device, re-coding, counts) %.1f ms, sampler + chain %.1f ms\n",
            1e3 * (t_b - t_a), 1e3 * (t_c - t_b), 1e3 * (now() - t_c));
  return h.release();
}
}  // namespace

#define BMG_TRY try {
#define BMG_CATCH                      \
  return 0;                            \
  }                                    \
  catch (const std::exception& e)      \
  {                                    \
    set_last_error(e.what());          \
    return 1;                          \
  }                                    \
  catch (...)                          \
  {                                    \
    set_last_error("unknown error");   \
    return 1;                          \
  }

BMG_API int bmg_sampler_create(const char* ini_path, int chain_index, int device, bmg_sampler** out)
{
  BMG_TRY
  BMG_REQUIRE(out != nullptr, "bmg_sampler_create: null argument");
  *out = reinterpret_cast<bmg_sampler*>(make(ini_path, chain_index, device, nullptr));
  BMG_CATCH
}
BMG_API int bmg_sampler_create_sharded(const char* ini_path, int chain_index, bmg_store* shard, const bmg_shard_comm* comm,
                                       bmg_sampler** out)
{
  BMG_TRY
  BMG_REQUIRE(out != nullptr && shard != nullptr && comm != nullptr, "bmg_sampler_create_sharded: null argument");
  BMG_REQUIRE(comm->world >= 1 && comm->rank >= 0 && comm->rank < comm->world && comm->snp_stride > 0,
              "bmg_sampler_create_sharded: invalid communicator");
  BMG_REQUIRE(comm->allgather != nullptr || comm->ctx != nullptr, "bmg_sampler_create_sharded: an all-gather callback or a shard group is required");
  Sampler::ShardComm cm;
  cm.world = comm->world; cm.rank = comm->rank; cm.stride = comm->snp_stride; cm.allgather = comm->allgather; cm.ctx = comm->ctx;
  const GroupHandle* gh = nullptr;
  if (comm->allgather == nullptr) {   // the lockstep chain over the group's native all-gather (peer memory, no host callback)
    gh = reinterpret_cast<const GroupHandle*>(comm->ctx);
    BMG_REQUIRE(group_world(gh->g) == comm->world && group_rank(gh->g) == comm->rank && group_stride(gh->g) == comm->snp_stride,
                "bmg_sampler_create_sharded: communicator and shard group disagree");
    cm.allgather = group_allgather;
    cm.ctx = gh->g;
  }
  *out = reinterpret_cast<bmg_sampler*>(make(ini_path, chain_index, -1, reinterpret_cast<Store*>(shard), &cm, gh));
  BMG_CATCH
}
BMG_API int bmg_group_create(bmg_store* shard, int world, int rank, int n_chains, int64_t snp_stride, const char* shm_name, bmg_group** out)
{
  BMG_TRY
  BMG_REQUIRE(out != nullptr && shard != nullptr, "bmg_group_create: null argument");
  Store* st = reinterpret_cast<Store*>(shard);
  std::unique_ptr<GroupHandle> gh(new GroupHandle());
  gh->g = group_create(st, world, rank, n_chains, snp_stride, shm_name);
  for (int i = 0; i < 6; ++i) gh->summaries[i] = st->summaries[i];
  if (world > 1) {
    Sampler::ShardComm cm;
    cm.world = world; cm.rank = rank; cm.stride = snp_stride; cm.allgather = group_allgather; cm.ctx = gh->g;
    gh->missing_cells = global_summaries(st, cm, gh->summaries);
  } else {
    gh->missing_cells = st->n_missing;
  }
  *out = reinterpret_cast<bmg_group*>(gh.release());
  BMG_CATCH
}
BMG_API int bmg_sampler_create_grouped(const char* ini_path, int chain_index, bmg_store* shard, bmg_group* g, bmg_sampler** out)
{
  BMG_TRY
  BMG_REQUIRE(out != nullptr && shard != nullptr && g != nullptr, "bmg_sampler_create_grouped: null argument");
  const GroupHandle* gh = reinterpret_cast<const GroupHandle*>(g);
  BMG_REQUIRE(chain_index == group_rank(gh->g) && chain_index < group_chains(gh->g),
              "bmg_sampler_create_grouped: chain c of a shard group lives on rank c (chain_index == rank < n_chains)");
  Sampler::ShardComm cm;
  cm.world = group_world(gh->g); cm.rank = group_rank(gh->g); cm.stride = group_stride(gh->g);
  cm.allgather = group_allgather; cm.ctx = gh->g; cm.group = gh->g;
  *out = reinterpret_cast<bmg_sampler*>(make(ini_path, chain_index, -1, reinterpret_cast<Store*>(shard), &cm, gh));
  BMG_CATCH
}
BMG_API int bmg_group_serve(bmg_group* g, int64_t n_rounds)
{
  BMG_TRY
  BMG_REQUIRE(g != nullptr && n_rounds >= 0, "bmg_group_serve: invalid argument");
  group_serve(reinterpret_cast<GroupHandle*>(g)->g, n_rounds);
  BMG_CATCH
}
BMG_API bmg_chain* bmg_group_scan_chain(bmg_group* g)
{
  return g ? reinterpret_cast<bmg_chain*>(group_scan_chain(reinterpret_cast<GroupHandle*>(g)->g)) : nullptr;
}
BMG_API int bmg_group_stats(bmg_group* g, double* out4)
{
  BMG_TRY
  BMG_REQUIRE(g != nullptr && out4 != nullptr, "bmg_group_stats: null argument");
  group_stats(reinterpret_cast<GroupHandle*>(g)->g, out4);
  BMG_CATCH
}
BMG_API int bmg_group_destroy(bmg_group* g)
{
  BMG_TRY
  delete reinterpret_cast<GroupHandle*>(g);
  BMG_CATCH
}
BMG_API int bmg_sampler_create_on_store(const char* ini_path, int chain_index, bmg_store* s, bmg_sampler** out)
{
  BMG_TRY
  BMG_REQUIRE(out != nullptr && s != nullptr, "bmg_sampler_create_on_store: null argument");
  *out = reinterpret_cast<bmg_sampler*>(make(ini_path, chain_index, -1, reinterpret_cast<Store*>(s)));
  BMG_CATCH
}
BMG_API int bmg_ini_lookup(const char* ini_path, const char* section, const char* key, const char* dflt, char* out, int out_len)
{
  BMG_TRY
  BMG_REQUIRE(ini_path && section && key && out && out_len > 0, "bmg_ini_lookup: invalid argument");
  IniFile ini(ini_path);
  BMG_REQUIRE(ini.parse_error() == 0, "Cannot load/parse configuration file.");   // the reference's message (src/options.hpp:83)
  const std::string v = ini.get(section, key, dflt ? dflt : "");
  std::snprintf(out, (size_t)out_len, "%s", v.c_str());
  BMG_CATCH
}
BMG_API int bmg_store_create_from_ini(const char* ini_path, int64_t snp_lo, int64_t snp_hi, int device, bmg_store** out)
{
  BMG_TRY
  BMG_REQUIRE(ini_path && out, "bmg_store_create_from_ini: null argument");
  const Options o(ini_path, /*quiet=*/true);
  if (snp_hi < 0) snp_hi = (int64_t)o.m_g;
  const Dataset d(o.n, o.m_g, o.m_e, o.file_fam, o.file_g, o.file_e, o.file_y, /*load_bed=*/false);
  Store* st = store_create_from_bed(o.file_g.c_str(), (int64_t)o.n, (int64_t)o.m_g, snp_lo, snp_hi, o.recode_g_to_minor_allele_count,
                                    device >= 0 ? device : o.device);
  try {
    store_set_phenotype(st, d.y.data(), d.e.data(), (int)d.m_e);
  } catch (...) {
    cudaSetDevice(st->device);
    delete st;
    throw;
  }
  *out = reinterpret_cast<bmg_store*>(st);
  BMG_CATCH
}
BMG_API int bmg_sampler_set_option(bmg_sampler* sp, const char* key, const char* value)
{
  BMG_TRY
  BMG_REQUIRE(key && value, "bmg_sampler_set_option: null argument");
  H(sp)->sampler->set_option(key, value);
  BMG_CATCH
}
BMG_API int bmg_sampler_begin(bmg_sampler* sp)
{
  BMG_TRY
  SamplerHandle* h = H(sp);
  if (h->chain_index == 0) h->sampler->print_prior();   // main.cpp:77
  if (!h->flat_done) { h->sampler->initialize_p_proposal_flat(); h->flat_done = true; }   // main.cpp:114
  h->sampler->begin();
  BMG_CATCH
}
BMG_API int bmg_sampler_run(bmg_sampler* sp, int64_t n_iter)
{
  BMG_TRY
  BMG_REQUIRE(n_iter >= 0, "bmg_sampler_run: negative iteration count");
  H(sp)->sampler->run(n_iter);
  BMG_CATCH
}
BMG_API int bmg_sampler_end(bmg_sampler* sp)
{
  BMG_TRY
  H(sp)->sampler->end();
  BMG_CATCH
}
BMG_API int bmg_sampler_stats(bmg_sampler* sp, double* out8)
{
  BMG_TRY
  BMG_REQUIRE(out8, "bmg_sampler_stats: null argument");
  H(sp)->sampler->stats(out8);
  BMG_CATCH
}
BMG_API int bmg_sampler_counters(bmg_sampler* sp, double* out, int n)
{
  BMG_TRY
  BMG_REQUIRE(out != nullptr && n >= 0, "bmg_sampler_counters: invalid argument");
  double v[12];
  H(sp)->sampler->counters(v);
  for (int i = 0; i < n && i < 12; ++i) out[i] = v[i];
  BMG_CATCH
}
BMG_API int bmg_sampler_inclusion_counts(bmg_sampler* sp, uint32_t* counts, int64_t* n_samples)
{
  BMG_TRY
  const int64_t ns = H(sp)->sampler->inclusion_counts(counts);
  if (n_samples) *n_samples = ns;
  BMG_CATCH
}
BMG_API bmg_store* bmg_sampler_store(bmg_sampler* sp) { return sp ? reinterpret_cast<bmg_store*>(reinterpret_cast<SamplerHandle*>(sp)->store) : nullptr; }
BMG_API bmg_chain* bmg_sampler_chain(bmg_sampler* sp)
{
  return sp ? reinterpret_cast<bmg_chain*>(reinterpret_cast<SamplerHandle*>(sp)->sampler->chain()) : nullptr;
}
BMG_API int bmg_sampler_destroy(bmg_sampler* sp)
{
  BMG_TRY
  delete reinterpret_cast<SamplerHandle*>(sp);
  BMG_CATCH
}
