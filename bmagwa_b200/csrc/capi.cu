// capi.cu -- the extern "C" boundary declared in include/bmagwa_b200.h (store + chain part).
#include <cstring>
#include "common.cuh"
#include "store.cuh"

namespace bmg {
std::atomic<uint64_t> g_launches{0}, g_h2d_bytes{0}, g_d2h_bytes{0};
static thread_local std::string t_last_error;
void set_last_error(const std::string& m) { t_last_error = m; }
}  // namespace bmg

using namespace bmg;

#define BMG_API extern "C" __attribute__((visibility("default")))

#define BMG_TRY try {
#define BMG_CATCH                                       \
  return 0;                                             \
  }                                                     \
  catch (const std::exception& e)                       \
  {                                                     \
    set_last_error(e.what());                           \
    return 1;                                           \
  }                                                     \
  catch (...)                                           \
  {                                                     \
    set_last_error("unknown error");                    \
    return 1;                                           \
  }

static Store* S(bmg_store* s)
{
  BMG_REQUIRE(s != nullptr, "null store handle");
  return reinterpret_cast<Store*>(s);
}
static const Store* S(const bmg_store* s)
{
  BMG_REQUIRE(s != nullptr, "null store handle");
  return reinterpret_cast<const Store*>(s);
}
static Chain* Cn(bmg_chain* c)
{
  BMG_REQUIRE(c != nullptr, "null chain handle");
  return reinterpret_cast<Chain*>(c);
}

BMG_API int bmg_abi_version(void) { return BMG_ABI_VERSION; }
BMG_API const char* bmg_last_error(void) { return t_last_error.c_str(); }
BMG_API int bmg_device_count(void)
{
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    set_last_error(std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    return -1;
  }
  return n;
}
BMG_API uint64_t bmg_launch_count(void) { return g_launches.load(); }
BMG_API void bmg_transfer_bytes(uint64_t* h2d, uint64_t* d2h)
{
  if (h2d) *h2d = g_h2d_bytes.load();
  if (d2h) *d2h = g_d2h_bytes.load();
}

// ---- store --------------------------------------------------------------------------------
BMG_API int bmg_store_create(const uint8_t* bed_payload, int payload_on_device, int64_t n, int64_t m_g, int64_t snp_lo,
                             int64_t snp_hi, int recode_to_minor, int device, bmg_store** out)
{
  BMG_TRY
  BMG_REQUIRE(out != nullptr && bed_payload != nullptr, "bmg_store_create: null argument");
  *out = reinterpret_cast<bmg_store*>(store_create(bed_payload, payload_on_device != 0, n, m_g, snp_lo, snp_hi,
                                                   recode_to_minor != 0, device));
  BMG_CATCH
}

BMG_API int bmg_store_create_from_bed(const char* bed_path, int64_t n, int64_t m_g, int64_t snp_lo, int64_t snp_hi,
                                      int recode_to_minor, int device, bmg_store** out)
{
  BMG_TRY
  BMG_REQUIRE(out != nullptr && bed_path != nullptr, "bmg_store_create_from_bed: null argument");
  *out = reinterpret_cast<bmg_store*>(store_create_from_bed(bed_path, n, m_g, snp_lo, snp_hi, recode_to_minor != 0, device));
  BMG_CATCH
}
BMG_API int bmg_store_destroy(bmg_store* s)
{
  BMG_TRY
  if (s) {
    Store* st = reinterpret_cast<Store*>(s);
    cudaSetDevice(st->device);
    for (PeerShard& p : st->peers)
      if (p.ipc_opened) cudaIpcCloseMemHandle(const_cast<uint32_t*>(p.codes));
    delete st;
  }
  BMG_CATCH
}

BMG_API int bmg_store_set_phenotype(bmg_store* s, const double* y, const double* e, int m_e)
{
  BMG_TRY
  BMG_REQUIRE(y && e, "bmg_store_set_phenotype: null argument");
  store_set_phenotype(S(s), y, e, m_e);
  BMG_CATCH
}

BMG_API int bmg_store_dims(const bmg_store* s, int64_t* n, int64_t* m_g, int64_t* snp_lo, int64_t* snp_hi, int* m_e,
                           int64_t* n_missing_cells)
{
  BMG_TRY
  const Store* st = S(s);
  if (n) *n = st->n;
  if (m_g) *m_g = st->m_g;
  if (snp_lo) *snp_lo = st->lo;
  if (snp_hi) *snp_hi = st->hi;
  if (m_e) *m_e = st->m_e;
  if (n_missing_cells) *n_missing_cells = st->n_missing;
  BMG_CATCH
}

BMG_API int bmg_store_counts(const bmg_store* s, int32_t* n1, int32_t* n2, int32_t* n_miss, uint8_t* swapped)
{
  BMG_TRY
  const Store* st = S(s);
  BMG_CUDA(cudaSetDevice(st->device));
  if (n1) bmg::copy_d2h_sync(n1, st->n1.p, st->m * sizeof(int32_t));
  if (n2) bmg::copy_d2h_sync(n2, st->n2.p, st->m * sizeof(int32_t));
  if (n_miss) bmg::copy_d2h_sync(n_miss, st->nmiss.p, st->m * sizeof(int32_t));
  if (swapped) bmg::copy_d2h_sync(swapped, st->swapped.p, st->m);
  BMG_CATCH
}

BMG_API int bmg_store_summaries(const bmg_store* s, double* out6)
{
  BMG_TRY
  BMG_REQUIRE(out6, "bmg_store_summaries: null argument");
  std::memcpy(out6, S(s)->summaries, 6 * sizeof(double));
  BMG_CATCH
}

BMG_API int bmg_store_moments(const bmg_store* s, double* xx)
{
  BMG_TRY
  const Store* st = S(s);
  BMG_REQUIRE(xx, "bmg_store_moments: null argument");
  BMG_CUDA(cudaSetDevice(st->device));
  bmg::copy_d2h_sync(xx, st->mom.p, 2 * st->m * sizeof(double));
  BMG_CATCH
}

BMG_API int bmg_store_missing(const bmg_store* s, int64_t* offsets, int64_t* idx, double* prior3)
{
  BMG_TRY
  const Store* st = S(s);
  BMG_CUDA(cudaSetDevice(st->device));
  if (offsets) std::memcpy(offsets, st->h_miss_off.data(), (st->m + 1) * sizeof(int64_t));
  if (idx && st->n_missing > 0) {
    std::vector<int32_t> tmp(st->n_missing);
    bmg::copy_d2h_sync(tmp.data(), st->miss_idx.p, st->n_missing * sizeof(int32_t));
    for (int64_t q = 0; q < st->n_missing; ++q) idx[q] = tmp[q];
  }
  if (prior3) {
    // Data::handle_missing_g (data.cpp:357-372): cumulative counts of 0/1/2 among the observed cells
    std::vector<int32_t> a(st->m), b(st->m), c(st->m);
    bmg::copy_d2h_sync(a.data(), st->n1.p, st->m * sizeof(int32_t));
    bmg::copy_d2h_sync(b.data(), st->n2.p, st->m * sizeof(int32_t));
    bmg::copy_d2h_sync(c.data(), st->nmiss.p, st->m * sizeof(int32_t));
    for (int64_t j = 0; j < st->m; ++j) {
      const double n0 = (double)(st->n - c[j] - a[j] - b[j]);
      prior3[3 * j] = n0;
      prior3[3 * j + 1] = n0 + a[j];
      prior3[3 * j + 2] = n0 + a[j] + b[j];
    }
  }
  BMG_CATCH
}

BMG_API int bmg_store_get_column(const bmg_store* s, int64_t snp, int type, double* out)
{
  BMG_TRY
  BMG_REQUIRE(out, "bmg_store_get_column: null argument");
  store_get_column(S(s), snp, type, nullptr, false, out, 0);
  BMG_CATCH
}

BMG_API int bmg_store_export(const bmg_store* s, void* ipc_handle64, int64_t* words_per_snp)
{
  BMG_TRY
  const Store* st = S(s);
  BMG_CUDA(cudaSetDevice(st->device));
  if (ipc_handle64) {
    cudaIpcMemHandle_t h;
    BMG_CUDA(cudaIpcGetMemHandle(&h, st->codes.p));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    std::memcpy(ipc_handle64, &h, 64);
  }
  if (words_per_snp) *words_per_snp = st->Wp;
  BMG_CATCH
}

BMG_API int bmg_store_attach_peer(bmg_store* s, const void* handle_or_ptr, int same_process, int64_t snp_lo, int64_t snp_hi)
{
  BMG_TRY
  Store* st = S(s);
  BMG_REQUIRE(handle_or_ptr != nullptr && snp_lo >= 0 && snp_hi > snp_lo && snp_hi <= st->m_g, "bmg_store_attach_peer: bad arguments");
  BMG_REQUIRE(snp_hi <= st->lo || snp_lo >= st->hi, "bmg_store_attach_peer: peer range overlaps the local shard");
  BMG_CUDA(cudaSetDevice(st->device));
  PeerShard p;
  p.lo = snp_lo; p.hi = snp_hi;
  if (same_process) {
    p.codes = reinterpret_cast<const uint32_t*>(handle_or_ptr);
    p.ipc_opened = false;
  } else {
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle_or_ptr, 64);
    void* ptr = nullptr;
    BMG_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    p.codes = reinterpret_cast<const uint32_t*>(ptr);
    p.ipc_opened = true;
  }
  st->peers.push_back(p);
  BMG_CATCH
}

// ---- chain --------------------------------------------------------------------------------
BMG_API int bmg_chain_create(bmg_store* s, bmg_chain** out)
{
  BMG_TRY
  BMG_REQUIRE(out, "bmg_chain_create: null argument");
  *out = reinterpret_cast<bmg_chain*>(chain_create(S(s)));
  BMG_CATCH
}
BMG_API int bmg_chain_destroy(bmg_chain* c)
{
  BMG_TRY
  chain_destroy(reinterpret_cast<Chain*>(c));
  BMG_CATCH
}
BMG_API int bmg_chain_sync(bmg_chain* c)
{
  BMG_TRY
  Chain* ch = Cn(c);
  BMG_CUDA(cudaSetDevice(ch->store->device));
  BMG_CUDA(cudaStreamSynchronize(ch->stream));
  BMG_CATCH
}
BMG_API void* bmg_chain_stream(bmg_chain* c) { return c ? (void*)reinterpret_cast<Chain*>(c)->stream : nullptr; }

BMG_API int bmg_chain_set_missing(bmg_chain* c, int64_t snp, const int8_t* vals, int64_t count)
{
  BMG_TRY
  BMG_REQUIRE(vals || count == 0, "bmg_chain_set_missing: null argument");
  chain_set_missing(Cn(c), snp, vals, count);
  BMG_CATCH
}
BMG_API int bmg_chain_set_missing_all(bmg_chain* c, const int8_t* vals, int64_t count)
{
  BMG_TRY
  BMG_REQUIRE(vals || count == 0, "bmg_chain_set_missing_all: null argument");
  chain_set_missing_all(Cn(c), vals, count);
  BMG_CATCH
}
BMG_API int bmg_chain_impute_from_prior(bmg_chain* c, const int64_t* loci, int k, uint64_t seed, uint64_t counter)
{
  BMG_TRY
  BMG_REQUIRE(loci || k == 0, "bmg_chain_impute_from_prior: null argument");
  chain_impute_from_prior(Cn(c), loci, k, seed, counter);
  BMG_CATCH
}
BMG_API int bmg_chain_get_cells(bmg_chain* c, const int64_t* loci, int k, const int32_t* rows, int64_t q, int8_t* out)
{
  BMG_TRY
  BMG_REQUIRE((loci && rows && out) || k == 0 || q == 0, "bmg_chain_get_cells: null argument");
  chain_get_cells(Cn(c), loci, k, rows, q, out);
  BMG_CATCH
}
BMG_API int bmg_chain_get_column(bmg_chain* c, int64_t snp, int type, double* out)
{
  BMG_TRY
  Chain* ch = Cn(c);
  BMG_REQUIRE(out, "bmg_chain_get_column: null argument");
  BMG_REQUIRE(ch->mv.base == ch->store->lo && ch->mv.m == ch->store->m,
              "bmg_chain_get_column: not available on a chain whose missing-call index covers a sharded data set");
  store_get_column(ch->store, snp, type, ch->miss_val.p, true, out, ch->stream);
  BMG_CATCH
}
BMG_API int bmg_chain_residual(bmg_chain* c, const int64_t* loci, const double* beta_e, const double* beta_g, int k,
                               double* stats9)
{
  BMG_TRY
  BMG_REQUIRE(beta_e && (k == 0 || (loci && beta_g)), "bmg_chain_residual: null argument");
  chain_residual(Cn(c), loci, beta_e, beta_g, k, stats9);
  BMG_CATCH
}
BMG_API int bmg_chain_residual_types(bmg_chain* c, const int64_t* loci, const int32_t* term_type, const double* beta_e,
                                     const double* beta_g, int k, double* stats9)
{
  BMG_TRY
  BMG_REQUIRE(beta_e && (k == 0 || (loci && beta_g && term_type)), "bmg_chain_residual_types: null argument");
  chain_residual(Cn(c), loci, beta_e, beta_g, k, stats9, term_type);
  BMG_CATCH
}
BMG_API int bmg_chain_scan_types(bmg_chain* c, const int64_t* loci, const int32_t* loci_type, const double* beta2,
                                 const double* tau2, int k, const bmg_scan_types_params* prm, double* p_r_host,
                                 double* p_r_types_host)
{
  BMG_TRY
  BMG_REQUIRE(prm && (k == 0 || (loci && loci_type && beta2 && tau2)), "bmg_chain_scan_types: null argument");
  chain_scan_types(Cn(c), loci, loci_type, beta2, tau2, k, prm, p_r_host, p_r_types_host);
  BMG_CATCH
}
BMG_API int bmg_chain_get_residual(bmg_chain* c, double* r)
{
  BMG_TRY
  Chain* ch = Cn(c);
  BMG_REQUIRE(r && ch->residual_valid, "bmg_chain_get_residual: no residual available");
  BMG_CUDA(cudaSetDevice(ch->store->device));
  bmg::copy_d2h(r, ch->r.p, ch->store->n * sizeof(double), ch->stream);
  BMG_CUDA(cudaStreamSynchronize(ch->stream));
  BMG_CATCH
}
BMG_API int bmg_chain_scan(bmg_chain* c, const int64_t* loci, const double* beta_g, const double* tau_g, int k,
                           const bmg_scan_params* prm, double* p_r_host)
{
  BMG_TRY
  BMG_REQUIRE(k == 0 || (loci && beta_g && tau_g), "bmg_chain_scan: null argument");
  chain_scan(Cn(c), loci, beta_g, tau_g, k, prm, p_r_host);
  BMG_CATCH
}
BMG_API int bmg_chain_scan_dots(bmg_chain* c, double* dot_host)
{
  BMG_TRY
  Chain* ch = Cn(c);
  chain_scan_dots(ch);
  if (dot_host) {
    // sum the per-chunk partials on the host side of the copy (tests / roofline probe only)
    const int64_t m = ch->store->m;
    std::vector<double> tmp((size_t)ch->last_chunks * m);
    bmg::copy_d2h(tmp.data(), ch->last_partial, tmp.size() * sizeof(double), ch->stream);
    BMG_CUDA(cudaStreamSynchronize(ch->stream));
    for (int64_t j = 0; j < m; ++j) {
      double s = 0.0;
      for (int q = 0; q < ch->last_chunks; ++q) s += tmp[(size_t)q * m + j];
      dot_host[j] = s;
    }
  }
  BMG_CATCH
}
BMG_API int bmg_chain_set_scan_variant(bmg_chain* c, int variant)
{
  BMG_TRY
  BMG_REQUIRE(variant >= 0 && variant <= 2, "bmg_chain_set_scan_variant: variant must be 0, 1 or 2");
  Cn(c)->scan_variant = variant;
  BMG_CATCH
}
BMG_API int bmg_chain_scan_kernel_time(bmg_chain* c, int enable, double* ms_total, int64_t* launches, int reset)
{
  BMG_TRY
  Chain* ch = Cn(c);
  BMG_CUDA(cudaSetDevice(ch->store->device));
  ch->time_scan = enable != 0;
  if (ms_total || launches) {
    BMG_CUDA(cudaStreamSynchronize(ch->stream));
    double tot = ch->scan_ms_done;
    for (size_t i = 0; i < ch->scan_ev_used; ++i) {
      float ms = 0.f;
      BMG_CUDA(cudaEventElapsedTime(&ms, ch->scan_ev[2 * i], ch->scan_ev[2 * i + 1]));
      tot += ms;
    }
    ch->scan_ms_done = tot;
    ch->scan_launches_done += (int64_t)ch->scan_ev_used;
    ch->scan_ev_used = 0;
    if (ms_total) *ms_total = ch->scan_ms_done;
    if (launches) *launches = ch->scan_launches_done;
  }
  if (reset) { ch->scan_ms_done = 0.0; ch->scan_launches_done = 0; ch->scan_ev_used = 0; }
  BMG_CATCH
}
BMG_API int bmg_chain_adapt(bmg_chain* c, int update_rao, int64_t n_rao_mean, int update_proposal, int64_t n_prop_mean,
                            double q_add_min, double q_rem_min)
{
  BMG_TRY
  chain_adapt(Cn(c), update_rao, n_rao_mean, update_proposal, n_prop_mean, q_add_min, q_rem_min);
  BMG_CATCH
}
BMG_API int bmg_chain_init_proposal_flat(bmg_chain* c, double value, double q_add_min, double q_rem_min)
{
  BMG_TRY
  chain_init_flat(Cn(c), value, q_add_min, q_rem_min);
  BMG_CATCH
}
BMG_API int bmg_chain_get_array(bmg_chain* c, int which, double* out)
{
  BMG_TRY
  Chain* ch = Cn(c);
  BMG_REQUIRE(out && which >= 0 && which <= 4, "bmg_chain_get_array: bad arguments");
  const double* src[5] = {ch->p_r.p, ch->p_rao.p, ch->p_proposal.p, ch->q_add.p, ch->q_rem.p};
  BMG_CUDA(cudaSetDevice(ch->store->device));
  bmg::copy_d2h(out, src[which], ch->mw * sizeof(double), ch->stream);
  BMG_CUDA(cudaStreamSynchronize(ch->stream));
  BMG_CATCH
}
BMG_API int bmg_chain_partial_cdf(bmg_chain* c, int64_t* n_blocks, int64_t* block_size, double* add_sums, double* rem_sums)
{
  BMG_TRY
  Chain* ch = Cn(c);
  BMG_CUDA(cudaSetDevice(ch->store->device));
  if (n_blocks) *n_blocks = ch->cdf_blocks;
  if (block_size) *block_size = ch->cdf_block;
  if (add_sums) bmg::copy_d2h(add_sums, ch->cdf_add.p, ch->cdf_blocks * sizeof(double), ch->stream);
  if (rem_sums) bmg::copy_d2h(rem_sums, ch->cdf_rem.p, ch->cdf_blocks * sizeof(double), ch->stream);
  BMG_CUDA(cudaStreamSynchronize(ch->stream));
  BMG_CATCH
}
BMG_API int bmg_chain_sample(bmg_chain* c, int which, double u01, int64_t* snp, double* total_w)
{
  BMG_TRY
  BMG_REQUIRE(snp, "bmg_chain_sample: null argument");
  chain_sample(Cn(c), which, u01, snp, total_w);
  BMG_CATCH
}
BMG_API int bmg_chain_set_zeroed(bmg_chain* c, int which, int64_t snp, int zeroed)
{
  BMG_TRY
  chain_set_zeroed(Cn(c), which, snp, zeroed);
  BMG_CATCH
}
BMG_API int bmg_chain_fill_zeroed(bmg_chain* c, int which, int zeroed)
{
  BMG_TRY
  chain_fill_zeroed(Cn(c), which, zeroed);
  BMG_CATCH
}
BMG_API int bmg_chain_column_stats(bmg_chain* c, const int64_t* cand, int m_c, const int64_t* loci, int k, double* xy,
                                   double* xe, double* xx_model, double* xx_cand)
{
  BMG_TRY
  BMG_REQUIRE(cand && (k == 0 || loci), "bmg_chain_column_stats: null argument");
  chain_column_stats(Cn(c), cand, m_c, loci, k, xy, xe, xx_model, xx_cand);
  BMG_CATCH
}
BMG_API int bmg_chain_probit_update(bmg_chain* c, const uint8_t* is_case, const double* u01, uint64_t seed, uint64_t counter,
                                    double* stats2)
{
  BMG_TRY
  chain_probit_update(Cn(c), is_case, u01, seed, counter, stats2);
  BMG_CATCH
}
BMG_API int bmg_chain_get_phenotype(bmg_chain* c, double* y_out)
{
  BMG_TRY
  Chain* ch = Cn(c);
  BMG_REQUIRE(y_out, "bmg_chain_get_phenotype: null argument");
  BMG_CUDA(cudaSetDevice(ch->store->device));
  bmg::copy_d2h(y_out, ch->y.p, ch->store->n * sizeof(double), ch->stream);
  BMG_CUDA(cudaStreamSynchronize(ch->stream));
  BMG_CATCH
}
