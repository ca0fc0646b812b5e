// store.cuh -- internal C++ interface between the modules of libbmagwa_b200.so
#pragma once
#include <memory>
#include <unordered_map>
#include <utility>
#include "common.cuh"
#include "../../include/bmagwa_b200.h"

namespace bmg {

// ---- store.cu
Store* store_create_from_bed(const char* path, int64_t n, int64_t m_g, int64_t lo, int64_t hi, bool recode, int device);
Store* store_create(const uint8_t* bed, bool on_device, int64_t n, int64_t m_g, int64_t lo, int64_t hi, bool recode,
                    int device);
void store_set_phenotype(Store* s, const double* y, const double* e, int m_e);
void store_get_column(const Store* s, int64_t snp, int type, const int8_t* miss_vals_dev, bool overlay,
                      double* out_host, cudaStream_t st);

// ---- chain state (scan.cu, colstats.cu, weights.cu, probit.cu)
// host-supplied all-gather over the ranks of a SNP-sharded chain (include/bmagwa_b200.h: bmg_allgather_fn)
typedef int (*AllGatherFn)(void* ctx, void* dev_buffer, int64_t elems_per_rank, int elem_bytes, void* cuda_stream);

// one column to (re)build in the overlay cache (overlay.cu)
struct PatchDesc {
  const uint32_t* src;   // store column
  uint32_t* dst;         // cache slot
  const int32_t* idx;    // individuals of the SNP's missing cells
  const int8_t* vals;    // their imputed values: mapped host staging, or the chain's device array
  int8_t* keep;          // where to store the values on the device (nullptr: they are already there)
  int64_t cnt;
  int type;              // 0 additive values; 1 H [x == 1], 2 D [x > 0], 3 R [x == 2] written as 0/1 fields
};

struct Group;   // group.cuh

// The missing-call index a chain works on: the store's own (one GPU), or the whole data set's when the packed genotypes
// are sharded -- the imputed values of ALL SNPs belong to the chain, whichever GPU holds a SNP's packed column
// (src/data_model.hpp:80-84: one miss_val per chain).  CSR over the SNPs [base, base + m); arrays on the chain's device.
struct MissView {
  int64_t base = 0, m = 0, n_missing = 0;
  const int64_t* off = nullptr;     // m + 1
  const int32_t* idx = nullptr;     // individuals of the missing cells, ascending per SNP
  const int32_t* n1 = nullptr;      // genotype counts per SNP of the view
  const int32_t* n2 = nullptr;
  const int32_t* nmiss = nullptr;
  const int64_t* h_off = nullptr;   // host mirror of off
  bool covers(int64_t snp) const { return snp >= base && snp < base + m; }
  int64_t count(int64_t snp) const { return h_off[snp - base + 1] - h_off[snp - base]; }
};
// owned arrays of a view over all SNPs of a sharded data set (built collectively, store.cu)
struct GlobalMissing {
  DevBuf<int64_t> off;
  DevBuf<int32_t> idx, n1, n2, nmiss;
  std::vector<int64_t> h_off;
  int64_t n_missing = 0, m_g = 0;
  MissView view() const
  {
    MissView v;
    v.base = 0; v.m = m_g; v.n_missing = n_missing; v.off = off.p; v.idx = idx.p; v.n1 = n1.p; v.n2 = n2.p; v.nmiss = nmiss.p;
    v.h_off = h_off.data();
    return v;
  }
};

struct Chain {
  Store* store = nullptr;
  // several chains over one SNP-sharded store (group.cu): this chain lives on its rank only, its weight arrays cover all
  // m_g SNPs, and its scan is served by every rank of the group
  Group* group = nullptr;
  cudaStream_t stream = nullptr;
  // Range of the per-SNP proposal / Rao-Blackwell arrays (p_r, p_rao, p_proposal, q_add, q_rem, zero flags, CDF
  // blocks).  Unsharded: mw = store->m, w_off = 0.  SNP-sharded chain: every rank keeps them for ALL m_g SNPs
  // (8 B per SNP each) and only the packed genotypes and the scan are sharded; the local shard's p_r lands at w_off
  // and is all-gathered (chain_scan), everything downstream is replicated and identical on every rank.
  int64_t mw = 0, w_off = 0, mw_alloc = 0;
  int world = 1, rank = 0;
  int64_t shard_stride = 0;
  AllGatherFn gather = nullptr;
  void* gather_ctx = nullptr;
  DevBuf<int32_t> inorder_own;              // in-order permutation over mw when it differs from the store's
  std::vector<int32_t> h_inorder_own;
  const int32_t* inorder_dev() const { return inorder_own.n ? inorder_own.p : store->inorder.p; }
  const std::vector<int32_t>& inorder_host() const { return h_inorder_own.empty() ? store->h_inorder : h_inorder_own; }
  int scan_variant = 2;   // 2 = integer tensor cores (default), 1 = fp64 + TMA staging, 0 = fp64 + direct loads
  // optional CUDA-event timing of the scan's reduction kernel (bench.py roofline): pairs of events
  bool time_scan = false;
  std::vector<cudaEvent_t> scan_ev;
  size_t scan_ev_used = 0;
  double scan_ms_done = 0.0;
  int64_t scan_launches_done = 0;
  // scan geometry (chosen once per chain from n and the SM count)
  int scan_warps = 0;        // warps per CTA
  int64_t scan_chunk_words = 0;  // words of each column one CTA covers
  int scan_chunks = 0;       // CTAs along the individual axis
  int scan_ctas_per_chunk = 0;
  // variant 2 (integer tensor cores, scan_imma.cu)
  bool imma_ready = false, imma_q_valid = false;
  int imma_warps = 0, imma_chunks = 0, imma_slices = 0;
  int imma_slices2 = 0;          // SNP slices of the two-residual kernel (0: not available for this geometry)
  int64_t imma_chunk_words = 0;
  DevBuf<unsigned char> imma_q;                 // residual limbs [word][8][16]
  DevBuf<int> imma_exp;                         // fixed-point exponent S
  DevBuf<double> imma_partial;                  // [imma_chunks][m]
  DevBuf<double> imma_partial_h;                // same for the heterozygote-indicator pass of the typed scan
  const double* last_partial = nullptr;         // per-chunk dots of the most recent scan
  int last_chunks = 0;
  // phenotype the chain works on (a copy of the store's y; the probit update overwrites it)
  DevBuf<double> y;
  DevBuf<uint8_t> is_case;
  bool have_case = false;
  // fitted values / residual
  DevBuf<double> yhat_e, yhat_g, r, r_scaled;   // n, n, n, 16*W (zero padded)
  DevBuf<double> red_partial;                   // block partials of the residual reductions
  DevBuf<double> red_out;                       // 16 doubles
  PinnedBuf<double> h_red;                      // pinned mirror
  double sum_r = 0.0;
  bool residual_valid = false;
  // per-chain imputed values of the missing cells (CSR of mv: the store's, or the whole data set's for sharded chains)
  MissView mv;
  std::unique_ptr<GlobalMissing> mv_own;        // lockstep sharded chain: its own copy of the global index
  DevBuf<int8_t> miss_val;
  DevBuf<double> miss_corr;                     // 3 per SNP of mv: (dot corr, sum val, sum val^2)
  DevBuf<double> miss_corr4;                    // typed scan: (sum_{val=1} r, sum_{val=2} r, #val=1, #val=2)
  DevBuf<double> p_r_types;                     // typed scan: m x n_types
  DevBuf<double> tm_d;                          // typed scan: model betas [k][2] | taus [k][2]
  DevBuf<int32_t> tm_i;                         // typed scan: model effect types
  PinnedBuf<double> tm_h;
  // packed columns with the imputed values filled in (overlay.cu)
  DevBuf<uint32_t> pc_cols;                     // pc_slots x store->Wp words
  int pc_slots = 0, pc_next = 0;
  unsigned int pc_seq = 0;
  std::vector<int64_t> pc_snp;                  // slot -> key (4 * local SNP + type), -1 free
  std::vector<unsigned int> pc_use;             // slot -> sequence number of the last request that used it
  std::unordered_map<int64_t, int> pc_map;      // 4 * local SNP + effect type -> slot
  std::vector<PatchDesc> pc_pending;
  PinnedBuf<int8_t> pc_h_vals;                  // mapped staging of new imputed values, read by k_patch_columns
  // scratch of chain_get_cells (values of a few (SNP, individual) cells for the missing-genotype Gibbs step)
  DevBuf<int64_t> gc_meta;                      // k column pointers, then k local SNP indices
  DevBuf<int32_t> gc_rows;
  DevBuf<int8_t> gc_out;
  PinnedBuf<int64_t> gc_h_meta;
  PinnedBuf<int32_t> gc_h_rows;
  PinnedBuf<int8_t> gc_h_out;
  // scan outputs and the per-SNP arrays Sampler keeps (sampler.hpp:201-258)
  DevBuf<double> dot_partial;                   // scan_chunks * m
  DevBuf<double> dot;                           // m
  DevBuf<double> p_r, p_rao, p_proposal, q_add, q_rem;
  DevBuf<double> tau_dev;                       // per-SNP tau upload (tau_mode 1)
  DevBuf<int64_t> loci_dev;
  DevBuf<double> beta_dev, taug_dev;
  PinnedBuf<double> h_stage;                    // pinned staging for small H2D/D2H
  PinnedBuf<int64_t> h_stage_i;
  // proposal weights: partial CDFs + zero flags
  int64_t cdf_block = 256;
  int64_t cdf_blocks = 0;
  DevBuf<double> cdf_add, cdf_rem;              // block sums over in-order positions (all items)
  DevBuf<double> cdf_eff_add, cdf_eff_rem;      // same, zeroed items excluded
  DevBuf<double> q_add_io, q_rem_io;            // weights in in-order layout (what the host mirror copies)
  DevBuf<uint8_t> zero_add, zero_rem;           // per local SNP
  DevBuf<double> sample_out;                    // {snp, total}
  PinnedBuf<double> h_sample;
  // column statistics scratch
  std::vector<int64_t> cs_involved;
  std::vector<const uint32_t*> cs_colp;
  DevBuf<double> cs_out;
  PinnedBuf<double> h_cs;
  DevBuf<int64_t> cs_idx;
  size_t cs_cap = 0;
  PinnedBuf<double> cs_map;                     // results written by the kernel straight into host memory
  PinnedBuf<unsigned int> cs_flag;              // completion sequence number (host-visible)
  DevBuf<unsigned int> cs_done;                 // CTA arrival counter
  unsigned int cs_seq = 0;
  bool cs_pending = false;
  // persistent column-statistics server (colstats.cu): one long-running kernel fed through a host mailbox
  bool server_enabled = true, server_running = false;   // BMG_COLSTATS_SERVER=0 / option colstats_server=0: one launch per move
  cudaStream_t server_stream = nullptr;
  PinnedBuf<uint4> server_mail;
  DevBuf<uint4> server_req;
  DevBuf<unsigned int> server_flag;
  int server_ctas = 0, server_seg_words = 0;
  const double* server_base_y = nullptr;
  const void* server_base_out = nullptr;
  int64_t server_requests = 0;
  bool server_counted = false;                  // among the chains of this process that share the device's SMs for their servers
  int server_failures = 0;                      // consecutive requests a server instance left unserved (see chain_column_stats_wait)
  int64_t server_fallbacks = 0;                 // requests repeated as an ordinary launch
  std::vector<unsigned char> cs_last_req;       // the pending request (a ColStatInline), kept for that repeat
  DevBuf<double> cs_part;                       // per-slice results kept on the device (many slices: summed by the last work item)
  DevBuf<unsigned int> cs_ticket;
  bool cs_p_reduced = false;
  int cs_p_mc = 0, cs_p_k = 0, cs_p_nseg = 0;
  unsigned int cs_p_seq = 0;
};

Chain* chain_create(Store* s);
void chain_destroy(Chain* c);
void chain_set_missing(Chain* c, int64_t snp, const int8_t* vals, int64_t count);
void chain_set_missing_many(Chain* c, const int64_t* snps, int count, const int8_t* const* vals);
int chain_overlay_columns(Chain* c, const int64_t* snps, int count, const uint32_t** out, const int8_t* const* host_vals = nullptr,
                          const int32_t* types = nullptr);
void chain_overlay_invalidate(Chain* c, const int64_t* keep, int k);
void chain_set_missing_all(Chain* c, const int8_t* vals, int64_t count);
void chain_impute_from_prior(Chain* c, const int64_t* loci, int k, uint64_t seed, uint64_t counter);
void chain_get_cells(Chain* c, const int64_t* loci, int k, const int32_t* rows, int64_t q, int8_t* out);
void chain_residual(Chain* c, const int64_t* loci, const double* beta_e, const double* beta_g, int k, double* stats9,
                    const int32_t* term_types = nullptr);
void chain_scan_types(Chain* c, const int64_t* loci, const int32_t* loci_type, const double* beta2, const double* tau2, int k,
                      const bmg_scan_types_params* prm, double* p_r_host, double* p_r_types_host);
void chain_scan_dots(Chain* c);
void chain_set_sharded(Chain* c, int world, int rank, int64_t stride, AllGatherFn fn, void* ctx);
void chain_set_group(Chain* c, Group* g);
void chain_bind_missing(Chain* c, const MissView& v);
// collective over the ranks of a sharded data set: genotype counts and missing-call index of ALL SNPs on every rank
GlobalMissing* build_global_missing(Store* s, int world, int rank, int64_t stride, AllGatherFn fn, void* ctx);
void scan_timer_begin(Chain* c, cudaStream_t st);
void scan_timer_end(Chain* c, cudaStream_t st);
// a chain that will send per-move requests (a sampler's chain): the servers of all such chains on one device split its SMs
void chain_expect_server(Chain* c);
void chain_forget_server(Chain* c);
void imma_launch_on(Chain* geom, const uint4* q, const int* scale_exp, double* out, bool het, cudaStream_t st, Chain* timed);
bool imma_launch2_on(Chain* geom, const uint4* q0, const int* scale_exp0, double* out0, const uint4* q1, const int* scale_exp1,
                     double* out1, cudaStream_t st, Chain* timed);
void chain_allgather(Chain* c, void* dev_buffer, int elem_bytes);
void imma_prepare(Chain* c);
void imma_quantize(Chain* c);
void imma_launch(Chain* c, bool het = false);
void chain_scan(Chain* c, const int64_t* loci, const double* beta_g, const double* tau_g, int k,
                const bmg_scan_params* prm, double* p_r_host);
void chain_adapt(Chain* c, int update_rao, int64_t n_rao_mean, int update_prop, int64_t n_prop_mean, double q_add_min,
                 double q_rem_min);
void chain_init_flat(Chain* c, double value, double q_add_min, double q_rem_min);
void chain_partial_cdf(Chain* c);
void chain_sample(Chain* c, int which, double u01, int64_t* snp, double* total);
void chain_set_zeroed(Chain* c, int which, int64_t snp, int flag);
void chain_fill_zeroed(Chain* c, int which, int flag);
void chain_column_stats(Chain* c, const int64_t* cand, int m_c, const int64_t* loci, int k, double* xy, double* xe,
                        double* xx_model, double* xx_cand);
void chain_column_stats_launch(Chain* c, const int64_t* cand, int m_c, const int64_t* loci, int k, double* xy, double* xe,
                               double* xx_model, double* xx_cand, bool launch_only);
void chain_server_stop(Chain* c);
void chain_column_stats_wait(Chain* c, double* xy, double* xe, double* xx_model, double* xx_cand);
void chain_probit_update(Chain* c, const uint8_t* is_case, const double* u01, uint64_t seed, uint64_t counter,
                         double* stats2, double* ez = nullptr);

}  // namespace bmg
