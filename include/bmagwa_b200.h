/* bmagwa_b200.h -- C ABI of the B200-native BMAGWA hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, int status
 * (0 = ok, non-zero = error, text via bmg_last_error()), no C++/torch types.  The reference has
 * no FFI of its own; each entry point below replaces the C++ member calls cited next to it
 * (paths relative to the reference tree).  A reference maintainer would bind these from
 * data.cpp / data_model.cpp / sampler.cpp / model.hpp as shown in INTEGRATION.md.
 *
 * Ownership and threading follow the reference (data.hpp:33-38, main.cpp:70-95):
 *   - a bmg_store is created once, is immutable afterwards and may be shared, read-only, by
 *     any number of chains/threads;
 *   - a bmg_chain owns all per-chain mutable device state (imputed genotypes, residual,
 *     p_r / p_rao / p_proposal, proposal weights) and one CUDA stream; calls on one chain must
 *     come from one host thread at a time;
 *   - every host pointer argument is ordinary (pageable or pinned) host memory unless the
 *     name says *_dev.
 * All SNP indices are GLOBAL (0 .. m_g-1); a store may hold only the shard [snp_lo, snp_hi).
 * Effect "type": 0 = A (additive), 1 = H, 2 = D, 3 = R   (data_model.hpp:41).
 */
#ifndef BMAGWA_B200_H
#define BMAGWA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BMG_ABI_VERSION 1

typedef struct bmg_store bmg_store;
typedef struct bmg_chain bmg_chain;
typedef struct bmg_sampler bmg_sampler;

/* ---- library ------------------------------------------------------------------------- */
int bmg_abi_version(void);
/* Thread-local text of the last error raised on the calling thread ("" if none). */
const char* bmg_last_error(void);
/* Number of CUDA devices visible; < 0 on error.  The library has NO CPU fallback: every compute
 * entry point fails with an error when no sm_100 device is present. */
int bmg_device_count(void);
/* Counter of kernels launched by this library in this process (for bench.py's gpu_launches). */
uint64_t bmg_launch_count(void);
/* Bytes this library has copied host->device / device->host in this process (bench.py's e2e). */
void bmg_transfer_bytes(uint64_t* h2d, uint64_t* d2h);

/* ---- genotype store: Data::Data + read_g + recode + handle_missing_g + var/mean --------
 * (data.hpp:45-72, data.cpp:245-273,324-376,403-434) and PrecomputedSNPCovariances
 * (precomputed_snp_covariances.hpp:58-131).
 *
 * bed_payload: the PLINK .bed bytes AFTER the 3-byte header, starting at SNP snp_lo, i.e.
 * (snp_hi - snp_lo) * ceil(n/4) bytes, SNP-major, 4 individuals per byte (data.cpp:40-54).
 * The payload is uploaded, re-coded on the device into the store's own packed 2-bit layout
 * (value-coded: 00=0, 01=1, 10=2; missing cells held in a sparse index), optionally swapped to
 * minor-allele counts, and per-SNP integer counts / moments are computed there.
 * payload_on_device != 0: bed_payload is a device pointer (data generated or staged on the GPU). */
int bmg_store_create(const uint8_t* bed_payload, int payload_on_device, int64_t n, int64_t m_g,
                     int64_t snp_lo, int64_t snp_hi, int recode_to_minor, int device, bmg_store** out);
/* Same store built straight from a PLINK .bed FILE (Data::read_g, data.cpp:245-273: header 0x6C 0x1B 0x01 then
 * SNP-major payload): the SNPs [snp_lo, snp_hi) are streamed to the device through two pinned staging buffers
 * (file read of block b+1 overlaps the H2D copy of block b); no host copy of the payload is kept.  Errors carry
 * the reference's messages ("BED file could not be opened", "BED file not recognised (magic number does not
 * match)", "BED file not in snp-major format", "Reading the BED file failed"). */
int bmg_store_create_from_bed(const char* bed_path, int64_t n, int64_t m_g, int64_t snp_lo, int64_t snp_hi,
                              int recode_to_minor, int device, bmg_store** out);
int bmg_store_destroy(bmg_store* s);

/* Phenotype and covariates (Data::_y, Data::_e; data.hpp:75-76).  e is n x m_e column-major and
 * INCLUDES the leading column of ones the reference adds (data.hpp:50,56), so m_e >= 1. */
int bmg_store_set_phenotype(bmg_store* s, const double* y, const double* e, int m_e);

int bmg_store_dims(const bmg_store* s, int64_t* n, int64_t* m_g, int64_t* snp_lo, int64_t* snp_hi,
                   int* m_e, int64_t* n_missing_cells);
/* Per local SNP: counts of genotype 1 and 2 among observed cells, number of missing cells, and
 * whether recode swapped it (any pointer may be NULL). */
int bmg_store_counts(const bmg_store* s, int32_t* n1, int32_t* n2, int32_t* n_miss, uint8_t* swapped);
/* Data::var_x / mean_x (data.cpp:403-434) and var_y / yy (data.hpp:67-70) over the LOCAL shard:
 * out = {sum over SNPs of per-SNP mean, number of SNPs in that sum, sum of per-SNP variance,
 * number in that sum, var_y, yy}; a multi-shard caller adds the first four across shards. */
int bmg_store_summaries(const bmg_store* s, double* out6);
/* Moment cache for types={A}: xx[2*j] = sum x, xx[2*j+1] = sum x^2 - (sum x)^2/n, missing = 0
 * (precomputed_snp_covariances.hpp:113-117).  j is local. */
int bmg_store_moments(const bmg_store* s, double* xx);
/* Missing index (Data::miss_loc / miss_prior, data.cpp:341-376): offsets[local m + 1], idx
 * (row indices, may be NULL to size), prior3 (3 cumulative counts per local SNP, may be NULL). */
int bmg_store_missing(const bmg_store* s, int64_t* offsets, int64_t* idx, double* prior3);
/* Data::get_genotypes_<type>(snp, v) (data.cpp:56-138): n doubles, missing = -1. */
int bmg_store_get_column(const bmg_store* s, int64_t snp, int type, double* out);
/* Device pointer / stride of the packed shard (for peers reading columns over NVLink);
 * ipc_handle receives the 64-byte cudaIpcMemHandle_t of the allocation. */
int bmg_store_export(const bmg_store* s, void* ipc_handle64, int64_t* words_per_snp);
/* Attach a peer shard [snp_lo, snp_hi) living on another GPU/process (opened from its IPC
 * handle, or a raw device pointer when same_process != 0). */
int bmg_store_attach_peer(bmg_store* s, const void* ipc_handle64_or_ptr, int same_process,
                          int64_t snp_lo, int64_t snp_hi);

/* ---- chain: DataModel + RaoBlackwellizer + the arrays Sampler keeps per chain ---------- */
int bmg_chain_create(bmg_store* s, bmg_chain** out);
int bmg_chain_destroy(bmg_chain* c);
/* Blocks until all work queued on the chain's stream is complete. */
int bmg_chain_sync(bmg_chain* c);
/* The chain's CUDA stream as a cudaStream_t value (for event timing by the caller). */
void* bmg_chain_stream(bmg_chain* c);

/* DataModel::miss_val (data_model.hpp:127-131): imputed values (0,1,2) of ALL missing cells of
 * local SNP `snp`, in miss_loc order. */
int bmg_chain_set_missing(bmg_chain* c, int64_t snp, const int8_t* vals, int64_t count);
/* DataModel::sample_missing (data_model.cpp:78-90, re-imputation of every SNP before a scan): the imputed values
 * of ALL missing cells of the local shard in one upload, in the order of bmg_store_missing's idx array. */
int bmg_chain_set_missing_all(bmg_chain* c, const int8_t* vals, int64_t count);
/* The same re-imputation drawn on the device (throughput mode): every missing cell of every local SNP that is NOT in
 * loci[0..k) gets a value from its SNP's prior (cumulative counts of 0/1/2 among the observed cells, data.cpp:357-372)
 * with a counter-based uniform keyed by (seed, counter, cell).  Not the chain's own stream, so traces differ from the
 * reference's while the stationary distribution is the same. */
int bmg_chain_impute_from_prior(bmg_chain* c, const int64_t* loci, int k, uint64_t seed, uint64_t counter);
/* DataModel::get_genotypes_<type>(snp, v) (data_model.cpp:30-72): overlay applied. */
int bmg_chain_get_column(bmg_chain* c, int64_t snp, int type, double* out);
/* The additive value (0, 1, 2; overlay applied) of a few cells: out[l*q + t] = SNP loci[l] at individual rows[t], in
 * bits 0-1; bit 2 is set when the cell is a missing call (its value is the chain's imputed one), which is
 * DataModel::genotype_missing (data_model.hpp:104-117).  This is what the missing-genotype Gibbs step reads as
 * current_model->x(i_miss, col) (sampler.cpp:304-449); the design matrix itself does not exist here, so the step
 * gathers the few cells it needs. */
int bmg_chain_get_cells(bmg_chain* c, const int64_t* loci, int k, const int32_t* rows, int64_t q, int8_t* out);

/* Fitted values and residual for the current model (what Model::compute_pve leaves in y_hat,
 * model.hpp:345-392, and RaoBlackwellizer takes r = y - y_hat from, sampler.cpp:48-49):
 * y_hat = E beta_e + sum_l beta_g[l] x_{loci[l]}, built on the device from packed columns.
 * stats9 (may be NULL) = {sum r, sum yhat_e, sum yhat_e^2, sum yhat_g, sum yhat_g^2, sum yhat,
 * sum yhat^2, xb'xb, (E beta_e - y)'xb}: the reductions behind compute_pve and
 * Prior::sample_alpha (prior.cpp:47-58). */
int bmg_chain_residual(bmg_chain* c, const int64_t* loci, const double* beta_e, const double* beta_g,
                       int k, double* stats9);
/* Copy the residual (n doubles) back, for tests. */
int bmg_chain_get_residual(bmg_chain* c, double* r);

typedef struct bmg_scan_params {
  double sigma2;        /* cmodel->sigma2                                   sampler.cpp:39   */
  double lmp_add;       /* log prior change of adding type A                sampler.cpp:56-59 */
  double lmp_rem;       /* same with one in-model SNP removed               sampler.cpp:61-73 */
  int32_t tau_mode;     /* 0 shared value, 1 per-SNP values drawn by the host in reference
                           order (parity mode), 2 per-SNP values drawn on the device with a
                           counter-based generator (throughput mode)           */
  double tau_shared;    /* inv_tau2_alpha2 when tau_mode == 0               sampler.cpp:78-86 */
  const double* tau_host; /* m_g(local) values when tau_mode == 1          sampler.cpp:99-106 */
  uint64_t tau_seed;    /* tau_mode == 2: key (seed, scan counter)                            */
  uint64_t tau_counter;
  double nu_tau2, s2_tau2, alpha2; /* prior of the per-SNP draw             prior.hpp:192-199 */
} bmg_scan_params;

/* RaoBlackwellizer::p_raoblackwell (sampler.hpp:73-74, sampler.cpp:32-261), single type A, over
 * the LOCAL shard, using the residual left by bmg_chain_residual.  In-model SNPs (loci, with
 * their beta and inv_tau2_alpha2) get the "residual without this SNP" treatment of
 * sampler.cpp:116-149.  p_r stays on the device; pass p_r_host != NULL to also copy it back. */
int bmg_chain_scan(bmg_chain* c, const int64_t* loci, const double* beta_g, const double* tau_g, int k,
                   const bmg_scan_params* prm, double* p_r_host);
/* ---- the scan with several effect types, or one type other than A (sampler.cpp:90-259) ------------------------------
 * Effect types: 0 A, 1 H, 2 D, 3 R, 4 AH (data_model.hpp:41); AH is a SNP with two terms (additive + heterozygous).
 * Fitted values of a model whose terms are typed columns: as bmg_chain_residual, with term_type[l] in 0..3 for every
 * term (an AH SNP contributes two terms: its SNP index twice, types 0 and 1). */
int bmg_chain_residual_types(bmg_chain* c, const int64_t* loci, const int32_t* term_type, const double* beta_e,
                             const double* beta_g, int k, double* stats9);

typedef struct bmg_scan_types_params {
  double sigma2;
  int32_t n_types;          /* effect types of the run ...                                      data_model.hpp:46-47 */
  int32_t types[5];         /* ... in increasing code order (the reference sorts model.types, options.hpp:254)      */
  double lmp_add[5];        /* by type code: log prior change of adding a SNP of that type      sampler.cpp:56-59    */
  double lmp_rem[25];       /* [5 * type of the in-model SNP + type]: the same with that SNP removed  sampler.cpp:61-73 */
  int32_t tau_mode;         /* 0: one value per TERM type (tau_shared); 1: per-SNP values drawn by the host in the
                               reference's order -- per SNP one draw per allowed term, terms in increasing code
                               (sampler.cpp:99-106)                                                                   */
  double tau_shared[4];     /* by term type A, H, D, R                                          sampler.cpp:78-86    */
  const double* tau_host;   /* tau_mode 1: m(local) x n_terms                                                          */
  int32_t reference_offsets;/* 1: read the moment cache exactly as the reference does.  Its offset_type[t] counts TERMS
                               although a term occupies two slots (precomputed_snp_covariances.hpp:73-83 vs :113-118), so
                               with several term types every type but the first reads (sum, variance) one slot early;
                               1 reproduces the reference's numbers, 0 uses the intended layout.                      */
} bmg_scan_types_params;

/* RaoBlackwellizer::p_raoblackwell for any configuration of effect types, over the local shard, using the residual left
 * by bmg_chain_residual[_types].  Model SNP l has effect type loci_type[l] (0..4), coefficients beta2[2l], beta2[2l+1] and
 * prior precisions tau2[2l], tau2[2l+1] (second entries used by AH only).  Both sums every type needs, sum_{x=1} r and
 * sum_{x=2} r, come from two passes of the tensor-core scan kernel over the packed store (additive and heterozygote
 * operands); p_r (m) always, p_r_types (m x n_types, the normalised type distribution of sampler.cpp:240-250) when
 * n_types > 1.  Either host pointer may be NULL. */
int bmg_chain_scan_types(bmg_chain* c, const int64_t* loci, const int32_t* loci_type, const double* beta2,
                         const double* tau2, int k, const bmg_scan_types_params* prm, double* p_r_host,
                         double* p_r_types_host);

/* Only the x_j . r reductions of the scan (dot[local m]); the roofline kernel in isolation. */
int bmg_chain_scan_dots(bmg_chain* c, double* dot_host);
/* Kernel variant used by bmg_chain_scan/_dots: 0 = fp64, direct vectorised global loads;
 * 1 = fp64, bulk-async (TMA) staging through shared memory; 2 = exact fixed-point limbs on the
 * integer tensor cores (IMMA) with TMA staging.  Default chosen by the library. */
int bmg_chain_set_scan_variant(bmg_chain* c, int variant);

/* CUDA-event timing of the scan's reduction kernel on the chain's stream (bench.py roofline).
 * enable != 0 turns recording on for later launches; ms_total / launches (may be NULL) receive the
 * sum of launch durations and the number of launches recorded so far (this synchronises the
 * stream); reset != 0 clears the totals afterwards. */
int bmg_chain_scan_kernel_time(bmg_chain* c, int enable, double* ms_total, int64_t* launches, int reset);

/* The scan epilogue of Sampler::sample (sampler.cpp:739-803): running means and weights.
 *   update_rao:      p_rao      = running mean of p_r (n_rao_mean samples so far)
 *   update_proposal: p_proposal = running mean of p_r (n_prop_mean samples so far), then
 *                    q_add = max(p_proposal, q_add_min), q_rem = max(1 - p_proposal, q_rem_min)
 * and the partial CDFs (block sums of q_add, q_rem over the in-order permutation, D5). */
int bmg_chain_adapt(bmg_chain* c, int update_rao, int64_t n_rao_mean, int update_proposal,
                    int64_t n_prop_mean, double q_add_min, double q_rem_min);
/* Sampler::initialize_p_proposal_flat (sampler.hpp:351-362) + the weight set-up of
 * sampler.cpp:594-598. */
int bmg_chain_init_proposal_flat(bmg_chain* c, double value, double q_add_min, double q_rem_min);
/* which: 0 p_r, 1 p_rao, 2 p_proposal, 3 q_add, 4 q_rem (local m doubles each). */
int bmg_chain_get_array(bmg_chain* c, int which, double* out);
/* Partial CDFs: number of blocks, block size (in in-order positions) and the per-block sums of
 * q_add / q_rem (either may be NULL); the last entry of a running sum of these is total_w(). */
int bmg_chain_partial_cdf(bmg_chain* c, int64_t* n_blocks, int64_t* block_size, double* add_sums,
                          double* rem_sums);
/* DiscreteDistribution::sample() (discrete_distribution.hpp:125-153) evaluated on the device:
 * which 0 = dd_add, 1 = dd_rem; u01 in [0,1); zeroed items are excluded.  Returns the sampled
 * global SNP and the total weight used. */
int bmg_chain_sample(bmg_chain* c, int which, double u01, int64_t* snp, double* total_w);
/* adddate / remdate (discrete_distribution.hpp:156-201): zero / un-zero an item. */
int bmg_chain_set_zeroed(bmg_chain* c, int which, int64_t snp, int zeroed);
/* Zero every item of dd_rem (sampler.cpp:601-605) / clear all flags. */
int bmg_chain_fill_zeroed(bmg_chain* c, int which, int zeroed);

/* Model::update_likelihood_on_add's data reductions (model.hpp:453-470) for m_c candidate SNPs
 * in one launch: xy[c] = x_c'y; xe[c*m_e + j] = x_c'E_j; xx_model[c*k + l] = x_c'x_{loci[l]};
 * xx_cand[c*m_c + d] = x_c'x_d (d = c gives x_c'x_c).  Genotype-by-genotype products are exact
 * integers (popcount arithmetic on the packed columns). */
int bmg_chain_column_stats(bmg_chain* c, const int64_t* cand, int m_c, const int64_t* loci, int k,
                           double* xy, double* xe, double* xx_model, double* xx_cand);

/* Probit latent-variable update (NEW; no reference counterpart):
 * z_i ~ N(yhat_i, 1) truncated to (0, inf) for cases and (-inf, 0] for controls, using the
 * fitted values left by bmg_chain_residual; the working phenotype of the chain becomes z.
 * is_case == NULL keeps the labels of the previous call.  u01 == NULL draws the uniforms on the
 * device (seed, counter); otherwise n host uniforms are used (parity with the oracle).
 * stats2 (may be NULL) = {sum z, sum z^2}. */
int bmg_chain_probit_update(bmg_chain* c, const uint8_t* is_case, const double* u01, uint64_t seed,
                            uint64_t counter, double* stats2);
int bmg_chain_get_phenotype(bmg_chain* c, double* y_out);

/* ---- host sampler: Options + Sampler::sample (options.hpp, sampler.cpp:551-880) --------
 * The C++ MH driver of the reference re-implemented over the entry points above; reads the
 * reference's INI file and writes the reference's output files (Appendix C of SURVEY.md). */
int bmg_sampler_create(const char* ini_path, int chain_index, int device, bmg_sampler** out);
/* Same, over a store that already exists (chains share it, main.cpp:54-76). */
int bmg_sampler_create_on_store(const char* ini_path, int chain_index, bmg_store* s, bmg_sampler** out);

/* One chain over a SNP-SHARDED store on the G GPUs of a box (the reference has no counterpart: its
 * only parallelism is one thread per chain, main.cpp:70-95).  One process per GPU; rank r holds the packed SNPs
 * [r*snp_stride, min(m_g, (r+1)*snp_stride)) in `shard` (bmg_store_create / _from_bed with that range, phenotype set,
 * every other rank's shard attached with bmg_store_attach_peer so that column statistics can read any SNP's packed
 * column over NVLink).  Every rank runs the SAME seeded sampler in lockstep: only the genotype scan is sharded; its
 * per-SNP result (8 bytes per SNP) is all-gathered through `allgather`, everything downstream is replicated and
 * bit-identical, so the chain equals the single-GPU chain draw for draw.  Data with missing genotype calls are handled:
 * every rank keeps the chain's imputed values for ALL SNPs over the whole data set's missing-call index, which the ranks
 * exchange once at creation through `allgather` (data_model.cpp:78-167).
 * allgather(ctx, dev_buffer, elems_per_rank, elem_bytes, cuda_stream): dev_buffer holds world*elems_per_rank elements;
 * rank r's block [r*elems_per_rank, (r+1)*elems_per_rank) is valid on entry, all blocks must be valid after the
 * work enqueued on cuda_stream (ncclAllGather in place, or torch.distributed.all_gather_into_tensor).  Returns 0 on
 * success.  Called from the thread inside bmg_sampler_create_sharded / bmg_sampler_run. */
typedef int (*bmg_allgather_fn)(void* ctx, void* dev_buffer, int64_t elems_per_rank, int elem_bytes, void* cuda_stream);
struct bmg_shard_comm {
  int world, rank;
  int64_t snp_stride;
  bmg_allgather_fn allgather;
  void* ctx;
};
int bmg_sampler_create_sharded(const char* ini_path, int chain_index, bmg_store* shard, const struct bmg_shard_comm* comm,
                               bmg_sampler** out);
/* SEVERAL chains over ONE SNP-sharded store (BASELINE configs[4]; the reference's chains are threads sharing one
 * `const Data*`, main.cpp:54-85).  A shard group is one process per GPU of a box: rank r holds the shard of
 * bmg_sampler_create_sharded (phenotype set, peers attached) and, for r < n_chains, the host sampler of chain r -- nothing
 * of a chain is replicated.  Per iteration a chain only talks to its own GPU (remote columns over the peer mappings).
 * At every scan all ranks meet (the chains of a group must run the same schedule: n_rao, iteration counts): each rank
 * scans its shard once per chain, and every chain pulls its dot products from all ranks through CUDA-IPC peer memory;
 * the ranks synchronise through a POSIX shared-memory segment `shm_name` ("/name", unique per job; created by rank 0 and
 * unlinked as soon as every rank has mapped it).  No host callback and no NCCL on the data path.  Every chain writes the
 * bytes of its single-GPU run.
 * bmg_group_create and bmg_group_destroy are collective (every rank calls them); chain_index of
 * bmg_sampler_create_grouped must equal the rank.  Ranks without a chain (n_chains <= rank) call bmg_group_serve(g, n)
 * to take part in the next n scans.  Passing comm->allgather == NULL and comm->ctx = a bmg_group* to
 * bmg_sampler_create_sharded runs the lockstep single chain over the group's native all-gather instead of a host one. */
typedef struct bmg_group bmg_group;
int bmg_group_create(bmg_store* shard, int world, int rank, int n_chains, int64_t snp_stride, const char* shm_name, bmg_group** out);
int bmg_sampler_create_grouped(const char* ini_path, int chain_index, bmg_store* shard, bmg_group* g, bmg_sampler** out);
int bmg_group_serve(bmg_group* g, int64_t n_rounds);
/* the chain-like handle this rank's share of the scans runs on (bmg_chain_scan_kernel_time, bmg_chain_stream) */
bmg_chain* bmg_group_scan_chain(bmg_group* g);
/* out[0..3] = {scan rounds this rank took part in, seconds this rank's chain spent in its scans (waiting for the other
 * chains at the barriers included), scans of this rank's chain, seconds in the barriers} */
int bmg_group_stats(bmg_group* g, double* out4);
int bmg_group_destroy(bmg_group* g);
/* This rank's shard of the data set an INI file names: SNPs [snp_lo, snp_hi) of datafiles.file_g streamed to `device`
 * (snp_hi < 0: up to sizes.m_g; device < 0: [b200] device), re-coded as datafiles.recode_g_to_minor_allele_count says, with
 * the phenotype and covariates of file_fam / file_y / file_e set -- what Data's constructor does for the whole file
 * (src/data.hpp:45-72, src/data.cpp:245-273,324-434), per shard.  The store is ready for bmg_store_export /
 * bmg_group_create / bmg_sampler_create_grouped (or, with the whole SNP range, bmg_sampler_create_on_store). */
int bmg_store_create_from_ini(const char* ini_path, int64_t snp_lo, int64_t snp_hi, int device, bmg_store** out);
/* Value of `key` in [section] of an INI file, read with the library's own parser (the inih rules the reference follows,
 * src/inih/ini.c:60-140: case-insensitive names, '=' or ':' separators, continuation lines, " ;" comments, 199-character
 * lines); `dflt` when absent.  Copies at most out_len - 1 characters.  For hosts that need n_threads / do_n_iter
 * before constructing samplers (src/main.cpp:47-52). */
int bmg_ini_lookup(const char* ini_path, const char* section, const char* key, const char* dflt, char* out, int out_len);
/* Overrides applied after the INI file, before bmg_sampler_begin.  Keys: "tau_rng" = host | device (per-SNP tau2 draws
 * of the scan from the chain's stream in reference order, or Philox on the device); "missing_rng" = host | device (the
 * re-imputation of missing calls before each scan likewise; defaults to tau_rng); "pip_burnin" = thinned samples to drop
 * before the running inclusion counts start (bmg_sampler_inclusion_counts); "basename" (output files);
 * "verbosity"; "reference_quirks" = 1 | 0 (keep the reference's stale-y_hat behaviour, model.hpp:345-392);
 * "scan_variant" = 2 | 1 | 0; "probit" = 1 (0/1 phenotype, Albert-Chib latent updates on the device; no reference
 * counterpart); "colstats_server" = 1 | 0 (serve the per-move column statistics from one persistent kernel fed through
 * a host mailbox instead of one launch per move); "gram_cache" = 1 | 0 (keep the device's x_j'y, x_j'E, x_j'x_l results on
 * the host so that a move whose SNPs were all seen before needs no device request; replaces the recomputation of
 * src/model.hpp:453-470, chain files are byte-identical either way).  Unknown keys are an error. */
int bmg_sampler_set_option(bmg_sampler* sp, const char* key, const char* value);
/* Opens output files, initialises the chain (sampler.cpp:592-620). */
int bmg_sampler_begin(bmg_sampler* sp);
/* Runs n_iter MCMC iterations (the loop body of sampler.cpp:626-834). */
int bmg_sampler_run(bmg_sampler* sp, int64_t n_iter);
/* Writes _rao.dat, _samplerstats.txt, closes files (sampler.cpp:836-879). */
int bmg_sampler_end(bmg_sampler* sp);
/* stats: {iterations done, accepted, model size, log likelihood, seconds in moves,
 * seconds in scans, scans done, seconds of the move time spent waiting for per-proposal column statistics}. */
int bmg_sampler_stats(bmg_sampler* sp, double* out8);
/* More counters of the same run (the reference's SamplerStats has no counterpart for these, src/samplerstats.hpp:33-128):
 * out[0..11] = {moves that needed column statistics, of which served from the host memo without a device request,
 * of which asked the device for a subset of their SNPs, device requests, SNP pairs held by the memo, requests the
 * persistent server left unserved and an ordinary launch repeated, seconds in move-0 delayed rejection, delayed-rejection
 * events, seconds in the scan epilogue (adaptation + weights to the host), seconds in the missing-genotype Gibbs step,
 * probit sweeps, reserved}.  n = number of doubles `out` holds (the first min(n, 12) are written). */
int bmg_sampler_counters(bmg_sampler* sp, double* out, int n);
/* Running MCMC inclusion counts: counts[j] (m_g entries, may be NULL) = number of thinned samples, after the first
 * "pip_burnin" of them, whose model contains SNP j; n_samples = how many samples were counted.  counts / n_samples is what
 * `bmagwa_postprocess.py mcmcpos basename m_g burnin 1` recomputes offline from _loci.dat and _modelsize.dat
 * (bmagwa_postprocess.py:79-123); chains are merged by averaging (ibid. :118-123; bmagwa_b200.postprocess.merge_chains
 * does it with one all-reduce across the ranks of a chain-per-GPU run). */
int bmg_sampler_inclusion_counts(bmg_sampler* sp, uint32_t* counts, int64_t* n_samples);
bmg_store* bmg_sampler_store(bmg_sampler* sp);
bmg_chain* bmg_sampler_chain(bmg_sampler* sp);
int bmg_sampler_destroy(bmg_sampler* sp);

#ifdef __cplusplus
}
#endif
#endif /* BMAGWA_B200_H */
