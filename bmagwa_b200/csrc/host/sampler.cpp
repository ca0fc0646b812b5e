// sampler.cpp -- see sampler.hpp.  Citations are to the reference tree (src/...).
#include "sampler.hpp"
#include <time.h>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <cstdlib>

namespace bmg {

namespace {
const double kLogHalf = -0.69314718055994528622676398299518041312694549560546875;  // sampler.hpp:39

double wall_seconds()
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

// Utils::sample_discrete_naive (utils.cpp:46-57)
int sample_discrete_naive(const double* cumsum, int m, ChainRng& rng)
{
  const double z = cumsum[m - 1];
  const double r = rng.u01() * z;
  for (int i = 0; i < m; ++i)
    if (r < cumsum[i]) return i;
  throw std::logic_error("sample_discrete_naive reached end, exiting");
}
// Utils::sample_discrete (utils.cpp:73-101): `level` bisection steps, then a linear search
size_t sample_discrete(const double* cumsum, size_t m, int level, ChainRng& rng)
{
  size_t a = 0, b = m - 1;
  const double z = cumsum[b];
  const double r = rng.u01() * z;
  while (level > 0) {
    const size_t c = (a + b) / 2;
    if (r < cumsum[c]) b = c; else a = c + 1;
    --level;
  }
  for (size_t i = a; i <= b; ++i)
    if (r < cumsum[i]) return i;
  throw std::logic_error("sample_discrete reached end");
}
// Utils::geometric_dist_cdf (utils.cpp:144-152): un-normalised cdf of a geometric truncated at maxsize
void geometric_dist_cdf(int maxsize, double p, std::vector<double>& values)
{
  values.assign(maxsize, 0.0);
  double q = 1 - p;
  const double qp = q;
  for (int i = 0; i < maxsize; ++i) { values[i] = 1 - q; q *= qp; }
}
}  // namespace

struct Sampler::Files {
  std::ofstream log, loci, modelsize, jumpdistance, log_likelihood, log_prior, move_type, move_size, pve, alpha, sigma2, beta;
  static void open(std::ofstream& f, const std::string& name, bool binary)
  {
    if (binary) f.open(name.c_str(), std::ios::binary); else f.open(name.c_str());
    if (!f.is_open()) throw std::runtime_error("Failed to open file: " + name);
  }
};

// ------------------------------------------------------------------------------------------------
// construction (sampler.hpp:107-282)
// ------------------------------------------------------------------------------------------------
Sampler::Sampler(const Options& opts, int chain_index, Store* store, const std::vector<double>& y, const std::vector<double>& e,
                 double var_y, double yy, double var_x, double mean_x, const ShardComm* comm)
: opt_(opts), basename_(opts.basename + std::to_string(chain_index)), seed_(opts.seeds.at(chain_index)), n_(opts.n),
  m_g_(opts.m_g), m_e_(opts.m_e + 1), n_rao_(opts.n_rao), n_rao_burnin_(opts.n_rao_burnin),
  n_sample_tau2_and_missing_(opts.n_sample_tau2_and_missing), thin_(opts.thin), verbosity_(opts.verbosity),
  flat_proposal_dist_(opts.flat_proposal_dist), adaptation_(opts.adaptation), save_beta_(opts.save_beta),
  tau_on_device_(opts.tau_rng == "device"), store_(store), rng_(opts.seeds.at(chain_index), (double)opts.n + opts.nu_sigma2),
  q_add_min_(1.0 / ((double)opts.m_g - opts.e_qg)), q_rem_min_(1.0 / opts.e_qg),
  max_move_size_((unsigned char)std::min(std::min(opts.m_g, (size_t)255), opts.max_move_size)),
  max_nbh_(opts.max_SNP_neighborhood_size), p_move_size_(opts.p_move_size), p_move_size_nbs_(opts.p_move_size_nbs),
  p_move_size_nbc_(opts.p_move_size_nbc), acpt_move_size_goal_(opts.p_move_size_acpt_goal)
{
  if (opts.sampler_type != 0) throw std::runtime_error("this build implements the PMV sampler only (sampler.type = PMV)");
  if (opts.types.size() != 1 || opts.types[0] != kA) throw std::runtime_error("this build implements effect type A only (model.types = A)");
  if (comm == nullptr && (store_->lo != 0 || store_->hi != (int64_t)m_g_))
    throw std::runtime_error("the host sampler needs the whole SNP range on its store (multi-GPU chains: bmg_sampler_create_sharded)");
  if (comm != nullptr) {
    for (int64_t snp : {(int64_t)0, (int64_t)m_g_ - 1}) (void)store_->column_ptr(snp);   // throws unless every shard is attached
  }
  yy_ = yy;
  adapt_p_move_size_ = opts.adapt_p_move_size && max_move_size_ > 1;
  delay_rejection_ = (unsigned char)std::min((size_t)max_move_size_, opts.delay_rejection);

  // prior scale parameters (sampler.hpp:173-193)
  double s2_sigma2 = opts.s2_sigma2 > 0 ? opts.s2_sigma2 : var_y * (1 - opts.R2mode_sigma2) * (opts.nu_sigma2 + 2) / opts.nu_sigma2;
  double s2_tau2 = opts.s2_tau2[kA] > 0
                       ? opts.s2_tau2[kA]
                       : opts.eh_tau2[kA] * (opts.nu_tau2[kA] - 2) / (opts.nu_tau2[kA] * (var_x + mean_x * mean_x) * (1 - opts.R2mode_sigma2));
  std::vector<double> inv_tau2_e(m_e_, opts.inv_tau2_e_val);
  inv_tau2_e[0] = opts.inv_tau2_e_const_val;
  prior_.reset(new Prior(n_, m_g_, m_e_, yy, opts.types_prior, opts.e_qg, opts.var_qg, inv_tau2_e, opts.nu_sigma2, s2_sigma2,
                         opts.nu_tau2[kA], s2_tau2, opts.mu_alpha, opts.use_individual_tau2));

  // covariate block of the Gram matrix, E'E and E'y (Model ctor, model.hpp:85-88)
  UpperMat exx;
  exx.resize((int)m_e_);
  std::vector<double> exy(m_e_, 0.0);
  for (size_t c = 0; c < m_e_; ++c) {
    const double* ec = &e[c * n_];
    for (size_t r = 0; r <= c; ++r) {
      const double* er = &e[r * n_];
      double s = 0.0;
      for (size_t i = 0; i < n_; ++i) s += er[i] * ec[i];
      exx((int)r, (int)c) = s;
    }
    double s = 0.0;
    for (size_t i = 0; i < n_; ++i) s += ec[i] * y[i];
    exy[c] = s;
  }
  current_.n_updates_add = &n_upd_add_; current_.n_updates_rem = &n_upd_rem_; current_.n_computations = &n_comp_;
  proposal_.n_updates_add = &n_upd_add_; proposal_.n_updates_rem = &n_upd_rem_; proposal_.n_computations = &n_comp_;
  current_.init((int)m_e_, exx, exy, prior_.get());
  proposal_.init((int)m_e_, exx, exy, prior_.get());
  pos_in_current_.assign(m_g_, -1);
  pos_in_proposal_.assign(m_g_, -1);
  incl_count_.assign(m_g_, 0);
  pip_burnin_ = (int64_t)opts.pip_burnin;

  chain_ = chain_create(store_);
  chain_expect_server(chain_);   // the chains of one device share its SMs for their column-statistics servers (colstats.cu)
  if (comm != nullptr && comm->group != nullptr) chain_set_group(chain_, comm->group);
  else if (comm != nullptr) chain_set_sharded(chain_, comm->world, comm->rank, comm->stride, comm->allgather, comm->ctx);
  dd_add_.init(&chain_->inorder_host(), (int)chain_->cdf_block);
  dd_rem_.init(&chain_->inorder_host(), (int)chain_->cdf_block);
  h_w_.alloc(m_g_);
  h_w2_[0].alloc(2 * m_g_); h_w2_[1].alloc(2 * m_g_);
  h_cdf_.assign(2 * chain_->cdf_blocks, 0.0);

  move_inds_add_.assign(max_move_size_, 0); move_inds_rem_.assign(max_move_size_, 0); move_inds_.assign(max_move_size_, 0);
  move_inds_map_.assign(max_move_size_, 0); move_isadd_.assign(max_move_size_, 0);
  for (int i = 0; i < 7; ++i) { n_acpt_moves_[i] = 0; n_moves_[i] = 0; }
  n_acpt_moves_[0] = n_acpt_moves_[1] = n_acpt_moves_[2] = 1;   // "prior" counts (sampler.hpp:236-238)
  n_moves_[0] = n_moves_[1] = n_moves_[2] = 2;
  const size_t half = ((size_t)max_move_size_ + 1) / 2;
  geometric_dist_cdf(max_move_size_, p_move_size_, q_p_move_size_);
  geometric_dist_cdf((int)half, p_move_size_nbs_, q_p_move_size_nbs_);
  geometric_dist_cdf(max_move_size_, p_move_size_nbc_, q_p_move_size_nbc_);
  r_move_size_sum_.assign(max_move_size_, 0.0); r_move_size_n_.assign(max_move_size_, 0.0); q_p0_move_size_.assign(max_move_size_, 0.0);
  if (delay_rejection_ > 0) {
    dr_model_probabilities_.assign((size_t)1 << delay_rejection_, 0.0);
    dr_q_add_.assign(delay_rejection_, 0.0); dr_q_rem_.assign(delay_rejection_, 0.0);
    dr_log_q_add_types_.assign(delay_rejection_, 0.0);
    dr_bit_to_normalized_order_.assign(delay_rejection_, 0);
  }
  y_host_ = &y; e_host_ = &e;
  if (chain_->mv.n_missing > 0) load_missing_index();   // the chain's index: the store's, or the whole data set's when sharded
  if (opts.probit) enable_probit();
}

// ------------------------------------------------------------------------------------------------
// missing genotypes: Data::miss_loc / miss_prior and DataModel::miss_val mirrored on the host
// ------------------------------------------------------------------------------------------------
void Sampler::load_missing_index()
{
  BMG_CUDA(cudaSetDevice(store_->device));
  const MissView& mv = chain_->mv;
  if (mv.base != 0 || mv.m != (int64_t)m_g_) throw std::runtime_error("the chain's missing-call index does not cover all SNPs");
  miss_.off.assign(mv.h_off, mv.h_off + m_g_ + 1);
  miss_.idx.resize((size_t)mv.n_missing);
  bmg::copy_d2h_sync(miss_.idx.data(), mv.idx, miss_.idx.size() * sizeof(int32_t));
  miss_.val.assign(miss_.idx.size(), 0);   // data_model.hpp:80-84; the chain's device copy starts at 0 as well
  std::vector<int32_t> n1(m_g_), n2(m_g_), nm(m_g_);
  bmg::copy_d2h_sync(n1.data(), mv.n1, m_g_ * sizeof(int32_t));
  bmg::copy_d2h_sync(n2.data(), mv.n2, m_g_ * sizeof(int32_t));
  bmg::copy_d2h_sync(nm.data(), mv.nmiss, m_g_ * sizeof(int32_t));
  miss_.prior3.resize(3 * m_g_);
  for (size_t j = 0; j < m_g_; ++j) {      // cumulative counts of 0/1/2 among the observed cells (data.cpp:357-372)
    const double c0 = (double)((int64_t)n_ - nm[j] - n1[j] - n2[j]);
    miss_.prior3[3 * j] = c0;
    miss_.prior3[3 * j + 1] = c0 + (double)n1[j];
    miss_.prior3[3 * j + 2] = c0 + (double)n1[j] + (double)n2[j];
  }
  have_missing_ = true;
}

std::vector<double> Sampler::draw_for_additions(const std::vector<uint32_t>& cand)
{
  std::vector<double> taus(cand.size());
  std::vector<int64_t> snps;
  std::vector<const int8_t*> vals;
  for (size_t i = 0; i < cand.size(); ++i) {
    const uint32_t snp = cand[i];
    if (miss_.count(snp) > 0) {   // DataModel::sample_missing_single (data_model.cpp:95-103)
      miss_.draw_from_prior(snp, rng_);
      snps.push_back(snp);
      vals.push_back(miss_.val.data() + miss_.off[snp]);
    }
    taus[i] = prior_->draw_inv_tau2_alpha2(rng_);
  }
  // one launch for the whole move: values to the device and the SNPs' packed columns with the values filled in
  if (!snps.empty()) chain_set_missing_many(chain_, snps.data(), (int)snps.size(), vals.data());
  return taus;
}

// Sampler::sample_missing (sampler.cpp:264-453): the arithmetic is in missing.hpp; here the few touched rows of the
// in-model columns are gathered from the device and the new imputed values are sent back
void Sampler::sample_missing()
{
  if (!have_missing_ || current_.size() == 0) return;
  const double t0 = wall_seconds();
  struct Toc { double& acc; double t0; ~Toc() { acc += wall_seconds() - t0; } } toc{gibbs_seconds_, t0};
  rows_missing_in_model(miss_, current_.loci, gibbs_rows_);
  if (gibbs_rows_.empty()) return;
  const int k = (int)current_.size();
  std::vector<int64_t> loci(current_.loci.begin(), current_.loci.end());
  gibbs_cells_.resize((size_t)k * gibbs_rows_.size());
  chain_get_cells(chain_, loci.data(), k, gibbs_rows_.data(), (int64_t)gibbs_rows_.size(), gibbs_cells_.data());
  const double* yv = y_host_->data();
  if (probit_) {   // the working phenotype is the latent z on the device
    y_work_.resize(n_);
    BMG_CUDA(cudaSetDevice(store_->device));
    bmg::copy_d2h(y_work_.data(), chain_->y.p, n_ * sizeof(double), chain_->stream);
    BMG_CUDA(cudaStreamSynchronize(chain_->stream));
    yv = y_work_.data();
  }
  gibbs_missing_in_model(current_, miss_, gibbs_rows_, gibbs_cells_.data(), yv, e_host_->data(), n_, yy_, rng_, gibbs_slot_);
  std::vector<int64_t> snps;
  std::vector<const int8_t*> vals;
  for (uint32_t snp : current_.loci)
    if (miss_.count(snp) > 0) { snps.push_back(snp); vals.push_back(miss_.val.data() + miss_.off[snp]); }
  chain_set_missing_many(chain_, snps.data(), (int)snps.size(), vals.data());
}

// ------------------------------------------------------------------------------------------------
// probit mode: Albert & Chib latent phenotype on the device (k_probit), sigma2 == 1
// ------------------------------------------------------------------------------------------------
void Sampler::enable_probit()
{
  if (probit_) return;
  if (current_.size() != 0) throw std::runtime_error("probit mode must be enabled before the chain starts");
  const std::vector<double>& y = *y_host_;
  const std::vector<double>& e = *e_host_;
  is_case_.resize(n_);
  std::vector<double> z(n_);
  size_t n_case = 0;
  for (size_t i = 0; i < n_; ++i) {
    if (y[i] != 0.0 && y[i] != 1.0) throw std::runtime_error("probit mode needs a 0/1 phenotype (file_y)");
    is_case_[i] = y[i] > 0.5 ? 1 : 0;
    n_case += is_case_[i];
    z[i] = is_case_[i] ? 0.7978845608028654 : -0.7978845608028654;   // E|N(0,1)|: start of the latent phenotype
  }
  if (n_case == 0 || n_case == n_) throw std::runtime_error("probit mode needs both cases and controls");
  double yy = 0.0;
  std::vector<double> exy(m_e_, 0.0);
  for (size_t i = 0; i < n_; ++i) yy += z[i] * z[i];
  for (size_t c = 0; c < m_e_; ++c) {
    double s = 0.0;
    for (size_t i = 0; i < n_; ++i) s += e[c * n_ + i] * z[i];
    exy[c] = s;
  }
  prior_->set_fixed_sigma2(yy);
  yy_ = yy;
  for (size_t c = 0; c < m_e_; ++c) { current_.xy[c] = exy[c]; proposal_.xy[c] = exy[c]; }
  current_.sigma2 = 1.0;
  current_.compute_log_likelihood();
  proposal_.assign(current_);
  BMG_CUDA(cudaSetDevice(store_->device));
  bmg::copy_h2d(chain_->y.p, z.data(), n_ * sizeof(double), chain_->stream);
  BMG_CUDA(cudaStreamSynchronize(chain_->stream));
  chain_->residual_valid = false;
  chain_->imma_q_valid = false;
  probit_ = true;
}

void Sampler::probit_sweep()
{
  const int k = (int)current_.size();
  std::vector<int64_t> loci(current_.loci.begin(), current_.loci.end());
  // fitted values of the CURRENT coefficients (fresh from sample_beta_sigma2), not the quirk-tracked ones of the scan
  chain_residual(chain_, loci.data(), current_.beta.data(), current_.beta.data() + m_e_, k, nullptr);
  double st[2];
  std::vector<double> ez(m_e_, 0.0);
  chain_probit_update(chain_, probit_labels_sent_ ? nullptr : is_case_.data(), nullptr, (uint64_t)seed_ * 0x9E3779B97F4A7C15ull + 17u,
                      ++probit_counter_, st, ez.data());
  probit_labels_sent_ = true;
  for (size_t c = 0; c < m_e_; ++c) current_.xy[c] = ez[c];
  for (int done = 0; done < k;) {   // X_gamma' z, 64 model columns per launch
    const int cnt = std::min(64, k - done);
    std::vector<double> xy(cnt);
    chain_column_stats(chain_, loci.data() + done, cnt, nullptr, 0, xy.data(), nullptr, nullptr, nullptr);
    for (int j = 0; j < cnt; ++j) current_.xy[m_e_ + done + j] = xy[j];
    done += cnt;
  }
  prior_->set_yy(st[1]);
  yy_ = st[1];
  cache_.bump_phenotype();   // every memoised x'z is stale
  ++n_probit_sweeps_;
}

Sampler::~Sampler()
{
  chain_destroy(chain_);
}

void Sampler::set_option(const std::string& key, const std::string& value)
{
  if (begun_) throw std::runtime_error("bmg_sampler_set_option: options are frozen once the sampler has begun");
  if (key == "tau_rng") {
    if (value != "host" && value != "device") throw std::runtime_error("tau_rng must be host or device");
    tau_on_device_ = value == "device";
  } else if (key == "missing_rng") {
    if (value != "host" && value != "device") throw std::runtime_error("missing_rng must be host or device");
    missing_on_device_ = value == "device";
  } else if (key == "pip_burnin") {
    pip_burnin_ = (int64_t)std::stoll(value);
    if (pip_burnin_ < 0) throw std::runtime_error("pip_burnin must be >= 0");
  } else if (key == "basename") {
    basename_ = value;
  } else if (key == "verbosity") {
    verbosity_ = (size_t)std::stoul(value);
  } else if (key == "reference_quirks") {
    reference_quirks_ = value != "0";
  } else if (key == "probit") {
    if (value != "0") enable_probit();
    else if (probit_) throw std::runtime_error("probit mode cannot be switched off once enabled");
  } else if (key == "gram_cache") {
    cache_.enabled = value != "0";
  } else if (key == "gram_cache_max_pairs") {   // slots of the pair table before it starts over (a power of two; tests)
    cache_.set_max_slots((size_t)std::stoull(value));
  } else if (key == "colstats_server") {
    chain_->server_enabled = value != "0";
    if (chain_->server_enabled) chain_expect_server(chain_); else chain_forget_server(chain_);
  } else if (key == "scan_variant") {
    chain_->scan_variant = std::stoi(value);
  } else {
    throw std::runtime_error("bmg_sampler_set_option: unknown key " + key);
  }
}

void Sampler::initialize_p_proposal_flat()
{
  chain_init_flat(chain_, prior_->e_g() / (double)m_g_, q_add_min_, q_rem_min_);
}

void Sampler::print_prior()
{
  prior_->print(basename_.substr(0, basename_.length() - 1) + "_prior.txt");
}

// ------------------------------------------------------------------------------------------------
// model copies and the proposal model's SNP -> term map
// ------------------------------------------------------------------------------------------------
void Sampler::copy_proposal_to_current()
{
  for (uint32_t s : current_.loci) pos_in_current_[s] = -1;
  current_.assign(proposal_);
  for (size_t i = 0; i < current_.loci.size(); ++i) pos_in_current_[current_.loci[i]] = (int32_t)i;
}
void Sampler::copy_current_to_proposal()
{
  for (uint32_t s : proposal_.loci) pos_in_proposal_[s] = -1;
  proposal_.assign(current_);
  for (size_t i = 0; i < proposal_.loci.size(); ++i) pos_in_proposal_[proposal_.loci[i]] = (int32_t)i;
}

// one device launch: statistics of every SNP the move wants to add against y, E, the current model and each other.
// begin_gram() returns as soon as the kernel is queued; finish_gram() collects the numbers, so host work that does
// not need them (the removals of the move) overlaps the device round trip.
void Sampler::begin_gram(const std::vector<uint32_t>& cand)
{
  gram_.m_e = (int)m_e_;
  gram_.k_cur = (int)current_.loci.size();
  gram_.m_c = (int)cand.size();
  gram_.cand = cand;
  gram_req_.clear();
  if (cand.empty()) return;
  const size_t m_c = cand.size(), k = current_.loci.size();
  gram_.xy.assign(m_c, 0.0);
  gram_.xe.assign(m_c * m_e_, 0.0);
  gram_.xm.assign(m_c * std::max<size_t>(1, k), 0.0);
  gram_.xc.assign(m_c * m_c, 0.0);
  // What the memo (gramcache.hpp) already holds is copied straight into gram_; a candidate with anything missing is
  // asked of the device, together with the other incomplete candidates, in one request.
  const bool memo = cache_.enabled;
  if (memo) {
    if (!cache_.ready()) cache_.init(m_g_, m_e_);
    ++cache_.requests;
    for (size_t c = 0; c < m_c; ++c) {
      const uint32_t snp = cand[c];
      bool complete = cacheable(snp) && cache_.have_snp(snp);
      if (complete) {
        gram_.xy[c] = cache_.xy(snp);
        const double* xe = cache_.xe(snp);
        for (size_t j = 0; j < m_e_; ++j) gram_.xe[c * m_e_ + j] = xe[j];
        gram_.xc[c * m_c + c] = cache_.xx(snp);
        for (size_t l = 0; l < k && complete; ++l)
          complete = cacheable(current_.loci[l]) && cache_.get_pair(snp, current_.loci[l], &gram_.xm[c * k + l]);
        for (size_t d = 0; d < m_c && complete; ++d)
          if (d != c) complete = cacheable(cand[d]) && cache_.get_pair(snp, cand[d], &gram_.xc[c * m_c + d]);
      }
      if (!complete) gram_req_.push_back((int)c);
    }
    if (gram_req_.empty()) { ++cache_.served; return; }
    if (gram_req_.size() < m_c) ++cache_.partial;
  } else {
    for (size_t c = 0; c < m_c; ++c) gram_req_.push_back((int)c);
  }
  const size_t m_r = gram_req_.size();
  gram_c64_.resize(m_r);
  for (size_t i = 0; i < m_r; ++i) gram_c64_[i] = (int64_t)cand[gram_req_[i]];
  gram_l64_.assign(current_.loci.begin(), current_.loci.end());
  req_xy_.assign(m_r, 0.0);
  req_xe_.assign(m_r * m_e_, 0.0);
  req_xm_.assign(m_r * std::max<size_t>(1, k), 0.0);
  req_xc_.assign(m_r * m_r, 0.0);
  const double t0 = wall_seconds();
  chain_column_stats_launch(chain_, gram_c64_.data(), (int)m_r, gram_l64_.data(), (int)k, req_xy_.data(), req_xe_.data(),
                            req_xm_.data(), req_xc_.data(), true);
  device_wait_seconds_ += wall_seconds() - t0;
  ++n_gram_requests_;
}
void Sampler::finish_gram()
{
  if (gram_req_.empty()) return;
  const double t0 = wall_seconds();
  chain_column_stats_wait(chain_, req_xy_.data(), req_xe_.data(), req_xm_.data(), req_xc_.data());
  device_wait_seconds_ += wall_seconds() - t0;
  const size_t m_c = gram_.cand.size(), k = (size_t)gram_.k_cur, m_r = gram_req_.size();
  const bool memo = cache_.enabled;
  for (size_t i = 0; i < m_r; ++i) {
    const size_t c = (size_t)gram_req_[i];
    const uint32_t snp = gram_.cand[c];
    gram_.xy[c] = req_xy_[i];
    for (size_t j = 0; j < m_e_; ++j) gram_.xe[c * m_e_ + j] = req_xe_[i * m_e_ + j];
    for (size_t l = 0; l < k; ++l) gram_.xm[c * k + l] = req_xm_[i * k + l];
    for (size_t j = 0; j < m_r; ++j) gram_.xc[c * m_c + (size_t)gram_req_[j]] = req_xc_[i * m_r + j];
    if (!memo) continue;
    const bool keep = cacheable(snp) && cache_.repeat_visitor(snp);
    if (keep) cache_.put_snp(snp, req_xy_[i], &req_xe_[i * m_e_], req_xc_[i * m_r + i]);
    for (size_t l = 0; l < k; ++l)
      if (keep && cacheable(current_.loci[l])) cache_.put_pair(snp, current_.loci[l], req_xm_[i * k + l]);
    for (size_t j = i + 1; j < m_r; ++j)
      if (keep && cacheable(gram_.cand[(size_t)gram_req_[j]])) cache_.put_pair(snp, gram_.cand[(size_t)gram_req_[j]], req_xc_[i * m_r + j]);
  }
  if (memo && m_r < m_c) {
    // products of a requested candidate c with a candidate d served from the memo: d was complete, so begin_gram copied its
    // product with every other candidate of the move into row d -- take it from there.  (Not from the table: filing the
    // results above may have made a full table start over.)
    for (size_t i = 0; i < m_r; ++i) {
      const size_t c = (size_t)gram_req_[i];
      for (size_t d = 0; d < m_c; ++d) {
        bool requested = false;
        for (size_t j = 0; j < m_r; ++j) requested = requested || (size_t)gram_req_[j] == d;
        if (!requested) gram_.xc[c * m_c + d] = gram_.xc[d * m_c + c];
      }
    }
  }
  gram_req_.clear();
}
void Sampler::fetch_gram(const std::vector<uint32_t>& cand)
{
  begin_gram(cand);
  finish_gram();
}

// Model::add_term for a candidate SNP of this move (model.hpp:453-470 supplies the column; here it comes from gram_)
void Sampler::add_to_proposal(uint32_t snp, double inv_tau2_alpha2)
{
  const int c = gram_.find(snp);
  if (c < 0) throw std::logic_error("add_to_proposal: statistics of the SNP were not fetched");
  const int cols = proposal_.cols();
  std::vector<double> col(cols + 1);
  for (int j = 0; j < (int)m_e_; ++j) col[j] = gram_.xe[(size_t)c * m_e_ + j];
  for (size_t i = 0; i < proposal_.loci.size(); ++i) {
    const uint32_t other = proposal_.loci[i];
    const int pc = pos_in_current_[other];
    if (pc >= 0) col[m_e_ + i] = gram_.xm[(size_t)c * gram_.k_cur + pc];
    else {
      const int d = gram_.find(other);
      if (d < 0) throw std::logic_error("add_to_proposal: proposal holds a SNP unknown to the move");
      col[m_e_ + i] = gram_.xc[(size_t)c * gram_.m_c + d];
    }
  }
  col[cols] = gram_.xc[(size_t)c * gram_.m_c + c];
  pos_in_proposal_[snp] = (int32_t)proposal_.loci.size();
  proposal_.add_term(snp, gram_.xy[c], col.data(), inv_tau2_alpha2);
}

// put a SNP of the CURRENT model (removed by this move) back into the proposal with its old prior precision
// (sampler.hpp:949-965): every statistic is already in the current model's Gram matrix or in gram_
void Sampler::readd_to_proposal(uint32_t snp)
{
  const int pc = pos_in_current_[snp];
  const int xc = (int)m_e_ + pc;
  const int cols = proposal_.cols();
  std::vector<double> col(cols + 1);
  for (int j = 0; j < (int)m_e_; ++j) col[j] = current_.gram(j, xc);
  for (size_t i = 0; i < proposal_.loci.size(); ++i) {
    const uint32_t other = proposal_.loci[i];
    const int po = pos_in_current_[other];
    if (po >= 0) col[m_e_ + i] = current_.gram((int)m_e_ + po, xc);
    else {
      const int d = gram_.find(other);
      if (d < 0) throw std::logic_error("readd_to_proposal: proposal holds a SNP unknown to the move");
      col[m_e_ + i] = gram_.xm[(size_t)d * gram_.k_cur + pc];
    }
  }
  col[cols] = current_.gram(xc, xc);
  pos_in_proposal_[snp] = (int32_t)proposal_.loci.size();
  proposal_.add_term(snp, current_.xy[xc], col.data(), current_.inv_tau2_alpha2[xc]);
}

void Sampler::remove_from_proposal(int model_ind)
{
  const uint32_t snp = proposal_.loci[model_ind];
  pos_in_proposal_[snp] = -1;
  for (size_t i = model_ind + 1; i < proposal_.loci.size(); ++i) --pos_in_proposal_[proposal_.loci[i]];
  proposal_.remove_term(model_ind);
}

// weights: device -> host mirrors (in-order layout + partial CDFs)
void Sampler::refresh_weights_from_device(bool first)
{
  Chain* c = chain_;
  BMG_CUDA(cudaSetDevice(store_->device));
  h_w_cur_ ^= 1;   // the CDFs keep reading the previous buffer until update() below
  double* hw = h_w2_[h_w_cur_].p;
  bmg::copy_d2h(hw, c->q_add_io.p, m_g_ * sizeof(double), c->stream);
  bmg::copy_d2h(hw + m_g_, c->q_rem_io.p, m_g_ * sizeof(double), c->stream);
  bmg::copy_d2h(h_cdf_.data(), c->cdf_add.p, c->cdf_blocks * sizeof(double), c->stream);
  bmg::copy_d2h(h_cdf_.data() + c->cdf_blocks, c->cdf_rem.p, c->cdf_blocks * sizeof(double), c->stream);
  BMG_CUDA(cudaStreamSynchronize(c->stream));
  if (first) {
    dd_add_.update(hw, h_cdf_.data(), true, std::vector<uint32_t>());
    dd_rem_.update(hw + m_g_, h_cdf_.data() + c->cdf_blocks, true, std::vector<uint32_t>());
    dd_rem_.zero_all();   // sampler.cpp:601-605
  } else {
    // zero flags are kept (discrete_distribution.hpp:263-314): dd_add has the model's SNPs zeroed,
    // dd_rem has everything but the model's SNPs zeroed
    dd_add_.update(hw, h_cdf_.data(), true, current_.loci);
    dd_rem_.update(hw + m_g_, h_cdf_.data() + c->cdf_blocks, false, current_.loci);
  }
}

void Sampler::compute_p_moves()  // sampler.cpp:455-515, PMV with one effect type
{
  const double p[7] = {0.7, 0.15, 0.15, 0.0, 0.0, 0.0, 0.0};
  for (int i = 0; i < 7; ++i) p_moves_[i] = p[i];
  p_moves_cumsum_[0] = p_moves_[0];
  for (int i = 1; i < 7; ++i) p_moves_cumsum_[i] = p_moves_cumsum_[i - 1] + p_moves_[i];
}

// ------------------------------------------------------------------------------------------------
// sample(): set-up, loop, tear-down (sampler.cpp:551-880)
// ------------------------------------------------------------------------------------------------

void Sampler::begin()
{
  if (begun_) throw std::runtime_error("bmg_sampler_begin called twice");
  if (!std::isfinite(current_.log_likelihood)) throw std::runtime_error("Sampler cannot start from non-finite likelihood.");
  files_.reset(new Files());
  Files& f = *files_;
  Files::open(f.log, basename_ + "_log.txt", false);
  f.log << std::setprecision(3) << std::fixed;
  Files::open(f.loci, basename_ + "_loci.dat", true);
  Files::open(f.modelsize, basename_ + "_modelsize.dat", true);
  Files::open(f.jumpdistance, basename_ + "_jumpdistance.dat", true);
  Files::open(f.log_likelihood, basename_ + "_log_likelihood.dat", true);
  Files::open(f.log_prior, basename_ + "_log_prior.dat", true);
  Files::open(f.move_type, basename_ + "_move_type.dat", true);
  Files::open(f.move_size, basename_ + "_move_size.dat", true);
  Files::open(f.pve, basename_ + "_pve.dat", true);
  Files::open(f.alpha, basename_ + "_alpha.dat", true);
  Files::open(f.sigma2, basename_ + "_sigma2.dat", true);
  if (save_beta_) Files::open(f.beta, basename_ + "_beta.dat", true);
  f.log << "output " << basename_ << std::endl
        << "seed " << seed_ << std::endl
        << "type (PMV = 0, NK = 1, KSC = 2, G = 3) " << 0 << std::endl
        << "----------------------------------------------------" << std::endl;

  copy_current_to_proposal();
  // proposal weights from the current p_proposal (sampler.cpp:594-605); the device already holds them
  refresh_weights_from_device(true);
  compute_p_moves();
  current_.sample_beta_sigma2(rng_);
  sample_missing();
  if (probit_) probit_sweep();
  prior_->sample_alpha_and_tau2(&current_, rng_);
  current_.compute_log_likelihood();
  copy_current_to_proposal();
  t_start_ = wall_seconds();
  begun_ = true;
}

void Sampler::run(int64_t do_n_iter)
{
  if (!begun_) throw std::runtime_error("bmg_sampler_run: call bmg_sampler_begin first");
  Files& f = *files_;
  sec_last_ = __builtin_ia32_rdtsc();
  const size_t end_iter = n_iter_ + (size_t)do_n_iter;
  for (size_t iter = n_iter_; iter < end_iter; ++iter) {
    mark(kSecOther);
    if ((iter + 1) % n_sample_tau2_and_missing_ == 0) {   // sampler.cpp:628-635
      sample_missing();
      if (probit_) probit_sweep();   // the latent phenotype moves with the same cadence as tau2 / missing genotypes
      prior_->sample_alpha_and_tau2(&current_, rng_);
      current_.compute_log_likelihood();
      copy_current_to_proposal();
      mark(kSecTauAlpha);
    }
    const unsigned char move = (unsigned char)sample_discrete_naive(p_moves_cumsum_, 7, rng_);
    unsigned char jumpdistance = 0;
    const double t0 = wall_seconds();
    switch (move) {
      case 0: jumpdistance = do_multistep_additions_and_removals(); break;
      case 1:
        if (current_.size() > 0 && current_.size() < (m_g_ - 1)) jumpdistance = do_switch_of_nearby_snps();
        break;
      case 2:
        if (current_.size() > 0 && m_g_ > 2) jumpdistance = do_statechange_of_nearby_snps();
        break;
      default: throw std::logic_error("Move is not in 0...2 for the PMV sampler with one effect type");
    }
    move_seconds_ += wall_seconds() - t0;
    if (move == 1) mark(kSecMove1);
    else if (move == 2) mark(kSecMove2);
    ++n_moves_[move];
    n_acpt_moves_[move] += jumpdistance > 0;
    n_accepted_ += jumpdistance > 0;

    if ((iter + 1) % thin_ == 0 || (iter + 2) % n_sample_tau2_and_missing_ == 0) { mark(kSecOther); current_.sample_beta_sigma2(rng_); mark(kSecBetaSigma); }

    f.jumpdistance.write(reinterpret_cast<const char*>(&jumpdistance), 1);
    f.move_type.write(reinterpret_cast<const char*>(&move), 1);
    f.move_size.write(reinterpret_cast<const char*>(&movesize_), 1);
    if ((iter + 1) % thin_ == 0) {   // sampler.cpp:691-728
      const uint32_t n_loci = (uint32_t)current_.size();
      f.modelsize.write(reinterpret_cast<const char*>(&n_loci), sizeof(n_loci));
      f.loci.write(reinterpret_cast<const char*>(current_.loci.data()), n_loci * sizeof(uint32_t));
      if (thinned_seen_++ >= pip_burnin_) {   // running MCMC inclusion counts (bmagwa_postprocess.py:94-103)
        ++incl_samples_;
        for (uint32_t snp : current_.loci) ++incl_count_[snp];
      }
      f.log_likelihood.write(reinterpret_cast<const char*>(&current_.log_likelihood), sizeof(double));
      const double log_prior = prior_->log_model((int)n_loci);
      f.log_prior.write(reinterpret_cast<const char*>(&log_prior), sizeof(double));
      f.sigma2.write(reinterpret_cast<const char*>(&current_.sigma2), sizeof(double));
      if (save_beta_) f.beta.write(reinterpret_cast<const char*>(current_.beta.data()), current_.beta.size() * sizeof(double));
      current_.compute_pve(n_, pves_);
      track_fitted_values();
      f.pve.write(reinterpret_cast<const char*>(pves_), 3 * sizeof(double));
      const double a = prior_->alpha();
      f.alpha.write(reinterpret_cast<const char*>(&a), sizeof(double));
    }
    mark(kSecOutput);
    if ((iter + 1) % n_rao_ == 0) { rao_block(); mark(kSecRao); }
    if (verbosity_ > 0 && (iter + 1) % verbosity_ == 0) {   // sampler.cpp:813-833
      f.log << "(" << (iter + 1) << ")"
            << " acc.rate " << (double)n_accepted_ / (iter + 1) << " acc.rate2 " << (double)n_acpt_moves_[1] / n_moves_[1]
            << " acc.rate3 " << (double)n_acpt_moves_[2] / n_moves_[2] << " p_move_size " << p_move_size_ << " model size "
            << current_.size() << " log p " << current_.log_likelihood + prior_->log_model((int)current_.size()) << " PVE "
            << pves_[0] << "/" << pves_[1] << "/" << pves_[2] << " sigma2 " << current_.sigma2;
      f.log << std::endl;
      f.log.flush();
    }
  }
  n_iter_ = end_iter;
  // the caller may now stay away for longer than the server's idle time-out; the next request restarts it
  chain_server_stop(chain_);
}

// y_hat as Model::compute_pve leaves it (model.hpp:345-392), kept as coefficients (see sampler.hpp)
void Sampler::track_fitted_values()
{
  const bool have_g = current_.size() > 0, have_e = m_e_ > 1;
  if (have_g || !have_e || !reference_quirks_) {
    fitted_.loci.assign(current_.loci.begin(), current_.loci.end());
    fitted_.beta_g.assign(current_.beta.begin() + m_e_, current_.beta.end());
    fitted_.beta_e.assign(current_.beta.begin(), current_.beta.begin() + m_e_);
  } else {
    // no SNP term: the reference adds E beta_e onto whatever y_hat held (zeros before the first call)
    if (fitted_.beta_e.size() != m_e_) fitted_.beta_e.assign(m_e_, 0.0);
    for (size_t j = 0; j < m_e_; ++j) fitted_.beta_e[j] += current_.beta[j];
  }
}

// the rao block of the loop (sampler.cpp:731-811)
void Sampler::rao_block()
{
  const double k_move_size = 1000.0 / (double)opt_.n_rao_burnin;   // sampler.cpp:558 (initial burn-in length)
  const bool do_scan = !flat_proposal_dist_ || n_rao_burnin_ <= 0;
  if (do_scan) {
    if (have_missing_) {   // DataModel::sample_missing (sampler.cpp:733): every SNP outside the model is imputed again from its prior
      const bool on_device = missing_on_device_ < 0 ? tau_on_device_ : missing_on_device_ != 0;
      if (on_device) {
        // the host mirror of SNPs outside the model goes stale; it is never read: a SNP entering the model gets fresh
        // values in draw_for_additions, and the Gibbs step reads its cells from the device
        std::vector<int64_t> in_model(current_.loci.begin(), current_.loci.end());
        chain_impute_from_prior(chain_, in_model.data(), (int)in_model.size(), (uint64_t)seed_ * 0x9E3779B97F4A7C15ull + 29u, ++impute_counter_);
      } else {
        miss_.draw_all_from_prior([this](size_t snp) { return pos_in_current_[snp] >= 0; }, rng_);
        chain_set_missing_all(chain_, miss_.val.data(), (int64_t)miss_.val.size());
      }
    }
    const double t0 = wall_seconds();
    const int k = (int)current_.size();
    std::vector<int64_t> loci(current_.loci.begin(), current_.loci.end());
    chain_residual(chain_, fitted_.loci.data(), fitted_.beta_e.data(), fitted_.beta_g.data(), (int)fitted_.loci.size(), nullptr);
    bmg_scan_params prm;
    std::memset(&prm, 0, sizeof(prm));
    prm.sigma2 = current_.sigma2;
    prm.lmp_add = (size_t)(k + 1) <= m_g_ ? prior_->log_change_on_add(k) : 0.0;   // sampler.cpp:56-60
    prm.lmp_rem = k > 0 ? prior_->log_change_on_add(k - 1) : 0.0;                 // sampler.cpp:61-73
    std::vector<double> tau_host;
    if (!prior_->use_individual_tau2) {
      prm.tau_mode = 0;
      prm.tau_shared = prior_->shared_inv_tau2_alpha2();
    } else if (tau_on_device_) {
      prm.tau_mode = 2;
      prm.tau_seed = seed_;
      prm.tau_counter = ++tau_counter_;
      prm.nu_tau2 = prior_->nu_tau2(); prm.s2_tau2 = prior_->s2_tau2(); prm.alpha2 = prior_->alpha2();
    } else {
      // one draw per SNP from the chain's own stream, in SNP order, in-model SNPs included (sampler.cpp:99-106)
      prm.tau_mode = 1;
      tau_host.resize(m_g_);
      for (size_t j = 0; j < m_g_; ++j) tau_host[j] = prior_->draw_inv_tau2_alpha2(rng_);
      prm.tau_host = tau_host.data();
    }
    chain_scan(chain_, loci.data(), current_.beta.data() + m_e_, current_.inv_tau2_alpha2.data() + m_e_, k, &prm, nullptr);
    if (prm.tau_mode == 1) BMG_CUDA(cudaStreamSynchronize(chain_->stream));   // tau_host must outlive the upload
    scan_seconds_ += wall_seconds() - t0;
    ++n_scans_;
  }
  const bool update_rao = n_rao_burnin_ <= 0;
  const int64_t n_rao_mean = p_rao_n_;
  if (update_rao) ++p_rao_n_; else --n_rao_burnin_;
  bool update_prop = false;
  const int64_t n_prop_mean = (int64_t)p_proposal_n_;
  if (adaptation_ || n_rao_burnin_ > 0) {
    update_prop = !flat_proposal_dist_;
    if (adapt_p_move_size_) {
      const double t = std::max(1.0, k_move_size * (double)p_proposal_n_);
      if (acpt_move_size_goal_ > 0) adapt_p_move_size_acptrate(2.0 + t);
      else adapt_p_move_size_jd_mb();
      geometric_dist_cdf(max_move_size_, p_move_size_, q_p_move_size_);
    }
    ++p_proposal_n_;
  }
  if (update_rao || update_prop) {
    const double t1 = wall_seconds();
    chain_adapt(chain_, update_rao ? 1 : 0, n_rao_mean, update_prop ? 1 : 0, n_prop_mean, q_add_min_, q_rem_min_);
    if (update_prop) refresh_weights_from_device(false);
    epilogue_seconds_ += wall_seconds() - t1;
  }
}

void Sampler::end()
{
  if (!begun_) return;
  if (getenv("BMG_TIMING")) std::cerr << "[bmg timing] end(): column-statistics server running = " << chain_->server_running << std::endl;
  chain_server_stop(chain_);
  Files& f = *files_;
  // samplerstats.print (samplerstats.hpp:92-113)
  {
    std::ofstream s((basename_ + "_samplerstats.txt").c_str());
    if (!s.is_open()) throw std::runtime_error("Failed to open file: " + basename_ + "_samplerstats.txt");
    struct timespec res;
    clock_getres(CLOCK_MONOTONIC, &res);
    s << "n_likelihood_updates_on_add " << n_upd_add_ << std::endl
      << "n_likelihood_updates_on_rem " << n_upd_rem_ << std::endl
      << "n_likelihood_computations " << n_comp_ << std::endl
      << "sampling_time_in_seconds " << wall_seconds() - t_start_ << std::endl
      << "mhmove_time_in_seconds " << move_seconds_ << std::endl
      << "rao_time_in_seconds " << scan_seconds_ << std::endl
      << "clock_resolution_in_seconds " << (double)res.tv_sec + (double)res.tv_nsec / 1e9 << std::endl
      << "p_movesize " << p_move_size_ << std::endl;
  }
  f.log << "Elapsed time: " << wall_seconds() - t_start_ << " seconds." << std::endl;
  if (getenv("BMG_TIMING"))
    std::cerr << "[bmg timing] iterations " << n_iter_ << " moves " << move_seconds_ << " s, of which column-stats wait "
              << device_wait_seconds_ << " s, move-0 delayed rejection " << dr_seconds_ << " s (" << n_dr_ << " events); scans "
              << scan_seconds_ << " s, scan epilogue (adapt + weights to the host, waits for the scan) " << epilogue_seconds_ << " s; missing-genotype Gibbs step "
              << gibbs_seconds_ << " s; column-statistics memo: " << cache_.requests << " moves asked, " << cache_.served
              << " needed no device trip, " << cache_.partial << " a subset, " << n_gram_requests_ << " device requests, "
              << cache_.pairs() << " pairs held; server fall-backs " << chain_->server_fallbacks << std::endl;
  if (getenv("BMG_TIMING")) {
    static const char* names[kSecCount] = {"propose", "request", "removals", "wait", "additions", "backward", "accept-copy", "reject-copy",
                                           "dr-readd", "dr-enumerate", "dr-proposal-probs", "dr-sample", "dr-apply", "move1", "move2",
                                           "beta/sigma2", "tau2/alpha", "output", "rao-block", "other"};
    uint64_t total = 0;
    for (int i = 0; i < kSecCount; ++i) total += sec_ticks_[i];
    std::cerr << "[bmg timing] host time by section (% of " << total << " ticks, " << n_iter_ << " iterations):";
    char num[32];
    for (int i = 0; i < kSecCount; ++i) {
      std::snprintf(num, sizeof num, " %.1f", 100.0 * (double)sec_ticks_[i] / (double)std::max<uint64_t>(1, total));
      std::cerr << " " << names[i] << num;
    }
    std::cerr << std::endl;
  }
  // _rao.dat (sampler.cpp:847-849): the running mean kept on the device
  p_rao_.assign(m_g_, 0.0);
  BMG_CUDA(cudaSetDevice(store_->device));
  bmg::copy_d2h(h_w_.p, chain_->p_rao.p, m_g_ * sizeof(double), chain_->stream);
  BMG_CUDA(cudaStreamSynchronize(chain_->stream));
  std::copy(h_w_.p, h_w_.p + m_g_, p_rao_.begin());
  {
    std::ofstream r;
    Files::open(r, basename_ + "_rao.dat", true);
    r.write(reinterpret_cast<const char*>(p_rao_.data()), m_g_ * sizeof(double));
  }
  files_.reset();
  begun_ = false;
}

int64_t Sampler::inclusion_counts(uint32_t* counts) const
{
  if (counts) std::copy(incl_count_.begin(), incl_count_.end(), counts);
  return incl_samples_;
}

void Sampler::stats(double* out8) const
{
  out8[0] = (double)n_iter_; out8[1] = (double)n_accepted_; out8[2] = (double)current_.size(); out8[3] = current_.log_likelihood;
  out8[4] = move_seconds_; out8[5] = scan_seconds_; out8[6] = (double)n_scans_; out8[7] = device_wait_seconds_;
}

void Sampler::counters(double* out12) const
{
  out12[0] = (double)cache_.requests; out12[1] = (double)cache_.served; out12[2] = (double)cache_.partial;
  out12[3] = (double)n_gram_requests_; out12[4] = (double)cache_.pairs(); out12[5] = (double)chain_->server_fallbacks;
  out12[6] = dr_seconds_; out12[7] = (double)n_dr_; out12[8] = epilogue_seconds_; out12[9] = gibbs_seconds_;
  out12[10] = (double)n_probit_sweeps_; out12[11] = 0.0;
}

// ------------------------------------------------------------------------------------------------
// move 0: multistep additions and removals with delayed rejection (sampler.hpp:899-1268)
// ------------------------------------------------------------------------------------------------
void Sampler::prepare_addrem(double& log_q_forward, unsigned char& n_removals, unsigned char ms)
{
  size_t max_adds = m_g_ - current_.size();
  size_t max_rems = current_.size();
  n_removals = 0;
  for (unsigned char i = 0; i < ms; ++i) {
    bool sample_add;
    if (max_rems == 0) sample_add = true;
    else if (max_adds == 0) sample_add = false;
    else {
      log_q_forward += kLogHalf;
      sample_add = rng_.u01() < 0.5;
    }
    if (sample_add) {
      --max_adds;
      move_isadd_[i] = 1;
      const double total = dd_add_.total();
      const uint32_t ind = dd_add_.sample(rng_.u01());
      move_inds_[i] = ind;
      move_inds_map_[i] = ind;
      log_q_forward += std::log(q_add(ind)) - std::log(total);
      dd_add_.zero(ind);
    } else {
      ++n_removals;
      --max_rems;
      move_isadd_[i] = 0;
      const double total = dd_rem_.total();
      const uint32_t ind = dd_rem_.sample(rng_.u01());
      move_inds_[i] = ind;
      move_inds_map_[i] = -1;
      log_q_forward += std::log(q_rem(ind)) - std::log(total);
      dd_rem_.zero(ind);
    }
  }
}

void Sampler::backward_prepare_addrem(double& log_q_backward, unsigned char ms)
{
  size_t max_adds = m_g_ - proposal_.size();
  size_t max_rems = proposal_.size();
  double w_add = dd_add_.total(), w_rem = dd_rem_.total();
  int last_rem_pos = ms - 1;
  for (unsigned char i = 0; i < ms; ++i) {
    if (max_adds > 0 && max_rems > 0) log_q_backward += kLogHalf;
    if (move_isadd_[i]) {   // undone by a removal; additions are removed in reverse order
      --max_rems;
      while (!move_isadd_[last_rem_pos]) --last_rem_pos;
      const size_t ind = (size_t)move_inds_map_[last_rem_pos];
      --last_rem_pos;
      log_q_backward += std::log(q_rem(ind)) - std::log(w_rem);
      w_rem -= q_rem(ind);
    } else {                // undone by an addition
      --max_adds;
      const size_t ind = (size_t)move_inds_map_[i];
      log_q_backward += std::log(q_add(ind)) - std::log(w_add);
      w_add -= q_add(ind);
    }
  }
}

void Sampler::do_addrem(double& log_q_forward, double& log_q_backward, double& log_mpc, unsigned char ms)
{
  (void)log_q_forward; (void)log_q_backward;   // only touched with several effect types
  int last_map_pos = ms - 1;
  std::vector<uint32_t> cand;
  for (unsigned char i = 0; i < ms; ++i)
    if (move_isadd_[i]) cand.push_back((uint32_t)move_inds_[i]);
  std::vector<double> taus;
  if (have_missing_) taus = draw_for_additions(cand);   // removals draw nothing, so the stream order is the reference's
  begin_gram(cand);
  mark(kSecRequest);
  for (unsigned char i = 0; i < ms; ++i) {   // removals first
    if (move_isadd_[i]) continue;
    const size_t ind = move_inds_[i];
    dd_add_.unzero((uint32_t)ind);
    while (move_inds_map_[last_map_pos] >= 0) --last_map_pos;
    move_inds_map_[last_map_pos] = (int64_t)ind;
    const int model_ind = pos_in_proposal_[ind];
    log_mpc += prior_->log_change_on_rem((int)proposal_.size());
    remove_from_proposal(model_ind);
  }
  mark(kSecRemovals);
  finish_gram();
  mark(kSecWait);
  size_t n_added = 0;
  for (unsigned char i = 0; i < ms; ++i) {   // then additions
    if (!move_isadd_[i]) continue;
    const size_t ind = move_inds_[i];
    dd_rem_.unzero((uint32_t)ind);
    log_mpc += prior_->log_change_on_add((int)proposal_.size());
    const double tau = have_missing_ ? taus[n_added] : prior_->draw_inv_tau2_alpha2(rng_);   // prepare_add_new_term (sampler.hpp:500-515)
    ++n_added;
    add_to_proposal((uint32_t)ind, tau);
  }
  mark(kSecAdditions);
}

void Sampler::undo_move0_flags()
{
  for (unsigned char i = 0; i < movesize_; ++i) {
    const uint32_t ind = (uint32_t)move_inds_[i];
    if (move_isadd_[i]) { dd_add_.unzero(ind); dd_rem_.zero(ind); }
    else { dd_add_.zero(ind); dd_rem_.unzero(ind); }
  }
}

unsigned char Sampler::do_multistep_additions_and_removals()
{
  movesize_ = (unsigned char)(sample_discrete_naive(q_p_move_size_.data(), max_move_size_, rng_) + 1);
  double log_q_forward = 0.0, log_q_backward = 0.0, log_mpc = 0.0;
  unsigned char ms_rem = 0;
  mark(kSecOther);
  prepare_addrem(log_q_forward, ms_rem, movesize_);
  mark(kSecPropose);
  do_addrem(log_q_forward, log_q_backward, log_mpc, movesize_);
  backward_prepare_addrem(log_q_backward, movesize_);
  mark(kSecBackward);
  double log_r = log_q_backward - log_q_forward;
  log_r += log_mpc + proposal_.log_likelihood - current_.log_likelihood;
  if (delay_rejection_ == 0) r_move_size_sum_[movesize_ - 1] += (log_r >= 0 ? 1.0 : std::exp(log_r));
  r_move_size_n_[movesize_ - 1] += 1.0;
  if (log_r >= 0 || std::log(rng_.u01()) <= log_r) {
    copy_proposal_to_current();
    if (delay_rejection_ != 0) r_move_size_sum_[movesize_ - 1] += 1.0;
    mark(kSecAcceptCopy);
    return movesize_;
  }
  if (movesize_ <= delay_rejection_ && movesize_ > 1) {
    const double t0 = wall_seconds();
    const unsigned char moved = delayed_rejection_move0(ms_rem, log_r, log_q_forward, log_q_backward);
    dr_seconds_ += wall_seconds() - t0;
    ++n_dr_;
    return moved;
  }
  copy_current_to_proposal();
  undo_move0_flags();
  mark(kSecRejectCopy);
  return 0;
}

// sampler.hpp:931-1098
unsigned char Sampler::delayed_rejection_move0(unsigned char ms_rem, double, double, double)
{
  const unsigned char ms = movesize_;
  size_t const_loci = current_.size() - ms_rem;
  const unsigned char ms_add = ms - ms_rem;
  double z_add = dd_add_.total(), z_rem = dd_rem_.total();
  const unsigned long newmodel_binary = (1ul << ms_add) - 1;
  for (unsigned char i = 0; i < ms; ++i) {
    const size_t ind = move_inds_[i];
    if (move_isadd_[i]) {   // totals for the model with none of the ms SNPs in it
      z_add += q_add(ind);
      z_rem -= q_rem(ind);
    } else {
      readd_to_proposal((uint32_t)ind);
    }
  }
  for (unsigned char i = 0; i < ms; ++i) {
    const size_t ind = (size_t)move_inds_map_[i];
    const size_t model_ind = (size_t)pos_in_proposal_[ind];
    dr_bit_to_normalized_order_[model_ind - const_loci] = i;
    dr_q_add_[i] = q_add(ind);
    dr_q_rem_[i] = q_rem(ind);
  }
  mark(kSecDrReadd);
  double max_log_model;
  double* P = dr_model_probabilities_.data();
  exh_.run(proposal_, (int)const_loci, (int)ms, P, max_log_model);
  mark(kSecDrEnumerate);
  compute_proposal_probs_for_exh_modelset(ms, dr_bit_to_normalized_order_.data(), dr_q_add_.data(), dr_q_rem_.data(), z_add, z_rem,
                                          const_loci, m_g_, P);
  mark(kSecDrProposal);
  const unsigned long nmodels = 1ul << ms, mask = nmodels - 1;
  double sum = 0.0;
  for (unsigned long i = 0; i < nmodels / 2; ++i) {
    const unsigned long j = (~i) & mask;
    const double a = P[j] - P[i];
    if (a > 0) { P[i] = std::exp(P[j] - max_log_model) * (1 - std::exp(-a)); P[j] = -1.0; }
    else { P[i] = std::exp(P[i] - max_log_model) * (1 - std::exp(a)); P[j] = 1.0; }
    P[i] += sum;
    sum = P[i];
  }
  unsigned long sampled = (unsigned long)sample_discrete(P, nmodels / 2, (int)ms / 2 - 3, rng_);
  if (P[(~sampled) & mask] < 0) sampled = (~sampled) & mask;
  mark(kSecDrSample);
  unsigned char moved = 0;
  if (sampled == ((~newmodel_binary) & mask)) {   // the current model: stay
    copy_current_to_proposal();
    undo_move0_flags();
    mark(kSecDrApply);
    return 0;
  }
  for (unsigned char b = 0; b < ms; ++b) {
    const bool in_sampled = (sampled >> b) & 1, in_new = (newmodel_binary >> b) & 1;
    const uint32_t ind = (uint32_t)move_inds_map_[dr_bit_to_normalized_order_[b]];
    if (in_sampled) {
      if (!in_new) { dd_add_.zero(ind); dd_rem_.unzero(ind); }
      else ++moved;
    } else {
      if (in_new) { dd_add_.unzero(ind); dd_rem_.zero(ind); }
      else ++moved;
    }
  }
  unsigned long removals = mask & sampled;
  for (unsigned char i = 0; i < ms; ++i) {
    if ((removals & 1) == 0) {
      remove_from_proposal((int)(const_loci + i));
      --const_loci;
    }
    removals >>= 1;
  }
  r_move_size_sum_[ms - 1] += (double)moved / (double)ms;
  copy_proposal_to_current();
  mark(kSecDrApply);
  return moved;
}

// ------------------------------------------------------------------------------------------------
// move 1: swap a model SNP with a nearby SNP (sampler.hpp:1271-1418)
// ------------------------------------------------------------------------------------------------
unsigned char Sampler::do_switch_of_nearby_snps()
{
  const size_t full_model_move = m_g_ - current_.size();
  const size_t half = ((size_t)max_move_size_ + 1) / 2;
  unsigned char maxmovesize = (unsigned char)std::min(half, (full_model_move + 1) / 2);
  maxmovesize = (unsigned char)std::min((size_t)maxmovesize, current_.size());
  movesize_ = (unsigned char)(sample_discrete_naive(q_p_move_size_nbs_.data(), maxmovesize, rng_) + 1);
  size_t nbh = std::max((size_t)1, (full_model_move - movesize_) / 2);
  nbh = std::min(max_nbh_, nbh);
  std::vector<int32_t>& cur = pos_in_current_;   // ">= 0" also marks SNPs already chosen for addition in this move
  for (unsigned char i = 0; i < movesize_; ++i) {
    const size_t model_ind_rem = (size_t)std::floor(rng_.u01() * (double)proposal_.size());
    const size_t ind_rem = proposal_.loci[model_ind_rem];
    move_inds_rem_[i] = ind_rem;
    size_t ind_add;
    int j = 0;
    int step = (int)std::floor(rng_.u01() * (double)nbh) + 1;
    if (rng_.u01() < 0.5) {   // walk up over SNPs that are not in the model
      ind_add = ind_rem;
      while (j < step) {
        ++ind_add;
        if (ind_add < m_g_) { if (cur[ind_add] < 0) ++j; }
        else break;
      }
      if (ind_add == m_g_) {  // reflected at the upper end
        ind_add = ind_rem;
        step = step + step - j;
        j = 0;
        while (j < step) {
          --ind_add;
          while (cur[ind_add] >= 0) --ind_add;
          ++j;
        }
      }
    } else {                  // walk down
      ind_add = ind_rem;
      if (ind_add > 0) {
        while (j < step) {
          --ind_add;
          if (ind_add > 0) { if (cur[ind_add] < 0) ++j; }
          else break;
        }
      }
      if (j < step) {
        if (cur[ind_add] < 0) ++j;
        if (j < step) {       // reflected at the lower end
          ind_add = ind_rem;
          step = step + step - j;
          j = 0;
          while (j < step) {
            ++ind_add;
            while (cur[ind_add] >= 0) ++ind_add;
            ++j;
          }
        }
      }
    }
    move_inds_add_[i] = ind_add;
    remove_from_proposal((int)model_ind_rem);
    cur[ind_add] = 0;         // temporary mark, reverted below
  }
  std::vector<uint32_t> cand(movesize_);
  for (unsigned char i = 0; i < movesize_; ++i) {
    cand[i] = (uint32_t)move_inds_add_[i];
    cur[move_inds_add_[i]] = -1;   // revert the marks before anything reads pos_in_current_ as a map
  }
  std::vector<double> taus;
  if (have_missing_) taus = draw_for_additions(cand);
  fetch_gram(cand);
  for (unsigned char i = 0; i < movesize_; ++i) {
    const double tau = have_missing_ ? taus[i] : prior_->draw_inv_tau2_alpha2(rng_);
    add_to_proposal((uint32_t)move_inds_add_[i], tau);
  }
  const double log_r = proposal_.log_likelihood - current_.log_likelihood;
  if (log_r >= 0 || std::log(rng_.u01()) <= log_r) {
    copy_proposal_to_current();
    for (unsigned char i = 0; i < movesize_; ++i) {
      dd_add_.unzero((uint32_t)move_inds_rem_[i]);
      dd_rem_.zero((uint32_t)move_inds_rem_[i]);
      dd_add_.zero((uint32_t)move_inds_add_[i]);
      dd_rem_.unzero((uint32_t)move_inds_add_[i]);
    }
    return (unsigned char)(2 * movesize_);
  }
  copy_current_to_proposal();
  return 0;
}

// ------------------------------------------------------------------------------------------------
// move 2: state change of SNPs near a model SNP (sampler.hpp:1421-1669)
// ------------------------------------------------------------------------------------------------
unsigned char Sampler::do_statechange_of_nearby_snps()
{
  const size_t maxmovesize = std::min(max_nbh_ * 2, std::min(m_g_ - 1, (size_t)max_move_size_));
  movesize_ = (unsigned char)(sample_discrete_naive(q_p_move_size_nbc_.data(), (int)maxmovesize, rng_) + 1);
  const size_t nbh = std::min(max_nbh_, m_g_ / 2);
  const size_t model_ind_c = (size_t)std::floor(rng_.u01() * (double)proposal_.size());
  const size_t ind_c = proposal_.loci[model_ind_c];
  std::vector<size_t> perm(2 * nbh);
  for (size_t i = 0; i < 2 * nbh; ++i) perm[i] = i;
  size_t wrap = ind_c + nbh;
  if (ind_c < nbh) wrap += (nbh - ind_c);
  else if (wrap > m_g_ - 1) wrap = m_g_ - 1;
  for (unsigned char i = 0; i < movesize_; ++i) {
    const size_t pick = (size_t)std::floor(rng_.u01() * (double)(nbh * 2 - i)) + i;
    std::swap(perm[i], perm[pick]);
  }
  unsigned char ms_rem = 0, ms_add = 0;
  for (unsigned char i = 0; i < movesize_; ++i) {
    size_t ind = perm[i] + ind_c + 1;
    if (ind > wrap) ind = ind_c - (ind - wrap);
    if (pos_in_proposal_[ind] < 0) move_inds_add_[ms_add++] = ind;
    else move_inds_rem_[ms_rem++] = ind;
  }
  std::vector<uint32_t> cand(ms_add);
  for (unsigned char i = 0; i < ms_add; ++i) cand[i] = (uint32_t)move_inds_add_[i];
  std::vector<double> taus;
  if (have_missing_) taus = draw_for_additions(cand);
  begin_gram(cand);
  double log_mpc = 0.0;
  for (unsigned char i = 0; i < ms_rem; ++i) {
    const int model_ind_rem = pos_in_proposal_[move_inds_rem_[i]];
    log_mpc += prior_->log_change_on_rem((int)proposal_.size());
    remove_from_proposal(model_ind_rem);
  }
  finish_gram();
  for (unsigned char i = 0; i < ms_add; ++i) {
    log_mpc += prior_->log_change_on_add((int)proposal_.size());
    const double tau = have_missing_ ? taus[i] : prior_->draw_inv_tau2_alpha2(rng_);
    add_to_proposal((uint32_t)move_inds_add_[i], tau);
  }
  double log_r = log_mpc + proposal_.log_likelihood - current_.log_likelihood;
  log_r += std::log((double)current_.size()) - std::log((double)proposal_.size());
  if (log_r >= 0 || std::log(rng_.u01()) <= log_r) {
    copy_proposal_to_current();
    for (unsigned char i = 0; i < ms_add; ++i) { dd_add_.zero((uint32_t)move_inds_add_[i]); dd_rem_.unzero((uint32_t)move_inds_add_[i]); }
    for (unsigned char i = 0; i < ms_rem; ++i) { dd_add_.unzero((uint32_t)move_inds_rem_[i]); dd_rem_.zero((uint32_t)move_inds_rem_[i]); }
    return movesize_;
  }
  if (!(movesize_ <= delay_rejection_ && movesize_ > 1)) {
    copy_current_to_proposal();
    return 0;
  }
  // delayed rejection over the 2^ms states of the touched SNPs (sampler.hpp:1547-1662)
  const unsigned char ms = movesize_;
  size_t const_loci = current_.size() - ms_rem;
  const unsigned long newmodel_binary = (1ul << ms_add) - 1;
  for (unsigned char i = 0; i < ms_rem; ++i) readd_to_proposal((uint32_t)move_inds_rem_[i]);
  double max_log_model;
  double* P = dr_model_probabilities_.data();
  exh_.run(proposal_, (int)const_loci, (int)ms, P, max_log_model);
  const unsigned long nmodels = 1ul << ms, mask = nmodels - 1;
  for (unsigned long i = 0; i < nmodels; ++i) P[i] -= std::log((double)(const_loci + (size_t)__builtin_popcountl(i)));
  double sum = 0.0;
  for (unsigned long i = 0; i < nmodels / 2; ++i) {
    const unsigned long j = (~i) & mask;
    const double a = P[j] - P[i];
    if (a > 0) { P[i] = std::exp(P[j] - max_log_model) * (1 - std::exp(-a)); P[j] = -1.0; }
    else { P[i] = std::exp(P[i] - max_log_model) * (1 - std::exp(a)); P[j] = 1.0; }
    P[i] += sum;
    sum = P[i];
  }
  unsigned long sampled = (unsigned long)sample_discrete(P, nmodels / 2, (int)ms / 2 - 3, rng_);
  if (P[(~sampled) & mask] < 0) sampled = (~sampled) & mask;
  unsigned char moved = 0;
  if (sampled == ((~newmodel_binary) & mask)) {
    copy_current_to_proposal();
    return 0;
  }
  for (unsigned char b = 0; b < ms; ++b) {
    const bool in_sampled = (sampled >> b) & 1, in_new = (newmodel_binary >> b) & 1;
    if (in_sampled && in_new) {          // an addition that is kept
      ++moved;
      const uint32_t ind = (uint32_t)move_inds_add_[b];
      dd_add_.zero(ind); dd_rem_.unzero(ind);
    } else if (!in_sampled && !in_new) { // a removal that is kept
      ++moved;
      const uint32_t ind = (uint32_t)move_inds_rem_[b - ms_add];
      dd_add_.unzero(ind); dd_rem_.zero(ind);
    }
  }
  unsigned long removals = mask & sampled;
  for (unsigned char i = 0; i < ms; ++i) {
    if ((removals & 1) == 0) {
      remove_from_proposal((int)(const_loci + i));
      --const_loci;
    }
    removals >>= 1;
  }
  copy_proposal_to_current();
  return moved;
}

// ------------------------------------------------------------------------------------------------
// move-size adaptation (sampler.hpp:1715-1769)
// ------------------------------------------------------------------------------------------------
void Sampler::adapt_p_move_size_acptrate(double t)
{
  const double acpt_rate = (double)n_acpt_moves_[0] / (double)n_moves_[0];
  p_move_size_ = p_move_size_ + (acpt_move_size_goal_ - acpt_rate) / t;
  p_move_size_ = std::max(std::min(p_move_size_, 0.99), 0.01);
  n_acpt_moves_[0] = 1;
  n_moves_[0] = 2;
}

void Sampler::adapt_p_move_size_jd_mb()
{
  const unsigned int max_ms = max_move_size_;
  double q = 1 - p_move_size_;
  double zp = 1 - std::pow(q, (double)max_ms);
  double q_p = p_move_size_ / zp;
  for (unsigned int i = 0; i < max_ms; ++i) { q_p0_move_size_[i] += q_p; q_p *= q; }
  double p_best = 0.01, p = 0.01, h_best = -INFINITY;
  const double step = 0.98 / 49.0;
  for (int j = 0; j < 50; ++j) {
    double h = 0.0;
    zp = 0.0;
    q = 1.0 - p;
    q_p = p;
    for (unsigned int i = 0; i < max_ms; ++i) {
      const double ratio = q_p / q_p0_move_size_[i];
      h += (i + 1) * r_move_size_sum_[i] * ratio;
      zp += r_move_size_n_[i] * ratio;
      q_p *= q;
    }
    h /= zp;
    if (h > h_best) { h_best = h; p_best = p; }
    p += step;
  }
  p_move_size_ = p_best;
}

}  // namespace bmg
