// model.hpp -- prior, linear-model state and proposal weights of one chain (host side, O(k^2)).
//
// What the reference keeps in bmagwa::Prior, Model, ExhModel and DiscreteDistribution
// (src/prior.hpp, src/prior.cpp, src/model.hpp, src/discrete_distribution.hpp), re-designed
// around the device store:
//   * a Model holds NO n x k design matrix.  The reference stores x as doubles (n x 1024 per
//     Model, three Models per chain) and copies it on every accept/reject; here the packed columns
//     stay in the device store and a Model is just the k x k Gram matrix, its Cholesky factor and a
//     few vectors, so copying a Model is O(k^2) bytes.  Everything the reference derives from
//     x (fitted-value variances for the PVE, x_b'x_b and e_b'x_b for the alpha update) is taken
//     from the Gram matrix instead.
//   * the proposal weights live on the device; the host keeps the in-order layout of the weights,
//     the per-block partial CDFs computed by the device and a Fenwick tree over blocks, which gives
//     the reference's in-order-traversal sampling semantics in O(log m + block).
// Only effect type A is supported by this build (types = A); the reference's other types are
// listed as "next" in SURVEY.md section 8(f).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>
#include "dense.hpp"
#include "options.hpp"
#include "rng.hpp"

namespace bmg {

constexpr int kMaxModelColumns = 1024;  // MAXMODELSIZE of the reference (model.hpp:26-28)

// ------------------------------------------------------------------------------------------------
// Prior (src/prior.hpp:38-290, src/prior.cpp:30-141), single effect type A
// ------------------------------------------------------------------------------------------------
struct Model;

class Prior {
 public:
  double n_plus_nu, minus_n_plus_nu_2, nus2_plus_yy, mu_alpha;
  // probit mode (no reference counterpart): the latent phenotype has sigma2 == 1, so the marginal
  // likelihood keeps -(y'y - v'v)/2 instead of integrating sigma2 out, and y'y changes with every latent sweep
  bool fixed_sigma2 = false;
  double residual_term(double syx_plus_vs2) const
  {
    return fixed_sigma2 ? -0.5 * syx_plus_vs2 : minus_n_plus_nu_2 * std::log(syx_plus_vs2);
  }
  void set_fixed_sigma2(double yy) { fixed_sigma2 = true; nus2_plus_yy = yy; }
  void set_yy(double yy) { nus2_plus_yy = (fixed_sigma2 ? 0.0 : nu_sigma2_ * s2_sigma2_) + yy; }
  bool use_individual_tau2;

  Prior(size_t n, size_t m_g, size_t m_e, double yy, const double* types_prior, double e_qg, double var_qg,
        const std::vector<double>& inv_tau2_e, double nu_sigma2, double s2_sigma2, double nu_tau2_A, double s2_tau2_A,
        double mu_alpha_, bool individual)
  : n_plus_nu((double)n + nu_sigma2), minus_n_plus_nu_2(-0.5 * ((double)n + nu_sigma2)),
    nus2_plus_yy(nu_sigma2 * s2_sigma2 + yy), mu_alpha(mu_alpha_), use_individual_tau2(individual), n_(n), m_g_(m_g),
    m_e_(m_e), inv_tau2_e_(inv_tau2_e), nu_sigma2_(nu_sigma2), s2_sigma2_(s2_sigma2), nu_tau2_(nu_tau2_A),
    s2_tau2_(s2_tau2_A), alpha_(std::max(0.5, mu_alpha_)), alpha2_(alpha_ * alpha_)
  {
    if (s2_sigma2 <= 0)
      throw std::runtime_error("Prior specification error: s2_sigma2 is not positive (Probably option s2_sigma2 is not "
                               "specified and using R2mode_sigma2 leads to the invalid value).");
    if (nu_sigma2 < 0) throw std::runtime_error("Prior specification error: nu_sigma2 is negative.");
    solve_betadist_params(e_qg, var_qg);
    for (int t = 0; t < 5; ++t) types_prior_[t] = types_prior[t];
    types_prior_sum_ = types_prior[kA];
    if (s2_tau2_ <= 0) throw std::runtime_error("Prior specification error: s2_tau2 for A is not positive.");
    if (nu_tau2_ <= 0) throw std::runtime_error("Prior specification error: nu_tau2 for A is not positive.");
    nus2_tau2_ = nu_tau2_ * s2_tau2_;
    const double tau2_init = nus2_tau2_ / std::max(0.1, nu_tau2_ - 2);
    inv_tau2_alpha2_ = 1.0 / (alpha2_ * tau2_init);
  }

  double alpha() const { return alpha_; }
  double alpha2() const { return alpha2_; }
  double nu_tau2() const { return nu_tau2_; }
  double s2_tau2() const { return s2_tau2_; }
  double shared_inv_tau2_alpha2() const { return inv_tau2_alpha2_; }
  const std::vector<double>& inv_tau2_e() const { return inv_tau2_e_; }
  double e_g() const { return g_a_ / (g_a_ + g_b_) * (double)m_g_; }

  // prior.hpp:144-167 with Ns = number of type-A terms = L.  The values depend on L only, so they are tabulated
  // on first use (the delayed-rejection enumeration asks for them 2^ms times per event).
  double log_change_on_add(int L) const
  {
    if (L >= 0 && L < (int)add_tab_.size() && add_tab_[L] == add_tab_[L]) return add_tab_[L];
    const double v = (std::log(g_a_ + L) - std::log(g_b_ + (double)m_g_ - L - 1)) +
                     (std::log(types_prior_[kA] + L) - std::log(types_prior_sum_ + L));
    if (L >= 0 && L < kMaxModelColumns + 2) {
      if ((int)add_tab_.size() <= L) add_tab_.resize(L + 1, NAN);
      add_tab_[L] = v;
    }
    return v;
  }
  double log_change_on_rem(int L) const
  {
    if (L >= 0 && L < (int)rem_tab_.size() && rem_tab_[L] == rem_tab_[L]) return rem_tab_[L];
    const double v = (std::log(g_b_ + (double)m_g_ - L) - std::log(g_a_ + L - 1)) +
                     (std::log(types_prior_sum_ + L - 1) - std::log(types_prior_[kA] + L - 1));
    if (L >= 0 && L < kMaxModelColumns + 2) {
      if ((int)rem_tab_.size() <= L) rem_tab_.resize(L + 1, NAN);
      rem_tab_[L] = v;
    }
    return v;
  }
  double log_model(int L) const
  {
    double lp = std::lgamma(types_prior_[kA] + (double)L);
    lp += std::lgamma(g_a_ + (double)L) + std::lgamma(g_b_ + (double)m_g_ - (double)L) - std::lgamma(types_prior_sum_ + (double)L);
    return lp;
  }
  double log_det_invQ_e() const
  {
    double s = 0.0;
    for (size_t i = 1; i < m_e_; ++i) s += std::log(inv_tau2_e_[i]);
    return s;
  }
  // prior.hpp:192-199
  double draw_inv_tau2_alpha2(ChainRng& rng) const
  {
    if (use_individual_tau2) return 1.0 / (alpha2_ * rng.sinvchi2(nu_tau2_, s2_tau2_));
    return inv_tau2_alpha2_;
  }
  void sample_alpha_and_tau2(Model* model, ChainRng& rng);  // prior.hpp:201-210
  void set_alpha(double val, Model* model);                 // prior.cpp:116-141

  // ---- several effect types (prior.hpp:60-183, prior.cpp:71-141).  Effect types 0 A, 1 H, 2 D, 3 R, 4 AH; an AH SNP has
  // two TERMS (A and H), so per-term hyper-parameters exist for 0..3 only.  configure_types() is called once, before the
  // chain starts; the type-A members above are left as they are (a model.types = A run never gets here).
  void configure_types(const std::vector<int>& types, const double* types_prior5, const double* nu_tau2_4, const double* s2_tau2_4)
  {
    typed_ = true;
    tp_sum_t_ = 0.0;
    for (int t = 0; t < 5; ++t) { allow_type_t_[t] = false; tp_t_[t] = NAN; }
    for (int t = 0; t < 4; ++t) { allow_term_t_[t] = false; nu_t_[t] = s2_t_[t] = nus2_t_[t] = inv_tau2_alpha2_t_[t] = NAN; }
    for (int t : types) {
      if (t < 0 || t > 4) throw std::runtime_error("Prior::configure_types: effect type out of range");
      allow_type_t_[t] = true;
      if (t == 4) allow_term_t_[0] = allow_term_t_[1] = true; else allow_term_t_[t] = true;
    }
    for (int t = 0; t < 5; ++t)
      if (allow_type_t_[t]) { tp_t_[t] = types_prior5[t]; tp_sum_t_ += types_prior5[t]; }
    static const char* names[4] = {"A", "H", "D", "R"};
    for (int t = 0; t < 4; ++t) {
      if (!allow_term_t_[t]) continue;
      nu_t_[t] = nu_tau2_4[t];
      s2_t_[t] = s2_tau2_4[t];
      if (s2_t_[t] <= 0) throw std::runtime_error(std::string("Prior specification error: s2_tau2 for ") + names[t] + " is not positive.");
      if (nu_t_[t] <= 0) throw std::runtime_error(std::string("Prior specification error: nu_tau2 for ") + names[t] + " is not positive.");
      nus2_t_[t] = nu_t_[t] * s2_t_[t];
      const double tau2_init = nus2_t_[t] / std::max(0.1, nu_t_[t] - 2);
      inv_tau2_alpha2_t_[t] = 1.0 / (alpha2_ * tau2_init);
    }
  }
  bool typed() const { return typed_; }
  bool allow_term(int t) const { return allow_term_t_[t]; }
  double shared_inv_tau2_alpha2(int term_type) const { return inv_tau2_alpha2_t_[term_type]; }
  // prior.hpp:144-167
  double log_change_on_add(const int* Ns, int L, int type) const
  {
    return (std::log(g_a_ + L) - std::log(g_b_ + (double)m_g_ - L - 1)) + (std::log(tp_t_[type] + Ns[type]) - std::log(tp_sum_t_ + L));
  }
  double log_change_on_rem(const int* Ns, int L, int type) const
  {
    return (std::log(g_b_ + (double)m_g_ - L) - std::log(g_a_ + L - 1)) + (std::log(tp_sum_t_ + L - 1) - std::log(tp_t_[type] + Ns[type] - 1));
  }
  // prior.hpp:160-167: the type of a SNP of the model changes
  double log_change_on_swi(const int* Ns, int type_add, int type_rem) const
  {
    return type_add == type_rem ? 0.0 : std::log(tp_t_[type_add] + Ns[type_add]) - std::log(tp_t_[type_rem] + Ns[type_rem] - 1);
  }
  // prior.hpp:169-183
  double log_model(const int* Ns) const
  {
    double lp = 0.0;
    int n_g = 0;
    for (int t = 0; t < 5; ++t)
      if (allow_type_t_[t]) { lp += std::lgamma(tp_t_[t] + (double)Ns[t]); n_g += Ns[t]; }
    lp += std::lgamma(g_a_ + (double)n_g) + std::lgamma(g_b_ + (double)m_g_ - (double)n_g) - std::lgamma(tp_sum_t_ + (double)n_g);
    return lp;
  }
  // prior.hpp:192-199 for a term of the given type
  double draw_inv_tau2_alpha2(int term_type, ChainRng& rng) const
  {
    if (use_individual_tau2) return 1.0 / (alpha2_ * rng.sinvchi2(nu_t_[term_type], s2_t_[term_type]));
    return inv_tau2_alpha2_t_[term_type];
  }
  // sample_alpha_and_tau2 for a model whose columns carry term types (Model::term_type)
  void sample_alpha_and_tau2_typed(Model* model, ChainRng& rng);

  void print(const std::string& filename) const
  {
    std::ofstream f(filename.c_str());
    if (!f.is_open()) throw std::runtime_error("Failed to open file: " + filename);
    static const char* names[5] = {"A", "H", "D", "R", "AH"};
    f << "a_omega = " << g_a_ << std::endl
      << "b_omega = " << g_b_ << std::endl
      << "nu_sigma2 = " << nu_sigma2_ << std::endl
      << "s2_sigma2 = " << s2_sigma2_ << std::endl
      << "mu_alpha = " << mu_alpha << std::endl;
    for (int t = 0; t < 5; ++t)
      f << "type_" << names[t] << " = " << (t == kA ? types_prior_[t] : NAN) << " (allowed? " << (t == kA) << ")" << std::endl;
    for (int t = 0; t < 4; ++t)
      f << "term_" << names[t] << " allowed = " << (t == kA) << std::endl
        << "nu_tau2_" << names[t] << " = " << (t == kA ? nu_tau2_ : NAN) << std::endl
        << "s2_tau2_" << names[t] << " = " << (t == kA ? s2_tau2_ : NAN) << std::endl;
  }

 private:
  size_t n_, m_g_, m_e_;
  double g_a_ = NAN, g_b_ = NAN;
  std::vector<double> inv_tau2_e_;
  double nu_sigma2_, s2_sigma2_;
  double nu_tau2_, s2_tau2_, nus2_tau2_;
  double alpha_, alpha2_;
  double inv_tau2_alpha2_;
  double types_prior_[5];
  double types_prior_sum_;
  mutable std::vector<double> add_tab_, rem_tab_;
  // several effect types (configure_types)
  bool typed_ = false;
  bool allow_type_t_[5] = {false, false, false, false, false}, allow_term_t_[4] = {false, false, false, false};
  double tp_t_[5] = {NAN, NAN, NAN, NAN, NAN}, tp_sum_t_ = 0.0;
  double nu_t_[4] = {NAN, NAN, NAN, NAN}, s2_t_[4] = {NAN, NAN, NAN, NAN}, nus2_t_[4] = {NAN, NAN, NAN, NAN};
  double inv_tau2_alpha2_t_[4] = {NAN, NAN, NAN, NAN};

  double sample_alpha(Model* model, ChainRng& rng);  // prior.cpp:30-69
  void sample_tau2(Model* model, ChainRng& rng);     // prior.cpp:71-114

  void solve_betadist_params(double e_q, double var_q)  // prior.hpp:273-290
  {
    const double m = (double)m_g_;
    if (m_g_ > 1) {
      const double z = (var_q - e_q * (1 - e_q)) / ((m - 1) * e_q);
      g_a_ = (z - 1) / (1 - m * z / e_q);
      g_b_ = (m / e_q - 1) * g_a_;
    } else {
      const double z = e_q / (1 - e_q);
      g_b_ = z / var_q / std::pow(1 + z, 3) - 1 / (1 + z);
      g_a_ = z * g_b_;
    }
    if (g_a_ <= 0 || g_b_ <= 0 || std::isnan(g_a_) || std::isnan(g_b_)) throw std::runtime_error("Invalid a or b.\n");
  }
};

// ------------------------------------------------------------------------------------------------
// Gram entries a move needs, gathered from ONE device launch (bmg_chain_column_stats) plus the
// current model's Gram matrix.  "Items" are: covariate columns, SNPs of the current model, and the
// candidate SNPs of the move.
// ------------------------------------------------------------------------------------------------
struct MoveGram {
  int m_e = 0, k_cur = 0, m_c = 0;          // covariate columns, SNPs in the current model, candidates
  std::vector<uint32_t> cand;               // candidate SNP ids
  std::vector<double> xy, xe, xm, xc;       // per candidate: x'y, x'E (m_e), x'X_model (k_cur), x'X_cand (m_c)
  int find(uint32_t snp) const
  {
    for (int c = 0; c < m_c; ++c) if (cand[c] == snp) return c;
    return -1;
  }
};

// ------------------------------------------------------------------------------------------------
// Model (src/model.hpp:46-582) without the design matrix
// ------------------------------------------------------------------------------------------------
struct Model {
  int m_e = 0;
  std::vector<uint32_t> loci;     // SNP of term i (column m_e + i)
  std::vector<uint8_t> term_type; // per column (covariates included) when effect types are in use: 0 A, 1 H, 2 D, 3 R; empty otherwise
  UpperMat xx, l;                 // X'X (upper) and the Cholesky factor of X'X + diag(inv_tau2_alpha2)
  std::vector<double> xy, v, inv_tau2_alpha2, beta, mu_beta;
  double sigma2 = 0.0, syx_plus_vs2 = 0.0, log_det_invQ = 0.0, log_det_invQ_plus_xx = 0.0;
  double log_likelihood = NAN;
  bool mu_beta_computed = false;
  const Prior* prior = nullptr;
  unsigned long* n_updates_add = nullptr;
  unsigned long* n_updates_rem = nullptr;
  unsigned long* n_computations = nullptr;

  int cols() const { return m_e + (int)loci.size(); }
  size_t size() const { return loci.size(); }

  // Model::operator= (model.hpp:115-162): O(k^2) bytes -- only the used upper triangles are copied
  void assign(const Model& o)
  {
    m_e = o.m_e;
    loci = o.loci;
    term_type = o.term_type;
    xx.copy_upper_from(o.xx);
    l.copy_upper_from(o.l);
    xy = o.xy; v = o.v; inv_tau2_alpha2 = o.inv_tau2_alpha2; beta = o.beta; mu_beta = o.mu_beta;
    sigma2 = o.sigma2; syx_plus_vs2 = o.syx_plus_vs2; log_det_invQ = o.log_det_invQ;
    log_det_invQ_plus_xx = o.log_det_invQ_plus_xx; log_likelihood = o.log_likelihood;
    mu_beta_computed = o.mu_beta_computed; prior = o.prior;
  }

  // Model::Model (model.hpp:49-113): covariate columns only; exx = E'E (upper), exy = E'y
  void init(int m_e_, const UpperMat& exx, const std::vector<double>& exy, const Prior* p)
  {
    m_e = m_e_;
    prior = p;
    loci.clear();
    xx.copy_upper_from(exx);
    l.resize(m_e);
    xy = exy;
    v.assign(m_e, 0.0);
    inv_tau2_alpha2 = p->inv_tau2_e();
    beta.assign(m_e, 0.0);
    mu_beta.assign(m_e, 0.0);
    compute_log_likelihood();
  }

  // model.hpp:199-237,558-576
  void compute_log_likelihood()
  {
    const int k = cols();
    if (k < 1) { log_likelihood = NAN; return; }
    if (n_computations) ++*n_computations;
    l.copy_upper_from(xx);
    double log_sum_2 = 0.0;
    l(0, 0) += inv_tau2_alpha2[0];
    for (int i = 1; i < k; ++i) { l(i, i) += inv_tau2_alpha2[i]; log_sum_2 += std::log(inv_tau2_alpha2[i]); }
    log_det_invQ = 0.5 * log_sum_2;
    if (!l.cholesky()) {
      log_likelihood = -INFINITY;
      syx_plus_vs2 = INFINITY;
      mu_beta_computed = false;
      return;
    }
    v = xy;
    l.solve_transposed(v.data(), k);
    double vv = 0.0;
    for (int i = 0; i < k; ++i) vv += v[i] * v[i];
    syx_plus_vs2 = prior->nus2_plus_yy - vv;
    log_det_invQ_plus_xx = 0.0;
    for (int i = 0; i < k; ++i) log_det_invQ_plus_xx += std::log(l(i, i));
    log_likelihood = log_det_invQ - log_det_invQ_plus_xx + prior->residual_term(syx_plus_vs2);
    mu_beta_computed = false;
  }

  // Model::add_term + update_likelihood_on_add (model.hpp:239-268,432-506).  newcol has cols()+1 entries:
  // X'x_new for the existing columns, then x_new'x_new.
  // term_type_val >= 0: the column's term type (0 A, 1 H, 2 D, 3 R) is recorded in term_type (runs with effect types)
  void add_term(uint32_t snp, double xy_new, const double* newcol, double inv_tau2_alpha2_val, int term_type_val = -1)
  {
    const int col = cols();
    if (col + 1 > kMaxModelColumns) throw std::runtime_error("Not enough memory for adding term.");
    if (term_type_val >= 0) {
      term_type.resize(col, 0);
      term_type.push_back((uint8_t)term_type_val);
    }
    loci.push_back(snp);
    mu_beta_computed = false;
    xy.push_back(xy_new);
    inv_tau2_alpha2.push_back(inv_tau2_alpha2_val);
    xx.resize(col + 1);
    for (int i = 0; i <= col; ++i) xx(i, col) = newcol[i];
    if (n_updates_add) ++*n_updates_add;
    v.resize(col + 1);
    if (!l.append(newcol, inv_tau2_alpha2_val)) {
      log_likelihood = -INFINITY;
      syx_plus_vs2 = INFINITY;
      return;
    }
    const double* lcol = l.col(col);
    double dotv = 0.0;
    for (int i = 0; i < col; ++i) dotv += lcol[i] * v[i];
    v[col] = (xy[col] - dotv) / lcol[col];
    syx_plus_vs2 -= v[col] * v[col];
    log_det_invQ += 0.5 * std::log(inv_tau2_alpha2_val);
    log_det_invQ_plus_xx += std::log(lcol[col]);
    log_likelihood = log_det_invQ - log_det_invQ_plus_xx + prior->residual_term(syx_plus_vs2);
  }

  // Model::remove_term + update_likelihood_on_remove (model.hpp:270-312,508-556)
  void remove_term(int model_ind)
  {
    const int x_ind = m_e + model_ind;
    loci.erase(loci.begin() + model_ind);
    if (!term_type.empty()) term_type.erase(term_type.begin() + x_ind);
    xy.erase(xy.begin() + x_ind);
    xx.remove_colrow(x_ind);
    const double inv_val = inv_tau2_alpha2[x_ind];
    inv_tau2_alpha2.erase(inv_tau2_alpha2.begin() + x_ind);
    if (n_updates_rem) ++*n_updates_rem;
    const int k = cols();
    if (x_ind == k) {  // last column: drop it
      log_det_invQ_plus_xx -= std::log(l(x_ind, x_ind));
      syx_plus_vs2 += v[x_ind] * v[x_ind];
      l.resize(k);
      v.resize(k);
    } else {
      l.remove(x_ind);
      v = xy;
      l.solve_transposed(v.data(), k);
      log_det_invQ_plus_xx = 0.0;
      for (int i = 0; i < k; ++i) log_det_invQ_plus_xx += std::log(l(i, i));
      double vv = 0.0;
      for (int i = 0; i < k; ++i) vv += v[i] * v[i];
      syx_plus_vs2 = prior->nus2_plus_yy - vv;
    }
    log_det_invQ -= 0.5 * std::log(inv_val);
    log_likelihood = log_det_invQ - log_det_invQ_plus_xx + prior->residual_term(syx_plus_vs2);
    mu_beta_computed = false;
  }

  void compute_mu_beta()  // model.hpp:314-324
  {
    if (mu_beta_computed) return;
    mu_beta = v;
    l.solve(mu_beta.data(), cols());
    mu_beta_computed = true;
  }
  void sample_beta_sigma2(ChainRng& rng)  // model.hpp:326-342, rand.hpp:144-168
  {
    sigma2 = prior->fixed_sigma2 ? 1.0 : rng.sinvchi2_fixed(syx_plus_vs2 / prior->n_plus_nu);
    compute_mu_beta();
    const int k = cols();
    beta.resize(k);
    for (int i = 0; i < k; ++i) beta[i] = rng.normal();
    l.solve(beta.data(), k);
    const double scale = std::sqrt(sigma2);
    for (int i = 0; i < k; ++i) beta[i] = scale * beta[i] + mu_beta[i];
  }

  double gram(int r, int c) const { return r <= c ? xx(r, c) : xx(c, r); }
  // b' G b over the column range [lo, hi)
  double quad(int lo, int hi) const
  {
    double s = 0.0;
    for (int c = lo; c < hi; ++c) {
      double t = 0.0;
      for (int r = lo; r < hi; ++r) t += gram(r, c) * beta[r];
      s += t * beta[c];
    }
    return s;
  }
  // sum of the fitted values of the column range: 1'X b = sum_c G(0,c) b_c (column 0 is the ones column)
  double fitted_sum(int lo, int hi) const
  {
    double s = 0.0;
    for (int c = lo; c < hi; ++c) s += gram(0, c) * beta[c];
    return s;
  }
  // Model::compute_pve (model.hpp:345-392) from the Gram matrix instead of the n-vectors
  void compute_pve(size_t n, double* pves) const
  {
    const int k = cols();
    const double dn = (double)n;
    auto var_of = [&](double sq, double sm) { return (sq - sm * sm / dn) / (dn - 1.0); };
    const bool have_e = m_e > 1, have_g = !loci.empty();
    pves[2] = have_e ? var_of(quad(0, m_e), fitted_sum(0, m_e)) : 0.0;
    pves[1] = have_g ? var_of(quad(m_e, k), fitted_sum(m_e, k)) : 0.0;
    if (!have_e && !have_g) { pves[0] = pves[1] = pves[2] = 0.0; return; }
    if (!have_e) pves[0] = pves[1];
    else if (!have_g) pves[0] = pves[2];
    else pves[0] = var_of(quad(0, k), fitted_sum(0, k));
    const double z = pves[0] + sigma2;
    pves[0] /= z; pves[1] /= z; pves[2] /= z;
  }
};

// ------------------------------------------------------------------------------------------------
// SNP-level view of a model whose SNPs have effect types (the reference's loci / type_inds / x_ind1 / x_ind2 / Ns,
// src/model.hpp:239-312): which columns of the Model belong to which SNP.  An AH SNP owns two consecutive columns (additive,
// then heterozygous); removing it drops the larger column first, as the reference does, so the Cholesky downdates run in the
// same order.  The Gram entries of the typed columns come from the caller (typed columns of the overlay cache).
// ------------------------------------------------------------------------------------------------
struct TypedTerms {
  std::vector<uint32_t> snp;       // SNPs in model order
  std::vector<uint8_t> type;       // effect type 0 A, 1 H, 2 D, 3 R, 4 AH
  std::vector<int> col1, col2;     // design-matrix columns of the SNP's term(s); col2 = -1 unless AH
  int Ns[5] = {0, 0, 0, 0, 0};     // SNPs per effect type

  size_t size() const { return snp.size(); }
  static int n_columns(int type_code) { return type_code == 4 ? 2 : 1; }
  // term type of the c-th column (0 or 1) of a SNP of this effect type
  static int term_type(int type_code, int c) { return type_code == 4 ? c : type_code; }
  // a SNP whose term(s) were appended to the Model at columns first_col (and first_col + 1 for AH)
  void add(uint32_t s, int type_code, int first_col)
  {
    snp.push_back(s);
    type.push_back((uint8_t)type_code);
    col1.push_back(first_col);
    col2.push_back(type_code == 4 ? first_col + 1 : -1);
    ++Ns[type_code];
  }
  // forgets SNP model_ind; cols_out receives the Model columns to remove, in the order to remove them; returns how many
  int remove(int model_ind, int cols_out[2])
  {
    const int t = type[model_ind], c1 = col1[model_ind], c2 = col2[model_ind];
    --Ns[t];
    snp.erase(snp.begin() + model_ind);
    type.erase(type.begin() + model_ind);
    col1.erase(col1.begin() + model_ind);
    col2.erase(col2.begin() + model_ind);
    const int xmove = t == 4 ? 2 : 1;
    for (size_t i = model_ind; i < snp.size(); ++i) { col1[i] -= xmove; if (col2[i] >= 0) col2[i] -= xmove; }
    if (t == 4) { cols_out[0] = std::max(c1, c2); cols_out[1] = std::min(c1, c2); return 2; }
    cols_out[0] = c1;
    return 1;
  }
};

inline void Prior::sample_tau2(Model* model, ChainRng& rng)
{
  const int k = (int)model->beta.size();
  const double alpha2_sigma2 = alpha2_ * model->sigma2;
  if (use_individual_tau2) {
    for (int i = (int)m_e_; i < k; ++i) {
      const double nu = nu_tau2_ + 1.0;
      const double s2 = (nus2_tau2_ + model->beta[i] * model->beta[i] / alpha2_sigma2) / nu;
      model->inv_tau2_alpha2[i] = 1.0 / (alpha2_ * rng.sinvchi2(nu, s2));
    }
  } else {
    double beta2sum = 0.0;
    size_t cnt = 0;
    for (int i = (int)m_e_; i < k; ++i) { beta2sum += model->beta[i] * model->beta[i]; ++cnt; }
    const double nu = nu_tau2_ + (double)cnt;
    const double s2 = (nus2_tau2_ + beta2sum / alpha2_sigma2) / nu;
    inv_tau2_alpha2_ = 1.0 / (alpha2_ * rng.sinvchi2(nu, s2));
    for (int i = (int)m_e_; i < k; ++i) model->inv_tau2_alpha2[i] = inv_tau2_alpha2_;
  }
  model->mu_beta_computed = false;
}

inline double Prior::sample_alpha(Model* model, ChainRng& rng)
{
  const int k = (int)model->beta.size();
  const int nterms = k - (int)m_e_;
  double a = alpha_;
  do {
    if (nterms > 0) {
      if ((a <= 0.0) || std::log(rng.u01()) <= -2.0 * mu_alpha * a) a = -a;
      // x_b'x_b and (E b_e - y)'x_b from the Gram matrix (prior.cpp:47-58 computes them from n-vectors)
      const double xbxb = model->quad((int)m_e_, k);
      double ebxb = 0.0;
      for (int c = (int)m_e_; c < k; ++c) {
        double t = 0.0;
        for (int r = 0; r < (int)m_e_; ++r) t += model->gram(r, c) * model->beta[r];
        ebxb += (t - model->xy[c]) * model->beta[c];
      }
      const double alpha_sigma2 = a * model->sigma2;
      const double var = 1.0 / (1.0 + xbxb / (a * alpha_sigma2));
      const double mu = var * (mu_alpha - ebxb / alpha_sigma2);
      a = std::sqrt(var) * rng.normal() + mu;
    } else {
      a = rng.normal() + mu_alpha;
    }
  } while (a == 0 || a * a == 0);
  model->mu_beta_computed = false;
  return a;
}

inline void Prior::set_alpha(double val, Model* model)
{
  const double old_alpha2 = alpha2_;
  alpha_ = val;
  alpha2_ = alpha_ * alpha_;
  const int k = (int)model->beta.size();
  if (use_individual_tau2) {
    for (int i = (int)m_e_; i < k; ++i) model->inv_tau2_alpha2[i] = model->inv_tau2_alpha2[i] * old_alpha2 / alpha2_;
  } else {
    inv_tau2_alpha2_ = inv_tau2_alpha2_ * old_alpha2 / alpha2_;
    for (int i = (int)m_e_; i < k; ++i) model->inv_tau2_alpha2[i] = inv_tau2_alpha2_;
  }
}

inline void Prior::sample_alpha_and_tau2(Model* model, ChainRng& rng)
{
  sample_tau2(model, rng);
  set_alpha(sample_alpha(model, rng), model);
}

// prior.cpp:71-141 with x_types: tau2 per term (individual) or per term TYPE (shared), then alpha and the rescaling
inline void Prior::sample_alpha_and_tau2_typed(Model* model, ChainRng& rng)
{
  const int k = (int)model->beta.size();
  const std::vector<uint8_t>& tt = model->term_type;
  if ((int)tt.size() != k) throw std::logic_error("sample_alpha_and_tau2_typed: the model carries no term types");
  const double alpha2_sigma2 = alpha2_ * model->sigma2;
  if (use_individual_tau2) {
    for (int i = (int)m_e_; i < k; ++i) {
      const double nu = nu_t_[tt[i]] + 1.0;
      const double s2 = (nus2_t_[tt[i]] + model->beta[i] * model->beta[i] / alpha2_sigma2) / nu;
      model->inv_tau2_alpha2[i] = 1.0 / (alpha2_ * rng.sinvchi2(nu, s2));
    }
  } else {
    double beta2sum[4] = {0.0, 0.0, 0.0, 0.0};
    size_t m_types[4] = {0, 0, 0, 0};
    for (int i = (int)m_e_; i < k; ++i) { beta2sum[tt[i]] += model->beta[i] * model->beta[i]; ++m_types[tt[i]]; }
    for (int t = 0; t < 4; ++t) {
      if (!allow_term_t_[t]) continue;
      const double nu = nu_t_[t] + (double)m_types[t];
      const double s2 = (nus2_t_[t] + beta2sum[t] / alpha2_sigma2) / nu;
      inv_tau2_alpha2_t_[t] = 1.0 / (alpha2_ * rng.sinvchi2(nu, s2));
    }
    for (int i = (int)m_e_; i < k; ++i) model->inv_tau2_alpha2[i] = inv_tau2_alpha2_t_[tt[i]];
  }
  model->mu_beta_computed = false;
  const double val = sample_alpha(model, rng);
  const double old_alpha2 = alpha2_;
  alpha_ = val;
  alpha2_ = alpha_ * alpha_;
  if (use_individual_tau2) {
    for (int i = (int)m_e_; i < k; ++i) model->inv_tau2_alpha2[i] = model->inv_tau2_alpha2[i] * old_alpha2 / alpha2_;
  } else {
    for (int t = 0; t < 4; ++t)
      if (allow_term_t_[t]) inv_tau2_alpha2_t_[t] = inv_tau2_alpha2_t_[t] * old_alpha2 / alpha2_;
    for (int i = (int)m_e_; i < k; ++i) model->inv_tau2_alpha2[i] = inv_tau2_alpha2_t_[tt[i]];
  }
}

// The delayed-rejection enumeration of sub-models lives in exhaustive.hpp (SubmodelEnumerator).

// ------------------------------------------------------------------------------------------------
// Proposal weights with in-order CDF semantics (src/discrete_distribution.hpp:64-330)
// ------------------------------------------------------------------------------------------------
class ProposalCdf {
 public:
  // order[pos] = item visited pos-th by the reference tree's in-order traversal; block = partial-CDF width
  void init(const std::vector<int32_t>* order, int block)
  {
    order_ = order;
    m_ = (int64_t)order->size();
    block_ = block;
    nb_ = (m_ + block - 1) / block;
    pos_of_.resize(m_);
    for (int64_t p = 0; p < m_; ++p) pos_of_[(*order)[p]] = (int32_t)p;
    w_own_.assign(m_, 0.0);
    w_ = w_own_.data();
    zeroed_.assign(m_, 0);
    fen_.assign(nb_ + 1, 0.0);
    eff_.assign(nb_, 0.0);
  }
  // new weights in IN-ORDER layout (w_inorder[pos]) with the device's per-block sums; zero flags kept
  // (DiscreteDistribution::update_weights, :263-314).  live_or_dead lists the currently zeroed items when
  // `mostly_live` (dd_add) or the currently non-zeroed items otherwise (dd_rem).
  void update(const double* w_inorder, const double* block_sums, bool mostly_live, const std::vector<uint32_t>& exceptions)
  {
    w_ = w_inorder;   // borrowed: the caller keeps the buffer alive and unchanged until the next update()
    if (mostly_live) {
      for (int64_t b = 0; b < nb_; ++b) eff_[b] = block_sums[b];
      for (uint32_t it : exceptions) eff_[pos_of_[it] / block_] -= w_[pos_of_[it]];
    } else {
      std::fill(eff_.begin(), eff_.end(), 0.0);
      for (uint32_t it : exceptions) eff_[pos_of_[it] / block_] += w_[pos_of_[it]];
    }
    rebuild();
  }
  void zero_all()
  {
    std::fill(zeroed_.begin(), zeroed_.end(), 1);
    std::fill(eff_.begin(), eff_.end(), 0.0);
    rebuild();
  }
  double weight(uint32_t item) const { return w_[pos_of_[item]]; }
  bool zeroed(uint32_t item) const { return zeroed_[item] != 0; }
  double total() const { return prefix(nb_); }
  void zero(uint32_t item)  // adddate
  {
    if (zeroed_[item]) return;
    zeroed_[item] = 1;
    const int32_t p = pos_of_[item];
    add(p / block_, -w_[p]);
  }
  void unzero(uint32_t item)  // remdate
  {
    if (!zeroed_[item]) return;
    zeroed_[item] = 0;
    const int32_t p = pos_of_[item];
    add(p / block_, w_[p]);
  }
  // sample(): first in-order position whose cumulative live weight exceeds u * total (:125-153)
  uint32_t sample(double u) const
  {
    const double r = u * total();
    // Fenwick descent: largest block index b with prefix(b) <= r
    int64_t b = 0;
    double acc = 0.0;
    int64_t step = 1;
    while (step * 2 <= nb_) step *= 2;
    for (; step > 0; step >>= 1) {
      const int64_t nx = b + step;
      if (nx <= nb_ && acc + fen_[nx] <= r) { b = nx; acc += fen_[nx]; }
    }
    if (b >= nb_) b = nb_ - 1, acc = prefix(b);
    // skip blocks that carry no live weight, then scan inside the block
    int64_t last_live = -1;
    for (int64_t blk = b; blk < nb_; ++blk) {
      const int64_t p0 = blk * block_, p1 = std::min(m_, p0 + block_);
      for (int64_t p = p0; p < p1; ++p) {
        const int32_t it = (*order_)[p];
        if (zeroed_[it]) continue;
        acc += w_[p];
        last_live = p;
        if (r < acc) return (uint32_t)it;
      }
    }
    if (last_live < 0) {  // rounding pushed r past the end of the live weight: last live item overall
      for (int64_t p = m_ - 1; p >= 0; --p)
        if (!zeroed_[(*order_)[p]]) { last_live = p; break; }
    }
    if (last_live < 0) throw std::logic_error("ProposalCdf::sample: every item is zeroed");
    return (uint32_t)(*order_)[last_live];
  }

 private:
  void rebuild()
  {
    std::fill(fen_.begin(), fen_.end(), 0.0);
    for (int64_t b = 0; b < nb_; ++b) {
      fen_[b + 1] += eff_[b];
      const int64_t parent = (b + 1) + ((b + 1) & -(b + 1));
      if (parent <= nb_) fen_[parent] += fen_[b + 1];
    }
  }
  void add(int64_t b, double d)
  {
    eff_[b] += d;
    for (int64_t i = b + 1; i <= nb_; i += i & -i) fen_[i] += d;
  }
  double prefix(int64_t b) const
  {
    double s = 0.0;
    for (int64_t i = b; i > 0; i -= i & -i) s += fen_[i];
    return s;
  }
  const std::vector<int32_t>* order_ = nullptr;
  int64_t m_ = 0, nb_ = 0;
  int block_ = 256;
  std::vector<int32_t> pos_of_;
  const double* w_ = nullptr;      // in-order layout (borrowed from the sampler's pinned staging buffer)
  std::vector<double> w_own_;      // zeros until the first update()
  std::vector<uint8_t> zeroed_; // by item
  std::vector<double> fen_, eff_;
};

}  // namespace bmg
