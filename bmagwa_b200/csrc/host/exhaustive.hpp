// exhaustive.hpp -- delayed rejection: the log posterior of all 2^ms sub-models of the SNPs a rejected move touched, and
// the probabilities with which the move would propose flipping all of them from each sub-model.  Written from the
// specification ("Delayed rejection"); the reference obtains the same numbers by walking a Gray-like
// sequence of adjacent column swaps of the whole factor (src/model.hpp:602-839, src/sampler.cpp:882-1049).  Plain host
// code, unit-tested on the CPU against from-scratch models and against the reference's own function
// (tests/test_cpu_host_model.py).
//
// Sub-model enumeration (SubmodelEnumerator).  The move's ms SNPs are the LAST terms of the model handed in.  With U the
// upper Cholesky factor of X'X + diag(tau) in that order and v = U^-T X'y, the trailing block T of U and the tail z of v
// describe the ms SNPs conditionally on the rest of the model, and only those (ms x ms numbers) are ever touched:
//   * INCLUDING the first remaining SNP is free: it contributes log T_00, z_0^2 and its prior terms, and the SNPs after
//     it are described, conditionally on it, by T[1:,1:] and z[1:] -- a pointer offset;
//   * EXCLUDING it deletes T's first column: T[:,1:] is upper Hessenberg and r-1 Givens rotations (applied to z as well)
//     make it triangular again, written into a scratch block of its own.
// A depth-first walk over "include / exclude the next SNP" therefore visits every sub-model exactly once at the price of
// about one rotation and two logarithms per sub-model, on data that stays in L1.  A SNP with two columns (effect type AH)
// is included / deleted as a pair.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstring>
#include <utility>
#include <vector>
#include "model.hpp"

namespace bmg {

namespace exhaustive_detail {
const double kLogHalf = -0.69314718055994528622676398299518041312694549560546875;  // log 1/2

// model prior along a path of inclusions (src/prior.hpp:144-167): one effect type / several
struct PriorWalkA {
  const Prior* p; int L; double acc;
  void add(int) { acc += p->log_change_on_add(L); ++L; }
};
struct PriorWalkTyped {
  const Prior* p; int Ns[5]; int L; double acc;
  void add(int t) { acc += p->log_change_on_add(Ns, L, t); ++Ns[t]; ++L; }
};
}  // namespace exhaustive_detail

class SubmodelEnumerator {
 public:
  // logp[b], b in [0, 2^ms): log marginal likelihood + log model prior of the sub-model holding the first const_loci SNPs
  // of `src` and those of its last ms SNPs whose bit is set in b (bit i = the i-th of them, in model order), relative to
  // the model prior of the sub-model b = 0.  max_log_model = the largest entry.
  void run(const Model& src, int const_loci, int ms, double* logp, double& max_log_model)
  {
    width_.assign(ms, 1);
    type_.assign(ms, 0);
    exhaustive_detail::PriorWalkA pw{src.prior, const_loci, 0.0};
    start(src, src.m_e + const_loci, ms, logp);
    walk(0, t_.data(), ncols_, z_.data(), ncols_, 0ul, half_logtau0_, sum_logdiag0_, s0_, pw);
    max_log_model = max_;
  }
  // the same for SNPs with effect types: an AH SNP (type 4) owns two adjacent columns
  void run_typed(const Model& src, const TypedTerms& terms, int const_loci, int ms, double* logp, double& max_log_model)
  {
    width_.resize(ms);
    type_.resize(ms);
    exhaustive_detail::PriorWalkTyped pw{src.prior, {0, 0, 0, 0, 0}, const_loci, 0.0};
    for (int i = 0; i < const_loci; ++i) ++pw.Ns[terms.type[i]];
    for (int i = 0; i < ms; ++i) { type_[i] = terms.type[const_loci + i]; width_[i] = TypedTerms::n_columns(type_[i]); }
    const int base = ms > 0 ? terms.col1[const_loci] : src.cols();
    start(src, base, ms, logp);
    walk(0, t_.data(), ncols_, z_.data(), ncols_, 0ul, half_logtau0_, sum_logdiag0_, s0_, pw);
    max_log_model = max_;
  }

 private:
  const Prior* prior_ = nullptr;
  int ms_ = 0, ncols_ = 0;
  double* logp_ = nullptr;
  double max_ = 0.0, half_logtau0_ = 0.0, sum_logdiag0_ = 0.0, s0_ = 0.0;
  std::vector<int> width_, type_, col0_;
  std::vector<double> t_, z_, half_logtau_;     // trailing block (column-major, ld = ncols_), tail of v, 0.5 log tau per column
  std::vector<std::vector<double>> scratch_;    // per depth: the block after a deletion, then its z

  void start(const Model& src, int base, int ms, double* logp)
  {
    prior_ = src.prior;
    ms_ = ms;
    logp_ = logp;
    col0_.resize(ms + 1);
    int c = 0;
    for (int i = 0; i < ms; ++i) { col0_[i] = c; c += width_[i]; }
    col0_[ms] = c;
    ncols_ = c;
    // the part of the model every sub-model shares
    double vv = 0.0, sld = 0.0, slt = prior_->log_det_invQ_e();
    for (int i = 0; i < base; ++i) { vv += src.v[i] * src.v[i]; sld += std::log(src.l(i, i)); }
    for (int i = src.m_e; i < base; ++i) slt += std::log(src.inv_tau2_alpha2[i]);
    s0_ = prior_->nus2_plus_yy - vv;
    sum_logdiag0_ = sld;
    half_logtau0_ = 0.5 * slt;
    // the move's SNPs given that part
    t_.assign((size_t)ncols_ * ncols_, 0.0);
    z_.resize(ncols_);
    half_logtau_.resize(ncols_);
    for (int q = 0; q < ncols_; ++q) {
      const double* col = src.l.col(base + q) + base;
      for (int r = 0; r <= q; ++r) t_[(size_t)q * ncols_ + r] = col[r];
      z_[q] = src.v[base + q];
      half_logtau_[q] = 0.5 * std::log(src.inv_tau2_alpha2[base + q]);
    }
    if ((int)scratch_.size() < ms + 1) scratch_.resize(ms + 1);
    max_ = half_logtau0_ - sum_logdiag0_ + prior_->residual_term(s0_);
    logp_[0] = max_;
  }

  // first column of the r x r upper-triangular block at (a, ld) removed in place; z likewise (its last entry is then void)
  static void delete_first_column(double* a, int ld, double* z, int r)
  {
    for (int q = 0; q + 1 < r; ++q) std::memcpy(a + (size_t)q * ld, a + (size_t)(q + 1) * ld, sizeof(double) * (size_t)(q + 2));
    for (int j = 0; j + 1 < r; ++j) {   // upper Hessenberg -> upper triangular, positive diagonal
      double* cj = a + (size_t)j * ld;
      const double x = cj[j], y = cj[j + 1];
      const double h = std::sqrt(x * x + y * y);
      const double c = x / h, s = y / h;
      cj[j] = h;
      for (int q = j + 1; q + 1 < r; ++q) {
        double* cq = a + (size_t)q * ld;
        const double u = cq[j], w = cq[j + 1];
        cq[j] = c * u + s * w;
        cq[j + 1] = c * w - s * u;
      }
      const double u = z[j], w = z[j + 1];
      z[j] = c * u + s * w;
      z[j + 1] = c * w - s * u;
    }
  }

  // SNPs d.. of the move are undecided; (a, ld, z) describe their r columns given everything included so far
  template <class PriorWalk>
  void walk(int d, const double* a, int ld, const double* z, int r, unsigned long mask, double half_logtau, double sum_logdiag,
            double s, PriorWalk pw)
  {
    if (d == ms_) return;
    const int w = width_[d];
    {   // with SNP d
      double hl = half_logtau, sl = sum_logdiag, s1 = s;
      for (int c = 0; c < w; ++c) {
        hl += half_logtau_[col0_[d] + c];
        sl += std::log(a[(size_t)c * ld + c]);
        s1 -= z[c] * z[c];
      }
      PriorWalk pw1 = pw;
      pw1.add(type_[d]);
      const unsigned long m1 = mask | (1ul << d);
      const double val = hl - sl + prior_->residual_term(s1) + pw1.acc;
      logp_[m1] = val;
      if (val > max_) max_ = val;
      walk(d + 1, a + (size_t)w * ld + w, ld, z + w, r - w, m1, hl, sl, s1, pw1);
    }
    if (d + 1 < ms_) {   // without it
      std::vector<double>& buf = scratch_[d + 1];
      if (buf.size() < (size_t)r * r + r) buf.resize((size_t)r * r + r);
      double* b = buf.data();
      double* bz = b + (size_t)r * r;
      for (int q = 0; q < r; ++q) std::memcpy(b + (size_t)q * r, a + (size_t)q * ld, sizeof(double) * (size_t)(q + 1));
      std::memcpy(bz, z, sizeof(double) * (size_t)r);
      for (int c = 0; c < w; ++c) delete_first_column(b, r, bz, r - c);
      walk(d + 1, b, r, bz, r - w, mask, half_logtau, sum_logdiag, s, pw);
    }
  }
};

// ------------------------------------------------------------------------------------------------
// log_prop_probs[b] += log q(b -> complement of b): the probability that move 0 proposes to flip ALL ms SNPs when the
// chain is at the sub-model b (src/sampler.cpp:982-1049 gives the reference's numbers).
//   bit j of b      SNP j of the move is in the model (the move would remove it); bit_to_normalized_order[j] = the
//                   position of that SNP in the order move 0 would handle the ms steps in
//   q_add / q_rem   proposal weights by normalised position; z_add / z_rem the totals of the sub-model b = 0
// Move 0 handles position p = 0, 1, ...: an addition is drawn with probability q_add[p] / (total add weight left), a
// removal step draws the removable SNPs in reverse position order, and a fair coin precedes every step at which both kinds
// are still possible.
// ------------------------------------------------------------------------------------------------
// The direct statement: one pass over the ms steps per sub-model; the logarithms of the ms weights are taken once and the
// normalising totals are multiplied up and logged once per sub-model.  Its 2 ms data-dependent branches per sub-model
// mispredict on the sub-model bits (about 100 ns per sub-model inside a chain, 20 % of a C2 chain's host time); it serves the
// corner the recurrences below leave out.
inline void compute_proposal_probs_stepwise(int n_inds, const unsigned char* bit_to_normalized_order, const double* q_add,
                                            const double* q_rem, double z_add, double z_rem, size_t const_loci, size_t m_g,
                                            double* log_prop_probs, const double* log_q_add_types = nullptr)
{
  char isadd[256];
  double lq_add[256], lq_rem[256];
  for (int j = 0; j < n_inds; ++j) { lq_add[j] = std::log(q_add[j]); lq_rem[j] = std::log(q_rem[j]); }
  const unsigned long nmodels = 1ul << n_inds;
  for (unsigned long i = 0; i < nmodels; ++i) {
    size_t max_adds = m_g - const_loci, max_rems = const_loci;
    double z_a = z_add, z_r = z_rem;
    for (int j = 0; j < n_inds; ++j) {
      const unsigned char nind = bit_to_normalized_order[j];
      if ((i >> j) & 1) {   // in the model: the move would remove it
        isadd[nind] = 0;
        z_a -= q_add[nind];
        z_r += q_rem[nind];
        ++max_rems;
        --max_adds;
      } else {
        isadd[nind] = 1;
      }
    }
    int last_rem_pos = n_inds - 1, n_half = 0;
    double sum_log_q = 0.0, prod_z = 1.0;
    for (int j = 0; j < n_inds; ++j) {
      if (max_adds > 0 && max_rems > 0) ++n_half;
      if (isadd[j]) {
        --max_adds;
        sum_log_q += lq_add[j];
        if (log_q_add_types) sum_log_q += log_q_add_types[j];   // several effect types: the type proposal of the addition
        prod_z *= z_a;
        z_a -= q_add[j];
      } else {
        --max_rems;
        while (isadd[last_rem_pos]) --last_rem_pos;
        sum_log_q += lq_rem[last_rem_pos];
        prod_z *= z_r;
        z_r -= q_rem[last_rem_pos];
        --last_rem_pos;
      }
    }
    log_prop_probs[i] += (double)n_half * exhaustive_detail::kLogHalf + sum_log_q - std::log(prod_z);
  }
}


// The usual case (more SNPs outside the model than the move could add, so an addition is always possible): the add steps
// and the removal steps normalise independently of each other.  With the in-model positions of a sub-model as a bit set u
// and the positions to add as its complement c,
//     the removal of the SNP at position p draws from    z_rem + (removal weights of u at positions <= p)
//     the addition at position p draws from              (z_add - all ms add weights) + (add weights of c at positions >= p)
// so the product over u's removals follows from the product for u without its HIGHEST position, and the product over c's
// additions from the product for c without its LOWEST position: two branch-free recurrences over 2^ms table entries and
// one logarithm per sub-model.
inline void compute_proposal_probs_for_exh_modelset(int n_inds, const unsigned char* bit_to_normalized_order, const double* q_add,
                                                    const double* q_rem, double z_add, double z_rem, size_t const_loci, size_t m_g,
                                                    double* log_prop_probs, const double* log_q_add_types = nullptr)
{
  if (n_inds > 20 || m_g < const_loci + (size_t)n_inds + 1) {
    compute_proposal_probs_stepwise(n_inds, bit_to_normalized_order, q_add, q_rem, z_add, z_rem, const_loci, m_g, log_prop_probs,
                                    log_q_add_types);
    return;
  }
  const unsigned long nmodels = 1ul << n_inds;
  struct Entry { double sum_r, prod_r, slq_r, sum_a, prod_a, slq_a; unsigned long index; };
  static thread_local std::vector<Entry> table;
  if (table.size() < nmodels) table.resize(nmodels);
  Entry* t = table.data();
  unsigned long bit_at[32];   // normalised position -> bit of the caller's sub-model index
  double lq_add[32], lq_rem[32], q_add_all = 0.0;
  for (int j = 0; j < n_inds; ++j) bit_at[bit_to_normalized_order[j]] = 1ul << j;
  for (int p = 0; p < n_inds; ++p) {
    lq_add[p] = std::log(q_add[p]) + (log_q_add_types ? log_q_add_types[p] : 0.0);
    lq_rem[p] = std::log(q_rem[p]);
    q_add_all += q_add[p];
  }
  const double z0 = z_add - q_add_all;
  t[0].sum_r = 0.0; t[0].prod_r = 1.0; t[0].slq_r = 0.0; t[0].sum_a = 0.0; t[0].prod_a = 1.0; t[0].slq_a = 0.0; t[0].index = 0;
  for (unsigned long u = 1; u < nmodels; ++u) {
    const int hi = 63 - __builtin_clzl(u), lo = __builtin_ctzl(u);
    const Entry& a = t[u ^ (1ul << hi)];   // u without its highest position
    const Entry& b = t[u & (u - 1)];       // u without its lowest position
    Entry& e = t[u];
    e.sum_r = a.sum_r + q_rem[hi];
    e.prod_r = a.prod_r * (z_rem + e.sum_r);
    e.slq_r = a.slq_r + lq_rem[hi];
    e.index = a.index | bit_at[hi];
    e.sum_a = b.sum_a + q_add[lo];
    e.prod_a = b.prod_a * (z0 + e.sum_a);
    e.slq_a = b.slq_a + lq_add[lo];
  }
  // a fair coin precedes every step at which a removal is still possible: all ms steps when the rest of the model holds
  // SNPs, otherwise the steps up to the sub-model's last in-model position
  for (unsigned long u = 0; u < nmodels; ++u) {
    const Entry& r = t[u];
    const Entry& a = t[(nmodels - 1) ^ u];
    const int n_half = const_loci > 0 ? n_inds : (u ? 64 - __builtin_clzl(u) : 0);
    log_prop_probs[r.index] += (double)n_half * exhaustive_detail::kLogHalf + (r.slq_r + a.slq_a) - std::log(r.prod_r * a.prod_a);
  }
}

}  // namespace bmg
