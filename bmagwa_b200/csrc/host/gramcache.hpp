// gramcache.hpp -- host-side memo of the per-proposal column statistics.
//
// The reference recomputes x_j'y, x_j'E and X_gamma'x_j from the n x k double design matrix every time a SNP is
// proposed (Model::update_likelihood_on_add, src/model.hpp:453-470).  In linear mode without missing calls these
// numbers are constants of the data: x_j'y and x_j'E depend on SNP j only, x_j'x_l on the unordered pair.  A move
// whose SNPs were all proposed before, against model SNPs they were already paired with, needs no device round trip.
// How often that happens is a property of the data and of the stage of the chain: moves 1/2 stay within
// +-max_SNP_neighborhood_size of the model's SNPs and repeat, whereas move 0 draws from a proposal that spreads about
// as much weight over the ~m_g null SNPs as over the model, so most of its additions are first-time visitors
// (measured at C2 during the first 6,500 iterations: 0.6 % of the moves served from here; profiles/round2_notes.md).
//
// Exactness: the entries are the device's own results (k_colstats_server / k_column_stats_inline), whose
// floating-point sums depend on the candidate column and the slice geometry only -- never on what else was in the
// request -- and genotype x genotype products are exact integers.  A chain with the cache therefore writes the
// same bytes as a chain without it (tests/test_gpu_chain.py).
//
// Not cached: SNPs with missing calls (their imputed values are redrawn at every proposal, src/data_model.cpp:95-103).
// The probit latent phenotype changes x'y at every sweep: bump_phenotype() invalidates those (and only those).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

namespace bmg {

class GramCache {
 public:
  bool enabled = true;
  uint64_t requests = 0, served = 0, partial = 0;   // moves that asked / needed no device trip / asked the device for a subset

  void init(size_t m_g, size_t m_e)
  {
    m_g_ = m_g; m_e_ = m_e;
    stamp_.assign(m_g, 0u);
    seen_.assign(m_g, 0);
    xy_.assign(m_g, 0.0);
    xx_.assign(m_g, 0.0);
    xe_.assign(m_g * m_e, 0.0);
    cap_ = 0; used_ = 0;
    keys_.clear(); vals_.clear();
  }
  bool ready() const { return !stamp_.empty(); }
  // Most SNPs a chain proposes are one-off visitors (the proposal spreads about as much weight over the ~m_g null SNPs
  // as over the model's neighbourhood), and filing their k products would cost more than it ever returns: results are
  // kept from a SNP's SECOND proposal on.
  bool repeat_visitor(uint32_t snp)
  {
    if (seen_[snp] < 255) ++seen_[snp];
    return seen_[snp] >= 2;
  }
  void bump_phenotype()
  {
    if (++epoch_ == 0) { epoch_ = 1; std::fill(stamp_.begin(), stamp_.end(), 0u); }
  }
  // x'y (and with it x'E, x'x) of the SNP for the CURRENT phenotype
  bool have_snp(uint32_t snp) const { return stamp_[snp] == epoch_; }
  double xy(uint32_t snp) const { return xy_[snp]; }
  double xx(uint32_t snp) const { return xx_[snp]; }
  const double* xe(uint32_t snp) const { return xe_.data() + (size_t)snp * m_e_; }
  void put_snp(uint32_t snp, double xy, const double* xe, double xx)
  {
    stamp_[snp] = epoch_;
    xy_[snp] = xy;
    xx_[snp] = xx;
    for (size_t j = 0; j < m_e_; ++j) xe_[(size_t)snp * m_e_ + j] = xe[j];
  }
  bool get_pair(uint32_t a, uint32_t b, double* out) const
  {
    if (cap_ == 0) return false;
    const uint64_t key = make_key(a, b);
    for (size_t i = hash(key) & (cap_ - 1);; i = (i + 1) & (cap_ - 1)) {
      if (keys_[i] == key) { *out = vals_[i]; return true; }
      if (keys_[i] == kEmpty) return false;
    }
  }
  void put_pair(uint32_t a, uint32_t b, double v)
  {
    if (2 * (used_ + 1) > cap_) grow();
    insert(make_key(a, b), v);
  }
  size_t pairs() const { return used_; }
  uint64_t restarts() const { return restarts_; }
  // slots of the pair table before it starts over: rounded down to a power of two, at least 16
  void set_max_slots(size_t slots)
  {
    size_t p = 16;
    while (2 * p <= slots) p *= 2;
    max_cap_ = p;
  }

 private:
  static constexpr uint64_t kEmpty = ~0ull;
  size_t max_cap_ = (size_t)1 << 24;   // 256 MB of keys + values; beyond that the pair table starts over
  size_t m_g_ = 0, m_e_ = 0;
  uint32_t epoch_ = 1;
  std::vector<uint32_t> stamp_;
  std::vector<uint8_t> seen_;
  std::vector<double> xy_, xx_, xe_;
  std::vector<uint64_t> keys_;
  std::vector<double> vals_;
  size_t cap_ = 0, used_ = 0;
  uint64_t restarts_ = 0;

  static uint64_t make_key(uint32_t a, uint32_t b) { return a < b ? ((uint64_t)a << 32) | b : ((uint64_t)b << 32) | a; }
  static size_t hash(uint64_t k)
  {
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (size_t)k;
  }
  void insert(uint64_t key, double v)
  {
    for (size_t i = hash(key) & (cap_ - 1);; i = (i + 1) & (cap_ - 1)) {
      if (keys_[i] == key) { vals_[i] = v; return; }
      if (keys_[i] == kEmpty) { keys_[i] = key; vals_[i] = v; ++used_; return; }
    }
  }
  void grow()
  {
    if (cap_ >= max_cap_) {   // start over: the chain refills what it still uses
      std::fill(keys_.begin(), keys_.end(), kEmpty);
      used_ = 0;
      ++restarts_;
      return;
    }
    const size_t new_cap = cap_ == 0 ? std::min(max_cap_, (size_t)1 << 16) : 2 * cap_;
    std::vector<uint64_t> ok;
    std::vector<double> ov;
    ok.swap(keys_); ov.swap(vals_);
    keys_.assign(new_cap, kEmpty);
    vals_.assign(new_cap, 0.0);
    const size_t old_cap = cap_;
    cap_ = new_cap; used_ = 0;
    for (size_t i = 0; i < old_cap; ++i)
      if (ok[i] != kEmpty) insert(ok[i], ov[i]);
  }
};

}  // namespace bmg
