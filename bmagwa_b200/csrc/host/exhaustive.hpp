// exhaustive.hpp -- delayed rejection: log probabilities of all 2^ms sub-models of the SNPs a rejected move touched, and the
// proposal probabilities of reaching each of them (src/sampler.cpp:882-1049).  Plain host code over ExhModel (model.hpp);
// kept apart from sampler.cpp so that it is unit-tested on the CPU (tests/test_cpu_host_model.py).
#pragma once
#include <cmath>
#include <cstddef>
#include <utility>
#include <vector>
#include "model.hpp"

namespace bmg {

namespace exhaustive_detail {
const double kLogHalf = -0.69314718055994528622676398299518041312694549560546875;  // sampler.hpp:39
}

// ------------------------------------------------------------------------------------------------
// exhaustive enumeration helpers (sampler.cpp:882-1049)
// ------------------------------------------------------------------------------------------------
// Exh: ExhModel (model.types = A) or TypedExhModel (several effect types) -- the walk over the sub-models is the same
template <class Exh>
inline void compute_exhaustive_modelset(size_t n_inds, Exh* exh, double* logp, double& max_log_model)
{
  std::vector<size_t> inds(n_inds);
  for (size_t i = 0; i < n_inds; ++i) inds[i] = i;
  size_t binary = 0;
  int model_size = 0;
  auto note = [&](double v) { logp[binary] = v; if (v > max_log_model) max_log_model = v; };
  logp[binary] = exh->log_prob();
  max_log_model = logp[binary];
  for (size_t i = 0; i < n_inds; ++i) {
    ++model_size;
    if (i > 1) { ++model_size; exh->update_on_add(); }
    binary = ((size_t)1 << model_size) - 1;
    note(exh->update_on_add());
    for (size_t j = 0; j < i; ++j) {   // walk the new variable to the left-most place
      --model_size;
      std::swap(inds[model_size - 1], inds[model_size]);
      binary &= ~((size_t)1 << inds[model_size]);
      note(exh->update_on_moveleft());
    }
    const size_t nmodels = ((size_t)1 << i) - i - 1;
    size_t j = 0, nK = 0;
    char do_lefts = 0;
    while (j < nmodels) {
      if (do_lefts < 2) {
        ++model_size;
        std::swap(inds[model_size], inds[model_size - 1]);
        binary |= ((size_t)1 << inds[model_size - 1]);
        note(exh->update_on_twonewswap());
        ++j;
        ++do_lefts;
      } else {
        ++nK;
        size_t K = 0;
        while (((nK >> K) & 1) == 0) ++K;   // 0,1,0,2,0,1,0,3,...
        for (size_t k = 0; k <= K; ++k) {
          --model_size;
          std::swap(inds[model_size - 1], inds[model_size]);
          binary &= ~((size_t)1 << inds[model_size]);
          note(exh->update_on_moveleft());
          ++j;
        }
        do_lefts = 0;
      }
    }
  }
}

inline void compute_proposal_probs_for_exh_modelset(int n_inds, const unsigned char* bit_to_normalized_order, const double* q_add,
                                             const double* q_rem, double z_add, double z_rem, size_t const_loci, size_t m_g,
                                             double* log_prop_probs, const double* log_q_add_types = nullptr)
{
  // sampler.cpp:982-1049.  Same sequence of factors as the reference; the logs of the ms weights are taken once
  // and the normalising totals are multiplied up and logged once per sub-model instead of once per step.
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

}  // namespace bmg
