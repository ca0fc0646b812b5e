// sampler.hpp -- the Metropolis-Hastings driver of one chain.
//
// Re-implementation of bmagwa::Sampler for the PMV sampler with effect type A
// (src/sampler.hpp:87-1780, src/sampler.cpp:264-1049): move 0 multistep additions/removals with
// delayed rejection, move 1 swap with a nearby SNP, move 2 state change of nearby SNPs, move-size
// adaptation, the tau2/alpha Gibbs steps, the Rao-Blackwell scan schedule and the reference's
// output files.  The random-number stream is consumed in the reference's order, so a fixed-seed
// chain reproduces the reference's accepted-move sequence until the first floating-point tie.
//
// Every O(n) or O(n m_g) quantity comes from the device through the chain handle:
//   bmg::chain_column_stats  ONE launch per move for all SNPs the move adds
//   bmg::chain_residual / chain_scan / chain_adapt  at every n_rao-th iteration
// and nothing in this file touches a genotype.
#pragma once
#include <cstdint>
#include <fstream>
#include <memory>
#include <string>
#include <vector>
#include "../store.cuh"
#include "dataset.hpp"
#include "exhaustive.hpp"
#include "gramcache.hpp"
#include "missing.hpp"
#include "model.hpp"
#include "options.hpp"
#include "rng.hpp"

namespace bmg {

class Sampler {
 public:
  // Takes ownership of nothing: store and dataset summaries must outlive the sampler.
  // SNP-sharded chain: `store` holds rank's SNP block (peers attached), the same sampler runs in lockstep on every rank
  // group != nullptr: one of several chains over the sharded store (group.cu) -- this chain lives on this rank only
  struct ShardComm { int world = 1, rank = 0; int64_t stride = 0; AllGatherFn allgather = nullptr; void* ctx = nullptr; Group* group = nullptr; };
  Sampler(const Options& opts, int chain_index, Store* store, const std::vector<double>& y, const std::vector<double>& e,
          double var_y, double yy, double var_x, double mean_x, const ShardComm* comm = nullptr);
  ~Sampler();

  void set_option(const std::string& key, const std::string& value);
  void initialize_p_proposal_flat();   // sampler.hpp:351-362
  void print_prior();                  // sampler.hpp:384-390
  void begin();                        // sampler.cpp:551-623 (everything before the loop)
  void run(int64_t n_iter);            // sampler.cpp:625-834
  void end();                          // sampler.cpp:836-879
  void stats(double* out8) const;
  void counters(double* out12) const;
  // inclusion counts over the thinned samples after pip_burnin (what bmagwa_postprocess.py mcmcpos recomputes offline
  // from _loci.dat / _modelsize.dat, bmagwa_postprocess.py:79-123)
  int64_t inclusion_counts(uint32_t* counts) const;
  Chain* chain() { return chain_; }
  Store* store() { return store_; }

 private:
  // ---- configuration (sampler.hpp:395-411)
  Options opt_;
  std::string basename_;
  uint32_t seed_;
  size_t n_, m_g_, m_e_;
  size_t n_iter_ = 0, n_accepted_ = 0, n_rao_, n_rao_burnin_, n_sample_tau2_and_missing_, thin_, verbosity_;
  bool flat_proposal_dist_, adaptation_, save_beta_;
  bool tau_on_device_ = false;
  // ---- state
  Store* store_;
  Chain* chain_ = nullptr;
  ChainRng rng_;
  std::unique_ptr<Prior> prior_;
  Model current_, proposal_;   // the reference's current_model / new_model
  SubmodelEnumerator exh_;   // delayed rejection: all sub-models of a rejected move's SNPs (exhaustive.hpp)
  std::vector<int32_t> pos_in_proposal_;   // model_inds of the proposal model: SNP -> term index or -1
  std::vector<int32_t> pos_in_current_;
  std::vector<double> p_rao_;              // fetched at the end only
  size_t p_proposal_n_ = 1;
  int p_rao_n_ = 0;
  double q_add_min_, q_rem_min_;
  ProposalCdf dd_add_, dd_rem_;
  PinnedBuf<double> h_w2_[2];              // pinned in-order weights, double-buffered: the proposal CDFs read them in place
  int h_w_cur_ = 0;
  PinnedBuf<double> h_w_;                  // pinned staging (p_rao at the end)
  std::vector<double> h_cdf_;
  // ---- moves (sampler.hpp:442-483)
  double p_moves_[7], p_moves_cumsum_[7];
  unsigned char max_move_size_;
  size_t max_nbh_;
  double p_move_size_, p_move_size_nbs_, p_move_size_nbc_;
  bool adapt_p_move_size_;
  std::vector<double> q_p_move_size_, q_p_move_size_nbs_, q_p_move_size_nbc_, r_move_size_sum_, r_move_size_n_, q_p0_move_size_;
  size_t n_acpt_moves_[7], n_moves_[7];
  double acpt_move_size_goal_;
  std::vector<size_t> move_inds_add_, move_inds_rem_, move_inds_;
  std::vector<int64_t> move_inds_map_;
  std::vector<char> move_isadd_;
  unsigned char movesize_ = 0;
  unsigned char delay_rejection_;
  std::vector<double> dr_model_probabilities_, dr_q_add_, dr_q_rem_, dr_log_q_add_types_;
  std::vector<unsigned char> dr_bit_to_normalized_order_;
  MoveGram gram_;
  // ---- book-keeping (samplerstats.hpp)
  unsigned long n_upd_add_ = 0, n_upd_rem_ = 0, n_comp_ = 0;
  double move_seconds_ = 0.0, scan_seconds_ = 0.0, device_wait_seconds_ = 0.0, dr_seconds_ = 0.0, epilogue_seconds_ = 0.0, gibbs_seconds_ = 0.0;
  size_t n_dr_ = 0;
  size_t n_scans_ = 0;
  double t_start_ = 0.0;
  double pves_[3] = {0, 0, 0};
  std::vector<uint32_t> incl_count_;   // per SNP: thinned samples (after pip_burnin_) with the SNP in the model
  int64_t incl_samples_ = 0, thinned_seen_ = 0, pip_burnin_ = 0;
  uint64_t tau_counter_ = 0;
  // what Model::compute_pve leaves in y_hat (model.hpp:345-392), tracked as coefficients: y_hat = X[loci] beta_g + E beta_e.
  // Reference quirk kept for drop-in parity: with covariates and NO SNP term in the model compute_pve does not reset
  // y_hat, it adds E beta_e onto the previous content; option reference_quirks=0 uses the fresh fitted values instead.
  struct FittedState { std::vector<int64_t> loci; std::vector<double> beta_g, beta_e; } fitted_;
  bool reference_quirks_ = true;
  void track_fitted_values();
  // ---- where the host time of an iteration goes (BMG_TIMING prints it): time-stamp-counter ticks per section
  enum Section { kSecPropose = 0, kSecRequest, kSecRemovals, kSecWait, kSecAdditions, kSecBackward, kSecAcceptCopy, kSecRejectCopy,
                 kSecDrReadd, kSecDrEnumerate, kSecDrProposal, kSecDrSample, kSecDrApply, kSecMove1, kSecMove2, kSecBetaSigma,
                 kSecTauAlpha, kSecOutput, kSecRao, kSecOther, kSecCount };
  uint64_t sec_ticks_[kSecCount] = {0}, sec_last_ = 0;
  void mark(Section s)
  {
    const uint64_t now = __builtin_ia32_rdtsc();
    sec_ticks_[s] += now - sec_last_;
    sec_last_ = now;
  }
  // ---- output files
  struct Files;
  std::unique_ptr<Files> files_;
  bool begun_ = false;

  // helpers
  double q_add(size_t snp) const { return dd_add_.weight((uint32_t)snp); }
  double q_rem(size_t snp) const { return dd_rem_.weight((uint32_t)snp); }
  void copy_proposal_to_current();
  void copy_current_to_proposal();
  void fetch_gram(const std::vector<uint32_t>& cand);
  void begin_gram(const std::vector<uint32_t>& cand);
  void finish_gram();
  std::vector<int64_t> gram_c64_, gram_l64_;
  // memo of the device's column statistics (gramcache.hpp): most moves after burn-in need no device round trip
  GramCache cache_;
  std::vector<int> gram_req_;                       // candidates of the pending move the device was asked for
  std::vector<double> req_xy_, req_xe_, req_xm_, req_xc_;
  uint64_t n_gram_requests_ = 0;
  bool cacheable(uint32_t snp) const { return !have_missing_ || miss_.count(snp) == 0; }
  void add_to_proposal(uint32_t snp, double inv_tau2_alpha2);
  void readd_to_proposal(uint32_t snp);
  void remove_from_proposal(int model_ind);
  void refresh_weights_from_device(bool first);
  void compute_p_moves();
  // missing genotypes (sampler.cpp:264-453, data_model.cpp:78-103): index + imputed values mirrored on the host
  // (missing.hpp), the device copy follows through bmg_chain_set_missing[_all]
  MissingCells miss_;
  bool have_missing_ = false;
  int missing_on_device_ = -1;         // re-imputation before scans: 0 host stream (parity), 1 device, -1 follow tau_rng
  uint64_t impute_counter_ = 0;
  double yy_ = 0.0;                    // y'y of the working phenotype
  std::vector<int32_t> gibbs_rows_;
  GibbsScratch gibbs_slot_;
  std::vector<int8_t> gibbs_cells_;
  std::vector<double> y_work_;         // probit mode: host copy of the latent phenotype for the Gibbs step
  void load_missing_index();
  void sample_missing();               // Sampler::sample_missing: Gibbs update of the in-model SNPs' missing cells
  // prepare_add_new_term (sampler.hpp:500-515) for the SNPs a move adds, in the move's order: imputed values from the
  // prior (uploaded before the move's statistics are taken), then the prior precision of the new term
  std::vector<double> draw_for_additions(const std::vector<uint32_t>& cand);
  // probit mode (no reference counterpart): y = 0/1 labels, the chain's phenotype is the latent z
  void enable_probit();
  void probit_sweep();       // z ~ N(X beta, 1) truncated by the labels (device), then y'y, E'z, X_gamma'z refreshed
  bool probit_ = false, probit_labels_sent_ = false;
  uint64_t probit_counter_ = 0;
  int64_t n_probit_sweeps_ = 0;
  std::vector<uint8_t> is_case_;
  const std::vector<double>* y_host_ = nullptr;
  const std::vector<double>* e_host_ = nullptr;
  void rao_block();
  // moves
  unsigned char do_multistep_additions_and_removals();
  void prepare_addrem(double& log_q_forward, unsigned char& n_removals, unsigned char ms);
  void backward_prepare_addrem(double& log_q_backward, unsigned char ms);
  void do_addrem(double& log_q_forward, double& log_q_backward, double& log_mpc, unsigned char ms);
  unsigned char do_switch_of_nearby_snps();
  unsigned char do_statechange_of_nearby_snps();
  void undo_move0_flags();
  unsigned char delayed_rejection_move0(unsigned char ms_rem, double log_r, double log_q_forward, double log_q_backward);
  void adapt_p_move_size_acptrate(double t);
  void adapt_p_move_size_jd_mb();
};

}  // namespace bmg
