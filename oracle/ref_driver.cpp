/* oracle/ref_driver.cpp -- TEST INFRASTRUCTURE, not product code.
 *
 * A C-ABI window onto the UNMODIFIED reference classes (headers are included
 * from /root/reference/src where they lie; this file is compiled with
 * -fno-access-control so it can reach the private members the reference's own
 * friend classes use).  It is linked with the reference objects into
 * oracle/_ref/libbmagwa_ref.so and is used
 *   - to pin oracle/oracle.c against the reference itself (tests/, CPU),
 *   - to generate the golden vectors under tests/golden/ (tests/golden/make_golden.py),
 *   - as the "reference" CPU baseline of bench.py (the scan loop and the sampler).
 * Nothing under bmagwa_b200/ links or loads this.
 */
#include <cstring>
#include <ctime>
#include <string>
#include <vector>
#include "options.hpp"
#include "data.hpp"
#include "data_model.hpp"
#include "prior.hpp"
#include "model.hpp"
#include "sampler.hpp"
#include "precomputed_snp_covariances.hpp"
#include "discrete_distribution.hpp"
#include "rand.hpp"
#include "utils.hpp"

using namespace bmagwa;

struct RefCtx {
  Options* opt0;
  Options* opt;
  Data* data;
  PrecomputedSNPCovariances* pre;
  Sampler* sampler;
  std::string err;
  bool shares_data;
};

static thread_local std::string g_err;

extern "C" {

const char* refd_last_error() { return g_err.c_str(); }

void refd_set_blas_threads(int n)
{
#ifdef BMAGWA_SHIM_OPENBLAS
  scipy_openblas_set_num_threads(n);
#else
  (void)n;
#endif
}

/* main.cpp:47-76 for chain `chain_index` */
void* refd_open(const char* ini, int chain_index)
{
  try {
    RefCtx* c = new RefCtx();
    c->opt0 = new Options(ini);
    c->data = new Data(c->opt0->n, c->opt0->m_g, c->opt0->m_e, c->opt0->file_fam,
                       c->opt0->file_g, c->opt0->recode_g_to_minor_allele_count,
                       c->opt0->file_e, c->opt0->file_y);
    DataModel* tmp = new DataModel(c->data, c->opt0->types);
    c->pre = new PrecomputedSNPCovariances(tmp);
    delete tmp;
    c->opt = c->opt0->clone(chain_index);
    c->sampler = new Sampler(*c->opt, c->data, c->pre);
    c->shares_data = false;
    return c;
  } catch (std::exception& e) {
    g_err = e.what();
    return NULL;
  }
}

/* a further chain over the SAME Data / moment cache, as main.cpp:70-76 does for n_threads > 1 */
void* refd_open_chain(void* parent, int chain_index)
{
  try {
    RefCtx* p = (RefCtx*)parent;
    RefCtx* c = new RefCtx();
    c->opt0 = p->opt0; c->data = p->data; c->pre = p->pre;
    c->opt = c->opt0->clone(chain_index);
    c->sampler = new Sampler(*c->opt, c->data, c->pre);
    c->shares_data = true;
    return c;
  } catch (std::exception& e) {
    g_err = e.what();
    return NULL;
  }
}

void refd_close(void* h)
{
  RefCtx* c = (RefCtx*)h;
  delete c->sampler;
  delete c->opt;
  if (!c->shares_data) {
    delete c->pre;
    delete c->data;
    delete c->opt0;
  }
  delete c;
}

/* number of iterations of the next sample() call (Sampler::do_n_iter, sampler.hpp:397) */
void refd_set_do_n_iter(void* h, long n) { ((RefCtx*)h)->sampler->do_n_iter = (size_t)n; }

/* sample() again on a sampler that already ran (sampler.cpp:625,835 continue from n_iter); wall seconds.
   ONLY valid while the model is still empty: sample() rebuilds dd_rem with every SNP zeroed (sampler.cpp:599-606),
   so re-entering with SNPs in the model corrupts the removal proposal and eventually crashes.  bench.py therefore
   never continues a chain; kept for the empty-model case only. */
double refd_continue_chain(void* h)
{
  RefCtx* c = (RefCtx*)h;
  struct timespec t0, t1;
  try {
    clock_gettime(CLOCK_MONOTONIC, &t0);
    c->sampler->sample();
    clock_gettime(CLOCK_MONOTONIC, &t1);
  } catch (std::exception& e) {
    g_err = e.what();
    return -1.0;
  }
  return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

/* ---- Data / DataModel (data.hpp:75-88, data_model.hpp:101-140) ---- */
void refd_sizes(void* h, long* n, long* m_g, long* m_e, long* n_types)
{
  RefCtx* c = (RefCtx*)h;
  *n = c->data->n; *m_g = c->data->m_g; *m_e = c->data->m_e;
  *n_types = c->sampler->data_model->n_types;
}

void refd_data_stats(void* h, double* out4)
{
  RefCtx* c = (RefCtx*)h;
  out4[0] = c->data->var_y(); out4[1] = c->data->var_x();
  out4[2] = c->data->mean_x(); out4[3] = c->data->yy();
}

void refd_y(void* h, double* out) { RefCtx* c = (RefCtx*)h; memcpy(out, c->data->y().data(), c->data->n * sizeof(double)); }

void refd_e(void* h, double* out)
{ /* n x m_e column-major */
  RefCtx* c = (RefCtx*)h;
  for (size_t j = 0; j < c->data->m_e; ++j)
    for (size_t i = 0; i < c->data->n; ++i) out[j * c->data->n + i] = c->data->e()(i, j);
}

double refd_get_genotype(void* h, long ind, long snp) { return ((RefCtx*)h)->data->get_genotype(ind, snp); }

/* type 0..3 = A,H,D,R; with_overlay=0 -> Data (missing = -1), 1 -> the chain's DataModel */
void refd_get_column(void* h, long snp, int type, int with_overlay, double* out)
{
  RefCtx* c = (RefCtx*)h;
  VectorView v(out, c->data->n);
  if (with_overlay) {
    DataModel* dm = c->sampler->data_model;
    (dm->*dm->get_genotypes[type])(snp, v);
  } else {
    switch (type) {
      case 0: c->data->get_genotypes_additive(snp, v); break;
      case 1: c->data->get_genotypes_heterozygous(snp, v); break;
      case 2: c->data->get_genotypes_dominant(snp, v); break;
      default: c->data->get_genotypes_recessive(snp, v); break;
    }
  }
}

/* returns the number of missing cells; idx_out (may be NULL) gets their row indices, prior3 the cumulative counts */
long refd_missing(void* h, long snp, long* idx_out, double* prior3)
{
  RefCtx* c = (RefCtx*)h;
  const size_t* ml = c->data->miss_loc()[snp];
  if (ml == NULL) return 0;
  if (idx_out) for (size_t i = 1; i <= ml[0]; ++i) idx_out[i - 1] = (long)ml[i];
  if (prior3) for (int k = 0; k < 3; ++k) prior3[k] = c->data->miss_prior()[snp][k];
  return (long)ml[0];
}

void refd_set_miss_val(void* h, long snp, long k, int val) { ((RefCtx*)h)->sampler->data_model->miss_val()[snp][k + 1] = (char)val; }
int refd_get_miss_val(void* h, long snp, long k) { return ((RefCtx*)h)->sampler->data_model->miss_val()[snp][k + 1]; }

/* ---- PrecomputedSNPCovariances (precomputed_snp_covariances.hpp:46-56) ---- */
int refd_moments_offset(void* h) { return ((RefCtx*)h)->pre->offset; }
void refd_moments(void* h, double* out)
{
  RefCtx* c = (RefCtx*)h;
  memcpy(out, c->pre->xx, sizeof(double) * c->pre->offset * c->data->m_g);
}
void refd_update_prexx_cov(void* h, long snp, double* inout) { ((RefCtx*)h)->sampler->data_model->update_prexx_cov(snp, inout); }

/* ---- Prior (prior.hpp:144-183) ---- */
double refd_prior_log_add(void* h, const int* Ns, int n_loci, int type) { return ((RefCtx*)h)->sampler->prior->compute_log_change_on_add(Ns, n_loci, (DataModel::ef_t)type); }
double refd_prior_log_rem(void* h, const int* Ns, int n_loci, int type) { return ((RefCtx*)h)->sampler->prior->compute_log_change_on_rem(Ns, n_loci, (DataModel::ef_t)type); }
double refd_prior_log_swi(void* h, const int* Ns, int type_add, int type_rem) { return ((RefCtx*)h)->sampler->prior->compute_log_change_on_swi(Ns, (DataModel::ef_t)type_add, (DataModel::ef_t)type_rem); }
double refd_prior_log_model(void* h, const int* Ns) { return ((RefCtx*)h)->sampler->prior->compute_log_model(Ns); }
void refd_prior_params(void* h, double* out)
{
  Prior* p = ((RefCtx*)h)->sampler->prior;
  out[0] = p->g_a; out[1] = p->g_b; out[2] = p->n_plus_nu; out[3] = p->nus2_plus_yy;
  out[4] = p->alpha_; out[5] = p->s2_sigma2; out[6] = p->nu_tau2[0]; out[7] = p->s2_tau2[0];
  out[8] = p->use_individual_tau2 ? NAN : p->inv_tau2_alpha2[0];
  out[9] = p->e_g();
}
/* per TERM type t = A,H,D,R: out[3t..] = {inv_tau2_alpha2[t] (shared-tau value), nu_tau2[t], s2_tau2[t]}; NaN where the
   term is not allowed (prior.hpp:88-119) */
void refd_prior_terms(void* h, double* out12)
{
  Prior* p = ((RefCtx*)h)->sampler->prior;
  for (int t = 0; t < 4; ++t) { out12[3 * t] = p->inv_tau2_alpha2[t]; out12[3 * t + 1] = p->nu_tau2[t]; out12[3 * t + 2] = p->s2_tau2[t]; }
}
void refd_prior_set_alpha(void* h, double a) { RefCtx* c = (RefCtx*)h; c->sampler->prior->set_alpha(a, c->sampler->current_model); }

/* ---- Model (model.hpp:199-312) on the sampler's current_model ---- */
void refd_model_add(void* h, long snp, int t_ind, const double* inv_tau2_alpha2) { ((RefCtx*)h)->sampler->current_model->add_term(snp, (char)t_ind, inv_tau2_alpha2); }
void refd_model_remove(void* h, int model_ind) { ((RefCtx*)h)->sampler->current_model->remove_term(model_ind); }
void refd_model_compute_loglik(void* h) { ((RefCtx*)h)->sampler->current_model->compute_log_likelihood(); }
double refd_model_loglik(void* h) { return ((RefCtx*)h)->sampler->current_model->log_likelihood_; }
int refd_model_size(void* h) { return (int)((RefCtx*)h)->sampler->current_model->size(); }
int refd_model_cols(void* h) { return (int)((RefCtx*)h)->sampler->current_model->x.cols(); }
/* what: 0 xx (upper, k x k col-major), 1 l, 2 xy, 3 v, 4 inv_tau2_alpha2, 5 beta, 6 scalars */
void refd_model_get(void* h, int what, double* out)
{
  Model* m = ((RefCtx*)h)->sampler->current_model;
  const size_t k = m->x.cols();
  switch (what) {
    case 0: for (size_t c = 0; c < k; ++c) for (size_t r = 0; r < k; ++r) out[c * k + r] = (r <= c) ? m->xx(r, c) : 0.0; break;
    case 1: for (size_t c = 0; c < k; ++c) for (size_t r = 0; r < k; ++r) out[c * k + r] = (r <= c) ? m->l(r, c) : 0.0; break;
    case 2: for (size_t i = 0; i < k; ++i) out[i] = m->xy(i); break;
    case 3: for (size_t i = 0; i < k; ++i) out[i] = m->v(i); break;
    case 4: for (size_t i = 0; i < k; ++i) out[i] = m->inv_tau2_alpha2(i); break;
    case 5: for (size_t i = 0; i < m->beta.length(); ++i) out[i] = m->beta(i); break;
    default:
      out[0] = m->sigma2; out[1] = m->syx_plus_vs2; out[2] = m->log_det_invQ;
      out[3] = m->log_det_invQ_plus_xx; out[4] = m->log_likelihood_;
  }
}
void refd_model_loci(void* h, unsigned* out) { Model* m = ((RefCtx*)h)->sampler->current_model; for (size_t i = 0; i < m->size(); ++i) out[i] = m->loci[i]; }
void refd_model_set_beta_sigma2(void* h, const double* beta, double sigma2)
{
  Model* m = ((RefCtx*)h)->sampler->current_model;
  m->beta.resize(m->x.cols());
  for (size_t i = 0; i < m->x.cols(); ++i) m->beta(i) = beta[i];
  m->sigma2 = sigma2;
}
void refd_model_sample_beta_sigma2(void* h) { RefCtx* c = (RefCtx*)h; c->sampler->current_model->sample_beta_sigma2(c->sampler->rng); }
void refd_model_compute_pve(void* h, double* pves3, double* y_hat_out)
{
  Model* m = ((RefCtx*)h)->sampler->current_model;
  m->compute_pve(pves3);
  if (y_hat_out) memcpy(y_hat_out, (const void*)m->y_hat.data(), m->y_hat.length() * sizeof(double));
}
void refd_sample_alpha_and_tau2(void* h)
{
  RefCtx* c = (RefCtx*)h;
  c->sampler->prior->sample_alpha_and_tau2(c->sampler->current_model, c->sampler->data_model->y(), c->sampler->rng);
}

/* Sampler::sample_missing (sampler.cpp:264-453) on the current model with the sampler's own random stream; new_model is
   brought level with current_model first, as it is whenever sample() calls it (sampler.cpp:266) */
void refd_sample_missing(void* h)
{
  Sampler* s = ((RefCtx*)h)->sampler;
  *s->new_model = *s->current_model;
  s->sample_missing();
}
/* DataModel::sample_missing (data_model.cpp:78-90): every SNP outside the current model, from its prior */
void refd_sample_missing_from_prior(void* h)
{
  Sampler* s = ((RefCtx*)h)->sampler;
  s->data_model->sample_missing(s->current_model->model_inds, s->rng);
}

/* ---- the all-SNP scan (sampler.cpp:32-261).  y_hat == NULL -> X*beta of the current model ---- */
void refd_scan(void* h, const double* y_hat_in, double* p_r, double* p_r_types /* m_g x n_types or NULL */)
{
  RefCtx* c = (RefCtx*)h;
  Sampler* s = c->sampler;
  const size_t n = c->data->n, m_g = c->data->m_g, nt = s->n_types;
  Vector y_hat(n);
  if (y_hat_in) memcpy((void*)y_hat.data(), y_hat_in, n * sizeof(double));
  else y_hat.set_to_product(s->current_model->x, s->current_model->beta, false);
  double** prt = NULL;
  if (nt > 1) { prt = new double*[m_g]; for (size_t i = 0; i < m_g; ++i) prt[i] = new double[nt]; }
  s->raob->p_raoblackwell(s, y_hat, p_r, prt);
  if (prt) {
    for (size_t i = 0; i < m_g; ++i) {
      if (p_r_types) memcpy(p_r_types + i * nt, prt[i], nt * sizeof(double));
      delete[] prt[i];
    }
    delete[] prt;
  }
}

/* wall-clock seconds for `reps` scans (CPU baseline of the genotype scan) */
double refd_scan_time(void* h, int reps)
{
  RefCtx* c = (RefCtx*)h;
  Sampler* s = c->sampler;
  const size_t n = c->data->n;
  Vector y_hat(n);
  y_hat.set_to_product(s->current_model->x, s->current_model->beta, false);
  std::vector<double> p_r(c->data->m_g);
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  for (int r = 0; r < reps; ++r) s->raob->p_raoblackwell(s, y_hat, &p_r[0], NULL);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}

/* ---- the chain itself (main.cpp:111-120); returns wall seconds of sample() ---- */
double refd_run_chain(void* h)
{
  RefCtx* c = (RefCtx*)h;
  struct timespec t0, t1;
  try {
    c->sampler->initialize_p_proposal_flat();
    clock_gettime(CLOCK_MONOTONIC, &t0);
    c->sampler->sample();
    clock_gettime(CLOCK_MONOTONIC, &t1);
  } catch (std::exception& e) {
    g_err = e.what();
    return -1.0;
  }
  return (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);
}
void refd_print_prior(void* h) { ((RefCtx*)h)->sampler->print_prior(); }

/* ---- Rand (rand.hpp:36-191) ---- */
void* refd_rng_new(unsigned seed, double nu) { return new Rand(seed, nu); }
void refd_rng_free(void* r) { delete (Rand*)r; }
double refd_rng_u01(void* r) { return ((Rand*)r)->rand_01(); }
double refd_rng_normal(void* r) { return ((Rand*)r)->rand_normal(); }
double refd_rng_sinvchi2_1(void* r, double s2) { return ((Rand*)r)->rand_sinvchi2(s2); }
double refd_rng_sinvchi2_2(void* r, double nu, double s2) { return ((Rand*)r)->rand_sinvchi2(nu, s2); }

/* ---- DiscreteDistribution (discrete_distribution.hpp:64-330) ---- */
struct RefDD { Rand* rng; DiscreteDistribution* dd; std::vector<double> w; };
void* refd_dd_new(const double* w, long m, unsigned seed)
{
  RefDD* d = new RefDD();
  d->w.assign(w, w + m);
  d->rng = new Rand(seed, 1.0);
  d->dd = new DiscreteDistribution(true, &d->w[0], m, *d->rng);
  return d;
}
void refd_dd_free(void* p) { RefDD* d = (RefDD*)p; delete d->dd; delete d->rng; delete d; }
long refd_dd_sample(void* p) { return (long)((RefDD*)p)->dd->sample(); }
void refd_dd_zero(void* p, long i) { ((RefDD*)p)->dd->adddate(i); }
void refd_dd_unzero(void* p, long i) { ((RefDD*)p)->dd->remdate(i); }
double refd_dd_total(void* p) { return ((RefDD*)p)->dd->total_w(); }
void refd_dd_update(void* p, const double* w)
{
  RefDD* d = (RefDD*)p;
  memcpy(&d->w[0], w, d->w.size() * sizeof(double));
  d->dd->update_weights(&d->w[0]);
}

/* ---- Cholesky kit (symmmatrix.cpp:144-265) on a free-standing k x k upper factor (col-major, ld=k) ---- */
int refd_chol(double* a, int k)
{
  SymmMatrix m(k, k);
  for (int c = 0; c < k; ++c) for (int r = 0; r <= c; ++r) m(r, c) = a[c * k + r];
  bool ok = m.cholesky();
  for (int c = 0; c < k; ++c) for (int r = 0; r <= c; ++r) a[c * k + r] = m(r, c);
  return ok ? 1 : 0;
}
void refd_chol_downdate(double* a, int k, int col_rem)
{ /* result is (k-1) x (k-1), written with ld = k-1 */
  SymmMatrix m(k, k);
  for (int c = 0; c < k; ++c) for (int r = 0; r <= c; ++r) m(r, c) = a[c * k + r];
  std::vector<double> ct(k + 1), st(k + 1);
  m.cholesky_downdate(col_rem, &ct[0], &st[0]);
  const int k1 = k - 1;
  for (int c = 0; c < k1; ++c) for (int r = 0; r < k1; ++r) a[c * k1 + r] = (r <= c) ? m(r, c) : 0.0;
}
void refd_chol_swapadj(double* a, int k, int col, double* v)
{
  SymmMatrix m(k, k);
  for (int c = 0; c < k; ++c) for (int r = 0; r <= c; ++r) m(r, c) = a[c * k + r];
  Vector vv(k);
  for (int i = 0; i < k; ++i) vv(i) = v ? v[i] : 0.0;
  m.cholesky_swapadj(col, &vv);
  for (int c = 0; c < k; ++c) for (int r = 0; r < k; ++r) a[c * k + r] = (r <= c) ? m(r, c) : 0.0;
  if (v) for (int i = 0; i < k; ++i) v[i] = vv(i);
}

/* ---- delayed rejection: proposal probabilities of the exhaustive model set (sampler.cpp:982-1049), one effect type ---- */
void refd_dr_proposal_probs(int n_inds, const unsigned char* bit_to_normalized_order, const double* q_add, const double* q_rem,
                            double z_add, double z_rem, long const_loci, long m_g, const double* log_q_add_types,
                            double* log_prop_probs)
{
  const bool use_types = log_q_add_types != NULL;   /* several effect types: the type proposal of each addition */
  const size_t cl = (size_t)const_loci, mg = (size_t)m_g;
  compute_proposal_probs_for_exh_modelset(use_types, n_inds, bit_to_normalized_order, q_add, q_rem, z_add, z_rem, cl, mg,
                                          log_q_add_types, log_prop_probs);
}

/* ---- utils (utils.cpp:144-152) ---- */
void refd_geometric_cdf(int maxsize, double p, double* out) { Utils::geometric_dist_cdf(maxsize, p, out); }
double refd_gammaln(double x) { return Utils::gammaln(x); }

} // extern "C"
