/* oracle/oracle.c -- TEST INFRASTRUCTURE.  Never linked into, loaded by, or
 * executed from the product (bmagwa_b200/); only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * A plain-C, single-threaded CPU restatement of the BMAGWA hot path
 * (SURVEY.md section 8): each function cites the reference file:line it
 * follows.  Paths are relative to /root/reference/.
 *
 * PINNING.  The deterministic functions are pinned (tests/test_oracle_*.py)
 *   - against the reference's own golden vectors and known answers
 *     (src/tests/data_tests.hpp:40-47,130-140,171-190; model_tests.hpp:68-89;
 *     raoblackwellizer_tests.hpp:69-92; discrete_distribution_tests.hpp:38-72;
 *     README.markdown:325-327), re-expressed in tests/golden/, and
 *   - against the unmodified reference compiled in oracle/_ref (ref_driver.cpp).
 * The random-variate functions (orc_rng_*) restate Boost.Random, which the
 * reference neither vendors nor pins: PARITY UNPINNED at the draw level (the
 * reference's tests check moments only, src/tests/rand_tests.hpp:30-176).
 * orc_probit_* has no reference counterpart at all: PARITY
 * UNPINNED, it is the CPU statement of our own definition.
 *
 * Conventions: genotype "type" 0..3 = A,H,D,R (data_model.hpp:41); 4 = AH.
 * bed = the PLINK payload after the 3-byte header, SNP-major, B=ceil(n/4)
 * bytes per SNP (data.cpp:245-273).  Matrices are column-major.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ */
/* 1. genotype store: decode, recode, missing index, summary moments   */
/* ------------------------------------------------------------------ */

/* data.cpp:36,40-54: 2-bit code -> additive value; 01 is "missing" (-1) */
static inline int code_of(const uint8_t* bed, long B, long snp, long i)
{
  return (bed[snp * B + (i >> 2)] >> (2 * (i & 3))) & 3;
}
static inline double additive_of_code(int c)
{
  return c == 0 ? 0.0 : (c == 1 ? -1.0 : (c == 2 ? 1.0 : 2.0));
}
/* genotype_tables.hpp:1,260,519,778: H = [x==1], D = [x>0], R = [x==2]; missing stays -1 */
static inline double typed_value(double a, int type)
{
  if (a < 0) return -1.0;
  switch (type) {
    case 0: return a;
    case 1: return a == 1.0 ? 1.0 : 0.0;
    case 2: return a > 0.0 ? 1.0 : 0.0;
    default: return a == 2.0 ? 1.0 : 0.0;
  }
}

ORC_API long orc_bytes_per_snp(long n) { return (n + 3) / 4; }

/* data.cpp:40-54 */
ORC_API double orc_get_genotype(const uint8_t* bed, long n, long snp, long i)
{
  return additive_of_code(code_of(bed, orc_bytes_per_snp(n), snp, i));
}

/* data.cpp:56-138 (the four get_genotypes_* unpackers), missing = -1 */
ORC_API void orc_decode_column(const uint8_t* bed, long n, long snp, int type, double* out)
{
  const long B = orc_bytes_per_snp(n);
  for (long i = 0; i < n; ++i) out[i] = typed_value(additive_of_code(code_of(bed, B, snp, i)), type);
}

/* data.cpp:324-339 + utils.cpp:33-44: swap 0<->2 when the allele frequency of the coded allele > 0.5 */
ORC_API void orc_recode_minor(uint8_t* bed, long n, long m_g, uint8_t* swapped /* m_g flags or NULL */)
{
  const long B = orc_bytes_per_snp(n);
  for (long s = 0; s < m_g; ++s) {
    double macount = 0;
    long ngenos = 0;
    for (long i = 0; i < n; ++i) {
      double g = additive_of_code(code_of(bed, B, s, i));
      if (g >= 0) { macount += g; ++ngenos; }
    }
    int sw = (macount / (2.0 * ngenos) > 0.5);
    if (swapped) swapped[s] = (uint8_t)sw;
    if (!sw) continue;
    for (long i = 0; i < n; ++i) {
      int c = code_of(bed, B, s, i);
      int nc = c == 0 ? 3 : (c == 3 ? 0 : c);
      uint8_t* byte = bed + s * B + (i >> 2);
      int sh = 2 * (i & 3);
      *byte = (uint8_t)((*byte & ~(3 << sh)) | (nc << sh));
    }
  }
}

/* data.cpp:341-376: per SNP the number of missing cells, their row indices (CSR), and the
 * cumulative counts of genotypes 0/1/2 among the observed cells.  offsets has m_g+1 entries.
 * Pass idx == NULL to only count. */
ORC_API void orc_missing_index(const uint8_t* bed, long n, long m_g, long* offsets, long* idx, double* prior3)
{
  const long B = orc_bytes_per_snp(n);
  long pos = 0;
  for (long s = 0; s < m_g; ++s) {
    offsets[s] = pos;
    double cnt[3] = {0, 0, 0};
    for (long i = 0; i < n; ++i) {
      int c = code_of(bed, B, s, i);
      if (c == 1) { if (idx) idx[pos] = i; ++pos; }
      else cnt[(int)additive_of_code(c)] += 1.0;
    }
    if (prior3) {
      prior3[3 * s] = cnt[0];
      prior3[3 * s + 1] = cnt[0] + cnt[1];
      prior3[3 * s + 2] = cnt[0] + cnt[1] + cnt[2];
    }
  }
  offsets[m_g] = pos;
}

/* data.cpp:403-434: mean over SNPs of the per-SNP mean and unbiased variance (missing excluded) */
ORC_API void orc_g_var_and_mean(const uint8_t* bed, long n, long m_g, double* mean, double* var)
{
  const long B = orc_bytes_per_snp(n);
  double tv = 0, tm = 0;
  long nm = 0, nv = 0;
  for (long s = 0; s < m_g; ++s) {
    double sq = 0, sm = 0;
    long ng = 0;
    for (long i = 0; i < n; ++i) {
      double g = additive_of_code(code_of(bed, B, s, i));
      if (g >= 0) { sq += g * g; sm += g; ++ng; }
    }
    if (ng > 1) { tm += sm / ng; tv += (sq - sm * sm / ng) / (ng - 1); ++nm; ++nv; }
    else if (ng == 1) { tm += sm / ng; ++nm; }
  }
  *var = tv / nv;
  *mean = tm / nm;
}

/* vector.cpp:112-123 (VectorView::var) and data.hpp:67 (yy) */
ORC_API double orc_var(const double* v, long n)
{
  double sq = 0, s = 0;
  for (long i = 0; i < n; ++i) { sq += v[i] * v[i]; s += v[i]; }
  return (sq - s * s / n) / (n - 1);
}
ORC_API double orc_dot(const double* a, const double* b, long n)
{
  double s = 0;
  for (long i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

/* data_model.cpp:30-72: typed column with the chain's imputed values at the missing cells.
 * miss_idx/miss_val describe THIS snp's missing cells (n_miss of them). */
ORC_API void orc_decode_column_overlay(const uint8_t* bed, long n, long snp, int type,
                                       const long* miss_idx, const int8_t* miss_val, long n_miss, double* out)
{
  orc_decode_column(bed, n, snp, type, out);
  for (long k = 0; k < n_miss; ++k) {
    int v = miss_val[k];
    double x = type == 0 ? (double)v : (type == 1 ? (double)(v == 1) : (type == 2 ? (double)(v > 0) : (double)(v == 2)));
    out[miss_idx[k]] = x;
  }
}

/* ------------------------------------------------------------------ */
/* 2. per-SNP moment cache                                             */
/* ------------------------------------------------------------------ */

/* layout helper, precomputed_snp_covariances.hpp:59-83.  allow_types[5] -> allow_terms[4], offset, offset_type[4] */
ORC_API int orc_moment_layout(const int* allow_types, int* allow_terms, int* offset_type)
{
  int n_types = 0;
  for (int t = 0; t < 5; ++t) n_types += allow_types[t] != 0;
  for (int t = 0; t < 4; ++t) allow_terms[t] = allow_types[t] != 0;
  if (allow_types[4]) { allow_terms[0] = 1; allow_terms[1] = 1; }
  int offset = 2 * n_types;
  if (allow_types[4]) {
    int extra = 1;
    if (allow_types[0] && allow_types[1]) extra += -2;
    if (!(allow_types[0] || allow_types[1])) extra += 2;
    offset += extra;
  }
  int i = 0;
  for (int t = 0; t < 4; ++t) {
    if (allow_terms[t]) { offset_type[t] = i; ++i; } else offset_type[t] = -1;
  }
  return offset;
}

/* precomputed_snp_covariances.hpp:95-131 with every missing cell = 0 (main.cpp:64-68).
 * NOTE (reference quirk kept): offset_type[] counts terms, so the slot of term t is
 * xx[j*offset + 2*rank(t)] only because the scan indexes pre_xx_tmp[offset_type] and
 * [offset_type+1] after the terms were written consecutively as pairs; for a single allowed
 * type this is (s, v) at [0],[1].  We write pairs consecutively, exactly as the loop does. */
ORC_API void orc_moments(const uint8_t* bed, long n, long m_g, const int* allow_types, double* xx)
{
  int allow_terms[4], offset_type[4];
  const int offset = orc_moment_layout(allow_types, allow_terms, offset_type);
  double* x = (double*)malloc(sizeof(double) * n);
  double* h = (double*)malloc(sizeof(double) * n);
  for (long j = 0; j < m_g; ++j) {
    double* loc = xx + j * offset;
    for (int t = 0; t < 4; ++t) {
      if (!allow_terms[t]) continue;
      orc_decode_column_overlay(bed, n, j, t, NULL, NULL, 0, x);
      for (long i = 0; i < n; ++i) if (x[i] < 0) x[i] = 0.0; /* missing imputed as 0 (data_model.hpp:80-84) */
      double s = 0;
      for (long i = 0; i < n; ++i) s += x[i];
      *loc++ = s;
      *loc++ = orc_dot(x, x, n) - s * s / (double)n;
    }
    if (allow_types[4]) {
      orc_decode_column(bed, n, j, 0, x);
      orc_decode_column(bed, n, j, 1, h);
      double sa = 0, sh = 0;
      for (long i = 0; i < n; ++i) { if (x[i] < 0) x[i] = 0; if (h[i] < 0) h[i] = 0; sa += x[i]; sh += h[i]; }
      *loc++ = orc_dot(x, h, n) - sa * sh / (double)n;
    }
  }
  free(x); free(h);
}

/* data_model.cpp:105-167: patch one SNP's cached moments for the currently imputed values */
ORC_API void orc_update_moments_for_missing(const int* allow_types, long n, const int8_t* miss_val, long n_miss, double* pre_xx)
{
  if (n_miss == 0) return;
  int allow_terms[4], offset_type[4];
  orc_moment_layout(allow_types, allow_terms, offset_type);
  double* loc = pre_xx;
  double sa_sh = 0.0;
  if (allow_types[4]) sa_sh = pre_xx[0] * pre_xx[2];
  for (int t = 0; t < 4; ++t) {
    if (!allow_terms[t]) continue;
    double sv = 0, sv2 = 0;
    for (long k = 0; k < n_miss; ++k) {
      int v = miss_val[k];
      double g = t == 0 ? (double)v : (t == 1 ? (double)(v == 1) : (t == 2 ? (double)(v > 0) : (double)(v == 2)));
      sv += g; sv2 += g * g;
    }
    double old_sum = *loc;
    *loc += sv; ++loc;
    *loc += sv2 - sv * (old_sum * 2.0 + sv) / (double)n; ++loc;
  }
  if (allow_types[4]) {
    *loc += (sa_sh - pre_xx[0] * pre_xx[2]) / (double)n;
    for (long k = 0; k < n_miss; ++k) if (miss_val[k] == 1) *loc += 1.0;
  }
}

/* ------------------------------------------------------------------ */
/* 3. model prior (prior.hpp:144-183,273-290)                          */
/* ------------------------------------------------------------------ */
typedef struct {
  double g_a, g_b, m_g;
  double types_prior[5];
  double types_prior_sum;
} orc_prior_t;

ORC_API int orc_prior_init(orc_prior_t* p, double m_g, double e_qg, double var_qg, const double* types_prior, const int* allow_types)
{
  p->m_g = m_g;
  if (m_g > 1) {
    double z = (var_qg - e_qg * (1 - e_qg)) / ((m_g - 1) * e_qg);
    p->g_a = (z - 1) / (1 - m_g * z / e_qg);
    p->g_b = (m_g / e_qg - 1) * p->g_a;
  } else {
    double z = e_qg / (1 - e_qg);
    p->g_b = z / var_qg / pow(1 + z, 3) - 1 / (1 + z);
    p->g_a = z * p->g_b;
  }
  p->types_prior_sum = 0;
  for (int t = 0; t < 5; ++t) {
    if (allow_types[t]) { p->types_prior[t] = types_prior[t]; p->types_prior_sum += types_prior[t]; }
    else p->types_prior[t] = NAN;
  }
  return !(p->g_a <= 0 || p->g_b <= 0 || isnan(p->g_a) || isnan(p->g_b));
}
ORC_API double orc_prior_log_add(const orc_prior_t* p, const int* Ns, int L, int type)
{
  return (log(p->g_a + L) - log(p->g_b + p->m_g - L - 1)) + (log(p->types_prior[type] + Ns[type]) - log(p->types_prior_sum + L));
}
ORC_API double orc_prior_log_rem(const orc_prior_t* p, const int* Ns, int L, int type)
{
  return (log(p->g_b + p->m_g - L) - log(p->g_a + L - 1)) + (log(p->types_prior_sum + L - 1) - log(p->types_prior[type] + Ns[type] - 1));
}
ORC_API double orc_prior_log_model(const orc_prior_t* p, const int* Ns, const int* allow_types)
{
  double lp = 0; int n_g = 0;
  for (int t = 0; t < 5; ++t) if (allow_types[t]) { lp += lgamma(p->types_prior[t] + (double)Ns[t]); n_g += Ns[t]; }
  return lp + lgamma(p->g_a + n_g) + lgamma(p->g_b + p->m_g - n_g) - lgamma(p->types_prior_sum + n_g);
}

/* ------------------------------------------------------------------ */
/* 4. the all-SNP Rao-Blackwell scan, single effect type A             */
/*    (sampler.cpp:32-261 with n_types == 1, types = {A})              */
/* ------------------------------------------------------------------ */
/* Inputs
 *   bed, n, m_g          packed genotypes (after optional recode)
 *   miss_off/idx/val     CSR of missing cells and the chain's imputed values (may be NULL)
 *   xx                   moment cache, 2 doubles per SNP (s, v), missing = 0
 *   y, y_hat             phenotype and fitted values (n)
 *   model_ind[m_g]       -1 or position k of the SNP in the model
 *   beta_g[k], tau_g[k]  coefficient and inv_tau2_alpha2 of in-model SNP k
 *   tau                  inv_tau2_alpha2 for SNPs: tau_stride==0 -> tau[0] shared, ==1 -> tau[j]
 *                        (individual mode: the value drawn for SNP j; in-model SNPs use tau_g)
 *   individual           use_individual_tau2
 *   sigma2, lmp_add, lmp_rem   sampler.cpp:39,52-76 (lmp_rem = log-prior change of adding type A
 *                        to the model with that SNP removed)
 * Outputs  p_r[m_g]; optional dot[m_g] = x_j . residual (before centring), rx[m_g] centred.
 */
ORC_API void orc_scan_A(const uint8_t* bed, long n, long m_g,
                        const long* miss_off, const long* miss_idx, const int8_t* miss_val,
                        const double* xx, const double* y, const double* y_hat,
                        const int* model_ind, const double* beta_g, const double* tau_g,
                        const double* tau, int tau_stride, int individual,
                        double sigma2, double lmp_add, double lmp_rem,
                        double* p_r, double* dot_out, double* rx_out)
{
  const int allow_types[5] = {1, 0, 0, 0, 0};
  double* r_cm = (double*)malloc(sizeof(double) * n);
  double* r_om = (double*)malloc(sizeof(double) * n);
  double* x = (double*)malloc(sizeof(double) * n);
  double sum = 0;
  for (long i = 0; i < n; ++i) { r_cm[i] = y[i] - y_hat[i]; }
  for (long i = 0; i < n; ++i) sum += r_cm[i];
  const double mr_cm = sum / n;                       /* sampler.cpp:48-49 */
  const double sigma2_times_2 = 2 * sigma2;

  for (long j = 0; j < m_g; ++j) {
    const long nm = miss_off ? miss_off[j + 1] - miss_off[j] : 0;
    const long* mi = nm ? miss_idx + miss_off[j] : NULL;
    const int8_t* mv = nm ? miss_val + miss_off[j] : NULL;
    double tau_j = tau[tau_stride ? j : 0];
    const double* residual; double mr, lmp;
    orc_decode_column_overlay(bed, n, j, 0, mi, mv, nm, x);
    for (long i = 0; i < n; ++i) if (x[i] < 0) x[i] = 0; /* cannot happen once overlaid; keeps NULL-overlay = 0 */
    if (model_ind[j] < 0) {                            /* sampler.cpp:109-115 */
      residual = r_cm; mr = mr_cm; lmp = lmp_add;
    } else {                                           /* sampler.cpp:116-149 */
      const int k = model_ind[j];
      if (individual) tau_j = tau_g[k];
      for (long i = 0; i < n; ++i) r_om[i] = r_cm[i];
      for (long i = 0; i < n; ++i) r_om[i] += beta_g[k] * x[i];
      double s2 = 0;
      for (long i = 0; i < n; ++i) s2 += r_om[i];
      residual = r_om; mr = s2 / n; lmp = lmp_rem;
    }
    double pre[2] = {xx[2 * j], xx[2 * j + 1]};        /* sampler.cpp:153-155 */
    orc_update_moments_for_missing(allow_types, n, mv, nm, pre);
    const double s = pre[0];
    const double det = pre[1] + tau_j;                 /* sampler.cpp:180-195 */
    const double sum_log_Q = -log(tau_j);
    const double d = orc_dot(x, residual, n);
    const double rx = d - s * mr;
    const double exp_term = (rx * rx) / det;
    double p = exp_term / sigma2_times_2 - 0.5 * (log(det) + sum_log_Q) + lmp;
    p = exp(p);                                        /* sampler.cpp:200-206 */
    p_r[j] = isfinite(p) ? p / (1 + p) : 1.0;
    if (dot_out) dot_out[j] = d;
    if (rx_out) rx_out[j] = rx;
  }
  free(r_cm); free(r_om); free(x);
}

/* scan epilogue in Sampler::sample (sampler.cpp:739-803): running means and proposal weights.
 * n_mean = number of samples already in `mean` (p_rao_n or p_proposal_n). */
ORC_API void orc_running_mean(double* mean, const double* p_r, long m_g, long n_mean)
{
  const double z1 = (double)(n_mean + 1), z2 = (double)n_mean / z1;
  for (long i = 0; i < m_g; ++i) mean[i] = z2 * mean[i] + p_r[i] / z1;
}
ORC_API void orc_proposal_weights(const double* p_proposal, long m_g, double q_add_min, double q_rem_min, double* q_add, double* q_rem)
{
  for (long i = 0; i < m_g; ++i) {
    q_add[i] = p_proposal[i] > q_add_min ? p_proposal[i] : q_add_min;   /* sampler.cpp:517-524 */
    double r = 1 - p_proposal[i];
    q_rem[i] = r > q_rem_min ? r : q_rem_min;                           /* sampler.cpp:799-801 */
  }
}

/* ------------------------------------------------------------------ */
/* 5. per-proposal column statistics (model.hpp:453-470)               */
/* ------------------------------------------------------------------ */
/* x_new = typed column of `snp`; xy = x_new.y; xxcol[c] = X[:,c].x_new for the ncols existing
 * columns (X col-major n x ncols doubles), xxcol[ncols] = x_new.x_new. */
ORC_API void orc_column_stats(const double* x_new, const double* y, const double* X, long n, long ncols, double* xy, double* xxcol)
{
  *xy = orc_dot(x_new, y, n);
  for (long c = 0; c < ncols; ++c) xxcol[c] = orc_dot(X + c * n, x_new, n);
  xxcol[ncols] = orc_dot(x_new, x_new, n);
}

/* ------------------------------------------------------------------ */
/* 6. Cholesky kit and marginal likelihood                             */
/* ------------------------------------------------------------------ */
/* upper-triangular factor U (k x k, col-major, leading dimension ld): A = U'U */
ORC_API int orc_chol_upper(double* a, int k, int ld)
{ /* symmmatrix.hpp:139-157 (dpotrf 'U'); column-by-column Cholesky-Crout */
  for (int j = 0; j < k; ++j) {
    double* cj = a + (long)j * ld;
    double d = cj[j];
    for (int i = 0; i < j; ++i) d -= cj[i] * cj[i];
    if (!(d > 0.0)) return 0;
    d = sqrt(d);
    cj[j] = d;
    for (int c = j + 1; c < k; ++c) {
      double* cc = a + (long)c * ld;
      double t = cc[j];
      for (int i = 0; i < j; ++i) t -= cj[i] * cc[i];
      cc[j] = t / d;
    }
  }
  return 1;
}
/* solve U' x = b in place (vector.cpp:165-178, transpose=true) */
ORC_API void orc_solve_ut(const double* u, int k, int ld, double* x)
{
  for (int j = 0; j < k; ++j) {
    const double* cj = u + (long)j * ld;
    double t = x[j];
    for (int i = 0; i < j; ++i) t -= cj[i] * x[i];
    x[j] = t / cj[j];
  }
}
/* solve U x = b in place (transpose=false) */
ORC_API void orc_solve_u(const double* u, int k, int ld, double* x)
{
  for (int j = k - 1; j >= 0; --j) {
    const double* cj = u + (long)j * ld;
    x[j] /= cj[j];
    for (int i = 0; i < j; ++i) x[i] -= x[j] * cj[i];
  }
}
/* symmmatrix.cpp:144-164: append column newcol (k+1 entries: X'x_new then x_new'x_new) with prior
 * precision tau; U has room for k+1 columns.  Returns 0 if not positive definite. */
ORC_API int orc_chol_append(double* u, int k, int ld, const double* newcol, double tau)
{
  double* c = u + (long)k * ld;
  for (int i = 0; i < k; ++i) c[i] = newcol[i];
  orc_solve_ut(u, k, ld, c);
  double d = newcol[k] + tau;
  double ss = 0;
  for (int i = 0; i < k; ++i) ss += c[i] * c[i];
  d -= ss;
  if (d <= 0) return 0;
  c[k] = sqrt(d);
  return 1;
}
static void givens(double* a, double* b, double* c, double* s)
{ /* BLAS drotg as called from dchex.f:229 and symmmatrix.cpp:237 */
  double roe = fabs(*a) > fabs(*b) ? *a : *b;
  double scale = fabs(*a) + fabs(*b), r, z;
  if (scale == 0.0) { *c = 1; *s = 0; r = 0; z = 0; }
  else {
    double ta = *a / scale, tb = *b / scale;
    r = scale * sqrt(ta * ta + tb * tb);
    if (roe < 0) r = -r;
    *c = *a / r; *s = *b / r; z = 1.0;
    if (fabs(*a) > fabs(*b)) z = *s;
    if (fabs(*b) >= fabs(*a) && *c != 0.0) z = 1.0 / *c;
  }
  *a = r; *b = z;
}
/* symmmatrix.cpp:166-195 + dchex.f job=2 with l=p: delete column `rem` (0-based) from the k x k
 * factor; the result is the (k-1) x (k-1) factor in the same storage (same ld).  Expressed as
 * "shift columns left, then re-triangularise the Hessenberg part with Givens rotations", which is
 * what the LINPACK routine does. */
ORC_API void orc_chol_delete(double* u, int k, int ld, int rem)
{
#define U(i, j) u[(long)(j) * ld + (i)]
  for (int j = rem; j < k - 1; ++j)
    for (int i = 0; i <= j + 1; ++i) U(i, j) = U(i, j + 1);
  for (int j = rem; j < k - 1; ++j) {
    double c, s;
    double a = U(j, j), b = U(j + 1, j);
    givens(&a, &b, &c, &s);
    U(j, j) = a; U(j + 1, j) = 0.0;
    for (int cidx = j + 1; cidx < k - 1; ++cidx) {
      double t = c * U(j, cidx) + s * U(j + 1, cidx);
      U(j + 1, cidx) = c * U(j + 1, cidx) - s * U(j, cidx);
      U(j, cidx) = t;
    }
  }
  /* the sign fix of symmmatrix.cpp:184-194 */
  for (int i = rem; i < k - 1; ++i)
    if (U(i, i) < 0) for (int j = i; j < k - 1; ++j) U(i, j) = -U(i, j);
#undef U
}
/* symmmatrix.cpp:220-265: swap adjacent columns col, col+1 of the factor and of v */
ORC_API void orc_chol_swapadj(double* u, int k, int ld, int col, double* v)
{
#define U(i, j) u[(long)(j) * ld + (i)]
  for (int i = 0; i < col + 2; ++i) { double t = U(i, col); U(i, col) = U(i, col + 1); U(i, col + 1) = t; }
  U(col + 1, col + 1) = 0.0;  /* the entry that came from below the diagonal */
  double c, s;
  /* after the swap column `col` has entries in rows col and col+1 */
  double a = U(col, col), b = U(col + 1, col);
  /* note: the reference swapped col+2 rows, so row col+1 of column col holds the old diagonal of col+1 */
  givens(&a, &b, &c, &s);
  if (a < 0) { a = -a; s = -s; c = -c; }
  U(col, col) = a; U(col + 1, col) = b; /* b now holds drotg's z; never read again (below diagonal) */
  for (int j = col + 1; j < k; ++j) {
    double t = U(col, j) * s - U(col + 1, j) * c;
    U(col, j) = U(col + 1, j) * s + U(col, j) * c;
    U(col + 1, j) = t;
  }
  if (v) {
    double t = v[col] * s - v[col + 1] * c;
    v[col] = v[col + 1] * s + v[col] * c;
    v[col + 1] = t;
  }
#undef U
}

/* model.hpp:199-237,558-576.  xx: k x k upper (col-major, ld), tau[k] prior precisions (tau[0] of the
 * constant excluded from the log), xy[k].  Work arrays u (k*ld) and v (k) are outputs.
 * Returns log marginal likelihood (up to the reference's constant); -inf if not PD. */
ORC_API double orc_log_marginal(const double* xx, const double* tau, const double* xy, int k, int ld,
                                double nus2_plus_yy, double n_plus_nu, double* u, double* v, double* s_out)
{
  for (int c = 0; c < k; ++c) for (int r = 0; r <= c; ++r) u[(long)c * ld + r] = xx[(long)c * ld + r];
  double log_sum_2 = 0;
  u[0] += tau[0];
  for (int i = 1; i < k; ++i) { u[(long)i * ld + i] += tau[i]; log_sum_2 += log(tau[i]); }
  const double log_det_invQ = 0.5 * log_sum_2;
  if (!orc_chol_upper(u, k, ld)) { if (s_out) *s_out = INFINITY; return -INFINITY; }
  for (int i = 0; i < k; ++i) v[i] = xy[i];
  orc_solve_ut(u, k, ld, v);
  double vv = 0;
  for (int i = 0; i < k; ++i) vv += v[i] * v[i];
  const double S = nus2_plus_yy - vv;
  double ld_sum = 0;
  for (int i = 0; i < k; ++i) ld_sum += log(u[(long)i * ld + i]);
  if (s_out) *s_out = S;
  return log_det_invQ - ld_sum + (-0.5 * n_plus_nu) * log(S);
}

/* model.hpp:345-392: proportions of variance explained from fitted parts */
ORC_API void orc_pve(const double* y_hat_e, const double* y_hat_g, long n, int have_e, int have_g, double sigma2, double* pves)
{
  double* t = (double*)malloc(sizeof(double) * n);
  pves[2] = have_e ? orc_var(y_hat_e, n) : 0.0;
  pves[1] = have_g ? orc_var(y_hat_g, n) : 0.0;
  if (!have_e && !have_g) { pves[0] = pves[1] = pves[2] = 0; free(t); return; }
  for (long i = 0; i < n; ++i) t[i] = (have_g ? y_hat_g[i] : 0.0) + y_hat_e[i];
  pves[0] = !have_e ? pves[1] : (!have_g ? pves[2] : orc_var(t, n));
  const double z = pves[0] + sigma2;
  pves[0] /= z; pves[1] /= z; pves[2] /= z;
  free(t);
}

/* ------------------------------------------------------------------ */
/* 7. proposal sampler with in-order-CDF semantics                     */
/*    (discrete_distribution.hpp:125-153,217-325)                      */
/* ------------------------------------------------------------------ */
/* The reference's threaded binary tree over heap indices 0..m-1 (children 2i+1, 2i+2) samples
 * by an in-order cumulative search.  order[pos] = heap index visited pos-th in order. */
ORC_API void orc_inorder_permutation(long m, long* order)
{
  long pos = 0, node = 0, top = 0;
  long* stack = (long*)malloc(sizeof(long) * 128);
  while (top > 0 || node < m) {
    while (node < m) { stack[top++] = node; node = 2 * node + 1; }
    node = stack[--top];
    order[pos++] = node;
    node = 2 * node + 2;
  }
  free(stack);
}
/* total weight of the non-zeroed items */
ORC_API double orc_dd_total(const double* w, const uint8_t* zeroed, long m)
{
  double t = 0;
  for (long i = 0; i < m; ++i) if (!zeroed[i]) t += w[i];
  return t;
}
/* sample(): first in-order position whose cumulative weight exceeds u*total; mirrors the
 * strict "r < cumulative" tests of discrete_distribution.hpp:131-152 (falls back to the last
 * non-zeroed item, as the tree's final "upright parent" return does). */
ORC_API long orc_dd_sample(const double* w, const uint8_t* zeroed, const long* order, long m, double u, double total)
{
  const double r = u * total;
  double cum = 0;
  long last = -1;
  for (long pos = 0; pos < m; ++pos) {
    const long i = order[pos];
    if (zeroed[i]) continue;
    cum += w[i];
    last = i;
    if (r < cum) return i;
  }
  return last;
}

/* ------------------------------------------------------------------ */
/* 8. small discrete helpers (utils.cpp:46-101,144-152)                */
/* ------------------------------------------------------------------ */
ORC_API int orc_sample_discrete_naive(const double* cumsum, int m, double u)
{
  const double r = u * cumsum[m - 1];
  for (int i = 0; i < m; ++i) if (r < cumsum[i]) return i;
  return -1;
}
ORC_API long orc_sample_discrete(const double* cumsum, long m, int level, double u)
{
  long a = 0, b = m - 1;
  const double r = u * cumsum[b];
  while (level > 0) { long c = (a + b) / 2; if (r < cumsum[c]) b = c; else a = c + 1; --level; }
  for (long i = a; i <= b; ++i) if (r < cumsum[i]) return i;
  return -1;
}
ORC_API void orc_geometric_cdf(int maxsize, double p, double* values)
{
  double q = 1 - p; const double qp = q;
  for (int i = 0; i < maxsize; ++i) { values[i] = 1 - q; q *= qp; }
}

/* ------------------------------------------------------------------ */
/* 9. RNG stream (rand.hpp:36-191 over Boost.Random; see header note)  */
/* ------------------------------------------------------------------ */
typedef struct {
  uint32_t mt[624]; int idx;
  /* persistent normal_distribution state of Rand::normal_sampler */
  double n_r1, n_rho; int n_valid;
  double sinv_nu;     /* Rand::_sinvchi2_nu */
} orc_rng_t;

ORC_API void orc_rng_seed(orc_rng_t* g, uint32_t seed, double sinvchi2_nu)
{
  g->mt[0] = seed;
  for (int i = 1; i < 624; ++i) g->mt[i] = 1812433253u * (g->mt[i - 1] ^ (g->mt[i - 1] >> 30)) + (uint32_t)i;
  g->idx = 624; g->n_valid = 0; g->n_r1 = g->n_rho = 0; g->sinv_nu = sinvchi2_nu;
}
static uint32_t mt_next(orc_rng_t* g)
{
  if (g->idx >= 624) {
    for (int i = 0; i < 624; ++i) {
      uint32_t yv = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
      g->mt[i] = g->mt[(i + 397) % 624] ^ (yv >> 1) ^ ((yv & 1u) ? 0x9908b0dfu : 0u);
    }
    g->idx = 0;
  }
  uint32_t yv = g->mt[g->idx++];
  yv ^= yv >> 11; yv ^= (yv << 7) & 0x9d2c5680u; yv ^= (yv << 15) & 0xefc60000u; yv ^= yv >> 18;
  return yv;
}
ORC_API size_t orc_rng_sizeof(void) { return sizeof(orc_rng_t); }
ORC_API double orc_rng_u01(orc_rng_t* g) { return (double)mt_next(g) * (1.0 / 4294967296.0); }
ORC_API double orc_rng_normal(orc_rng_t* g)
{
  if (!g->n_valid) {
    g->n_r1 = orc_rng_u01(g);
    double r2 = orc_rng_u01(g);
    g->n_rho = sqrt(-2.0 * log(1.0 - r2));
    g->n_valid = 1;
  } else g->n_valid = 0;
  const double pi = 3.14159265358979323846;
  return g->n_rho * (g->n_valid ? cos(2.0 * pi * g->n_r1) : sin(2.0 * pi * g->n_r1));
}
static double rng_exp1(orc_rng_t* g) { return -log(1.0 - orc_rng_u01(g)); }
ORC_API double orc_rng_gamma(orc_rng_t* g, double alpha)
{
  const double pi = 3.14159265358979323846;
  if (alpha == 1.0) return rng_exp1(g);
  if (alpha > 1.0) {
    for (;;) {
      double yv = tan(pi * orc_rng_u01(g));
      double xv = sqrt(2.0 * alpha - 1.0) * yv + alpha - 1.0;
      if (xv <= 0.0) continue;
      if (orc_rng_u01(g) > (1.0 + yv * yv) * exp((alpha - 1.0) * log(xv / (alpha - 1.0)) - sqrt(2.0 * alpha - 1.0) * yv)) continue;
      return xv;
    }
  }
  const double p = exp(1.0) / (alpha + exp(1.0));
  for (;;) {
    double u = orc_rng_u01(g), yv = rng_exp1(g), xv, q;
    if (u < p) { xv = exp(-yv / alpha); q = p * exp(-xv); }
    else { xv = 1.0 + yv; q = p + (1.0 - p) * pow(xv, alpha - 1.0); }
    if (u >= q) continue;
    return xv;
  }
}
/* rand.hpp:61-78 (nu fixed at construction) and :85-97 */
ORC_API double orc_rng_sinvchi2_fixed(orc_rng_t* g, double s2)
{
  double val = -1.0;
  while (val <= 0 || !isfinite(val)) val = g->sinv_nu * s2 / (2.0 * orc_rng_gamma(g, 0.5 * g->sinv_nu));
  return val;
}
ORC_API double orc_rng_sinvchi2(orc_rng_t* g, double nu, double s2)
{
  double val = -1.0;
  while (val <= 0 || !isfinite(val)) val = nu * s2 / (2.0 * orc_rng_gamma(g, 0.5 * nu));
  return val;
}

/* ------------------------------------------------------------------ */
/* 10. probit latent update (NEW, no reference counterpart) */
/* ------------------------------------------------------------------ */
/* z_i ~ N(mu_i, 1) truncated to (0, inf) if case_i else (-inf, 0], by inverse CDF from a uniform
 * u_i in (0,1):  case: z = mu + Phi^-1( Phi(-mu) + u (1 - Phi(-mu)) ); control: z = mu + Phi^-1( u Phi(-mu) ).
 * Phi via erfc; Phi^-1 by Newton refinement of Acklam's rational approximation. */
static double phi_cdf(double x) { return 0.5 * erfc(-x * 0.70710678118654752440); }
static double phi_inv(double p)
{
  static const double a[] = {-3.969683028665376e+01, 2.209460984245205e+02, -2.759285104469687e+02, 1.383577518672690e+02, -3.066479806614716e+01, 2.506628277459239e+00};
  static const double b[] = {-5.447609879822406e+01, 1.615858368580409e+02, -1.556989798598866e+02, 6.680131188771972e+01, -1.328068155288572e+01};
  static const double c[] = {-7.784894002430293e-03, -3.223964580411365e-01, -2.400758277161838e+00, -2.549732539343734e+00, 4.374664141464968e+00, 2.938163982698783e+00};
  static const double d[] = {7.784695709041462e-03, 3.224671290700398e-01, 2.445134137142996e+00, 3.754408661907416e+00};
  double x, q, r;
  if (p <= 0) return -INFINITY;
  if (p >= 1) return INFINITY;
  if (p < 0.02425) { q = sqrt(-2 * log(p)); x = (((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) / ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1); }
  else if (p <= 1 - 0.02425) { q = p - 0.5; r = q * q; x = (((((a[0] * r + a[1]) * r + a[2]) * r + a[3]) * r + a[4]) * r + a[5]) * q / (((((b[0] * r + b[1]) * r + b[2]) * r + b[3]) * r + b[4]) * r + 1); }
  else { q = sqrt(-2 * log(1 - p)); x = -(((((c[0] * q + c[1]) * q + c[2]) * q + c[3]) * q + c[4]) * q + c[5]) / ((((d[0] * q + d[1]) * q + d[2]) * q + d[3]) * q + 1); }
  for (int it = 0; it < 2; ++it) { /* Halley steps to full double precision */
    double e = phi_cdf(x) - p;
    double uu = e * sqrt(2 * 3.14159265358979323846) * exp(x * x / 2);
    x = x - uu / (1 + x * uu / 2);
  }
  return x;
}
ORC_API void orc_probit_latent(const double* mu, const uint8_t* is_case, const double* u, long n, double* z)
{
  for (long i = 0; i < n; ++i) {
    /* s = mu (case) or -mu (control); t = Phi^-1(u Phi(s)) taken in the tail where its argument is small */
    const double sgn = is_case[i] ? 1.0 : -1.0, s = sgn * mu[i];
    const double p = u[i] * phi_cdf(s);
    const double t = p <= 0.5 ? phi_inv(p) : -phi_inv((1.0 - u[i]) + u[i] * phi_cdf(-s));
    z[i] = mu[i] - sgn * t;
  }
}
