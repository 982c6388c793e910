/*
 * msb_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 * See msb_oracle.h for scope, provenance and the parity status
 * ("parity unpinned" at the `distributions` boundary; pinned against the
 * reference's vendor/stats.py closed forms and scipy).
 *
 * Build: gcc -O2 -std=c11 -ffp-contract=off -fno-fast-math -pthread -shared -fPIC
 * (-ffp-contract=off matters: the sampler below must round exactly like the
 * device sampler, which uses explicit __fmul_rn/__fadd_rn/__fmaf_rn.)
 */
#include "msb_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ---- runtime types: include/microscopes/common/type_info.h:10-34 ---------- */
static const size_t k_prim_size[11] = {1, 1, 1, 2, 2, 4, 4, 8, 8, 4, 8};

size_t orc_prim_size(int prim) { return (prim >= 0 && prim < 11) ? k_prim_size[prim] : 0; }

/* runtime_cast::cast<T>, runtime_type.hpp:145-166: read the stored primitive,
 * convert to the model's Value type (here: widened to double, exact for every
 * primitive except 64-bit ints above 2^53, which no model Value can hold anyway) */
double orc_cast(const uint8_t *px, int prim) {
  switch (prim) {
    case 0: return (double)(*(const uint8_t *)px != 0);
    case 1: return (double)*(const int8_t *)px;
    case 2: return (double)*(const uint8_t *)px;
    case 3: { int16_t v; memcpy(&v, px, 2); return (double)v; }
    case 4: { uint16_t v; memcpy(&v, px, 2); return (double)v; }
    case 5: { int32_t v; memcpy(&v, px, 4); return (double)v; }
    case 6: { uint32_t v; memcpy(&v, px, 4); return (double)v; }
    case 7: { int64_t v; memcpy(&v, px, 8); return (double)v; }
    case 8: { uint64_t v; memcpy(&v, px, 8); return (double)v; }
    case 9: { float v; memcpy(&v, px, 4); return (double)v; }
    case 10: { double v; memcpy(&v, px, 8); return v; }
    default: return 0.0;
  }
}

/* ---- flat field layouts: distributions.hpp:21-56 (scalar), :165-199 (dd) --- */
size_t orc_hp_size(const orc_model *m) {
  switch (m->family) {
    case ORC_BB: return 2;                                 /* alpha beta */
    case ORC_BNB: return 3;                                /* alpha beta r */
    case ORC_BBNC: return 2;                               /* alpha beta (src/models/bbnc.cpp) */
    case ORC_GP: return 2;                                 /* alpha inv_beta */
    case ORC_NICH: return 4;                               /* mu kappa sigmasq nu */
    case ORC_DD: case ORC_DM: return m->dim;               /* alphas[dim] */
    case ORC_NIW: return (size_t)m->dim * m->dim + m->dim + 2; /* mu[d] kappa psi[d*d] nu */
    default: return 0;
  }
}
size_t orc_ss_size(const orc_model *m) {
  switch (m->family) {
    case ORC_BB: return 2;                                 /* heads tails */
    case ORC_BNB: return 2;                                /* count sum */
    case ORC_BBNC: return 3;                               /* p heads tails */
    case ORC_GP: return 3;                                 /* count sum log_prod */
    case ORC_NICH: return 3;                               /* count mean count_times_variance */
    case ORC_DD: return (size_t)m->dim + 1;                /* count_sum counts[dim] */
    case ORC_DM: return (size_t)m->dim + 1;                /* counts[dim] ratio (src/models/dm.cpp) */
    case ORC_NIW: return (size_t)m->dim * m->dim + m->dim + 1; /* count sum_x[d] sum_xxT[d*d] */
    default: return 0;
  }
}

/* ---- closed forms, double ------------------------------------------------- */
/* bb: log((x ? alpha+heads : beta+tails) / (alpha+beta+heads+tails))   SURVEY 8a */
static double bb_score64(const double *hp, const double *ss, double x) {
  double a = hp[0] + ss[0], b = hp[1] + ss[1];
  return log((x != 0.0 ? a : b) / (a + b));
}
/* dd: log((alpha_x + counts[x]) / (sum alpha + count_sum)) */
static double dd_score64(unsigned dim, const double *hp, const double *ss, double x) {
  double asum = 0.0;
  for (unsigned i = 0; i < dim; i++) asum += hp[i];
  long xi = (long)x;
  if (xi < 0 || xi >= (long)dim) return NAN;
  return log((hp[xi] + ss[1 + xi]) / (asum + ss[0]));
}
/* bnb (beta-negative-binomial, hp = alpha beta r, ss = count sum): posterior Beta(a, b), a = alpha + r count,
 * b = beta + sum; predictive lgamma(r+x) - lgamma(r) - lgamma(x+1) + lbeta(a + r, b + x) - lbeta(a, b) */
static double bnb_score64(const double *hp, const double *ss, double x) {
  double a = hp[0] + hp[2] * ss[0], b = hp[1] + ss[1], r = hp[2];
  return lgamma(r + x) - lgamma(r) - lgamma(x + 1.0) + (lgamma(a + r) + lgamma(b + x) - lgamma(a + r + b + x)) -
         (lgamma(a) + lgamma(b) - lgamma(a + b));
}
/* gp: a = alpha+sum, b = inv_beta+count:
 * lgamma(a+x) - lgamma(a) - lgamma(x+1) + a log b - (a+x) log(1+b) */
static double gp_score64(const double *hp, const double *ss, double x) {
  double a = hp[0] + ss[1], b = hp[1] + ss[0];
  return lgamma(a + x) - lgamma(a) - lgamma(x + 1.0) + a * log(b) - (a + x) * log1p(b);
}
/* nich: Student-t posterior predictive */
static void nich_post64(const double *hp, const double *ss, double *mu, double *kappa, double *sigmasq, double *nu) {
  double n = ss[0], mean = ss[1], ctv = ss[2];
  double mu1 = hp[0] - mean;
  *kappa = hp[1] + n;
  *mu = (hp[1] * hp[0] + mean * n) / *kappa;
  *nu = hp[3] + n;
  *sigmasq = (hp[3] * hp[2] + ctv + (n * hp[1] * mu1 * mu1) / *kappa) / *nu;
}
static double nich_score64(const double *hp, const double *ss, double x) {
  double mu, kappa, sigmasq, nu;
  nich_post64(hp, ss, &mu, &kappa, &sigmasq, &nu);
  double lambda = kappa / ((kappa + 1.0) * sigmasq);
  double t = x - mu;
  return lgamma(0.5 * nu + 0.5) - lgamma(0.5 * nu) + 0.5 * log(lambda / (M_PI * nu)) -
         (0.5 * nu + 0.5) * log1p(lambda * t * t / nu);
}

/* niw: multivariate Student-t, same decomposition as the reference's in-tree
 * multivariate_t_loglik (microscopes/common/vendor/stats.py:235-243):
 * dof = nu' - d + 1, loc = mu', scale = psi' (kappa'+1) / (kappa' dof). */
typedef struct niw_prep {
  unsigned d;
  double dof, c0;
  double *mu; /* d */
  double *L;  /* d*d lower Cholesky factor of the scale matrix, row-major */
} niw_prep;

static int chol_lower(double *A, unsigned d) { /* in place, row-major, lower */
  for (unsigned j = 0; j < d; j++) {
    double s = A[j * d + j];
    for (unsigned k = 0; k < j; k++) s -= A[j * d + k] * A[j * d + k];
    if (!(s > 0.0)) return -1;
    double ljj = sqrt(s);
    A[j * d + j] = ljj;
    for (unsigned i = j + 1; i < d; i++) {
      double t = A[i * d + j];
      for (unsigned k = 0; k < j; k++) t -= A[i * d + k] * A[j * d + k];
      A[i * d + j] = t / ljj;
    }
    for (unsigned i = 0; i < j; i++) A[i * d + j] = 0.0;
  }
  return 0;
}

static int niw_prepare(unsigned d, const double *hp, const double *ss, niw_prep *p) {
  const double *mu0 = hp, kappa0 = hp[d], *psi0 = hp + d + 1, nu0 = hp[d + 1 + (size_t)d * d];
  double n = ss[0];
  const double *sx = ss + 1, *sxx = ss + 1 + d;
  double kn = kappa0 + n, nun = nu0 + n;
  p->d = d;
  p->mu = (double *)malloc(sizeof(double) * d);
  p->L = (double *)malloc(sizeof(double) * d * d);
  for (unsigned i = 0; i < d; i++) p->mu[i] = (kappa0 * mu0[i] + sx[i]) / kn;
  p->dof = nun - (double)d + 1.0;
  double scale = (kn + 1.0) / (kn * p->dof);
  for (unsigned i = 0; i < d; i++)
    for (unsigned j = 0; j < d; j++)
      p->L[i * d + j] = (psi0[i * d + j] + sxx[i * d + j] + kappa0 * mu0[i] * mu0[j] - kn * p->mu[i] * p->mu[j]) * scale;
  if (chol_lower(p->L, d) != 0) return -1;
  double logdiag = 0.0;
  for (unsigned i = 0; i < d; i++) logdiag += log(p->L[i * d + i]);
  p->c0 = lgamma(0.5 * (p->dof + d)) - lgamma(0.5 * p->dof) - 0.5 * d * log(p->dof * M_PI) - logdiag;
  return 0;
}
static void niw_free(niw_prep *p) { free(p->mu); free(p->L); p->mu = p->L = NULL; }

static double niw_score_prepared(const niw_prep *p, const double *x) {
  unsigned d = p->d;
  double q = 0.0;
  double ybuf[256];
  double *y = d <= 256 ? ybuf : (double *)malloc(sizeof(double) * d);
  for (unsigned i = 0; i < d; i++) { /* forward substitution L y = x - mu */
    double t = x[i] - p->mu[i];
    for (unsigned k = 0; k < i; k++) t -= p->L[i * d + k] * y[k];
    y[i] = t / p->L[i * d + i];
    q += y[i] * y[i];
  }
  if (y != ybuf) free(y);
  return p->c0 - 0.5 * (p->dof + d) * log1p(q / p->dof);
}
static double niw_score64(unsigned d, const double *hp, const double *ss, const double *x) {
  niw_prep p;
  if (niw_prepare(d, hp, ss, &p) != 0) { niw_free(&p); return NAN; }
  double s = niw_score_prepared(&p, x);
  niw_free(&p);
  return s;
}

/* ---- marginal likelihood of a group's data: group::score_data (models/base.hpp:28, forwarded at
 * models/distributions.hpp:287-291).  Closed forms of the conjugate families; by the chain rule each equals
 * the sum of the sequential predictives score_value(x_i | x_<i), which is how tests/test_oracle.py pins them
 * against the (golden-pinned) predictive. ------------------------------------------------------------- */
static double lbeta(double a, double b) { return lgamma(a) + lgamma(b) - lgamma(a + b); }
static double lmvgamma(double a, unsigned d) { /* log of the multivariate gamma function Gamma_d(a) */
  double s = 0.25 * d * (d - 1.0) * log(M_PI);
  for (unsigned j = 0; j < d; j++) s += lgamma(a - 0.5 * j);
  return s;
}
static double logdet_spd(const double *A, unsigned d) { /* log |A| by Cholesky; NAN if not positive definite */
  double *L = (double *)malloc(sizeof(double) * d * d);
  memcpy(L, A, sizeof(double) * d * d);
  double s = NAN;
  if (chol_lower(L, d) == 0) {
    s = 0.0;
    for (unsigned i = 0; i < d; i++) s += 2.0 * log(L[i * d + i]);
  }
  free(L);
  return s;
}
double orc_score_data(const orc_model *m, const double *hp, const double *ss) {
  switch (m->family) {
    case ORC_BB: return lbeta(hp[0] + ss[0], hp[1] + ss[1]) - lbeta(hp[0], hp[1]);
    case ORC_DD: {
      double asum = 0.0, s = 0.0;
      for (unsigned i = 0; i < m->dim; i++) { asum += hp[i]; s += lgamma(hp[i] + ss[1 + i]) - lgamma(hp[i]); }
      return s + lgamma(asum) - lgamma(asum + ss[0]);
    }
    case ORC_DM: { /* dm.cpp:79-95 */
      double score = ss[m->dim], asum = 0.0, csum = 0.0;
      for (unsigned i = 0; i < m->dim; i++) { asum += hp[i]; csum += ss[i]; score += lgamma(ss[i] + hp[i]) - lgamma(hp[i]); }
      return score + lgamma(asum) - lgamma(asum + csum);
    }
    case ORC_BBNC: { /* bbnc.cpp:61-73: Beta(alpha, beta) density of p + Bernoulli likelihood of (heads, tails) */
      double p = ss[0];
      if (p < 0.0 || p > 1.0) return -INFINITY;
      return (hp[0] - 1.0) * log(p) + (hp[1] - 1.0) * log1p(-p) - lbeta(hp[0], hp[1]) + ss[1] * log(p) + ss[2] * log1p(-p);
    }
    case ORC_BNB: { /* the part (count, sum) determine: lbeta(a_n, b_n) - lbeta(alpha, beta) */
      double a = hp[0] + hp[2] * ss[0], b = hp[1] + ss[1];
      return lbeta(a, b) - lbeta(hp[0], hp[1]);
    }
    case ORC_GP: { /* prior Gamma(alpha, rate inv_beta); ss = count, sum, log_prod = sum log x! */
      double a = hp[0] + ss[1], b = hp[1] + ss[0];
      return lgamma(a) - lgamma(hp[0]) + hp[0] * log(hp[1]) - a * log(b) - ss[2];
    }
    case ORC_NICH: {
      double mu, kappa, sigmasq, nu, n = ss[0];
      nich_post64(hp, ss, &mu, &kappa, &sigmasq, &nu);
      return lgamma(0.5 * nu) - lgamma(0.5 * hp[3]) + 0.5 * log(hp[1] / kappa) + 0.5 * hp[3] * log(hp[3] * hp[2]) -
             0.5 * nu * log(nu * sigmasq) - 0.5 * n * log(M_PI);
    }
    case ORC_NIW: {
      unsigned d = m->dim;
      const double *mu0 = hp, kappa0 = hp[d], *psi0 = hp + d + 1, nu0 = hp[d + 1 + (size_t)d * d];
      double n = ss[0];
      const double *sx = ss + 1, *sxx = ss + 1 + d;
      double kn = kappa0 + n, nun = nu0 + n;
      double *psin = (double *)malloc(sizeof(double) * d * d);
      for (unsigned i = 0; i < d; i++)
        for (unsigned j = 0; j < d; j++) {
          double mi = (kappa0 * mu0[i] + sx[i]) / kn, mj = (kappa0 * mu0[j] + sx[j]) / kn;
          psin[i * d + j] = psi0[i * d + j] + sxx[i * d + j] + kappa0 * mu0[i] * mu0[j] - kn * mi * mj;
        }
      double r = -0.5 * n * d * log(M_PI) + lmvgamma(0.5 * nun, d) - lmvgamma(0.5 * nu0, d) +
                 0.5 * nu0 * logdet_spd(psi0, d) - 0.5 * nun * logdet_spd(psin, d) + 0.5 * d * log(kappa0 / kn);
      free(psin);
      return r;
    }
    default: return NAN;
  }
}

/* group_manager<T>::score_assignment, include/microscopes/common/group_manager.hpp:250-272, statement by
 * statement in float (libm logf where upstream has fast_log): the CRP probability of the partition in entity
 * order.  assign[i] >= 0 for all i. */
float orc_score_assignment(const int64_t *assign, size_t n, float alpha) {
  if (n == 0) return 0.f;
  int64_t maxg = 0;
  for (size_t i = 0; i < n; i++) if (assign[i] > maxg) maxg = assign[i];
  size_t *counts = (size_t *)calloc((size_t)maxg + 1, sizeof(size_t));
  counts[assign[0]] = 1;
  float sum = 0.f;
  for (size_t i = 1; i < n; i++) {
    const int64_t gid = assign[i];
    const int found = counts[gid] != 0;
    const float numer = !found ? alpha : (float)counts[gid];
    const float denom = (float)i + alpha;
    sum += logf(numer / denom);
    counts[gid]++;
  }
  free(counts);
  return sum;
}
/* the same quantity in closed form, double: depends on the partition only through the group sizes and on
 * which group holds entity 0 -- sum_g [log alpha (unless g holds entity 0) + lgamma(n_g)] - sum_{i=1}^{n-1} log(i + alpha) */
double orc_score_assignment64(const int64_t *assign, size_t n, double alpha) {
  if (n == 0) return 0.0;
  int64_t maxg = 0;
  for (size_t i = 0; i < n; i++) if (assign[i] > maxg) maxg = assign[i];
  size_t *counts = (size_t *)calloc((size_t)maxg + 1, sizeof(size_t));
  for (size_t i = 0; i < n; i++) counts[assign[i]]++;
  double s = 0.0;
  for (int64_t g = 0; g <= maxg; g++)
    if (counts[g]) s += (g == assign[0] ? 0.0 : log(alpha)) + lgamma((double)counts[g]);
  s -= lgamma((double)n + alpha) - lgamma(1.0 + alpha);
  free(counts);
  return s;
}

/* dm: src/models/dm.cpp:38-76, statement by statement; prec 64 in double, prec 32 in float with lgammaf where
 * upstream has fast_lgamma.  ss = counts[dim], ratio. */
static double dm_score64(unsigned dim, const double *hp, const double *ss, const double *x) {
  double score = 0.0, x_sum = 0.0, a_sum = 0.0, n_sum = 0.0;
  for (unsigned i = 0; i < dim; i++) {
    double xi = x[i], ai = hp[i], ni = ss[i];
    x_sum += xi; a_sum += ai; n_sum += ni;
    double e = ai + ni;
    score += lgamma(e + xi) - lgamma(e);
    score -= lgamma(xi + 1.0);
  }
  score += lgamma(x_sum + 1.0);
  score += lgamma(a_sum + n_sum) - lgamma(a_sum + n_sum + x_sum);
  return score;
}
static float dm_score32(unsigned dim, const double *hp, const double *ss, const double *x) {
  float score = 0.f, a_sum = 0.f;
  unsigned x_sum = 0, n_sum = 0;
  for (unsigned i = 0; i < dim; i++) {
    unsigned xi = (unsigned)x[i], ni = (unsigned)ss[i];
    float ai = (float)hp[i];
    x_sum += xi; a_sum += ai; n_sum += ni;
    float e = ai + ni;
    score += lgammaf(e + xi) - lgammaf(e);
    score -= lgammaf(xi + 1);
  }
  score += lgammaf(x_sum + 1);
  score += lgammaf(a_sum + n_sum) - lgammaf(a_sum + n_sum + x_sum);
  return score;
}
static double dm_ratio_term(unsigned dim, const double *x, int prec) { /* dm.cpp:9-21: lgamma(sum + 1) - sum lgamma(x_i + 1) */
  double s = 0.0, tot = 0.0;
  for (unsigned i = 0; i < dim; i++) { tot += x[i]; s -= prec == 32 ? (double)lgammaf((float)x[i] + 1.f) : lgamma(x[i] + 1.0); }
  return s + (prec == 32 ? (double)lgammaf((float)tot + 1.f) : lgamma(tot + 1.0));
}

/* ---- fp32 restatements: float arithmetic, libm logf/lgammaf.  Upstream uses
 * table-driven fast_log / fast_lgamma whose error is not reproducible here. -- */
static float bb_score32(const double *hp, const double *ss, double x) {
  float a = (float)hp[0] + (float)ss[0], b = (float)hp[1] + (float)ss[1];
  return logf((x != 0.0 ? a : b) / (a + b));
}
static float dd_score32(unsigned dim, const double *hp, const double *ss, double x) {
  float asum = 0.f;
  for (unsigned i = 0; i < dim; i++) asum += (float)hp[i];
  long xi = (long)x;
  if (xi < 0 || xi >= (long)dim) return NAN;
  return logf(((float)hp[xi] + (float)ss[1 + xi]) / (asum + (float)ss[0]));
}
static float bnb_score32(const double *hp, const double *ss, double x) {
  float a = (float)hp[0] + (float)hp[2] * (float)ss[0], b = (float)hp[1] + (float)ss[1], r = (float)hp[2], xf = (float)x;
  float s = lgammaf(r + xf) - lgammaf(r) - lgammaf(xf + 1.f);
  s += lgammaf(a + r) + lgammaf(b + xf) - lgammaf(a + r + b + xf);
  s -= lgammaf(a) + lgammaf(b) - lgammaf(a + b);
  return s;
}
static float gp_score32(const double *hp, const double *ss, double x) {
  float a = (float)hp[0] + (float)ss[1], b = (float)hp[1] + (float)ss[0];
  float xf = (float)x;
  float s = lgammaf(a + xf) - lgammaf(a) - lgammaf(xf + 1.f);
  s += a * logf(b) - (a + xf) * logf(1.f + b);
  return s;
}
static float nich_score32(const double *hp, const double *ss, double x) {
  float n = (float)ss[0], mean = (float)ss[1], ctv = (float)ss[2];
  float mu0 = (float)hp[0], k0 = (float)hp[1], s0 = (float)hp[2], nu0 = (float)hp[3];
  float mu1 = mu0 - mean;
  float kappa = k0 + n;
  float mu = (k0 * mu0 + mean * n) / kappa;
  float nu = nu0 + n;
  float sigmasq = 1.f / nu * (nu0 * s0 + ctv + (n * k0 * mu1 * mu1) / kappa);
  float lambda = kappa / ((kappa + 1.f) * sigmasq);
  float t = (float)x - mu;
  float s = lgammaf(0.5f * nu + 0.5f) - lgammaf(0.5f * nu) + 0.5f * logf(lambda / ((float)M_PI * nu));
  s += (-0.5f * nu - 0.5f) * logf(1.f + (lambda * t * t) / nu);
  return s;
}

double orc_score_value(const orc_model *m, const double *hp, const double *ss, const double *x, int prec) {
  switch (m->family) {
    case ORC_BB: return prec == 32 ? (double)bb_score32(hp, ss, x[0]) : bb_score64(hp, ss, x[0]);
    case ORC_DD: return prec == 32 ? (double)dd_score32(m->dim, hp, ss, x[0]) : dd_score64(m->dim, hp, ss, x[0]);
    case ORC_BNB: return prec == 32 ? (double)bnb_score32(hp, ss, x[0]) : bnb_score64(hp, ss, x[0]);
    case ORC_BBNC: /* bbnc.cpp:46-53: value ? log(p) : log(1. - p), p a float member */
      if (prec == 32) { float p = (float)ss[0]; return (double)(x[0] != 0.0 ? logf(p) : logf((float)(1. - p))); }
      return x[0] != 0.0 ? log(ss[0]) : log1p(-ss[0]);
    case ORC_GP: return prec == 32 ? (double)gp_score32(hp, ss, x[0]) : gp_score64(hp, ss, x[0]);
    case ORC_NICH: return prec == 32 ? (double)nich_score32(hp, ss, x[0]) : nich_score64(hp, ss, x[0]);
    case ORC_NIW: {
      double s = niw_score64(m->dim, hp, ss, x);
      return prec == 32 ? (double)(float)s : s;
    }
    case ORC_DM: return prec == 32 ? (double)dm_score32(m->dim, hp, ss, x) : dm_score64(m->dim, hp, ss, x);
    default: return NAN;
  }
}

/* ---- add_value / remove_value --------------------------------------------- */
/* integer fields are exact in either mode; nich (mean, count_times_variance)
 * follows the upstream Welford recurrences [R]; prec=32 rounds those fields to
 * float after every step like the upstream float members do. */
static double rnd(double v, int prec) { return prec == 32 ? (double)(float)v : v; }

void orc_add_value(const orc_model *m, const double *hp, double *ss, const double *x, int prec) {
  (void)hp;
  switch (m->family) {
    case ORC_BB: ss[x[0] != 0.0 ? 0 : 1] += 1.0; break;
    case ORC_DD: ss[0] += 1.0; ss[1 + (long)x[0]] += 1.0; break;
    case ORC_BNB: ss[0] += 1.0; ss[1] += x[0]; break;
    case ORC_BBNC: ss[x[0] != 0.0 ? 1 : 2] += 1.0; break;   /* bbnc.cpp:21-30 */
    case ORC_DM: /* dm.cpp:9-21 */
      for (unsigned i = 0; i < m->dim; i++) ss[i] += x[i];
      ss[m->dim] = rnd(ss[m->dim] + dm_ratio_term(m->dim, x, prec), prec);
      break;
    case ORC_GP:
      ss[0] += 1.0; ss[1] += x[0];
      ss[2] = rnd(ss[2] + rnd(prec == 32 ? (double)lgammaf((float)x[0] + 1.f) : lgamma(x[0] + 1.0), prec), prec);
      break;
    case ORC_NICH: {
      ss[0] += 1.0;
      double delta = rnd(x[0] - ss[1], prec);
      ss[1] = rnd(ss[1] + rnd(delta / ss[0], prec), prec);
      ss[2] = rnd(ss[2] + rnd(delta * rnd(x[0] - ss[1], prec), prec), prec);
      break;
    }
    case ORC_NIW: {
      unsigned d = m->dim;
      ss[0] += 1.0;
      for (unsigned i = 0; i < d; i++) ss[1 + i] = rnd(ss[1 + i] + x[i], prec);
      for (unsigned i = 0; i < d; i++)
        for (unsigned j = 0; j < d; j++)
          ss[1 + d + i * d + j] = rnd(ss[1 + d + i * d + j] + rnd(x[i] * x[j], prec), prec);
      break;
    }
    default: break;
  }
}

void orc_remove_value(const orc_model *m, const double *hp, double *ss, const double *x, int prec) {
  (void)hp;
  switch (m->family) {
    case ORC_BB: ss[x[0] != 0.0 ? 0 : 1] -= 1.0; break;
    case ORC_DD: ss[0] -= 1.0; ss[1 + (long)x[0]] -= 1.0; break;
    case ORC_BNB: ss[0] -= 1.0; ss[1] -= x[0]; break;
    case ORC_BBNC: ss[x[0] != 0.0 ? 1 : 2] -= 1.0; break;   /* bbnc.cpp:32-44 */
    case ORC_DM: /* dm.cpp:23-36 */
      for (unsigned i = 0; i < m->dim; i++) ss[i] -= x[i];
      ss[m->dim] = rnd(ss[m->dim] - dm_ratio_term(m->dim, x, prec), prec);
      break;
    case ORC_GP:
      ss[0] -= 1.0; ss[1] -= x[0];
      ss[2] = rnd(ss[2] - rnd(prec == 32 ? (double)lgammaf((float)x[0] + 1.f) : lgamma(x[0] + 1.0), prec), prec);
      break;
    case ORC_NICH: {
      double total = rnd(ss[1] * ss[0], prec);
      double delta = rnd(x[0] - ss[1], prec);
      ss[0] -= 1.0;
      if (ss[0] == 0.0) ss[1] = 0.0;
      else ss[1] = rnd(rnd(total - x[0], prec) / ss[0], prec);
      if (ss[0] <= 1.0) ss[2] = 0.0;
      else ss[2] = rnd(ss[2] - rnd(delta * rnd(x[0] - ss[1], prec), prec), prec);
      break;
    }
    case ORC_NIW: {
      unsigned d = m->dim;
      ss[0] -= 1.0;
      for (unsigned i = 0; i < d; i++) ss[1 + i] = rnd(ss[1 + i] - x[i], prec);
      for (unsigned i = 0; i < d; i++)
        for (unsigned j = 0; j < d; j++)
          ss[1 + d + i * d + j] = rnd(ss[1 + d + i * d + j] - rnd(x[i] * x[j], prec), prec);
      break;
    }
    default: break;
  }
}

/* ---- batched K x D loop --------------------------------------------------- */
typedef struct layout {
  size_t *off, *moff, *hpoff, *ssoff;
  size_t rowsize, maskrowsize, HP, SS;
} layout;

static void layout_init(layout *L, const orc_model *models, const orc_type *types, size_t D) {
  L->off = (size_t *)malloc(sizeof(size_t) * 4 * (D + 1));
  L->moff = L->off + (D + 1); L->hpoff = L->moff + (D + 1); L->ssoff = L->hpoff + (D + 1);
  size_t o = 0, mo = 0, ho = 0, so = 0;
  for (size_t d = 0; d < D; d++) { /* runtime_type.hpp:123-134 */
    L->off[d] = o; L->moff[d] = mo; L->hpoff[d] = ho; L->ssoff[d] = so;
    o += (size_t)types[d].n * orc_prim_size(types[d].prim);
    mo += types[d].n;
    ho += orc_hp_size(&models[d]);
    so += orc_ss_size(&models[d]);
  }
  L->rowsize = o; L->maskrowsize = mo; L->HP = ho; L->SS = so;
}
static void layout_free(layout *L) { free(L->off); }

static int cell_masked(const uint8_t *mrow, const layout *L, const orc_type *types, size_t d) {
  if (!mrow) return 0; /* value_accessor::anymasked, runtime_value.hpp:34-44 */
  for (unsigned i = 0; i < types[d].n; i++)
    if (mrow[L->moff[d] + i]) return 1;
  return 0;
}
static void cell_value(const uint8_t *row, const layout *L, const orc_type *types, size_t d, double *x) {
  size_t ps = orc_prim_size(types[d].prim);
  for (unsigned i = 0; i < types[d].n; i++) x[i] = orc_cast(row + L->off[d] + i * ps, types[d].prim);
}

typedef struct score_job {
  const orc_model *models; size_t D; const double *hp; const double *ss; size_t K;
  const double *logprior; const uint8_t *data; const uint8_t *mask; const orc_type *types;
  const layout *L; const niw_prep *prep; unsigned maxn;
  size_t row_lo, lo, hi; int prec; double *out64; float *out32;
} score_job;

static void *score_worker(void *arg) {
  const score_job *j = (const score_job *)arg;
  const layout *L = j->L;
  const size_t D = j->D, K = j->K;
  double *x = (double *)malloc(sizeof(double) * j->maxn);
  for (size_t i = j->lo; i < j->hi; i++) {
    const uint8_t *row = j->data + L->rowsize * i;
    const uint8_t *mrow = j->mask ? j->mask + L->maskrowsize * i : NULL;
    for (size_t k = 0; k < K; k++) {
      if (j->out32) { /* float accumulation in feature order, like the reference's float score */
        float s = (float)j->logprior[k];
        for (size_t d = 0; d < D; d++) {
          if (cell_masked(mrow, L, j->types, d)) continue;
          cell_value(row, L, j->types, d, x);
          if (j->models[d].family == ORC_NIW) s += (float)niw_score_prepared(&j->prep[k * D + d], x);
          else s += (float)orc_score_value(&j->models[d], j->hp + L->hpoff[d], j->ss + k * L->SS + L->ssoff[d], x, 32);
        }
        j->out32[(i - j->row_lo) * K + k] = s;
      } else {
        double s = j->logprior[k];
        for (size_t d = 0; d < D; d++) {
          if (cell_masked(mrow, L, j->types, d)) continue;
          cell_value(row, L, j->types, d, x);
          if (j->models[d].family == ORC_NIW) s += niw_score_prepared(&j->prep[k * D + d], x);
          else s += orc_score_value(&j->models[d], j->hp + L->hpoff[d], j->ss + k * L->SS + L->ssoff[d], x, j->prec);
        }
        j->out64[(i - j->row_lo) * K + k] = s;
      }
    }
  }
  free(x);
  return NULL;
}

static void score_rows_impl(const orc_model *models, size_t D, const double *hp, const double *ss, size_t K,
                            const double *logprior, const uint8_t *data, const uint8_t *mask,
                            const orc_type *types, size_t row_lo, size_t row_hi, int prec, int nthreads,
                            double *out64, float *out32) {
  layout L;
  layout_init(&L, models, types, D);
  /* per (group, niw feature) factorisation, computed once */
  niw_prep *prep = NULL;
  int has_niw = 0;
  for (size_t d = 0; d < D; d++) has_niw |= (models[d].family == ORC_NIW);
  if (has_niw) {
    prep = (niw_prep *)calloc(K * D, sizeof(niw_prep));
    for (size_t k = 0; k < K; k++)
      for (size_t d = 0; d < D; d++)
        if (models[d].family == ORC_NIW)
          niw_prepare(models[d].dim, hp + L.hpoff[d], ss + k * L.SS + L.ssoff[d], &prep[k * D + d]);
  }
  unsigned maxn = 1;
  for (size_t d = 0; d < D; d++) if (types[d].n > maxn) maxn = types[d].n;
  if (nthreads < 1) nthreads = 1;
  if ((size_t)nthreads > row_hi - row_lo) nthreads = (int)(row_hi - row_lo ? row_hi - row_lo : 1);
  score_job *jobs = (score_job *)calloc((size_t)nthreads, sizeof(score_job));
  pthread_t *tids = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
  size_t nrows = row_hi - row_lo;
  for (int t = 0; t < nthreads; t++) {
    score_job *j = &jobs[t];
    j->models = models; j->D = D; j->hp = hp; j->ss = ss; j->K = K; j->logprior = logprior;
    j->data = data; j->mask = mask; j->types = types; j->L = &L; j->prep = prep; j->maxn = maxn;
    j->row_lo = row_lo; j->prec = prec; j->out64 = out64; j->out32 = out32;
    j->lo = row_lo + nrows * (size_t)t / (size_t)nthreads;
    j->hi = row_lo + nrows * (size_t)(t + 1) / (size_t)nthreads;
    if (t + 1 < nthreads) pthread_create(&tids[t], NULL, score_worker, j);
  }
  score_worker(&jobs[nthreads - 1]);
  for (int t = 0; t + 1 < nthreads; t++) pthread_join(tids[t], NULL);
  free(jobs); free(tids);
  if (prep) {
    for (size_t i = 0; i < K * D; i++) if (prep[i].mu) niw_free(&prep[i]);
    free(prep);
  }
  layout_free(&L);
}

void orc_score_rows(const orc_model *models, size_t D, const double *hp, const double *ss, size_t K,
                    const double *logprior, const uint8_t *data, const uint8_t *mask,
                    const orc_type *types, size_t row_lo, size_t row_hi, int prec, int nthreads,
                    double *out) {
  score_rows_impl(models, D, hp, ss, K, logprior, data, mask, types, row_lo, row_hi, prec, nthreads, out, NULL);
}
void orc_score_rows_f32(const orc_model *models, size_t D, const double *hp, const double *ss, size_t K,
                        const double *logprior, const uint8_t *data, const uint8_t *mask,
                        const orc_type *types, size_t row_lo, size_t row_hi, int nthreads, float *out) {
  score_rows_impl(models, D, hp, ss, K, logprior, data, mask, types, row_lo, row_hi, 32, nthreads, NULL, out);
}

void orc_update_rows(const orc_model *models, size_t D, const double *hp, double *ss, size_t K,
                     double *group_counts, const uint8_t *data, const uint8_t *mask, const orc_type *types,
                     size_t row_lo, size_t row_hi, const int32_t *assign_old, const int32_t *assign_new,
                     int prec) {
  (void)K;
  layout L;
  layout_init(&L, models, types, D);
  unsigned maxn = 1;
  for (size_t d = 0; d < D; d++) if (types[d].n > maxn) maxn = types[d].n;
  double *x = (double *)malloc(sizeof(double) * maxn);
  for (size_t i = row_lo; i < row_hi; i++) {
    int32_t a = assign_old ? assign_old[i - row_lo] : -1, b = assign_new ? assign_new[i - row_lo] : -1;
    if (a == b) continue;
    const uint8_t *row = data + L.rowsize * i;
    const uint8_t *mrow = mask ? mask + L.maskrowsize * i : NULL;
    if (a >= 0 && group_counts) group_counts[a] -= 1.0; /* group_manager.hpp:235-248 */
    if (b >= 0 && group_counts) group_counts[b] += 1.0; /* group_manager.hpp:218-233 */
    for (size_t d = 0; d < D; d++) {
      if (cell_masked(mrow, &L, types, d)) continue;
      cell_value(row, &L, types, d, x);
      if (a >= 0) orc_remove_value(&models[d], hp + L.hpoff[d], ss + (size_t)a * L.SS + L.ssoff[d], x, prec);
      if (b >= 0) orc_add_value(&models[d], hp + L.hpoff[d], ss + (size_t)b * L.SS + L.ssoff[d], x, prec);
    }
  }
  free(x);
  layout_free(&L);
}

/* ---- sampler: util.hpp:125-156 -------------------------------------------- */
/* msb_expf: round-to-nearest range reduction, degree-6 polynomial
 * evaluated with fused multiply-adds, two-step power-of-two scaling.  Every
 * operation is a single correctly rounded IEEE-754 binary32 operation, so the
 * device version (explicit __f*_rn intrinsics) produces the same bits. */
float orc_expf(float x) {
  if (!(x >= -104.0f)) return x != x ? x : 0.0f;
  if (x > 88.0f) x = 88.0f;
  float kf = rintf(x * 1.44269504f);
  float r = fmaf(kf, -0.693145752f, x);
  r = fmaf(kf, -1.42860677e-6f, r);
  float p = 1.9875691500e-4f;
  p = fmaf(p, r, 1.3981999507e-3f);
  p = fmaf(p, r, 8.3334519073e-3f);
  p = fmaf(p, r, 4.1665795894e-2f);
  p = fmaf(p, r, 1.6666665459e-1f);
  p = fmaf(p, r, 5.0000001201e-1f);
  float r2 = r * r;
  p = fmaf(p, r2, r);
  p = p + 1.0f;
  int k = (int)kf;
  int k1 = k / 2, k2 = k - k1;
  union { uint32_t u; float f; } a, b;
  a.u = (uint32_t)(k1 + 127) << 23;
  b.u = (uint32_t)(k2 + 127) << 23;
  return (p * a.f) * b.f;
}

int64_t orc_sample_discrete_log(const float *scores, size_t K, float u) {
  if (K == 0) return -1;
  float m = scores[0]; /* scores_to_probs, util.hpp:125-136 */
  for (size_t k = 1; k < K; k++) if (scores[k] > m) m = scores[k];
  double acc_d = 0.0; /* std::accumulate(..., 0.) accumulates in double */
  for (size_t k = 0; k < K; k++) acc_d += (double)orc_expf(scores[k] - m);
  const float acc = (float)acc_d;
  float dart = u; /* sample_discrete, util.hpp:145-156 */
  for (size_t k = 0; k < K; k++) {
    float p = orc_expf(scores[k] - m) / acc;
    dart -= p;
    if (dart <= 0.f) return (int64_t)k;
  }
  return (int64_t)K - 1;
}

void orc_sample_rows(const float *scores, size_t nrows, size_t K, size_t ld, const float *u, int32_t *out) {
  for (size_t i = 0; i < nrows; i++) out[i] = (int32_t)orc_sample_discrete_log(scores + i * ld, K, u[i]);
}

/* The same walk with the C library's expf, i.e. exactly what the reference executes (util.hpp:131 calls expf from
 * <cmath>).  The device cannot run glibc's expf, so the bit-exact contract is stated against orc_expf above; this
 * variant exists to MEASURE how often the two exponentials lead to a different draw on given scores (bench.py's
 * parity record, tests/test_oracle.py). */
int64_t orc_sample_discrete_log_libm(const float *scores, size_t K, float u) {
  if (K == 0) return -1;
  float m = scores[0];
  for (size_t k = 1; k < K; k++) if (scores[k] > m) m = scores[k];
  double acc_d = 0.0;
  for (size_t k = 0; k < K; k++) acc_d += (double)expf(scores[k] - m);
  const float acc = (float)acc_d;
  float dart = u;
  for (size_t k = 0; k < K; k++) {
    float p = expf(scores[k] - m) / acc;
    dart -= p;
    if (dart <= 0.f) return (int64_t)k;
  }
  return (int64_t)K - 1;
}
void orc_sample_rows_libm(const float *scores, size_t nrows, size_t K, size_t ld, const float *u, int32_t *out) {
  for (size_t i = 0; i < nrows; i++) out[i] = (int32_t)orc_sample_discrete_log_libm(scores + i * ld, K, u[i]);
}

/* Philox4x32-10 (Salmon et al., SC'11): key = seed, counter = (row lo, row hi, sweep lo, sweep hi) */
void orc_philox_raw(uint64_t seed, uint64_t row, uint64_t sweep, uint32_t out[4]) {
  uint32_t c0 = (uint32_t)row, c1 = (uint32_t)(row >> 32), c2 = (uint32_t)sweep, c3 = (uint32_t)(sweep >> 32);
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
float orc_philox_u01(uint64_t seed, uint64_t row, uint64_t sweep) {
  uint32_t r[4];
  orc_philox_raw(seed, row, sweep, r);
  return (float)(r[0] >> 8) * 5.9604644775390625e-8f; /* 2^-24, in [0,1) */
}
void orc_philox_u01_rows(uint64_t seed, uint64_t row_lo, size_t n, uint64_t sweep, float *out) {
  for (size_t i = 0; i < n; i++) out[i] = orc_philox_u01(seed, row_lo + i, sweep);
}

/* ---- group::sample_value (models/base.hpp:29; distributions.hpp:293-298; bbnc.cpp:75-83; dm.cpp:100-111) ----------
 * Draws from the posterior predictive; upstream draws a parameter from the posterior and a value from the likelihood,
 * whose marginal law is the same.  Draw i reads the Philox blocks (key = seed, counter = (counter + i, TAG | block)).
 * bb / bbnc / dd: inverse CDF; gp / bnb: walk of the predictive pmf from 0 by its ratio recurrence;
 * nich: Student-t from a normal and a gamma draw (Box-Muller, Marsaglia-Tsang); niw: multivariate Student-t through
 * the Cholesky factor of the scale matrix; dm: not implemented upstream either (returns -1). */
#define ORC_DRAW_TAG 0x53414d5000000000ull
#define ORC_DRAW_WALK_CAP (1u << 26)
typedef struct draw_stream { uint64_t seed, idx; uint32_t blk, buf[4]; int pos; } draw_stream;
static uint32_t ds_next(draw_stream *r) {
  if (r->pos == 4) { orc_philox_raw(r->seed, r->idx, ORC_DRAW_TAG | r->blk, r->buf); r->blk++; r->pos = 0; }
  return r->buf[r->pos++];
}
static double ds_u53(draw_stream *r) {
  uint32_t a = ds_next(r) >> 5, b = ds_next(r) >> 6;
  return ((double)a * 67108864.0 + (double)b) * (1.0 / 9007199254740992.0);
}
static double ds_normal(draw_stream *r) {
  double u1 = ds_u53(r), u2 = ds_u53(r);
  return sqrt(-2.0 * log1p(-u1)) * cos(6.283185307179586 * u2);
}
static double ds_gamma(draw_stream *r, double k) {
  double boost = 1.0;
  if (k < 1.0) { boost = pow(1.0 - ds_u53(r), 1.0 / k); k += 1.0; }
  double d = k - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
  for (int it = 0; it < 1000; it++) {
    double z = ds_normal(r);
    double v = 1.0 + c * z;
    if (v <= 0.0) continue;
    v = v * v * v;
    double u = 1.0 - ds_u53(r);
    if (log(u) < 0.5 * z * z + d - d * v + d * log(v)) return boost * d * v;
  }
  return boost * d;
}
int orc_sample_value(const orc_model *m, const double *hp, const double *ss, uint64_t seed, uint64_t counter, size_t n,
                     double *out) {
  niw_prep prep;
  unsigned d = m->dim;
  if (m->family == ORC_DM) return -1;
  if (m->family == ORC_NIW && niw_prepare(d, hp, ss, &prep) != 0) { niw_free(&prep); return -2; }
  for (size_t i = 0; i < n; i++) {
    draw_stream rs = {seed, counter + i, 0, {0, 0, 0, 0}, 4};
    switch (m->family) {
      case ORC_BB: out[i] = ds_u53(&rs) < (hp[0] + ss[0]) / (hp[0] + hp[1] + ss[0] + ss[1]) ? 1.0 : 0.0; break;
      case ORC_BBNC: out[i] = ds_u53(&rs) < ss[0] ? 1.0 : 0.0; break;
      case ORC_DD: {
        double tot = 0.0;
        for (unsigned c = 0; c < d; c++) tot += hp[c] + ss[1 + c];
        double t = ds_u53(&rs) * tot;
        unsigned x = d - 1;
        for (unsigned c = 0; c < d; c++) { t -= hp[c] + ss[1 + c]; if (t < 0.0) { x = c; break; } }
        out[i] = (double)x;
      } break;
      case ORC_GP: case ORC_BNB: {
        double a, b, r = 0.0, p, ib = 0.0;
        if (m->family == ORC_GP) {
          a = hp[0] + ss[1]; b = hp[1] + ss[0];
          p = exp(a * (log(b) - log1p(b)));
          ib = 1.0 / (1.0 + b);
        } else {
          r = hp[2]; a = hp[0] + r * ss[0]; b = hp[1] + ss[1];
          p = exp(lgamma(a + r) + lgamma(a + b) - lgamma(a + r + b) - lgamma(a));
        }
        double t = ds_u53(&rs);
        uint32_t x = 0;
        while (t >= p && x < ORC_DRAW_WALK_CAP) {
          t -= p;
          double xd = (double)x;
          p *= m->family == ORC_GP ? (a + xd) / (xd + 1.0) * ib : (r + xd) / (xd + 1.0) * ((b + xd) / (a + r + b + xd));
          x++;
          if (p == 0.0) break;
        }
        out[i] = (double)x;
      } break;
      case ORC_NICH: {
        double mu, kappa, sigmasq, nu;
        nich_post64(hp, ss, &mu, &kappa, &sigmasq, &nu);
        double lambda = kappa / ((kappa + 1.0) * sigmasq);
        double g = ds_gamma(&rs, 0.5 * nu);
        double z = ds_normal(&rs);
        double tv = z * sqrt(0.5 * nu / g);
        out[i] = mu + tv / (sqrt(lambda / nu) * sqrt(nu));
      } break;
      case ORC_NIW: {
        double *o = out + i * (size_t)d;
        double g = ds_gamma(&rs, 0.5 * prep.dof);
        double s = sqrt(0.5 * prep.dof / g);
        for (unsigned a = 0; a < d; a++) o[a] = prep.mu[a];
        for (unsigned j = 0; j < d; j++) {
          double z = ds_normal(&rs) * s;
          for (unsigned a = j; a < d; a++) o[a] += prep.L[a * d + j] * z;
        }
      } break;
      default: return -1;
    }
  }
  if (m->family == ORC_NIW) niw_free(&prep);
  return 0;
}
