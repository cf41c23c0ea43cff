// oracle/shim/gsl/gsl_randist.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// Distributions used by the reference, following GSL's
// published algorithms: Marsaglia-Tsang gamma (with the a<1 boost),
// beta from two gammas (Johnk for a,b<=1), multinomial by conditional
// binomials, and GSL's gamma/beta density formulas.  The Gaussian used inside
// the gamma sampler is a polar Box-Muller draw instead of GSL's ziggurat, so
// random STREAMS differ from real GSL while DISTRIBUTIONS are identical --
// chain parity is distributional.
#ifndef PA_SHIM_GSL_RANDIST_H
#define PA_SHIM_GSL_RANDIST_H
#include <math.h>
#include <stddef.h>
#include "gsl_rng.h"

static inline double gsl_ran_gaussian(gsl_rng* r, double sigma) {
  double x, y, r2;
  do {
    x = -1 + 2 * gsl_rng_uniform_pos(r);
    y = -1 + 2 * gsl_rng_uniform_pos(r);
    r2 = x * x + y * y;
  } while (r2 > 1.0 || r2 == 0);
  return sigma * y * sqrt(-2.0 * log(r2) / r2);
}

static inline double gsl_ran_gamma(gsl_rng* r, double a, double b) {
  if (a < 1) {
    double u = gsl_rng_uniform_pos(r);
    return gsl_ran_gamma(r, 1.0 + a, b) * pow(u, 1.0 / a);
  }
  double x, v, u;
  double d = a - 1.0 / 3.0;
  double c = (1.0 / 3.0) / sqrt(d);
  while (1) {
    do {
      x = gsl_ran_gaussian(r, 1.0);
      v = 1.0 + c * x;
    } while (v <= 0);
    v = v * v * v;
    u = gsl_rng_uniform_pos(r);
    if (u < 1 - 0.0331 * x * x * x * x) break;
    if (log(u) < 0.5 * x * x + d * (1 - v + log(v))) break;
  }
  return b * d * v;
}

static inline double gsl_ran_beta(gsl_rng* r, double a, double b) {
  if (a <= 1.0 && b <= 1.0) {
    double U, V, X, Y;
    while (1) {  // Johnk's algorithm
      U = gsl_rng_uniform_pos(r);
      V = gsl_rng_uniform_pos(r);
      X = pow(U, 1.0 / a);
      Y = pow(V, 1.0 / b);
      if ((X + Y) <= 1.0) {
        if (X + Y > 0) return X / (X + Y);
        double logX = log(U) / a, logY = log(V) / b;
        double logM = logX > logY ? logX : logY;
        logX -= logM; logY -= logM;
        return exp(logX - log(exp(logX) + exp(logY)));
      }
    }
  }
  double x1 = gsl_ran_gamma(r, a, 1.0);
  double x2 = gsl_ran_gamma(r, b, 1.0);
  return x1 / (x1 + x2);
}

static inline unsigned int gsl_ran_binomial(gsl_rng* r, double p, unsigned int n) {
  unsigned int k = 0;
  for (unsigned int i = 0; i < n; i++) if (gsl_rng_uniform(r) < p) k++;
  return k;
}

static inline void gsl_ran_multinomial(gsl_rng* r, size_t K, unsigned int N, const double p[], unsigned int n[]) {
  double norm = 0.0, sum_p = 0.0;
  unsigned int sum_n = 0;
  for (size_t k = 0; k < K; k++) norm += p[k];
  for (size_t k = 0; k < K; k++) {
    if (p[k] > 0.0) n[k] = gsl_ran_binomial(r, p[k] / (norm - sum_p), N - sum_n);
    else n[k] = 0;
    sum_p += p[k];
    sum_n += n[k];
  }
}

/* exponential with MEAN mu (GSL's parameterisation), used by PhyloAcc-GT's coalescent times (genetree.cpp:414,1004) */
static inline double gsl_ran_exponential(gsl_rng* r, double mu) {
  double u;
  do { u = gsl_rng_uniform(r); } while (u <= 0.0);
  return -mu * std::log(u);
}

static inline double gsl_ran_gamma_pdf(double x, double a, double b) {
  if (x < 0) return 0;
  if (x == 0) return (a == 1) ? 1 / b : 0;
  if (a == 1) return exp(-x / b) / b;
  return exp((a - 1) * log(x / b) - x / b - lgamma(a)) / b;
}

static inline double gsl_ran_beta_pdf(double x, double a, double b) {
  if (x < 0 || x > 1) return 0;
  double gab = lgamma(a + b), ga = lgamma(a), gb = lgamma(b);
  if (x == 0.0 || x == 1.0) {
    if (a > 1.0 && b > 1.0) return 0.0;
    return exp(gab - ga - gb) * pow(x, a - 1) * pow(1 - x, b - 1);
  }
  return exp(gab - ga - gb + log(x) * (a - 1) + log1p(-x) * (b - 1));
}
#endif
