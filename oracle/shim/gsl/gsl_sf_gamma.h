// oracle/shim/gsl/gsl_sf_gamma.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
#ifndef PA_SHIM_GSL_SF_GAMMA_H
#define PA_SHIM_GSL_SF_GAMMA_H
#include <math.h>
static inline double gsl_sf_lngamma(double x) { return lgamma(x); }
static inline double gsl_sf_lnbeta(double a, double b) { return lgamma(a) + lgamma(b) - lgamma(a + b); }
#endif
