// oracle/shim/gsl/gsl_cdf.h -- TEST INFRASTRUCTURE: included by the reference, nothing used.
#ifndef PA_SHIM_gsl_cdf
#define PA_SHIM_gsl_cdf
#define GSL_EDOM 1
#endif
