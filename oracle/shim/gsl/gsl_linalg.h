// oracle/shim/gsl/gsl_linalg.h -- TEST INFRASTRUCTURE: included by the reference, nothing used.
#ifndef PA_SHIM_gsl_linalg
#define PA_SHIM_gsl_linalg
#define GSL_EDOM 1
#endif
