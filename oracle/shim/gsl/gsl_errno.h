// oracle/shim/gsl/gsl_errno.h -- TEST INFRASTRUCTURE: included by the reference, nothing used.
#ifndef PA_SHIM_gsl_errno
#define PA_SHIM_gsl_errno
#define GSL_EDOM 1
#endif
