// oracle/shim/gsl/gsl_rng.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// Minimal stand-in for the GSL random-number API used by the unmodified
// reference sources.  gsl_rng_default is mt19937 with the
// 2002 seeding routine, as in GSL; gsl_rng_uniform() = genrand_int32 / 2^32.
#ifndef PA_SHIM_GSL_RNG_H
#define PA_SHIM_GSL_RNG_H
#include <stdlib.h>
#include <stdint.h>

typedef struct { int dummy; } gsl_rng_type;
static const gsl_rng_type pa_shim_mt19937 = {0};
static const gsl_rng_type* const gsl_rng_default = &pa_shim_mt19937;
static const gsl_rng_type* const gsl_rng_mt19937 = &pa_shim_mt19937;

typedef struct {
  uint32_t mt[624];
  int mti;
} gsl_rng;

static inline void gsl_rng_set(gsl_rng* r, unsigned long int s) {
  if (s == 0) s = 4357;
  r->mt[0] = (uint32_t)(s & 0xffffffffUL);
  for (int i = 1; i < 624; i++)
    r->mt[i] = (uint32_t)(1812433253UL * (r->mt[i - 1] ^ (r->mt[i - 1] >> 30)) + (uint32_t)i);
  r->mti = 624;
}
static inline gsl_rng* gsl_rng_alloc(const gsl_rng_type* t) {
  (void)t;
  gsl_rng* r = (gsl_rng*)malloc(sizeof(gsl_rng));
  gsl_rng_set(r, 0);
  return r;
}
static inline void gsl_rng_free(gsl_rng* r) { free(r); }

static inline uint32_t gsl_rng_get(gsl_rng* r) {
  const uint32_t UPPER = 0x80000000u, LOWER = 0x7fffffffu;
  uint32_t* mt = r->mt;
  if (r->mti >= 624) {
    int kk;
    for (kk = 0; kk < 624 - 397; kk++) {
      uint32_t y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
      mt[kk] = mt[kk + 397] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    for (; kk < 623; kk++) {
      uint32_t y = (mt[kk] & UPPER) | (mt[kk + 1] & LOWER);
      mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    uint32_t y = (mt[623] & UPPER) | (mt[0] & LOWER);
    mt[623] = mt[396] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    r->mti = 0;
  }
  uint32_t k = mt[r->mti++];
  k ^= (k >> 11);
  k ^= (k << 7) & 0x9d2c5680u;
  k ^= (k << 15) & 0xefc60000u;
  k ^= (k >> 18);
  return k;
}
static inline double gsl_rng_uniform(gsl_rng* r) { return gsl_rng_get(r) / 4294967296.0; }
static inline double gsl_rng_uniform_pos(gsl_rng* r) {
  double x;
  do { x = gsl_rng_uniform(r); } while (x == 0);
  return x;
}
static inline unsigned long int gsl_rng_uniform_int(gsl_rng* r, unsigned long int n) {
  return (unsigned long int)(gsl_rng_uniform(r) * (double)n);
}
#endif
