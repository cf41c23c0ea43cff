/* oracle/oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C restatement of the floating-point part of the reference's PhyloAcc-ST
 * per-element chain (log-space, column by column, exactly as the reference
 * computes it).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load this library; the product never links it.
 *
 * Parity status: PINNED.  The deterministic functions (P(t), pruning,
 * emissions, Z messages, Eval2) are checked against dumps of the unmodified
 * reference (oracle/_ref/ref_probe) by tests/test_oracle_pinning.py; the
 * sampler is checked distributionally against runs of the reference binary
 * built by oracle/Makefile (fixtures under tests/golden/).
 *
 * Randomness: the reference draws from one mt19937 stream per element
 * (bpp_c.hpp:146-147); exact GSL streams cannot be reproduced without GSL, so
 * chain parity with the reference is distributional.  This
 * oracle instead defines a counter-based Philox4x32-10 draw schedule (see
 * pa_or_uniform) that the CUDA engine follows too, which makes oracle-vs-CUDA
 * chains comparable draw for draw.
 */
#ifndef PA_ORACLE_H
#define PA_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pa_or_element {
  int32_t N, S, GG;            /* nodes (original numbering), species, columns after trimming */
  const int8_t* tg;            /* [GG][N] initial Tg: tips 0..3, -1 ambiguous, 4 gap, 5 other; internal -1 or 4 */
  const uint8_t* masks;        /* [GG][S] allowed-base bit masks of the tips */
  const uint8_t* missing;      /* [N] */
  int32_t root;
  int32_t n_nodes;
  const int32_t* nodes;        /* post-order of the pruned tree (bpp_c.hpp:395) */
  const int32_t* children2;    /* [N][2] */
  const int32_t* parent2;      /* [N], parent2[root] = N */
  const double* distances2;    /* [N] */
  const double* distances;     /* [N] un-pruned branch lengths (null P(t) cache of BPP) */
  int32_t n_upper_c;
  const int32_t* upper_c;
  int32_t n_upper_conserve_c;
  const int32_t* upper_conserve_c;
  /* substitution model: eigenvec = V^T, eigeninv = inv(V^T) (bpp.hpp:357-362), row-major */
  double eigenval[4], eigenvec[16], eigeninv[16], log_pi[4];
  /* parameters (main.cpp:49-90) */
  double cprior_a, cprior_b, nprior_a, nprior_b;
  double prior_l_a, prior_l_b, prior_l2_a, prior_l2_b, prior_g_a, prior_g_b;
  double init_crate, init_nrate, init_lrate, init_lrate2, init_grate;
  double cub, nlb, consToMis;
  int32_t ropt, num_burn, num_mcmc, adaptive_freq;
} pa_or_element;

typedef struct pa_or_trace {   /* caller-allocated, T = num_burn + num_mcmc rows */
  double* loglik;              /* [T] trace_loglik */
  double* full_loglik;         /* [T] trace_full_loglik */
  double* n_rate; double* c_rate; double* g_rate; double* l_rate; double* l2_rate; /* [T] each */
  int8_t* Z;                   /* [T][N] */
  int8_t* max_Z;               /* [N] */
  double max_loglik;
  int32_t max_m;
  int32_t accept_n, accept_c;  /* in: carried acceptance counters (bpp_c.hpp:123-124); out: final values */
  int32_t failure;
} pa_or_trace;

/* log P(t*r)^T in the reference's (child,parent) layout, row-major [16] (bpp.hpp:370-397) */
void pa_or_log_tm(const pa_or_element* el, double t, double rate, double* out16);

/* replayed log-likelihood: initMCMC TM caches + Update_Tg(false) + log_exp_sum (bpp_c.cpp:131-236,647-671) */
double pa_or_loglik(const pa_or_element* el, int resZ, const int8_t* Z, double n_rate, double c_rate, double* persite);

/* getEmission + upward part of Update_Z_subtree for a given history (bpp_c2.cpp:355-390, bpp_c.cpp:1186-1212,1223).
 * tg_hist [GG][N] holds sampled bases for internal nodes. emission/back are [N][3]. */
void pa_or_emission_back(const pa_or_element* el, int resZ, const int8_t* Z, double n_rate, double c_rate, double g_rate,
                         double l_rate, double l2_rate, const int8_t* tg_hist, double* emission, double* back);

/* full Gibbs chain for one model (bpp_c.cpp:131-343 initMCMC, 346-519 Gibbs) with the Philox draw schedule */
int pa_or_chain(const pa_or_element* el, int resZ, uint32_t seed, uint32_t chain_key, pa_or_trace* tr);

/* Eval2 marginal-likelihood estimate from a trace (bpp_c2.cpp:14-353) */
double pa_or_eval2(const pa_or_element* el, int resZ, const pa_or_trace* tr, int32_t* failure);

/* draw schedule: Philox4x32-10, key = (seed, chain_key), counter = (idx, sub, iter, stream); returns 4 words */
void pa_or_philox(uint32_t seed, uint32_t chain_key, uint32_t idx, uint32_t sub, uint32_t iter, uint32_t stream, uint32_t out[4]);

/* samplers on top of the schedule (exposed for unit tests) */
double pa_or_gamma(uint32_t seed, uint32_t chain_key, uint32_t idx, uint32_t sub0, uint32_t iter, uint32_t stream, double a, double b);

double pa_or_gamma_pdf(double x, double a, double b);
double pa_or_lnbeta(double a, double b);

#ifdef __cplusplus
}
#endif
#endif
