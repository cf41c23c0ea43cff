/* oracle/oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See oracle.h.
 *
 * Literal, log-space, column-by-column restatement of the reference's
 * PhyloAcc-ST per-element chain.  File:line citations refer to
 * /root/reference/src/PhyloAcc-ST/.
 */
#include "oracle.h"

#define _GNU_SOURCE
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#define NB 4
#define NEG_INF (-INFINITY)

/* ------------------------------------------------------------------ Philox4x32-10 draw schedule */
static inline void philox_round(uint32_t c[4], const uint32_t k[2]) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
  const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
  const uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

void pa_or_philox(uint32_t seed, uint32_t chain_key, uint32_t idx, uint32_t sub, uint32_t iter, uint32_t stream, uint32_t out[4]) {
  uint32_t c[4] = {idx, sub, iter, stream};
  uint32_t k[2] = {seed, chain_key};
  for (int r = 0; r < 10; r++) {
    philox_round(c, k);
    k[0] += 0x9E3779B9u;
    k[1] += 0xBB67AE85u;
  }
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

/* streams */
enum { ST_HIST = 0, ST_Z = 1, ST_RATE = 2, ST_TRANS = 3 };

static inline double u01(uint32_t w) { return ((double)w + 0.5) * 2.3283064365386963e-10; } /* (0,1), 32-bit like mt19937 */

/* Marsaglia-Tsang gamma(shape a, scale b) (the algorithm GSL's gsl_ran_gamma uses), driven by schedule calls
 * (idx, sub0 + attempt): words 0,1 -> Box-Muller normal, word 2 -> acceptance uniform, word 3 of the first call -> a<1 boost. */
double pa_or_gamma(uint32_t seed, uint32_t chain_key, uint32_t idx, uint32_t sub0, uint32_t iter, uint32_t stream, double a, double b) {
  uint32_t w[4];
  double boost = 1.0;
  pa_or_philox(seed, chain_key, idx, sub0, iter, stream, w);
  if (a < 1.0) {
    boost = pow(u01(w[3]), 1.0 / a);
    a += 1.0;
  }
  const double d = a - 1.0 / 3.0;
  const double c = (1.0 / 3.0) / sqrt(d);
  for (uint32_t attempt = 0;; attempt++) {
    if (attempt > 0) pa_or_philox(seed, chain_key, idx, sub0 + attempt, iter, stream, w);
    if (attempt > 1000) return NAN; /* non-finite parameters: never spin */
    const double x = sqrt(-2.0 * log(u01(w[0]))) * cos(6.283185307179586 * u01(w[1]));
    double v = 1.0 + c * x;
    if (v <= 0) continue;
    v = v * v * v;
    const double u = u01(w[2]);
    if (u < 1 - 0.0331 * x * x * x * x) return b * d * v * boost;
    if (log(u) < 0.5 * x * x + d * (1 - v + log(v))) return b * d * v * boost;
  }
}

static double sched_beta(uint32_t seed, uint32_t key, uint32_t which, uint32_t iter, double a, double b) {
  const double x1 = pa_or_gamma(seed, key, which, 0, iter, ST_TRANS, a, 1.0);
  const double x2 = pa_or_gamma(seed, key, which, 4096, iter, ST_TRANS, b, 1.0);
  return x1 / (x1 + x2);
}

/* first index k with u*total < cumsum_k, weights summed in index order; falls back to the last positive weight */
static int categorical(const double* p, int K, double u) {
  double tot = 0;
  for (int k = 0; k < K; k++) tot += p[k];
  const double x = u * tot;
  double acc = 0;
  int last = -1;
  for (int k = 0; k < K; k++) {
    if (p[k] > 0) last = k;
    acc += p[k];
    if (x < acc) return k;
  }
  return last;
}

/* ------------------------------------------------------------------ densities (GSL formulas) */
double pa_or_gamma_pdf(double x, double a, double b) { /* gsl_ran_gamma_pdf */
  if (x < 0) return 0;
  if (x == 0) return (a == 1) ? 1 / b : 0;
  if (a == 1) return exp(-x / b) / b;
  return exp((a - 1) * log(x / b) - x / b - lgamma(a)) / b;
}
double pa_or_lnbeta(double a, double b) { return lgamma(a) + lgamma(b) - lgamma(a + b); }

/* ------------------------------------------------------------------ log-space helpers (bpp.hpp:427-485) */
/* log_multi: tm is (child i, parent j) row-major; returns res[j] = log sum_i exp(tm(i,j) + y[i]) with ONE max over the 4x4 */
static void log_multi(const double* tm, const double* y, double* res) {
  double x[16];
  double mx = NEG_INF;
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      x[i * 4 + j] = tm[i * 4 + j] + y[i];
      if (x[i * 4 + j] > mx) mx = x[i * 4 + j];
    }
  if (mx == NEG_INF) {
    for (int j = 0; j < 4; j++) res[j] = NEG_INF;
    return;
  }
  for (int j = 0; j < 4; j++) {
    double s = 0;
    for (int i = 0; i < 4; i++) s += exp(x[i * 4 + j] - mx);
    res[j] = log(s) + mx;
  }
}

/* log_multi2 on 3x3 (bpp.hpp:447-460): per-column max */
static void log_multi2(const double* t, const double* y, double* res) {
  for (int j = 0; j < 3; j++) {
    double col[3], mx = NEG_INF, s = 0;
    for (int i = 0; i < 3; i++) {
      col[i] = t[i * 3 + j] + y[i];
      if (col[i] > mx) mx = col[i];
    }
    for (int i = 0; i < 3; i++) s += exp(col[i] - mx);
    res[j] = log(s) + mx;
  }
}

static double log_exp_sum(const double* v, int n) {
  double mx = NEG_INF, s = 0;
  for (int i = 0; i < n; i++)
    if (v[i] > mx) mx = v[i];
  if (mx == NEG_INF) return NEG_INF;
  for (int i = 0; i < n; i++) s += exp(v[i] - mx);
  return log(s) + mx;
}

static void log_sample(const double* v, int n, double* p) { /* unnormalised probabilities */
  double mx = v[0];
  for (int i = 1; i < n; i++)
    if (v[i] > mx) mx = v[i];
  for (int i = 0; i < n; i++) p[i] = exp(v[i] - mx);
}

/* log(eigeninv * (diag(exp(eigenval*t*rate)) * eigenvec)) (bpp.hpp:375-380) */
void pa_or_log_tm(const pa_or_element* el, double t, double rate, double* out) {
  double x[16];
  for (int m = 0; m < 4; m++) {
    const double e = exp(el->eigenval[m] * t * rate);
    for (int b = 0; b < 4; b++) x[m * 4 + b] = el->eigenvec[m * 4 + b] * e;
  }
  for (int a = 0; a < 4; a++)
    for (int b = 0; b < 4; b++) {
      double s = 0;
      for (int m = 0; m < 4; m++) s += el->eigeninv[a * 4 + m] * x[m * 4 + b];
      out[a * 4 + b] = log(s);
    }
}

static void identity_log_tm(double* out) {
  for (int i = 0; i < 16; i++) out[i] = NEG_INF;
  for (int i = 0; i < 4; i++) out[i * 4 + i] = 0;
}

/* ------------------------------------------------------------------ chain state */
typedef struct {
  const pa_or_element* el;
  int N, S, GG, n_nodes, n_int, root;
  const int32_t* nodes;
  int32_t* internal;       /* internal_nodes (post-order) */
  double* lambda;          /* [GG][N][4] */
  int8_t* Tg;              /* [GG][N] */
  int8_t* Z;
  int8_t* fixZ;
  double *tm_null, *tm_cons, *tm_neut; /* [N][16] */
  double* tm_int;          /* [N][9] (child i, parent j) */
  double* back;            /* [N][3] log_prob_back */
  double* emis;            /* [N][3] log_emission */
  double* ambig_null;      /* [GG][S][4] */
  double prop_c, prop_n;
  int accept_n, accept_c;
  int failure;
  uint32_t seed, key;
} chain_t;

#define LAM(st, g, s) ((st)->lambda + (((size_t)(g) * (st)->N + (s)) * 4))
#define TG(st, g, s) ((st)->Tg[(size_t)(g) * (st)->N + (s)])

static chain_t* chain_new(const pa_or_element* el) {
  chain_t* st = (chain_t*)calloc(1, sizeof(chain_t));
  st->el = el;
  st->N = el->N; st->S = el->S; st->GG = el->GG; st->n_nodes = el->n_nodes; st->root = el->root;
  st->nodes = el->nodes;
  const int N = st->N, GG = st->GG, S = st->S;
  st->internal = (int32_t*)malloc(sizeof(int32_t) * N);
  st->n_int = 0;
  for (int i = 0; i < el->n_nodes; i++)
    if (el->nodes[i] >= S) st->internal[st->n_int++] = el->nodes[i];
  st->lambda = (double*)calloc((size_t)GG * N * 4, sizeof(double));
  st->Tg = (int8_t*)malloc((size_t)GG * N);
  memcpy(st->Tg, el->tg, (size_t)GG * N);
  st->Z = (int8_t*)malloc(N);
  st->fixZ = (int8_t*)malloc(N);
  st->tm_null = (double*)calloc((size_t)N * 16, sizeof(double));
  st->tm_cons = (double*)calloc((size_t)N * 16, sizeof(double));
  st->tm_neut = (double*)calloc((size_t)N * 16, sizeof(double));
  st->tm_int = (double*)calloc((size_t)N * 9, sizeof(double));
  st->back = (double*)calloc((size_t)N * 3, sizeof(double));
  st->emis = (double*)calloc((size_t)N * 3, sizeof(double));
  st->ambig_null = (double*)calloc((size_t)GG * S * 4, sizeof(double));
  /* tips (bpp_c.hpp:208-264): lambda = 0 on allowed bases, -inf otherwise */
  for (int g = 0; g < GG; g++)
    for (int s = 0; s < S; s++) {
      const uint8_t m = el->masks[(size_t)g * S + s];
      for (int b = 0; b < 4; b++) LAM(st, g, s)[b] = (m >> b) & 1 ? 0.0 : NEG_INF;
    }
  /* missing-node emission constants (bpp_c.hpp:448-474), nconsToMis = 1 */
  for (int s = 0; s < N; s++)
    if (el->missing[s]) {
      st->emis[s * 3 + 2] = log(1.0);
      st->emis[s * 3 + 1] = log(el->consToMis);
      st->emis[s * 3 + 0] = log(1.0);
    }
  return st;
}

static void chain_free(chain_t* st) {
  free(st->internal); free(st->lambda); free(st->Tg); free(st->Z); free(st->fixZ);
  free(st->tm_null); free(st->tm_cons); free(st->tm_neut); free(st->tm_int); free(st->back); free(st->emis);
  free(st->ambig_null); free(st);
}

/* BPP ctor null cache (bpp.hpp:366-397) followed by initMCMC (bpp_c.cpp:131-343).  Z_in == NULL -> iter==0 default Z. */
static void init_mcmc(chain_t* st, int resZ, const int8_t* Z_in, double n_rate, double c_rate, double g_rate, double l_rate, double l2_rate) {
  const pa_or_element* el = st->el;
  const int N = st->N, S = st->S, GG = st->GG;
  st->failure = 0;
  st->prop_c = 1.2;
  st->prop_n = (double)1 / GG * 10;
  if (Z_in == NULL) {
    for (int s = 0; s < N; s++) st->Z[s] = -1;
    for (int i = 0; i < st->n_nodes; i++) st->Z[st->nodes[i]] = 1;
    st->Z[st->root] = 0;
  } else {
    memcpy(st->Z, Z_in, N);
  }
  for (int s = 0; s < N; s++) st->fixZ[s] = -1;
  for (int i = 0; i < el->n_upper_c; i++) st->fixZ[el->upper_c[i]] = 1;
  st->fixZ[st->root] = 1;

  /* bpp.log_cache_TM_null: N-1 branches, identity when t <= 1e-10; entry N-1 stays all-zero (bpp.hpp:366) */
  for (int s = 0; s < N; s++) {
    double* tm = st->tm_null + (size_t)s * 16;
    if (s == N - 1) { memset(tm, 0, sizeof(double) * 16); continue; }
    if (el->distances[s] > 1e-10) pa_or_log_tm(el, el->distances[s], 1.0, tm);
    else identity_log_tm(tm);
  }
  for (int i = 0; i < st->n_nodes - 1; i++) { /* bpp_c.cpp:187-228 */
    const int s = st->nodes[i];
    const double t = el->distances2[s];
    if (t > 1e-10) {
      if (t != el->distances[s]) pa_or_log_tm(el, t, 1.0, st->tm_null + (size_t)s * 16);
      pa_or_log_tm(el, t, n_rate, st->tm_neut + (size_t)s * 16);
      pa_or_log_tm(el, t, c_rate, st->tm_cons + (size_t)s * 16);
    } else {
      memcpy(st->tm_cons + (size_t)s * 16, st->tm_null + (size_t)s * 16, sizeof(double) * 16);
      memcpy(st->tm_neut + (size_t)s * 16, st->tm_null + (size_t)s * 16, sizeof(double) * 16);
    }
  }
  memset(st->back, 0, sizeof(double) * N * 3);
  for (int s = S; s < N; s++)
    for (int g = 0; g < GG; g++) memset(LAM(st, g, s), 0, sizeof(double) * 4);
  /* getEmission_ambig (bpp_c.cpp:116-128) */
  for (int s = 0; s < S; s++) {
    if (el->missing[s]) continue;
    for (int g = 0; g < GG; g++) {
      if (TG(st, g, s) != -1) continue;
      log_multi(st->tm_null + (size_t)s * 16, LAM(st, g, s), st->ambig_null + ((size_t)g * S + s) * 4);
    }
  }
  if (resZ == 2) {
    for (int i = 0; i < el->n_upper_conserve_c; i++) st->fixZ[el->upper_conserve_c[i]] = 1;
  } else if (resZ == 0) {
    for (int i = 0; i < st->n_nodes - 1; i++) st->fixZ[st->nodes[i]] = 1;
  }
  for (int i = 0; i < st->n_nodes; i++) { /* bpp_c.cpp:301-332 */
    const int s = st->nodes[i];
    double* t = st->tm_int + (size_t)s * 9;
    if (st->fixZ[s] == 1) {
      t[0 * 3 + 1] = NEG_INF; t[1 * 3 + 1] = 0; t[2 * 3 + 1] = NEG_INF;
      t[0 * 3 + 0] = log(1 - g_rate); t[1 * 3 + 0] = log(g_rate); t[2 * 3 + 0] = NEG_INF;
    } else {
      t[0 * 3 + 0] = log(1 - g_rate); t[1 * 3 + 0] = log(g_rate); t[2 * 3 + 0] = NEG_INF;
      const double y = 1 - l_rate;
      t[0 * 3 + 1] = NEG_INF; t[1 * 3 + 1] = log(y); t[2 * 3 + 1] = log(1 - y);
      t[0 * 3 + 2] = NEG_INF; t[1 * 3 + 2] = log(l2_rate); t[2 * 3 + 2] = log(1 - l2_rate);
    }
  }
}

static const double* tm_for(const chain_t* st, int chi) {
  if (st->Z[chi] == 2) return st->tm_neut + (size_t)chi * 16;
  if (st->Z[chi] == 0) return st->tm_null + (size_t)chi * 16;
  return st->tm_cons + (size_t)chi * 16;
}

/* Update_Tg(g, visited, false) (bpp_c.cpp:647-671) */
static void update_tg_up(chain_t* st, int g, const uint8_t* visited) {
  const pa_or_element* el = st->el;
  double res[4];
  for (int q = 0; q < st->n_int; q++) {
    const int s = st->internal[q];
    if (visited[s] || el->missing[s] || TG(st, g, s) >= NB) continue;
    for (int cc = 0; cc < 2; cc++) {
      const int chi = el->children2[s * 2 + cc];
      if (el->missing[chi] || TG(st, g, chi) >= NB) continue;
      log_multi(tm_for(st, chi), LAM(st, g, chi), res);
      for (int b = 0; b < 4; b++) LAM(st, g, s)[b] += res[b];
    }
  }
  if (!visited[st->root])
    for (int b = 0; b < 4; b++) LAM(st, g, st->root)[b] += el->log_pi[b];
}

/* Update_Tg(g, ., true) (bpp_c.cpp:672-733): draw (site g, internal index q) uses word q&3 of schedule call (g, q>>2, m, ST_HIST) */
static void update_tg_sample(chain_t* st, int g, uint32_t m) {
  const pa_or_element* el = st->el;
  double p[4], v[4];
  uint32_t w[4];
  int have_grp = -1;
  {
    const int q = st->n_int - 1;
    pa_or_philox(st->seed, st->key, (uint32_t)g, (uint32_t)(q >> 2), m, ST_HIST, w);
    have_grp = q >> 2;
    log_sample(LAM(st, g, st->root), 4, p);
    TG(st, g, st->root) = (int8_t)categorical(p, 4, u01(w[q & 3]));
  }
  for (int q = st->n_int - 2; q >= 0; q--) {
    const int s = st->internal[q];
    const int par = el->parent2[s];
    if (el->missing[s] || TG(st, g, s) >= NB) continue;
    if ((q >> 2) != have_grp) {
      pa_or_philox(st->seed, st->key, (uint32_t)g, (uint32_t)(q >> 2), m, ST_HIST, w);
      have_grp = q >> 2;
    }
    const double* tm = tm_for(st, s);
    const int xp = TG(st, g, par);
    for (int i = 0; i < 4; i++) v[i] = tm[i * 4 + xp] + LAM(st, g, s)[i];
    log_sample(v, 4, p);
    TG(st, g, s) = (int8_t)categorical(p, 4, u01(w[q & 3]));
  }
}

/* getEmission (bpp_c2.cpp:355-390) */
static void get_emission(chain_t* st) {
  const pa_or_element* el = st->el;
  const int S = st->S;
  double v[4];
  for (int i = 0; i < st->n_nodes - 1; i++) {
    const int s = st->nodes[i];
    const int p = el->parent2[s];
    if (el->missing[s]) continue;
    double e0 = 0, e1 = 0, e2 = 0;
    const double* tn = st->tm_null + (size_t)s * 16;
    const double* tc = st->tm_cons + (size_t)s * 16;
    const double* tu = st->tm_neut + (size_t)s * 16;
    for (int g = 0; g < st->GG; g++) {
      const int x = TG(st, g, s);
      if (x >= NB) continue;
      const int xp = TG(st, g, p);
      if (x == -1) {
        e0 += st->ambig_null[((size_t)g * S + s) * 4 + xp];
        for (int b = 0; b < 4; b++) v[b] = tc[b * 4 + xp] + LAM(st, g, s)[b];
        e1 += log_exp_sum(v, 4);
        for (int b = 0; b < 4; b++) v[b] = tu[b * 4 + xp] + LAM(st, g, s)[b];
        e2 += log_exp_sum(v, 4);
      } else {
        e0 += tn[x * 4 + xp];
        e1 += tc[x * 4 + xp];
        e2 += tu[x * 4 + xp];
      }
    }
    st->emis[s * 3 + 0] = e0; st->emis[s * 3 + 1] = e1; st->emis[s * 3 + 2] = e2;
  }
}

/* upward half of Update_Z_subtree (bpp_c.cpp:1185-1223) */
static void z_messages(chain_t* st) {
  const pa_or_element* el = st->el;
  double res[3];
  get_emission(st);
  for (int i = 0; i < st->n_nodes; i++) {
    const int s = st->nodes[i];
    for (int k = 0; k < 3; k++) st->back[s * 3 + k] = st->emis[s * 3 + k];
    if (st->fixZ[s] == 1) st->back[s * 3 + 2] = NEG_INF;
    else if (st->fixZ[s] == 2) st->back[s * 3 + 0] = NEG_INF;
  }
  for (int i = 0; i < st->n_nodes - 1; i++) {
    const int s = st->nodes[i];
    const int p = el->parent2[s];
    log_multi2(st->tm_int + (size_t)s * 9, st->back + s * 3, res);
    for (int k = 0; k < 3; k++) st->back[p * 3 + k] += res[k];
  }
  /* prior_z = log(0.5, 0.5, 0) (bpp_c.hpp:197-200) */
  st->back[st->root * 3 + 0] += log(0.5);
  st->back[st->root * 3 + 1] += log(0.5);
  st->back[st->root * 3 + 2] += NEG_INF;
}

/* Update_Z_subtree (bpp_c.cpp:1182-1293); draw for nodes[k] uses word 0 of call (k, 0, m, ST_Z); returns #changed, list in changed[] */
static int update_z(chain_t* st, uint32_t m, int* changed) {
  const pa_or_element* el = st->el;
  double p[3], v[3];
  uint32_t w[4];
  int nch = 0;
  z_messages(st);
  {
    log_sample(st->back + st->root * 3, 3, p);
    pa_or_philox(st->seed, st->key, (uint32_t)(st->n_nodes - 1), 0, m, ST_Z, w);
    const int nn = categorical(p, 3, u01(w[0]));
    const int old = st->Z[st->root];
    st->Z[st->root] = (int8_t)nn;
    if (old != nn) changed[nch++] = st->root;
  }
  for (int k = st->n_nodes - 2; k >= 0; k--) {
    const int s = st->nodes[k];
    const int par = el->parent2[s];
    const int old = st->Z[s];
    const double* t = st->tm_int + (size_t)s * 9;
    for (int i = 0; i < 3; i++) v[i] = t[i * 3 + st->Z[par]] + st->back[s * 3 + i];
    log_sample(v, 3, p);
    pa_or_philox(st->seed, st->key, (uint32_t)k, 0, m, ST_Z, w);
    st->Z[s] = (int8_t)categorical(p, 3, u01(w[0]));
    if (old != st->Z[s]) changed[nch++] = s;
  }
  return nch;
}

/* getUpdateNode(changedZ, visited) (bpp_c.cpp:620-642) */
static void update_node_changed(chain_t* st, const int* changed, int nch, uint8_t* visited) {
  const pa_or_element* el = st->el;
  for (int i = 0; i < st->N; i++) visited[i] = 1;
  for (int k = nch - 1; k >= 0; k--) {
    int j = el->parent2[changed[k]];
    while (j != st->N) {
      if (!visited[j]) break;
      visited[j] = 0;
      for (int g = 0; g < st->GG; g++) memset(LAM(st, g, j), 0, sizeof(double) * 4);
      j = el->parent2[j];
    }
  }
}

/* getUpdateNode(neut, visited) (bpp_c.cpp:836-858) */
static void update_node_class(chain_t* st, int neut, uint8_t* visited) {
  const pa_or_element* el = st->el;
  for (int i = 0; i < st->n_nodes; i++) {
    const int s = st->nodes[i];
    if (el->missing[s]) continue;
    if (!neut && (st->Z[s] == 0 || st->Z[s] == 2)) continue;
    if (neut && (st->Z[s] == 1 || st->Z[s] == 0)) continue;
    int j = el->parent2[s];
    while (j != st->N) {
      if (!visited[j]) break;
      visited[j] = 0;
      j = el->parent2[j];
    }
  }
}

/* log_f_Xz (bpp_c.cpp:860-928) */
static double log_f_xz(chain_t* st, const uint8_t* visited, double* lambda_tmp, int neut, double propos) {
  const pa_or_element* el = st->el;
  const int N = st->N;
  double transition[16], res[4];
  for (int q = 0; q < st->n_int; q++) {
    const int s = st->internal[q];
    if (visited[s]) continue;
    for (int cc = 0; cc < 2; cc++) {
      const int chi = el->children2[s * 2 + cc];
      if (el->missing[chi]) continue;
      const double* tr;
      if (!neut) {
        if (st->Z[chi] == 1) { pa_or_log_tm(el, el->distances2[chi], propos, transition); tr = transition; }
        else if (st->Z[chi] == 2) tr = st->tm_neut + (size_t)chi * 16;
        else tr = st->tm_null + (size_t)chi * 16;
      } else {
        if (st->Z[chi] == 2) { pa_or_log_tm(el, el->distances2[chi], propos, transition); tr = transition; }
        else if (st->Z[chi] == 1) tr = st->tm_cons + (size_t)chi * 16;
        else tr = st->tm_null + (size_t)chi * 16;
      }
      for (int g = 0; g < st->GG; g++)
        if (TG(st, g, chi) < NB) {
          log_multi(tr, lambda_tmp + ((size_t)g * N + chi) * 4, res);
          for (int b = 0; b < 4; b++) lambda_tmp[((size_t)g * N + s) * 4 + b] += res[b];
        }
    }
  }
  double result = 0;
  for (int g = 0; g < st->GG; g++) {
    double* lr = lambda_tmp + ((size_t)g * N + st->root) * 4;
    if (!visited[st->root])
      for (int b = 0; b < 4; b++) lr[b] += el->log_pi[b];
    result += log_exp_sum(lr, 4);
  }
  return result;
}

/* sample_rate (bpp_c.cpp:930-1176) with M = 1, adaptive = true, adaptive_factor = 0.5.
 * Schedule: MH uniform r = word 0 of call (which, 0, m, ST_RATE); gamma draws use calls (which, 1.., m, ST_RATE). */
static double sample_rate(chain_t* st, int resZ, double old_rate, int neut, const uint8_t* visited, double* loglik_old, uint32_t m,
                          double trace_c_m, double trace_n_m1) {
  const pa_or_element* el = st->el;
  const int N = st->N;
  const uint32_t which = neut ? 0u : 1u;
  double cur_r = old_rate, proposal = 0, r, MH_ratio, loglik_new;
  const double prior_a = neut ? el->nprior_a : el->cprior_a, prior_b = neut ? el->nprior_b : el->cprior_b;
  const double adaptive_factor = 0.5;

  if (visited[st->root]) {
    if (resZ != 0) {
      cur_r = pa_or_gamma(st->seed, st->key, which, 1, m, ST_RATE, prior_a, prior_b);
      if (neut) {
        if (el->ropt == 1 && cur_r < el->nlb) cur_r = old_rate;
        else if (el->ropt == 2 && cur_r < trace_c_m) cur_r = old_rate;
      } else {
        if (el->ropt == 1 && cur_r > el->cub) cur_r = old_rate;
        else if (el->ropt == 2 && cur_r > trace_n_m1) cur_r = old_rate;
      }
    } else cur_r = old_rate;
  } else {
    double* lambda_tmp = (double*)calloc((size_t)st->GG * N * 4, sizeof(double));
    for (int i = 0; i < st->n_nodes - 1; i++) {
      const int s = st->nodes[i];
      if (!visited[s]) continue;
      for (int g = 0; g < st->GG; g++) memcpy(lambda_tmp + ((size_t)g * N + s) * 4, LAM(st, g, s), sizeof(double) * 4);
    }
    do { /* mm loop with M = 1; 'continue' in the reference == leave */
      uint32_t w[4];
      pa_or_philox(st->seed, st->key, which, 0, m, ST_RATE, w);
      r = u01(w[0]);
      if (neut) {
        proposal = pa_or_gamma(st->seed, st->key, which, 1, m, ST_RATE, cur_r / st->prop_n, st->prop_n);
        if (el->ropt == 1 && proposal < el->nlb) break;
        else if (el->ropt == 2 && proposal < trace_c_m) break;
      } else {
        const double u = r * (st->prop_c - 1 / st->prop_c) + 1 / st->prop_c;
        proposal = cur_r * u;
        if (el->ropt == 1 && proposal > el->cub) break;
        else if (el->ropt == 2 && proposal > trace_n_m1) break;
      }
      if (proposal < 1e-12 || proposal > 1e8) break;
      loglik_new = log_f_xz(st, visited, lambda_tmp, neut, proposal);
      if (neut)
        MH_ratio = loglik_new - *loglik_old + log(pa_or_gamma_pdf(proposal, prior_a, prior_b)) - log(pa_or_gamma_pdf(cur_r, prior_a, prior_b)) +
                   log(pa_or_gamma_pdf(cur_r, proposal / st->prop_n, st->prop_n)) - log(pa_or_gamma_pdf(proposal, cur_r / st->prop_n, st->prop_n));
      else
        MH_ratio = loglik_new - *loglik_old + log(pa_or_gamma_pdf(proposal, prior_a, prior_b)) - log(pa_or_gamma_pdf(cur_r, prior_a, prior_b)) +
                   log(cur_r) - log(proposal);
      if (log(r) < MH_ratio) {
        cur_r = proposal;
        *loglik_old = loglik_new;
        for (int q = 0; q < st->n_int; q++) {
          const int s = st->internal[q];
          if (visited[s]) continue;
          for (int g = 0; g < st->GG; g++) memcpy(LAM(st, g, s), lambda_tmp + ((size_t)g * N + s) * 4, sizeof(double) * 4);
        }
      }
    } while (0);
    free(lambda_tmp);
  }
  if (old_rate != cur_r) {
    for (int i = 0; i < st->n_nodes - 1; i++) {
      const int s = st->nodes[i];
      if (el->distances2[s] < 1e-10) continue;
      pa_or_log_tm(el, el->distances2[s], cur_r, (neut ? st->tm_neut : st->tm_cons) + (size_t)s * 16);
    }
    if (neut) st->accept_n += 1; else st->accept_c += 1;
  }
  if ((int)(m + 1) % el->adaptive_freq == 0) {
    const int k = (int)(m + 1) / el->adaptive_freq;
    const double step = fmin(adaptive_factor, 1 / sqrt((double)k));
    double scale_adj;
    if (neut) {
      if ((double)st->accept_n / el->adaptive_freq > 0.44) scale_adj = exp(step);
      else if ((double)st->accept_n / el->adaptive_freq < 0.23) scale_adj = exp(-step);
      else scale_adj = 1;
      st->prop_n = st->prop_n * scale_adj;
      if (st->prop_n > 1) st->prop_n = 1;
      st->accept_n = 0;
    } else {
      if ((double)st->accept_c / el->adaptive_freq > 0.44) scale_adj = exp(step);
      else if ((double)st->accept_c / el->adaptive_freq < 0.23) scale_adj = exp(-step);
      else scale_adj = 1;
      st->prop_c = pow(st->prop_c, scale_adj);
      if (st->prop_c < 1.0001) st->prop_c = 1.0001;
      st->accept_c = 0;
    }
  }
  return cur_r;
}

static void set_tm_int(chain_t* st, double gr, double lr, double lr2) { /* bpp_c.cpp:573-599 */
  for (int i = 0; i < st->n_nodes; i++) {
    const int s = st->nodes[i];
    double* t = st->tm_int + (size_t)s * 9;
    if (st->fixZ[s] == 1) {
      t[1 * 3 + 1] = 0; t[2 * 3 + 1] = NEG_INF;
      t[0 * 3 + 0] = log(1 - gr); t[1 * 3 + 0] = log(gr); t[2 * 3 + 0] = NEG_INF;
    } else {
      t[0 * 3 + 0] = log(1 - gr); t[1 * 3 + 0] = log(gr); t[2 * 3 + 0] = NEG_INF;
      const double y = 1 - lr;
      t[1 * 3 + 1] = log(y); t[2 * 3 + 1] = log(1 - y);
      t[1 * 3 + 2] = log(lr2); t[2 * 3 + 2] = log(1 - lr2);
    }
  }
}

static void count_transitions(const pa_or_element* el, const int32_t* nodes, int n_nodes, const int8_t* Z, const int8_t* fixZ, double nZ[9], double nZfree[9]) {
  memset(nZ, 0, sizeof(double) * 9);
  for (int i = 0; i < n_nodes - 1; i++) {
    const int s = nodes[i];
    nZ[Z[s] * 3 + Z[el->parent2[s]]] += 1;
  }
  memcpy(nZfree, nZ, sizeof(double) * 9);
  for (int i = 0; i < n_nodes - 1; i++) {
    const int s = nodes[i];
    if (fixZ[s] == 1) nZfree[Z[s] * 3 + Z[el->parent2[s]]] -= 1;
  }
}

/* sample_transition (bpp_c.cpp:522-613) */
static void sample_transition(chain_t* st, uint32_t m, double* gr, double* lr, double* lr2) {
  const pa_or_element* el = st->el;
  double nZ[9], nZf[9];
  count_transitions(el, st->nodes, st->n_nodes, st->Z, st->fixZ, nZ, nZf);
  *gr = sched_beta(st->seed, st->key, 0, m, el->prior_g_a + nZ[1 * 3 + 0] + nZ[2 * 3 + 0], el->prior_g_b + nZ[0]);
  *lr = sched_beta(st->seed, st->key, 1, m, el->prior_l_a + nZf[2 * 3 + 1], el->prior_l_b + nZf[1 * 3 + 1]);
  if (el->prior_l2_a == 0) *lr2 = 0;
  else *lr2 = sched_beta(st->seed, st->key, 2, m, el->prior_l2_a + nZf[1 * 3 + 2], el->prior_l2_b + nZf[2 * 3 + 2]);
  set_tm_int(st, *gr, *lr, *lr2);
}

/* ------------------------------------------------------------------ public: replay */
double pa_or_loglik(const pa_or_element* el, int resZ, const int8_t* Z, double n_rate, double c_rate, double* persite) {
  chain_t* st = chain_new(el);
  init_mcmc(st, resZ, Z, n_rate, c_rate, el->init_grate, el->init_lrate, el->init_lrate2);
  uint8_t* visited = (uint8_t*)calloc(st->N, 1);
  double tot = 0;
  for (int g = 0; g < st->GG; g++) {
    update_tg_up(st, g, visited);
    const double v = log_exp_sum(LAM(st, g, st->root), 4);
    if (persite) persite[g] = v;
    tot += v;
  }
  free(visited);
  chain_free(st);
  return tot;
}

void pa_or_emission_back(const pa_or_element* el, int resZ, const int8_t* Z, double n_rate, double c_rate, double g_rate, double l_rate,
                         double l2_rate, const int8_t* tg_hist, double* emission, double* back) {
  chain_t* st = chain_new(el);
  init_mcmc(st, resZ, Z, n_rate, c_rate, g_rate, l_rate, l2_rate);
  uint8_t* visited = (uint8_t*)calloc(st->N, 1);
  for (int g = 0; g < st->GG; g++) update_tg_up(st, g, visited);
  for (int g = 0; g < st->GG; g++)
    for (int s = st->S; s < st->N; s++)
      if (TG(st, g, s) < NB && tg_hist[(size_t)g * st->N + s] >= 0) TG(st, g, s) = tg_hist[(size_t)g * st->N + s];
  set_tm_int(st, g_rate, l_rate, l2_rate);
  z_messages(st);
  memcpy(emission, st->emis, sizeof(double) * st->N * 3);
  memcpy(back, st->back, sizeof(double) * st->N * 3);
  free(visited);
  chain_free(st);
}

/* ------------------------------------------------------------------ public: chain (Gibbs, bpp_c.cpp:346-519) */
int pa_or_chain(const pa_or_element* el, int resZ, uint32_t seed, uint32_t chain_key, pa_or_trace* tr) {
  chain_t* st = chain_new(el);
  const int N = st->N, T = el->num_burn + el->num_mcmc;
  st->seed = seed; st->key = chain_key;
  st->accept_n = tr->accept_n; st->accept_c = tr->accept_c;
  init_mcmc(st, resZ, NULL, el->init_nrate, el->init_crate, el->init_grate, el->init_lrate, el->init_lrate2);
  tr->n_rate[0] = el->init_nrate; tr->c_rate[0] = el->init_crate;
  tr->l_rate[0] = el->init_lrate; tr->l2_rate[0] = el->init_lrate2; tr->g_rate[0] = el->init_grate;
  uint8_t* visited = (uint8_t*)calloc(N, 1);
  uint8_t* vis_n = (uint8_t*)malloc(N);
  uint8_t* vis_c = (uint8_t*)malloc(N);
  int* changed = (int*)malloc(sizeof(int) * N);
  for (int g = 0; g < st->GG; g++) update_tg_up(st, g, visited);
  double log_lik_old = 0;
  double max_loglik = NEG_INF;
  int max_m = -1;
  int m;
  for (m = 0; m < T; m++) {
    /* MonitorChain (bpp_c.cpp:741-834) */
    memcpy(tr->Z + (size_t)m * N, st->Z, N);
    if (m == 0) {
      double s = 0;
      for (int g = 0; g < st->GG; g++) s += log_exp_sum(LAM(st, g, st->root), 4);
      tr->loglik[0] = s;
      log_lik_old = s;
    } else tr->loglik[m] = log_lik_old;
    tr->full_loglik[m] = tr->loglik[m] + 0.0;
    tr->full_loglik[m] += log(pa_or_gamma_pdf(tr->c_rate[m], el->cprior_a, el->cprior_b));
    if (resZ != 0) tr->full_loglik[m] += log(pa_or_gamma_pdf(tr->n_rate[m], el->nprior_a, el->nprior_b));
    if (m >= el->num_burn && tr->full_loglik[m] > max_loglik) {
      max_loglik = tr->full_loglik[m];
      memcpy(tr->max_Z, st->Z, N);
      max_m = m;
    }
    if (st->failure || m == T - 1) break;

    for (int g = 0; g < st->GG; g++) update_tg_sample(st, g, (uint32_t)m);
    const int nch = update_z(st, (uint32_t)m, changed);
    update_node_changed(st, changed, nch, visited);
    if (nch > 0) {
      log_lik_old = 0;
      for (int g = 0; g < st->GG; g++) {
        update_tg_up(st, g, visited);
        log_lik_old += log_exp_sum(LAM(st, g, st->root), 4);
      }
    }
    memset(vis_n, 1, N);
    memset(vis_c, 1, N);
    update_node_class(st, 1, vis_n);
    update_node_class(st, 0, vis_c);
    tr->n_rate[m + 1] = sample_rate(st, resZ, tr->n_rate[m], 1, vis_n, &log_lik_old, (uint32_t)m, tr->c_rate[m], 0.0);
    tr->c_rate[m + 1] = sample_rate(st, resZ, tr->c_rate[m], 0, vis_c, &log_lik_old, (uint32_t)m, tr->c_rate[m], tr->n_rate[m + 1]);
    sample_transition(st, (uint32_t)m, &tr->g_rate[m + 1], &tr->l_rate[m + 1], &tr->l2_rate[m + 1]);
  }
  tr->max_loglik = max_loglik;
  tr->max_m = max_m;
  tr->accept_n = st->accept_n; tr->accept_c = st->accept_c;
  tr->failure = st->failure;
  free(visited); free(vis_n); free(vis_c); free(changed);
  chain_free(st);
  return 0;
}

/* ------------------------------------------------------------------ public: Eval2 (bpp_c2.cpp:14-353) */
static double prior_z_config(const pa_or_element* el, const int8_t* cfgZ, const int8_t* fixZ) {
  double nZ[9], nZf[9];
  count_transitions(el, el->nodes, el->n_nodes, cfgZ, fixZ, nZ, nZf);
  double priorZ = pa_or_lnbeta(el->prior_g_a + nZ[1 * 3 + 0] + nZ[2 * 3 + 0], el->prior_g_b + nZ[0]) - pa_or_lnbeta(el->prior_g_a, el->prior_g_b);
  priorZ += pa_or_lnbeta(el->prior_l_a + nZf[2 * 3 + 1], el->prior_l_b + nZf[1 * 3 + 1]) - pa_or_lnbeta(el->prior_l_a, el->prior_l_b);
  if (el->prior_l2_a > 0)
    priorZ += pa_or_lnbeta(el->prior_l2_a + nZf[1 * 3 + 2], el->prior_l2_b + nZf[2 * 3 + 2]) - pa_or_lnbeta(el->prior_l2_a, el->prior_l2_b);
  const double prior_z[3] = {log(0.5), log(0.5), NEG_INF};
  priorZ += prior_z[cfgZ[el->root]];
  return priorZ;
}

double pa_or_eval2(const pa_or_element* el, int resZ, const pa_or_trace* tr, int32_t* failure) {
  const int N = el->N, num_burn = el->num_burn, num_mcmc = el->num_mcmc, T = num_burn + num_mcmc;
  const int8_t* Max_Z = tr->max_Z;
  /* fixZ of this model (initMCMC) */
  int8_t* fixZ = (int8_t*)malloc(N);
  for (int s = 0; s < N; s++) fixZ[s] = -1;
  for (int i = 0; i < el->n_upper_c; i++) fixZ[el->upper_c[i]] = 1;
  fixZ[el->root] = 1;
  if (resZ == 2) for (int i = 0; i < el->n_upper_conserve_c; i++) fixZ[el->upper_conserve_c[i]] = 1;
  else if (resZ == 0) for (int i = 0; i < el->n_nodes - 1; i++) fixZ[el->nodes[i]] = 1;

  int* uncertain = (int*)malloc(sizeof(int) * N);
  int n_unc = 0;
  for (int s = 0; s < N; s++) {
    int cnt = 0;
    for (int i = num_burn; i < T; i++)
      if (tr->Z[(size_t)i * N + s] == Max_Z[s]) cnt++;
    if (cnt < num_mcmc) uncertain[n_unc++] = s;
  }
  int cap = 16, numConf = 1;
  const int8_t** configZ = (const int8_t**)malloc(sizeof(int8_t*) * cap);
  int* num = (int*)calloc(cap, sizeof(int));
  double* var_c = (double*)calloc(cap, sizeof(double));
  double* mean_c = (double*)calloc(cap, sizeof(double));
  double* var_n = (double*)calloc(cap, sizeof(double));
  double* mean_n = (double*)calloc(cap, sizeof(double));
  double* cov = (double*)calloc(cap, sizeof(double));
  double* sum_full = (double*)calloc(cap, sizeof(double));
  configZ[0] = Max_Z;
  for (int i = T - 1; i >= num_burn; i--) {
    const int8_t* zi = tr->Z + (size_t)i * N;
    int j = 0;
    for (; j < numConf; j++) {
      int k = 0;
      for (; k < n_unc; k++)
        if (zi[uncertain[k]] != configZ[j][uncertain[k]]) break;
      if (k == n_unc) break;
    }
    if (j == numConf) {
      if (numConf == cap) {
        cap *= 2;
        configZ = (const int8_t**)realloc(configZ, sizeof(int8_t*) * cap);
#define GROW(p, T_) p = (T_*)realloc(p, sizeof(T_) * cap); memset(p + cap / 2, 0, sizeof(T_) * (cap / 2))
        GROW(num, int); GROW(var_c, double); GROW(mean_c, double); GROW(var_n, double); GROW(mean_n, double); GROW(cov, double); GROW(sum_full, double);
#undef GROW
      }
      configZ[numConf] = zi;
      numConf++;
    }
    num[j]++;
    var_c[j] += pow(tr->c_rate[i], 2);
    mean_c[j] += tr->c_rate[i];
    if (resZ != 0) {
      var_n[j] += pow(tr->n_rate[i], 2);
      mean_n[j] += tr->n_rate[i];
      cov[j] += tr->c_rate[i] * tr->n_rate[i];
    }
    sum_full[j] += tr->full_loglik[i] / num_mcmc;
  }
  double VB = 0, logdet;
  int effective_sum = 0;
  for (int i = 0; i < numConf; i++) {
    const double w = (double)num[i] / num_mcmc;
    if (w < 0.02) continue;
    const double priorZ = prior_z_config(el, configZ[i], fixZ);
    sum_full[i] += num[i] * priorZ / num_mcmc;
    if (resZ == 0) {
      logdet = log(var_c[i] - pow(mean_c[i], 2) / num[i]) - log(num[i]);
    } else {
      const double vc = var_c[i] - pow(mean_c[i], 2) / num[i];
      const double vn = var_n[i] - pow(mean_n[i], 2) / num[i];
      const double cv = cov[i] - mean_n[i] * mean_c[i] / num[i];
      logdet = log(vc * vn - cv * cv) - 2 * log(num[i]);
    }
    if (isnan(logdet)) continue;
    effective_sum += num[i];
    VB -= w * log(num[i]);
    VB += sum_full[i];
    VB += 0.5 * logdet * num[i] / num_mcmc;
  }
  if (effective_sum == 0) {
    double varn = 0, varc = 0, cv = 0, meann = 0, meanc = 0;
    for (int i = 0; i < numConf; i++) {
      varc += var_c[i]; meanc += mean_c[i];
      if (resZ != 0) { meann += mean_n[i]; varn += var_n[i]; cv += cov[i]; }
    }
    const double rn = meann / num_mcmc, rc = meanc / num_mcmc;
    /* P(X | Z_i, r*) by pruning at the posterior-mean rates (bpp_c2.cpp:220-292); note: no t<=1e-10 special case here */
    chain_t* st = chain_new(el);
    init_mcmc(st, resZ, Max_Z, el->init_nrate, el->init_crate, el->init_grate, el->init_lrate, el->init_lrate2);
    /* the reference keeps the TM caches of the finished chain; they are overwritten for every branch below,
       except neut under M0 which is never read for M0 (no Z==2).  */
    for (int k = 0; k < st->n_nodes - 1; k++) {
      const int s = st->nodes[k];
      pa_or_log_tm(el, el->distances2[s], rc, st->tm_cons + (size_t)s * 16);
      if (resZ != 0) pa_or_log_tm(el, el->distances2[s], rn, st->tm_neut + (size_t)s * 16);
    }
    uint8_t* visited = (uint8_t*)calloc(N, 1);
    for (int i = 0; i < numConf; i++) {
      const double priorZ = prior_z_config(el, configZ[i], fixZ);
      memcpy(st->Z, configZ[i], N);
      for (int s = el->S; s < N; s++)
        for (int g = 0; g < st->GG; g++) memset(LAM(st, g, s), 0, sizeof(double) * 4);
      double s_ = 0;
      for (int g = 0; g < st->GG; g++) {
        update_tg_up(st, g, visited);
        s_ += log_exp_sum(LAM(st, g, st->root), 4);
      }
      sum_full[i] = s_ + priorZ;
    }
    free(visited);
    chain_free(st);
    VB = log_exp_sum(sum_full, numConf);
    VB += log(pa_or_gamma_pdf(rc, el->cprior_a, el->cprior_b));
    if (resZ != 0) VB += log(pa_or_gamma_pdf(rn, el->nprior_a, el->nprior_b));
    if (resZ == 0) {
      varc -= pow(meanc, 2) / num_mcmc;
      logdet = log(varc) - log(num_mcmc);
      VB += 0.5 * logdet + 0.5 * log(2 * M_PI);
    } else {
      varn -= pow(meann, 2) / num_mcmc;
      varc -= pow(meanc, 2) / num_mcmc;
      cv -= meann * meanc / num_mcmc;
      logdet = log(varc * varn - cv * cv) - 2 * log(num_mcmc);
      VB += 0.5 * logdet + log(2 * M_PI);
    }
  } else {
    if (resZ != 0) VB = VB * ((double)num_mcmc / effective_sum) + log(num_mcmc) + log(2 * M_PI) + 1;
    else VB = VB * ((double)num_mcmc / effective_sum) + log(num_mcmc) + 0.5 * (log(2 * M_PI) + 1);
  }
  if (failure) *failure = isnan(VB) ? 1 : 0;
  free(fixZ); free(uncertain); free(configZ); free(num); free(var_c); free(mean_c); free(var_n); free(mean_n); free(cov); free(sum_full);
  return VB;
}
