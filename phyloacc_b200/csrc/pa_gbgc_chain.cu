// pa_gbgc_chain.cu -- the PhyloAcc-ST-GBGC chain on the device: BPP_C::Gibbs of src/PhyloAcc-ST-GBGC/SRC/bpp_c.cpp:259-345.
//
// One CTA (160 threads) per (element, model) chain, persistent grid pulling chains from an atomic queue.  Unlike PhyloAcc-ST the GC-biased
// sampler draws no substitution histories: an iteration works on the EXPECTED (child base, parent base) pairs of every branch
//   1. MonitorChain                                                                      bpp_c.cpp:598-652
//   2. expected pairs at the current state (downward pass over compressed site patterns)  Update_Tg(g,.,.,true)   :530-583
//   3. six (Z, B) emissions per branch + 6-state upward messages + the ELBO term          initial_prob_back       bpp_c2.cpp:354-407
//   4. three joint MH moves on (gBGC bias, accelerated selection, conserved selection)    sample_rates            bpp_c.cpp:881-958
//      with (Z, B) integrated out: five new rate matrices Q o f(s + b B), P(t) of every
//      branch under each (uniformisation + squaring, pa_gbgc.cuh), emissions, messages    integrate_ZB            :1765-1795
//   5. top-down draws of (Z, B), the exact likelihood at the drawn state (upward pass),
//      ELBO at the new state, one MH decision for the whole iteration                     update_rates, Update_ZB :1079-1131, 1667-1763
// Random numbers: Philox4x32-10 keyed by (seed, element number * 4 + model), counter (index, sub, iteration, PA_ST_GBGC); sub 0 = the
// three rate moves (index = move; words = bias / accelerated / conserved proposal, acceptance), sub 1 = the (Z, B) draw of node `index`,
// sub 2 = the iteration's final acceptance.  oracle/gbgc.py (GbgcChain) restates the chain on the same schedule.
// Layout: P(t) logs of the five rate-dependent classes live in three rotating per-CTA buffers in global memory (committed / working /
// proposal: accepting a move is a pointer swap), class 0 (background, no gBGC) is computed once per chain; partials [node][pattern][4] in
// the CTA's scratch; everything else in shared memory.  fp64 throughout.
#include <cuda_runtime.h>

#include <algorithm>
#include <string>
#include <vector>

#include "pa_device.cuh"
#include "pa_gbgc.cuh"
#include "pa_internal.h"

namespace pa {

namespace {

#ifndef PA_GBGC_BD
#define PA_GBGC_BD 160
#endif
// threads per CTA (a multiple of 32, at least 96); A/B knob together with PA_GBGC_CTAS (CTAs per SM).  160 x 3 at 128 registers: elements of
// up to 160 site patterns (most 200-bp elements) take one pattern tile per pass; measured 128 x 3: 422.6 k, 160 x 3: 432.3 k, 192 x 3 and
// 256 x 2 (before the last instruction cuts): +1 % / -3 % against 128 x 3
constexpr int kCBD = PA_GBGC_BD;
constexpr int kNW = kCBD / 32;
// phase clock (thread 0): adds the cycles since the previous mark to phase k
#define PH(k) do { if (tid == 0) { const long long t__ = clock64(); s_ph[k] += t__ - s_t0; s_t0 = t__; } } while (0)
enum { PH_SETUP = 0, PH_MONITOR, PH_DOWN, PH_INIT_EMIS, PH_INIT_MSG, PH_BUILD, PH_EMIS, PH_MSG, PH_DRAW, PH_UP, PH_ELBO, PH_COMMIT };
#define PA_ST_GBGC 4u

struct GbgcChainLaunch {
  const uint8_t* blobs; const int64_t* blob_off;
  int n_chains, N, n_max, slots_max, P_max, T;
  int codes_cap;              // bytes of shared memory for a staged copy of an element's packed tip codes (0: always read them from the blob)
  const int32_t* chain_elem;  // element index in the engine's batch
  const int32_t* chain_resz;  // the reference's resZ: 0 = M0, 2 = M1, 1 = M2
  const int8_t* fixZ;         // [n_chains][N] original numbering (1: class 2 forbidden)
  const int8_t* upper;        // [n_chains][N] outgroup lineage: conserved stays conserved
  pa_gbgc_params prm;
  double pi[4];
  double Q0[16];
  unsigned char* scratch; size_t stride;  // per resident CTA
  int* queue;
  double* tr;       // [n_chains][5][T]: log P(X | Z, B, r), log P(X, Z, B, r), conserved / accelerated selection, gBGC bias
  uint8_t* tr_zb;   // [n_chains][T][n_max] local node order: Z | B << 2
  double* head;     // [n_chains][8]: log P(constraints), accepted rate sweeps, accepted iterations
  unsigned long long* counters;  // [16]: node-pattern updates of the upward passes, (branch, class) P(t) builds, chain iterations, branch-pattern updates of the
                                 // downward passes, then SM cycles per phase (thread 0 of every CTA)
};

constexpr int kPow = 19;
constexpr int kPwStride = (kPow + 1) * 16 + 2;  // doubles between the power tables of two classes: the lanes of a warp read up to six classes at the same
                                                 // (power, entry) -- a stride of 320 would put them all in one bank
__constant__ double c_inv[kPow + 1] = {0.0, 1.0, 1.0 / 2, 1.0 / 3, 1.0 / 4, 1.0 / 5, 1.0 / 6, 1.0 / 7, 1.0 / 8, 1.0 / 9, 1.0 / 10, 1.0 / 11, 1.0 / 12, 1.0 / 13,
                                       1.0 / 14, 1.0 / 15, 1.0 / 16, 1.0 / 17, 1.0 / 18, 1.0 / 19};

// doubles of the phase-shared region: pruning stack [slots][4][BD] + its scale counts | transposition tiles of the downward pass [warps][16][33] |
// power basis of the P(t) builds [6][20][16]
__host__ __device__ inline size_t uni_doubles(int n_max, int slots_max) {
  const size_t a = (size_t)slots_max * 4 * kCBD + ((size_t)slots_max * kCBD + 1) / 2, b = kNW * 16 * 33, c = 6 * ((kPow + 1) * 16 + 2);
  return a > b ? (a > c ? a : c) : (b > c ? b : c);
}

struct Sh {  // shared-memory views of one chain
  double *Pm, *uni, *red, *emis, *emis_t, *back, *Qc, *scal, *msg, *lsum, *pw, *mu;
  int *stack_sc, *msc;
  uint4* up_ch;   // upward program: per internal node (Sethi-Ullman order) both children resolved: {node | kind << 16, tip row or stack slot} x 2
  uint2* up_self; // ... and the node itself: {node | stack slot << 16, arena index}
  int n_upprog;
  int16_t *child, *tiprow, *su, *intidx, *parent, *up_nodes, *up_off, *down_nodes, *down_off, *dn_node, *dn_src, *dn_par;
  uint8_t *flags, *fix, *upr, *codes_s;
  int8_t *slot, *Z, *B, *Zn, *Bn;
};

enum { SC_LG0 = 0, SC_LG1, SC_LL0, SC_LL1, SC_SELC, SC_SELN, SC_GBC, SC_PSELC, SC_PSELN, SC_PGBC, SC_LOGLIK, SC_LOGLIK_OLD, SC_ELBO_OLD, SC_LOGPZ,
       SC_LOGPZC, SC_FULL, SC_LL_PROP, SC_G, SC_L, SC_KC, SC_KN, SC_KG, SC_LB0, SC_LB1, SC_LC, SC_LN, SC_LCM, SC_LNM, SC_N };

// log TM_Int (child class a, parent class j) of a branch (GBGC/SRC/bpp.hpp:284-316): (0,0) = 1-g, (1,0) = g, (1,1) = 1-l, (2,1) = l, (2,2) = 1;
// on the outgroup lineage (1,1) = 1 and (2,1) = 0
__device__ __forceinline__ double lt_val(const double* scal, int a, int j, bool up) {
  if (j == 0) return a == 0 ? scal[SC_LG0] : (a == 1 ? scal[SC_LG1] : -CUDART_INF);
  if (j == 1) return a == 1 ? (up ? 0.0 : scal[SC_LL0]) : (a == 2 ? (up ? -CUDART_INF : scal[SC_LL1]) : -CUDART_INF);
  return a == 2 ? 0.0 : -CUDART_INF;
}

// message of child c to its parent's class j: log sum_k exp(lt[k % 3][j] + back[c][k]) with ONE maximum over the whole 6 x 3 table
// (BPP::log_multi, GBGC/SRC/bpp.hpp)
__device__ __forceinline__ double child_msg(const Sh& s, int c, int j) {
  const bool up = s.upr[c] != 0;
  double mx = -CUDART_INF;
  for (int k = 0; k < 6; k++) {
    const double b = s.back[c * 6 + k];
#pragma unroll
    for (int jj = 0; jj < 3; jj++) mx = fmax(mx, lt_val(s.scal, k % 3, jj, up) + b);
  }
  if (!(mx > -CUDART_INF)) return -CUDART_INF;
  double sum = 0.0;
  for (int k = 0; k < 6; k++) sum += exp(lt_val(s.scal, k % 3, j, up) + s.back[c * 6 + k] - mx);
  return log(sum) + mx;
}

// 6-state upward messages by height levels; src = emissions [n][6] (the root's row is taken as zero); result in s.back
__device__ __noinline__ void messages(const Sh& s, const double* src, int n, int n_up, bool init_pass, int tid) {
  for (int t = tid; t < n * 6; t += kCBD) s.back[t] = (t >= (n - 1) * 6) ? 0.0 : src[t];
  __syncthreads();
  for (int h = 1; h < n_up; h++) {
    const int lo = s.up_off[h], cnt = s.up_off[h + 1] - lo;
    for (int t = tid; t < cnt * 3; t += kCBD) {
      const int p = s.up_nodes[lo + t / 3], j = t % 3;
      const int c0 = s.child[2 * p], c1 = s.child[2 * p + 1];
      if (c0 < 0) continue;
      double m0 = child_msg(s, c0, j), m1 = child_msg(s, c1, j);
      if (init_pass) { m0 -= 0.69314718055994530942; m1 -= 0.69314718055994530942; }
      s.back[p * 6 + j] = (s.back[p * 6 + j] + m0) + m1;
      s.back[p * 6 + 3 + j] = (s.back[p * 6 + 3 + j] + m0) + m1;
    }
    __syncthreads();
  }
}

// linear T of a branch: j = parent class
__device__ __forceinline__ double lt_lin(const double* scal, int a, int j, bool up) {
  if (j == 0) return a == 0 ? 1.0 - scal[SC_G] : (a == 1 ? scal[SC_G] : 0.0);
  if (j == 1) return a == 1 ? (up ? 1.0 : 1.0 - scal[SC_L]) : (a == 2 ? (up ? 0.0 : scal[SC_L]) : 0.0);
  return a == 2 ? 1.0 : 0.0;
}

// The same 6-state upward pass in LINEAR space: per node the six class weights exp(emission - max) times the children's messages,
// renormalised by a power of two; the log of the total travels as (sum of the maxima) + (sum of the exponents) ln 2.  No transcendental
// inside the level loop (6 exp per node up front, in parallel; one log at the root), the levels run on one warp without block barriers.
// s.back receives the LINEAR weights w[node][6] (what the top-down draws need), *ll = log_prob_back[root][0].
// Returns false when a finite emission underflows against its node's maximum (> 700 nats apart) or the root's class 0 vanishes: the
// caller then repeats the pass in log space (messages()).
__device__ bool messages_lin(const Sh& s, const double* src, int n, int n_up, int tid, double* ll) {
  bool bad = false;
  for (int i = tid; i < n; i += kCBD) {
    double e[6], mE = -CUDART_INF;
    const bool is_root = i == n - 1;
#pragma unroll
    for (int k = 0; k < 6; k++) { e[k] = is_root ? 0.0 : src[i * 6 + k]; mE = fmax(mE, e[k]); }
#pragma unroll
    for (int k = 0; k < 6; k++) {
      const double d = e[k] - mE;
      if (d < -700.0 && e[k] > -CUDART_INF) bad = true;
      e[k] = exp(d);
      s.back[i * 6 + k] = e[k];
    }
    if (!(mE > -CUDART_INF)) bad = true;
    s.lsum[i] = mE;
    s.msc[i] = 0;
    if (s.child[2 * i] < 0) {  // tips: their messages are final
      const bool up = s.upr[i] != 0;
      const double a0 = e[0] + e[3], a1 = e[1] + e[4], a2 = e[2] + e[5];
      s.msg[i * 3] = lt_lin(s.scal, 0, 0, up) * a0 + lt_lin(s.scal, 1, 0, up) * a1;
      s.msg[i * 3 + 1] = lt_lin(s.scal, 1, 1, up) * a1 + lt_lin(s.scal, 2, 1, up) * a2;
      s.msg[i * 3 + 2] = a2;
    }
  }
  if (__syncthreads_or(bad ? 1 : 0)) return false;
  int bad_w = 0;
  if (tid < 32) {
    for (int h = 1; h < n_up; h++) {
      const int lo = s.up_off[h], cnt = s.up_off[h + 1] - lo;
      for (int t = tid; t < cnt; t += 32) {
        const int i = s.up_nodes[lo + t];
        const int c0 = s.child[2 * i], c1 = s.child[2 * i + 1];
        if (c0 < 0) continue;
        double w[6], mx = 0.0;
#pragma unroll
        for (int k = 0; k < 6; k++) {
          w[k] = (s.back[i * 6 + k] * s.msg[c0 * 3 + k % 3]) * s.msg[c1 * 3 + k % 3];
          mx = fmax(mx, w[k]);
        }
        int ex = 0;
        if (mx > 0.0) {
          if (mx < 0x1p-200) {  // renormalise by an exact power of two, and only when the weights have drifted far down
            ex = ((__double2hiint(mx) >> 20) & 0x7ff) - 1023;
            const double up2 = __hiloint2double((1023 - ex) << 20, 0);  // 2^-ex
#pragma unroll
            for (int k = 0; k < 6; k++) w[k] *= up2;
          }
        } else {
          bad_w = 1;
        }
#pragma unroll
        for (int k = 0; k < 6; k++) s.back[i * 6 + k] = w[k];
        s.msc[i] = (ex + s.msc[c0]) + s.msc[c1];
        s.lsum[i] = (s.lsum[i] + s.lsum[c0]) + s.lsum[c1];
        const bool up = s.upr[i] != 0;
        const double a0 = w[0] + w[3], a1 = w[1] + w[4], a2 = w[2] + w[5];
        s.msg[i * 3] = lt_lin(s.scal, 0, 0, up) * a0 + lt_lin(s.scal, 1, 0, up) * a1;
        s.msg[i * 3 + 1] = lt_lin(s.scal, 1, 1, up) * a1 + lt_lin(s.scal, 2, 1, up) * a2;
        s.msg[i * 3 + 2] = a2;
      }
      __syncwarp();
    }
    if (tid == 0) {
      const int r = n - 1;
      if (!(s.back[r * 6] > 0.0)) bad_w = 1;
      else *ll = (log(s.back[r * 6]) + (double)s.msc[r] * 0.69314718055994530942) + s.lsum[r];
    }
  }
  return !__syncthreads_or(bad_w);
}

// log P(X | Z, B, r): the upward pass at the matrices in s.Pm over the element's site patterns (thread = pattern, partials on a shared-memory
// stack, every internal node's partial also written to the arena for the downward pass).  Same arithmetic, in the same order, as
// prune_pattern<false, true> (pa_device.cuh), but driven by a per-chain RESOLVED program (s.up_ch / s.up_self, built once per chain): one
// broadcast 16-byte load tells a step what both children are and where their data lies, instead of a chain of dependent look-ups
// (order -> children -> flags -> tip row / slot).
__device__ double upward_pass(const Sh& s, const ElemHeader* h, const uint8_t* codes, const uint16_t* weights, double* gpart, int16_t* gsc,
                              const double* pi, int tid) {
  const int P = h->P, Ppad = h->Ppad, half = Ppad >> 1, root = h->n - 1, np = s.n_upprog;
  double* stack = s.uni;
  double acc = 0.0;
  for (int p = tid; p < Ppad; p += kCBD) {
    if (p >= P) continue;
    double root_l[4] = {1.0, 1.0, 1.0, 1.0};
    int root_sc = 0;
    for (int k = 0; k < np; k++) {
      const uint4 d = s.up_ch[k];
      const uint2 me = s.up_self[k];
      double out[4] = {1.0, 1.0, 1.0, 1.0};
      int sc = 0;
#pragma unroll
      for (int cc = 0; cc < 2; cc++) {
        const uint32_t cx = cc ? d.z : d.x, cy = cc ? d.w : d.y;
        const uint32_t kind = cx >> 16;
        if (kind == 0) continue;  // missing child
        const double* Pc = s.Pm + (cx & 0xffffu) * 16;
        if (kind == 1) {
          const uint32_t byte = codes[(size_t)cy * half + (p >> 1)];
          const uint32_t code = (byte >> ((p & 1) * 4)) & 15u;
          if (code == 15u) continue;  // gap / N
          if ((code & (code - 1)) == 0) {
            const int bb = __ffs(code) - 1;
#pragma unroll
            for (int j = 0; j < 4; j++) out[j] *= Pc[j * 4 + bb];
          } else {
#pragma unroll
            for (int j = 0; j < 4; j++) {
              double m = 0;
#pragma unroll
              for (int bb = 0; bb < 4; bb++) m += ((code >> bb) & 1u) ? Pc[j * 4 + bb] : 0.0;
              out[j] *= m;
            }
          }
        } else {
          double Lc[4];
#pragma unroll
          for (int i = 0; i < 4; i++) Lc[i] = stack[(cy * 4 + i) * kCBD + tid];
          sc += s.stack_sc[cy * kCBD + tid];
#pragma unroll
          for (int j = 0; j < 4; j++) out[j] *= ((Pc[j * 4 + 0] * Lc[0] + Pc[j * 4 + 1] * Lc[1]) + Pc[j * 4 + 2] * Lc[2]) + Pc[j * 4 + 3] * Lc[3];
        }
      }
      const double mx = fmax(fmax(out[0], out[1]), fmax(out[2], out[3]));
      if (mx < PA_SCALE_THRESH && mx > 0.0) {
#pragma unroll
        for (int j = 0; j < 4; j++) out[j] *= PA_SCALE_UP;
        sc += 1;
      }
      const uint32_t sl = me.x >> 16;
#pragma unroll
      for (int j = 0; j < 4; j++) stack[(sl * 4 + j) * kCBD + tid] = out[j];
      s.stack_sc[sl * kCBD + tid] = sc;
      double2* g2 = reinterpret_cast<double2*>(gpart + ((size_t)me.y * Ppad + p) * 4);
      g2[0] = make_double2(out[0], out[1]);
      g2[1] = make_double2(out[2], out[3]);
      gsc[(size_t)me.y * Ppad + p] = (int16_t)sc;
      if ((int)(me.x & 0xffffu) == root) {
#pragma unroll
        for (int j = 0; j < 4; j++) root_l[j] = out[j];
        root_sc = sc;
      }
    }
    const double lik = ((pi[0] * root_l[0] + pi[1] * root_l[1]) + pi[2] * root_l[2]) + pi[3] * root_l[3];
    acc += (double)weights[p] * (log(lik) - root_sc * PA_LOG_SCALE);
  }
  const double v = block_sum(acc, s.red, tid, kCBD);
  __syncthreads();
  return v;
}

__device__ __forceinline__ const double* class_log(const double* P0, const double* buf, int i, int k) {
  return k == 0 ? P0 + (size_t)i * 16 : buf + ((size_t)i * 5 + (k - 1)) * 16;
}

__device__ __forceinline__ double pair_dot(const double* __restrict__ lp, const double* __restrict__ post) {
  double sum = 0.0;
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) {
      const double c = post[a * 4 + b];
      if (c != 0.0) sum += lp[b * 4 + a] * c;  // log P(parent b -> child a) x expected pairs (child a, parent b)
    }
  return sum;
}

// emissions of classes k_lo..5 of every branch into dst; class 0 (and missing branches) are copied from `base` when k_lo == 1
__device__ void emissions(const Sh& s, const GbgcChainLaunch& L, double* dst, const double* base, const double* P0, const double* buf,
                          const double* post, int n, int k_lo, int tid) {
  const double l_b0 = s.scal[SC_LB0], l_b1 = s.scal[SC_LB1], l_c = s.scal[SC_LC], l_n = s.scal[SC_LN];  // logs of the chain's constants
  for (int t = tid; t < 6 * (n - 1); t += kCBD) {
    const int i = t / 6, k = t - 6 * i;
    const bool fix1 = s.fix[i] == 1;
    double e;
    if (s.flags[i] & PA_F_MISSING) {
      if (base) e = base[t];
      else {
        e = 0.0;
        if (!(s.flags[s.parent[i]] & PA_F_MISSING)) e = (k % 3 == 1) ? s.scal[SC_LCM] : s.scal[SC_LNM];
        if (fix1 && k % 3 == 2) e = -CUDART_INF;
      }
    } else if (fix1 && k % 3 == 2) {
      e = -CUDART_INF;
    } else if (k < k_lo) {
      e = base[t];
    } else {
      e = pair_dot(class_log(P0, buf, i, k), post + (size_t)i * 16) + (k < 3 ? l_b0 : l_b1) + (k % 3 == 1 ? l_c : l_n);
    }
    dst[t] = e;
  }
}

// sum over the non-missing branches of <log P of the branch's class, expected pairs>  (BPP_C::ELBO, GBGC/SRC/bpp_c.cpp:840-878)
// The same sum taken from emissions already in shared memory: emis[i][k] = <log P_k, pairs> + log prior(B) + log(1 - loss), so the branch's term
// is the emission of its class minus the two constants -- no second pass over the log tables in global memory.  `emis` must belong to the
// same rates and the same expected pairs as the tables the term is defined on (s.emis after initial_prob_back / after the rate moves).
__device__ double elbo_from_emissions(const Sh& s, const GbgcChainLaunch& L, const double* emis, const int8_t* Z, const int8_t* B, int n, int tid) {
  const double l_b0 = s.scal[SC_LB0], l_b1 = s.scal[SC_LB1], l_c = s.scal[SC_LC], l_n = s.scal[SC_LN];  // logs of the chain's constants
  double acc = 0.0;
  for (int i = tid; i < n - 1; i += kCBD)
    if (!(s.flags[i] & PA_F_MISSING)) {
      const int k = Z[i] + 3 * B[i];
      acc += (emis[i * 6 + k] - (k < 3 ? l_b0 : l_b1)) - (k % 3 == 1 ? l_c : l_n);
    }
  return block_sum(acc, s.red, tid, kCBD);
}

__device__ double elbo_term(const Sh& s, const int8_t* Z, const int8_t* B, const double* P0, const double* buf, const double* post, int n, int tid) {
  double acc = 0.0;
  for (int i = tid; i < n - 1; i += kCBD)
    if (!(s.flags[i] & PA_F_MISSING)) acc += pair_dot(class_log(P0, buf, i, Z[i] + 3 * B[i]), post + (size_t)i * 16);
  return block_sum(acc, s.red, tid, kCBD);
}

// log P(t) of classes k_lo..5 of every branch at (sel_c, sel_n, gbc) into P0 (class 0) / buf (classes 1..5); row-major P(i -> j).
// Uniformisation with a power basis SHARED by the branches of a class: with mu = max |Q_ii| and A = I + Q / mu (stochastic), the powers
// A^1..A^19 are built once per class (16 lanes per class, one entry each), and a branch of length t is
//   P(t) = [ sum_j Poisson(j; x) A^j ]^(2^q),  x = mu t / 2^q <= 1/2,
// i.e. a linear combination of the stored powers (16 FMA per term, as many terms as the Poisson tail needs) and q squarings -- against
// 19 + q matrix products per branch before.  Every term is non-negative: no cancellation, nothing depends on the spectrum.
// emis_dst != nullptr: the emissions of the built classes are taken from the matrices while they are in registers (the proposal's
// getEmission_update), every other entry (class 0, missing branches, forbidden classes) is copied from emis_base.
__device__ void build_logs(const Sh& s, const GbgcChainLaunch& L, const ElemView& ev, double* P0, double* buf, double sel_c, double sel_n, double gbc, int n,
                           int k_lo, int tid, double* emis_dst = nullptr, const double* emis_base = nullptr, const double* post = nullptr) {
  if (emis_dst)
    for (int t = tid; t < 6 * (n - 1); t += kCBD) {
      const int i = t / 6, k = t - 6 * i;
      if ((s.flags[i] & PA_F_MISSING) || k < k_lo) emis_dst[t] = emis_base[t];
      else if (s.fix[i] == 1 && k % 3 == 2) emis_dst[t] = -CUDART_INF;
    }
  const double l_b0 = s.scal[SC_LB0], l_b1 = s.scal[SC_LB1], l_c = s.scal[SC_LC], l_n = s.scal[SC_LN];  // logs of the chain's constants
  if (tid >= k_lo && tid < 6) {
    double Q[16];
    gbgc_rate(L.Q0, tid % 3 == 0 ? 0.0 : (tid % 3 == 1 ? sel_c : sel_n), tid >= 3 ? gbc : 0.0, Q);
    const double mu = fmax(fmax(-Q[0], -Q[5]), fmax(-Q[10], -Q[15]));
    s.mu[tid] = mu;
#pragma unroll
    for (int e = 0; e < 16; e++) s.pw[tid * kPwStride + 16 + e] = (mu > 0.0 ? Q[e] / mu : 0.0) + (((e >> 2) == (e & 3)) ? 1.0 : 0.0);  // A^1
  }
  __syncthreads();
  if (tid < 96) {  // 16 lanes per class: lane (r, c) owns entry [r][c] of every power
    const int k = tid >> 4, e = tid & 15, r = e >> 2, c = e & 3;
    const double* A = s.pw + k * kPwStride + 16;
    const double a0 = A[c], a1 = A[4 + c], a2 = A[8 + c], a3 = A[12 + c];
    for (int j = 2; j <= kPow; j++) {
      if (k >= k_lo) {
        const double* prev = s.pw + k * kPwStride + (j - 1) * 16 + r * 4;
        s.pw[k * kPwStride + j * 16 + e] = ((prev[0] * a0 + prev[1] * a1) + prev[2] * a2) + prev[3] * a3;
      }
      __syncwarp();
    }
  }
  __syncthreads();
  const int nk = 6 - k_lo;
  for (int t = tid; t < nk * (n - 1); t += kCBD) {
    const int i = t / nk, k = k_lo + (t - nk * i);
    if (s.flags[i] & PA_F_MISSING) continue;
    if (s.fix[i] == 1 && k % 3 == 2) continue;  // never read (update_cache_TM's own skip, GBGC/SRC/bpp_c.cpp:1655)
    double P16[16];
    const double mu = s.mu[k], tt = ev.dist[i];
    if (!(mu > 0.0) || !(tt > 0.0)) {
#pragma unroll
      for (int e = 0; e < 16; e++) P16[e] = ((e >> 2) == (e & 3)) ? 1.0 : 0.0;
    } else {
      int ksq = 0;
      double x = mu * tt;
      while (x > 0.5 && ksq < 60) { x *= 0.5; ksq++; }
      double wgt = exp(-x);
#pragma unroll
      for (int e = 0; e < 16; e++) P16[e] = ((e >> 2) == (e & 3)) ? wgt : 0.0;
      const double* pw = s.pw + k * kPwStride;
      for (int j = 1; j <= kPow; j++) {
        wgt *= x * c_inv[j];
        const double* Aj = pw + j * 16;
#pragma unroll
        for (int e = 0; e < 16; e++) P16[e] += wgt * Aj[e];
        if (wgt < 1e-19) break;  // x <= 1/2: the remaining Poisson tail is below 1e-19 of a row sum
      }
      double nxt[16];
      for (int q = 0; q < ksq; q++) {
        matmul4(P16, P16, nxt);
#pragma unroll
        for (int e = 0; e < 16; e++) P16[e] = nxt[e];
      }
    }
    double* dst = k == 0 ? P0 + (size_t)i * 16 : buf + ((size_t)i * 5 + (k - 1)) * 16;
#pragma unroll
    for (int e = 0; e < 16; e++) { P16[e] = log_fast(P16[e]); dst[e] = P16[e]; }
    if (emis_dst) emis_dst[i * 6 + k] = pair_dot(P16, post + (size_t)i * 16) + (k < 3 ? l_b0 : l_b1) + (k % 3 == 1 ? l_c : l_n);
  }
  __syncthreads();
}

// log of the Gamma(a, scale b) density at x with K = lgamma(a) + log(b) precomputed (the reference takes log(gsl_ran_gamma_pdf(x, a, b)))
__device__ __forceinline__ double lgp(double x, double a, double b, double K) {
  if (!(x > 0.0)) return (x == 0.0 && a == 1.0) ? -K : -CUDART_INF;
  return (a == 1.0 ? 0.0 : (a - 1.0) * log(x / b)) - x / b - K;
}

__device__ __forceinline__ int categorical_n(const double* w, int K, double u) {
  double tot = 0.0;
  for (int k = 0; k < K; k++) tot += w[k];
  const double x = u * tot;
  double acc = 0.0;
  for (int k = 0; k < K; k++) {
    acc += w[k];
    if (x < acc) return k;
  }
  for (int k = K - 1; k >= 0; k--)
    if (w[k] > 0.0) return k;
  return -1;
}

// deterministic sum over warp 0 (values of the other warps are ignored): returns the total on every lane of warp 0
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

#ifndef PA_GBGC_CTAS
#define PA_GBGC_CTAS 3
#endif
__global__ void __launch_bounds__(kCBD, PA_GBGC_CTAS) gbgc_chain_kernel(GbgcChainLaunch L) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_chain, s_flag, s_bufA, s_bufW;
  __shared__ long long s_ph[12], s_t0;
  __shared__ int s_ntask, s_nbranch, s_ndown, s_nup;
  __shared__ double s_pi[4];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_max = L.n_max, slots_max = L.slots_max, N = L.N, T = L.T;
  Sh s;
  {
    double* d = reinterpret_cast<double*>(smem_raw);
    s.Pm = d; d += (size_t)n_max * 16;
    s.up_ch = reinterpret_cast<uint4*>(d); d += (size_t)(n_max / 2 + 1) * 2;
    s.up_self = reinterpret_cast<uint2*>(d); d += (size_t)(n_max / 2 + 1);
    s.uni = d;  // pruning stack [slots][4][BD] + scale counts | power basis of the P(t) builds
    const size_t uni_d = uni_doubles(n_max, slots_max);
    s.stack_sc = reinterpret_cast<int*>(d + (size_t)slots_max * 4 * kCBD);
    d += uni_d;
    s.red = d; d += 34;
    s.emis = d; d += (size_t)n_max * 6;
    s.emis_t = d; d += (size_t)n_max * 6;
    s.back = d; d += (size_t)n_max * 6;
    s.Qc = d; d += 96;
    s.scal = d; d += SC_N;
    s.msg = d; d += (size_t)n_max * 3;
    s.lsum = d; d += n_max;
    s.pw = s.uni;  // the power basis of build_logs lives in the phase-shared region (no pruning pass runs during the rate moves)
    s.mu = d; d += 6;
    s.msc = reinterpret_cast<int*>(d); d += (n_max + 1) / 2;
    s.codes_s = reinterpret_cast<uint8_t*>(d); d += (L.codes_cap + 7) / 8;
    int16_t* h16 = reinterpret_cast<int16_t*>(d);
    s.child = h16; h16 += 2 * n_max;
    s.tiprow = h16; h16 += n_max;
    s.su = h16; h16 += n_max;
    s.intidx = h16; h16 += n_max;
    s.parent = h16; h16 += n_max;
    s.up_nodes = h16; h16 += n_max;
    s.up_off = h16; h16 += n_max + 2;
    s.down_nodes = h16; h16 += n_max;
    s.down_off = h16; h16 += n_max + 2;
    s.dn_node = h16; h16 += n_max;
    s.dn_src = h16; h16 += n_max;
    s.dn_par = h16; h16 += n_max;
    uint8_t* b8 = reinterpret_cast<uint8_t*>(h16);
    s.flags = b8; b8 += n_max;
    s.fix = b8; b8 += n_max;
    s.upr = b8; b8 += n_max;
    s.slot = reinterpret_cast<int8_t*>(b8); b8 += n_max;
    s.Z = reinterpret_cast<int8_t*>(b8); b8 += n_max;
    s.B = reinterpret_cast<int8_t*>(b8); b8 += n_max;
    s.Zn = reinterpret_cast<int8_t*>(b8); b8 += n_max;
    s.Bn = reinterpret_cast<int8_t*>(b8); b8 += n_max;
  }
  // per-CTA scratch
  unsigned char* sc = L.scratch + (size_t)blockIdx.x * L.stride;
  double* P0 = reinterpret_cast<double*>(sc);                     // [n_max][16]
  double* bufs = P0 + (size_t)n_max * 16;                         // [3][n_max][5][16]
  const size_t buf_d = (size_t)n_max * 80;
  double* postb = bufs + 3 * buf_d;                               // [2][n_max][16]
  double* gacc = postb + (size_t)2 * n_max * 16;                  // [warps][n_max][16] pair sums of each warp's patterns
  double* gpart = gacc + (size_t)kNW * n_max * 16;                  // [n_max][P_max][4]
  int16_t* gsc = reinterpret_cast<int16_t*>(gpart + (size_t)n_max * L.P_max * 4);
  const pa_gbgc_params& pr = L.prm;
  if (tid < 4) s_pi[tid] = L.pi[tid];
  if (tid < 12) s_ph[tid] = 0;
  if (tid == 0) s_t0 = clock64();

  for (;;) {
    __syncthreads();
    if (tid == 0) s_chain = atomicAdd(L.queue, 1);
    __syncthreads();
    const int ch = s_chain;
    if (ch >= L.n_chains) break;
    const int64_t off = L.blob_off[L.chain_elem[ch]];
    if (off < 0) continue;  // filtered element
    const ElemView ev = make_view(L.blobs + off);
    const ElemHeader* h = ev.h;
    const int n = h->n, ni = h->ni, root = n - 1, n_up = h->n_up, n_down = h->n_down, P = h->P, Ppad = h->Ppad;
    const int resz = L.chain_resz[ch];
    const uint32_t key = (uint32_t)h->elem_id * 4u + (uint32_t)resz;
    double* tr = L.tr + (size_t)ch * 5 * T;
    uint8_t* trzb = L.tr_zb + (size_t)ch * T * n_max;
    for (int i = tid; i < n; i += kCBD) {
      s.child[2 * i] = ev.child[2 * i];
      s.child[2 * i + 1] = ev.child[2 * i + 1];
      s.tiprow[i] = ev.tiprow[i];
      s.intidx[i] = ev.intidx[i];
      s.parent[i] = ev.parent[i];
      s.flags[i] = ev.flags[i];
      s.slot[i] = ev.slot[i];
      s.up_nodes[i] = ev.up_nodes[i];
      s.down_nodes[i] = ev.down_nodes[i];
      const int o = ev.orig[i];
      s.fix[i] = (uint8_t)(L.fixZ[(size_t)ch * N + o] == 1 ? 1 : 0);
      s.upr[i] = (uint8_t)(L.upper[(size_t)ch * N + o] ? 1 : 0);
      s.Z[i] = 0; s.B[i] = 0; s.Zn[i] = 0; s.Bn[i] = 0;
    }
    for (int i = tid; i < ni; i += kCBD) s.su[i] = ev.su[i];
    for (int i = tid; i <= n_up; i += kCBD) s.up_off[i] = ev.up_off[i];
    for (int i = tid; i <= n_down; i += kCBD) s.down_off[i] = ev.down_off[i];
    if (tid == 0) {
      s.scal[SC_LG0] = log(1.0 - pr.grate); s.scal[SC_LG1] = log(pr.grate);
      s.scal[SC_LL0] = log(1.0 - pr.lrate); s.scal[SC_LL1] = log(pr.lrate);
      s.scal[SC_SELC] = pr.sel_c0; s.scal[SC_SELN] = pr.sel_n0; s.scal[SC_GBC] = pr.gbc0;
      s.scal[SC_G] = pr.grate; s.scal[SC_L] = pr.lrate;
      // constants of the three log Gamma densities: lgamma(a) + log(b)
      s.scal[SC_KC] = lgamma(pr.cprior_a) + log(pr.cprior_b); s.scal[SC_KN] = lgamma(pr.nprior_a) + log(pr.nprior_b);
      s.scal[SC_KG] = lgamma(pr.gbprior_a) + log(pr.gbprior_b);
      // log prior of the gBGC indicator (absent / present) and log(1 - loss) of a conserved / non-conserved branch
      s.scal[SC_LB0] = log(1.0 - pr.gbprior_c); s.scal[SC_LB1] = log(pr.gbprior_c);
      s.scal[SC_LC] = log(1.0 - pr.cons_to_mis); s.scal[SC_LN] = log(1.0 - pr.ncons_to_mis);
      s.scal[SC_LCM] = log(pr.cons_to_mis); s.scal[SC_LNM] = log(pr.ncons_to_mis);
      s_bufA = 0; s_bufW = 0;
    }
    __syncthreads();
    if (warp == 0) {  // work counters: (branch, class) matrices per rate move, non-missing branches
      int nt = 0, nb = 0;
      for (int i = lane; i < n - 1; i += 32)
        if (!(s.flags[i] & PA_F_MISSING)) { nb++; nt += s.fix[i] == 1 ? 3 : 5; }
      nt = __reduce_add_sync(0xffffffffu, nt);
      nb = __reduce_add_sync(0xffffffffu, nb);
      if (lane == 0) { s_ntask = nt; s_nbranch = nb; }
    }
    // the packed tip codes never change during a chain: staged once in shared memory when they fit (rows of Ppad / 2 bytes, a multiple of 16)
    const int half = Ppad >> 1;
    const bool staged = (size_t)h->S * half <= (size_t)L.codes_cap;
    if (staged) {
      const uint32_t* src = reinterpret_cast<const uint32_t*>(ev.codes);
      uint32_t* dst = reinterpret_cast<uint32_t*>(s.codes_s);
      for (int t = tid; t < (h->S * half) >> 2; t += kCBD) dst[t] = __ldg(src + t);
    }
    const uint8_t* codes = staged ? s.codes_s : ev.codes;
    // program of the downward pass: the non-missing branches, parents before children; source of the branch's own upward partial
    // (tip row as -(row + 1), or the internal node's arena index) and of its parent's posterior (-1: the parent is the branch of the
    // previous step, its posterior is still in registers)
    if (tid == 0) {
      int k = 0, prev = root;
      for (int i = n - 2; i >= 0; i--) {
        if (s.flags[i] & PA_F_MISSING) continue;
        s.dn_node[k] = (int16_t)i;
        s.dn_src[k] = (s.flags[i] & PA_F_TIP) ? (int16_t)(-(s.tiprow[i] + 1)) : s.intidx[i];
        s.dn_par[k] = (s.parent[i] == prev) ? (int16_t)-1 : s.intidx[s.parent[i]];
        prev = i;
        k++;
      }
      s_ndown = k;
      // program of the upward pass (see upward_pass)
      int q = 0;
      for (int kk = 0; kk < ni; kk++) {
        const int x = s.su[kk];
        if (x < 0) break;
        uint32_t w4[4];
        for (int cc = 0; cc < 2; cc++) {
          const int chd = s.child[2 * x + cc];
          const uint8_t f = s.flags[chd];
          const uint32_t kind = (f & PA_F_MISSING) ? 0u : ((f & PA_F_TIP) ? 1u : 2u);
          w4[2 * cc] = (uint32_t)chd | (kind << 16);
          w4[2 * cc + 1] = kind == 1u ? (uint32_t)s.tiprow[chd] : (uint32_t)s.slot[chd];
        }
        s.up_ch[q] = make_uint4(w4[0], w4[1], w4[2], w4[3]);
        s.up_self[q] = make_uint2((uint32_t)x | ((uint32_t)s.slot[x] << 16), (uint32_t)s.intidx[x]);
        q++;
      }
      s_nup = q;
    }
    // log P(constraints): the 6-state pass of initMCMC with "- log 2" per message (GBGC/SRC/bpp_c.cpp:232-254)
    for (int t = tid; t < n * 6; t += kCBD) s.emis_t[t] = (s.fix[t / 6] == 1 && (t % 3) == 2) ? -CUDART_INF : 0.0;
    __syncthreads();
    messages(s, s.emis_t, n, n_up, true, tid);
    if (tid == 0) s.scal[SC_LOGPZC] = s.back[root * 6];
    // the six classes' log P(t) at the initial rates; P(t) of the initial classes (Z = 0, B = 0 everywhere)
    build_logs(s, L, ev, P0, bufs, pr.sel_c0, pr.sel_n0, pr.gbc0, n, 0, tid);
    auto load_Pm = [&](const int8_t* Zs, const int8_t* Bs, const double* buf) {
      for (int t = tid; t < (n - 1) * 16; t += kCBD) {
        const int i = t >> 4;
        if (s.flags[i] & PA_F_MISSING) continue;
        s.Pm[t] = exp(class_log(P0, buf, i, Zs[i] + 3 * Bs[i])[t & 15]);
      }
      __syncthreads();
    };
    auto upward = [&]() -> double { return upward_pass(s, h, codes, ev.weights, gpart, gsc, s_pi, tid); };
    // (Z, B) integrated out: log_prob_back[root][0] into s.scal[SC_LL_PROP]; true = s.back holds LINEAR class weights, false = logs
    auto integrate = [&](const double* src) -> bool {
      const bool lin = messages_lin(s, src, n, n_up, tid, &s.scal[SC_LL_PROP]);
      if (!lin) {
        messages(s, src, n, n_up, false, tid);
        if (tid == 0) s.scal[SC_LL_PROP] = s.back[root * 6];
        __syncthreads();
      }
      return lin;
    };
    load_Pm(s.Z, s.B, bufs);
    s.n_upprog = s_nup;  // written by thread 0 above, several barriers ago
    {
      const double ll0 = upward();
      if (tid == 0) {
        s.scal[SC_LOGLIK] = ll0;
        double pz = 0.0;
        for (int i = n - 2; i >= 0; i--) pz += lt_val(s.scal, 0, 0, s.upr[i] != 0);
        s.scal[SC_LOGPZ] = pz - s.scal[SC_LOGPZC];
        s.scal[SC_FULL] = 0.0;
      }
    }
    __syncthreads();
    PH(PH_SETUP);
    bool move = true;
    int cur = 0, nxt = 0;  // expected-pair buffers: post_cur / post_next
    int acc_rates = 0, acc_zb = 0;
    for (int m = 0; m < T; m++) {
      // ---- 1. MonitorChain
      if (warp == 0) {
        if (move) {
          int c0 = 0, c1 = 0, c2 = 0, c3 = 0, gb = 0, ungb = 0;
          for (int i = lane; i < n - 1; i += 32) {
            if (s.flags[s.parent[i]] & PA_F_MISSING) continue;
            const bool mis = (s.flags[i] & PA_F_MISSING) != 0;
            const int idx = 2 * (mis ? 1 : 0) + (s.Z[i] == 1 ? 1 : 0);
            c0 += idx == 0; c1 += idx == 1; c2 += idx == 2; c3 += idx == 3;
            if (!mis) { gb += s.B[i] ? 1 : 0; ungb += s.B[i] ? 0 : 1; }
          }
          c0 = __reduce_add_sync(0xffffffffu, c0); c1 = __reduce_add_sync(0xffffffffu, c1); c2 = __reduce_add_sync(0xffffffffu, c2);
          c3 = __reduce_add_sync(0xffffffffu, c3); gb = __reduce_add_sync(0xffffffffu, gb); ungb = __reduce_add_sync(0xffffffffu, ungb);
          if (lane == 0) {
            double v = s.scal[SC_LOGLIK] + s.scal[SC_LOGPZ];
            v += lgp(-s.scal[SC_SELC], pr.cprior_a, pr.cprior_b, s.scal[SC_KC]);
            if (resz != 0) v += lgp(s.scal[SC_SELN] + pr.nprior_c, pr.nprior_a, pr.nprior_b, s.scal[SC_KN]);
            v += s.scal[SC_LN] * c0 + s.scal[SC_LC] * c1 + s.scal[SC_LNM] * c2 + s.scal[SC_LCM] * c3;
            v += s.scal[SC_LB0] * ungb + lgp(s.scal[SC_GBC], pr.gbprior_a, pr.gbprior_b, s.scal[SC_KG]) + s.scal[SC_LB1] * gb;
            s.scal[SC_FULL] = v;
          }
        }
        __syncwarp();
        if (lane == 0) {
          tr[m] = s.scal[SC_LOGLIK];
          tr[T + m] = s.scal[SC_FULL];
          tr[2 * T + m] = s.scal[SC_SELC];
          tr[3 * T + m] = s.scal[SC_SELN];
          tr[4 * T + m] = s.scal[SC_GBC];
        }
      }
      for (int i = tid; i < n; i += kCBD) trzb[(size_t)m * n_max + i] = (uint8_t)(s.Z[i] | (s.B[i] << 2));
      PH(PH_MONITOR);
      if (m == T - 1) break;
      // ---- 2. expected (child, parent) pairs at the current state (only when the state moved): parents before children, an internal
      //         node's partial is overwritten by the posterior of its base.  Thread = pattern; the step of branch k + 1 has its operands
      //         (its own upward partial or tip code, its parent's posterior) requested before branch k is computed, and a parent that was
      //         the previous step hands its posterior over in registers: the global-memory latency is off the dependent chain.  The 16
      //         pair sums of a warp's 32 patterns go to that warp's own slice of `gacc` (no block barrier inside the pass).
      if (move) {
        nxt = (m == 0) ? 0 : 1 - cur;
        double* ga = gacc + (size_t)warp * n_max * 16;
        double* rt = s.uni + warp * (16 * 33);  // the pruning stack is idle during this pass
        const int nd = s_ndown;
        int tile = 0;
        for (int p0 = 0; p0 < Ppad; p0 += kCBD, tile++) {
          if (p0 + warp * 32 >= P) break;  // none of this warp's patterns in this tile or a later one
          const int p = p0 + tid;
          const bool act = p < P;
          const double w = act ? (double)ev.weights[p] : 0.0;
          double qprev[4] = {0.0, 0.0, 0.0, 0.0};
          if (act) {
            double* g = gpart + ((size_t)s.intidx[root] * Ppad + p) * 4;
            const double r0 = s_pi[0] * g[0], r1 = s_pi[1] * g[1], r2 = s_pi[2] * g[2], r3 = s_pi[3] * g[3];
            const double tot = ((r0 + r1) + r2) + r3;
            qprev[0] = r0 / tot; qprev[1] = r1 / tot; qprev[2] = r2 / tot; qprev[3] = r3 / tot;
            g[0] = qprev[0]; g[1] = qprev[1]; g[2] = qprev[2]; g[3] = qprev[3];
          }
          double LvN[4] = {0.0, 0.0, 0.0, 0.0}, qpN[4] = {0.0, 0.0, 0.0, 0.0};
          int nodeN = 0, srcN = 0, parN = -1;
          auto fetch = [&](int k) {
            nodeN = s.dn_node[k]; srcN = s.dn_src[k]; parN = s.dn_par[k];
            if (act) {
              if (srcN < 0) {
                const uint32_t byte = codes[(size_t)(-srcN - 1) * half + (p >> 1)];
                const uint32_t code = (byte >> ((p & 1) * 4)) & 15u;
#pragma unroll
                for (int a = 0; a < 4; a++) LvN[a] = ((code >> a) & 1u) ? 1.0 : 0.0;
              } else {
                const double2* g2 = reinterpret_cast<const double2*>(gpart + ((size_t)srcN * Ppad + p) * 4);
                const double2 x = g2[0], y = g2[1];
                LvN[0] = x.x; LvN[1] = x.y; LvN[2] = y.x; LvN[3] = y.y;
              }
              if (parN >= 0) {
                const double2* q2 = reinterpret_cast<const double2*>(gpart + ((size_t)parN * Ppad + p) * 4);
                const double2 x = q2[0], y = q2[1];
                qpN[0] = x.x; qpN[1] = x.y; qpN[2] = y.x; qpN[3] = y.y;
              }
            }
          };
          if (nd > 0) fetch(0);
          for (int k = 0; k < nd; k++) {
            const int i = nodeN, src = srcN;
            const bool fwd = parN < 0;
            double Lv[4], qp[4];
#pragma unroll
            for (int a = 0; a < 4; a++) { Lv[a] = LvN[a]; qp[a] = fwd ? qprev[a] : qpN[a]; }
            double old = 0.0;
            if (tile > 0 && lane < 16) old = ga[i * 16 + lane];
            if (k + 1 < nd) fetch(k + 1);
            // 16 weighted pairs of this pattern into the warp's transposition tile [16][33], then lane (e, h) adds up element e of the patterns
            // 16 h .. 16 h + 15 in order and the two halves meet in one shuffle: 16 STS + 16 LDS + 16 DADD instead of a 31-shuffle butterfly
            // an unambiguous tip (one base a*): ratio[b] PL[b][a*] = qp[b], so its pairs are w qp[b] in row a*, nothing else to compute
            const bool tip1 = act && src < 0 && ((Lv[0] + Lv[1]) + (Lv[2] + Lv[3])) == 1.0;
            if (tip1) {
#pragma unroll
              for (int b = 0; b < 4; b++) {
                const double v = qp[b] * w;
#pragma unroll
                for (int a = 0; a < 4; a++) rt[(a * 4 + b) * 33 + lane] = Lv[a] * v;
              }
            } else if (act) {
              const double* Pi = s.Pm + i * 16;
              double PL[16], ratio[4];
#pragma unroll
              for (int e = 0; e < 16; e++) PL[e] = Pi[e] * Lv[e & 3];  // [parent base b][child base a]
#pragma unroll
              for (int b = 0; b < 4; b++) {
                const double mm = ((PL[b * 4] + PL[b * 4 + 1]) + PL[b * 4 + 2]) + PL[b * 4 + 3];
                ratio[b] = (mm > 0.0) ? div_fast(qp[b], mm) : 0.0;
              }
#pragma unroll
              for (int a = 0; a < 4; a++) qprev[a] = ((ratio[0] * PL[a] + ratio[1] * PL[4 + a]) + ratio[2] * PL[8 + a]) + ratio[3] * PL[12 + a];
              if (src >= 0) {
                double2* g2 = reinterpret_cast<double2*>(gpart + ((size_t)src * Ppad + p) * 4);
                g2[0] = make_double2(qprev[0], qprev[1]);
                g2[1] = make_double2(qprev[2], qprev[3]);
              }
#pragma unroll
              for (int b = 0; b < 4; b++) {
                const double rw = ratio[b] * w;  // weighted by the pattern's column count
#pragma unroll
                for (int a = 0; a < 4; a++) rt[(a * 4 + b) * 33 + lane] = rw * PL[b * 4 + a];  // [child base a][parent base b]
              }
            } else {
#pragma unroll
              for (int e = 0; e < 16; e++) rt[e * 33 + lane] = 0.0;
            }
            __syncwarp();
            double v1 = 0.0;
            {
              const double* col = rt + (lane & 15) * 33 + (lane & 16);
#pragma unroll
              for (int q = 0; q < 16; q++) v1 += col[q];
            }
            v1 += __shfl_xor_sync(0xffffffffu, v1, 16);
            if (lane < 16) ga[i * 16 + lane] = old + v1;
            __syncwarp();
          }
        }
        __syncthreads();
        double* pn = postb + (size_t)nxt * n_max * 16;
        const int nw = min(kNW, (P + 31) >> 5);  // warps that own patterns, in fixed order
        for (int t = tid; t < (n - 1) * 16; t += kCBD) {
          double v = 0.0;
          if (!(s.flags[t >> 4] & PA_F_MISSING)) {
            v = gacc[t];
            for (int ww = 1; ww < nw; ww++) v += gacc[(size_t)ww * n_max * 16 + t];
          }
          pn[t] = v;
        }
        __syncthreads();
      }
      PH(PH_DOWN);
      if (m == 0) cur = nxt;
      const double* post = postb + (size_t)cur * n_max * 16;
      // ---- 3. initial_prob_back: emissions of the six classes, messages, ELBO at the current classes
      const double* bufA = bufs + (size_t)s_bufA * buf_d;
      emissions(s, L, s.emis, nullptr, P0, bufA, post, n, 0, tid);
      __syncthreads();
      PH(PH_INIT_EMIS);
      integrate(s.emis);
      PH(PH_INIT_MSG);
      const double elbo_old = elbo_from_emissions(s, L, s.emis, s.Z, s.B, n, tid);
      if (tid == 0) {
        s.scal[SC_LOGLIK_OLD] = s.scal[SC_LL_PROP];
        s.scal[SC_ELBO_OLD] = elbo_old;
        s.scal[SC_PSELC] = s.scal[SC_SELC]; s.scal[SC_PSELN] = s.scal[SC_SELN]; s.scal[SC_PGBC] = s.scal[SC_GBC];  // cur_* of sample_rates
        s_bufW = s_bufA;
      }
      __syncthreads();
      // ---- 4. sample_rates: three joint moves with (Z, B) integrated out
      bool any_moved = false;
      for (int mm = 0; mm < 3; mm++) {
        const uint4 w4 = philox4(pr.seed, key, (uint32_t)mm, 0u, (uint32_t)m, PA_ST_GBGC);
        const double cur_c = s.scal[SC_PSELC], cur_n = s.scal[SC_PSELN], cur_g = s.scal[SC_PGBC];
        double new_g = cur_g + (u01(w4.x) - 0.5) * pr.prop_gb * 2;
        if (new_g < 0.01) new_g = cur_g;
        double new_n = cur_n;
        if (resz != 0) {
          const double u = u01(w4.y) * (pr.prop_n - 1 / pr.prop_n) + 1 / pr.prop_n;
          new_n = (cur_n + pr.nprior_c) * u - pr.nprior_c;
        }
        const double new_c = cur_c * (u01(w4.z) * (pr.prop_c - 1 / pr.prop_c) + 1 / pr.prop_c);
        const int bA = s_bufA, bW = s_bufW;
        const int bP = (bA != 0 && bW != 0) ? 0 : ((bA != 1 && bW != 1) ? 1 : 2);
        double* bufP = bufs + (size_t)bP * buf_d;
        __syncthreads();
        PH(PH_ELBO);
        build_logs(s, L, ev, P0, bufP, new_c, new_n, new_g, n, 1, tid, s.emis_t, s.emis, post);
        PH(PH_BUILD);
        integrate(s.emis_t);
        PH(PH_MSG);
        if (tid == 0) {
          const double ll_p = s.scal[SC_LL_PROP];
          double ratio = ll_p - s.scal[SC_LOGLIK_OLD] + lgp(new_g, pr.gbprior_a, pr.gbprior_b, s.scal[SC_KG]) - lgp(cur_g, pr.gbprior_a, pr.gbprior_b, s.scal[SC_KG]);
          if (resz != 0)
            ratio += lgp(new_n + pr.nprior_c, pr.nprior_a, pr.nprior_b, s.scal[SC_KN]) - lgp(cur_n + pr.nprior_c, pr.nprior_a, pr.nprior_b, s.scal[SC_KN]);
          ratio += lgp(-new_c, pr.cprior_a, pr.cprior_b, s.scal[SC_KC]) - lgp(-cur_c, pr.cprior_a, pr.cprior_b, s.scal[SC_KC]) - log(-new_c) + log(-cur_c);
          const bool acc = log(u01(w4.w)) < ratio;
          s_flag = acc ? 1 : 0;
          if (acc) {
            s.scal[SC_PSELC] = new_c; s.scal[SC_PSELN] = new_n; s.scal[SC_PGBC] = new_g;
            s.scal[SC_LOGLIK_OLD] = ll_p;
            s_bufW = bP;
          }
        }
        __syncthreads();
        if (s_flag) {
          any_moved = true;
          for (int t = tid; t < (n - 1) * 6; t += kCBD) s.emis[t] = s.emis_t[t];
        }
        __syncthreads();
      }
      acc_rates += any_moved ? 1 : 0;
      // ---- 5. Update_ZB + update_rates
      const double* bufW = bufs + (size_t)s_bufW * buf_d;
      PH(PH_COMMIT);
      const bool lin = integrate(s.emis);
      PH(PH_MSG);
      for (int d = 1; d < n_down; d++) {
        const int lo = s.down_off[d], cnt = s.down_off[d + 1] - lo;
        for (int t = tid; t < cnt; t += kCBD) {
          const int i = s.down_nodes[lo + t], par = s.parent[i];
          const int zp = s.Zn[par];
          const double u = u01(philox4(pr.seed, key, (uint32_t)i, 1u, (uint32_t)m, PA_ST_GBGC).x);
          const bool mis = (s.flags[i] & PA_F_MISSING) != 0;
          const bool up = s.upr[i] != 0;
          int zn = s.Z[i], bn = s.B[i];
          if (zp != 2) {
            double lw[6];
            if (lin) {
              for (int k = 0; k < 6; k++) lw[k] = lt_lin(s.scal, k % 3, zp, up) * s.back[i * 6 + k];
            } else {
              double mx = -CUDART_INF;
              for (int k = 0; k < 6; k++) {
                lw[k] = lt_val(s.scal, k % 3, zp, up) + s.back[i * 6 + k];
                mx = fmax(mx, lw[k]);
              }
              for (int k = 0; k < 6; k++) lw[k] = exp(lw[k] - mx);
            }
            const int nn = categorical_n(lw, 6, u);
            if (nn >= 0) {
              zn = nn % 3;
              if (!mis) bn = nn > 2 ? 1 : 0;
            }
          } else {
            zn = 2;
            double lw[2] = {s.back[i * 6 + 2], s.back[i * 6 + 5]};
            if (!lin) {
              const double mx = fmax(lw[0], lw[1]);
              lw[0] = exp(lw[0] - mx); lw[1] = exp(lw[1] - mx);
            }
            const int nn = categorical_n(lw, 2, u);
            if (nn >= 0 && !mis) bn = nn > 0 ? 1 : 0;
          }
          s.Zn[i] = (int8_t)zn;
          s.Bn[i] = (int8_t)bn;
        }
        __syncthreads();
      }
      PH(PH_DRAW);
      load_Pm(s.Zn, s.Bn, bufW);
      const double ll_new = upward();
      PH(PH_UP);
      const double elbo_new = elbo_from_emissions(s, L, s.emis, s.Zn, s.Bn, n, tid);
      if (tid == 0) {
        const double ratio = ll_new - s.scal[SC_LOGLIK] + s.scal[SC_ELBO_OLD] - elbo_new;
        const bool acc = log(u01(philox4(pr.seed, key, 0u, 2u, (uint32_t)m, PA_ST_GBGC).x)) < ratio;
        s_flag = acc ? 1 : 0;
        if (acc) {
          s.scal[SC_LOGLIK] = ll_new;
          s.scal[SC_SELC] = s.scal[SC_PSELC]; s.scal[SC_SELN] = s.scal[SC_PSELN]; s.scal[SC_GBC] = s.scal[SC_PGBC];
          s_bufA = s_bufW;
        }
      }
      __syncthreads();
      PH(PH_ELBO);
      move = s_flag != 0;
      if (move) {
        acc_zb++;
        for (int i = tid; i < n; i += kCBD) { s.Z[i] = s.Zn[i]; s.B[i] = s.Bn[i]; }
        __syncthreads();
        if (warp == 0) {  // prior_Z (GBGC/SRC/bpp_c2.cpp:484-498)
          double pz = 0.0;
          for (int i = lane; i < n - 1; i += 32) pz += lt_val(s.scal, s.Z[i], s.Z[s.parent[i]], s.upr[i] != 0);
          pz = warp_sum(pz);
          if (lane == 0) s.scal[SC_LOGPZ] = pz - s.scal[SC_LOGPZC];
        }
      } else {
        for (int i = tid; i < n; i += kCBD) { s.Zn[i] = s.Z[i]; s.Bn[i] = s.B[i]; }
        // the partials in the arena now belong to the rejected state; they are not read again before the next accepted state rewrites them
      }
      __syncthreads();
      cur = nxt;
      PH(PH_COMMIT);
      if (tid == 0 && L.counters) {
        atomicAdd(&L.counters[0], (unsigned long long)ni * (unsigned long long)P);
        atomicAdd(&L.counters[1], 3ull * (unsigned long long)s_ntask);
        atomicAdd(&L.counters[2], 1ull);
        if (move) atomicAdd(&L.counters[3], (unsigned long long)s_nbranch * (unsigned long long)P);
      }
    }
    if (tid == 0) {
      double* hd = L.head + (size_t)ch * 8;
      hd[0] = s.scal[SC_LOGPZC];
      hd[1] = (double)acc_rates;
      hd[2] = (double)acc_zb;
    }
  }
  if (tid == 0 && L.counters)
    for (int k = 0; k < 12; k++) atomicAdd(&L.counters[4 + k], (unsigned long long)s_ph[k]);
}

size_t chain_smem_bytes(int n_max, int slots_max, int codes_cap) {
  const size_t uni_d = uni_doubles(n_max, slots_max);
  size_t b = ((size_t)n_max * 16 + (size_t)(n_max / 2 + 1) * 3 + uni_d + 34 + (size_t)n_max * 18 + 96 + SC_N + (size_t)n_max * 4 + 6 + (n_max + 1) / 2 + (codes_cap + 7) / 8) * 8;
  b += ((size_t)n_max * 13 + 4) * 2 + (size_t)n_max * 8;  // thirteen int16 arrays (two with 2 spare entries), eight byte arrays
  return (b + 63) & ~size_t(63);
}

}  // namespace

size_t gbgc_chain_scratch_stride(int n_max, int P_max) {
  const size_t d = (size_t)n_max * 16 + 3 * (size_t)n_max * 80 + 2 * (size_t)n_max * 16 + (size_t)kNW * n_max * 16 + (size_t)n_max * P_max * 4;
  return ((d * 8 + (size_t)n_max * P_max * 2) + 255) & ~size_t(255);
}

// One launch of the chain kernel = one WAVE of chains, asynchronous: gbgc_chains_launch uploads the wave's inputs, launches the kernel on the
// wave's own stream and returns; gbgc_chains_wait blocks until the wave is done and copies its results to the host (traces [n][5][T], states
// [n][T][n_max] in the local node order of each element, heads [n][8]; counters are ADDED to counters_out).  The caller keeps two waves in
// flight so that the host-side summaries of one wave run under the kernel of the next (Engine::gbgc_run).
int gbgc_chains_launch(int device, int num_sms, const uint8_t* d_blobs, const int64_t* d_blob_off, int N, int n_max, int slots_max, int P_max, const double* pi,
                       const double* Q0, int n_chains, const int32_t* chain_elem, const int32_t* chain_resz, const int8_t* fixZ, const int8_t* upper,
                       const pa_gbgc_params* prm, GbgcWave* w) {
  // `w` may come back from an earlier wave of the same run: its stream, events and device buffers are reused when they are large enough
  // (cudaMalloc / cudaFree synchronise the device and would serialise the waves)
  w->n_chains = 0;
  if (n_chains <= 0) return PA_OK;
  cudaError_t err = cudaSetDevice(device);
  const size_t nc = (size_t)n_chains;
  const int T = prm->num_burn + prm->num_mcmc;
  // tip codes of one element: at most (n_max + 1) / 2 rows of P_max / 2 bytes; staged in shared memory up to 12 KB
  const int codes_cap = (int)std::min<size_t>(12288, ((size_t)((n_max + 1) / 2) * (size_t)(P_max / 2) + 15) & ~size_t(15));
  const size_t smem = chain_smem_bytes(n_max, slots_max, codes_cap);
  int per_sm = 0;
  if (err == cudaSuccess) err = cudaFuncSetAttribute(gbgc_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err == cudaSuccess) err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gbgc_chain_kernel, kCBD, smem);
  if (err == cudaSuccess && per_sm < 1) {
    set_error("gbgc_chains: the tree does not fit the chain kernel's shared memory");
    return PA_ERR_ARG;
  }
  int grid = (int)std::min<size_t>(nc, (size_t)std::max(1, num_sms) * std::max(1, per_sm));
  const size_t stride = gbgc_chain_scratch_stride(n_max, P_max);
  {  // one very long element inflates every CTA's scratch: fewer resident CTAs rather than an out-of-memory failure (three wave slots share the device)
    size_t free_b = 0, total_b = 0;
    if (err == cudaSuccess && !w->d_scr && cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && stride * (size_t)grid > free_b / 5)
      grid = (int)std::max<size_t>(1, (free_b / 5) / stride);
    if (w->d_scr && w->cap_scr < stride * (size_t)grid) grid = (int)std::max<size_t>(1, w->cap_scr / stride);
  }
  const size_t b_in = nc * 8 + 2 * nc * N + 64, b_tr = nc * 5 * T * 8, b_zb = nc * T * n_max, b_head = nc * 8 * 8, b_ctr = 18 * 8;
  w->n_chains = n_chains; w->b_tr = b_tr; w->b_zb = b_zb; w->b_head = b_head;
  if (err == cudaSuccess && !w->stream) err = cudaStreamCreateWithFlags(&w->stream, cudaStreamNonBlocking);
  auto ensure = [&](unsigned char*& p, size_t& cap, size_t need) {
    if (err != cudaSuccess || (p && cap >= need)) return;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    err = cudaMalloc(&p, need);
    if (err == cudaSuccess) cap = need;
  };
  ensure(w->d_in, w->cap_in, b_in);
  ensure(w->d_scr, w->cap_scr, stride * grid);
  ensure(w->d_out, w->cap_out, b_tr + b_head + b_ctr + b_zb);
  GbgcChainLaunch L;
  L.blobs = d_blobs; L.blob_off = d_blob_off; L.n_chains = n_chains; L.N = N; L.n_max = n_max; L.slots_max = slots_max; L.P_max = P_max; L.T = T;
  L.codes_cap = codes_cap;
  int32_t* di = reinterpret_cast<int32_t*>(w->d_in);
  L.chain_elem = di; L.chain_resz = di + nc;
  int8_t* d8 = reinterpret_cast<int8_t*>(di + 2 * nc);
  L.fixZ = d8; L.upper = d8 + nc * N;
  L.prm = *prm;
  for (int k = 0; k < 4; k++) L.pi[k] = pi[k];
  for (int k = 0; k < 16; k++) L.Q0[k] = Q0[k];
  L.scratch = w->d_scr; L.stride = stride;
  L.tr = reinterpret_cast<double*>(w->d_out);
  L.head = L.tr + nc * 5 * T;
  L.counters = reinterpret_cast<unsigned long long*>(L.head + nc * 8);
  L.queue = reinterpret_cast<int*>(L.counters + 16);
  L.tr_zb = reinterpret_cast<uint8_t*>(L.counters + 18);
  w->dev_tr = L.tr; w->dev_head = L.head; w->dev_ctr = L.counters; w->dev_zb = L.tr_zb;
  auto up = [&](const void* dst, const void* src, size_t bytes) {
    if (err == cudaSuccess) err = cudaMemcpyAsync(const_cast<void*>(dst), src, bytes, cudaMemcpyHostToDevice, w->stream);
  };
  up(L.chain_elem, chain_elem, nc * 4); up(L.chain_resz, chain_resz, nc * 4); up(L.fixZ, fixZ, nc * N); up(L.upper, upper, nc * N);
  if (err == cudaSuccess) err = cudaMemsetAsync(w->d_out, 0, b_tr + b_head + b_ctr, w->stream);
  if (err == cudaSuccess && !w->e0) err = cudaEventCreate(&w->e0);
  if (err == cudaSuccess && !w->e1) err = cudaEventCreate(&w->e1);
  if (err == cudaSuccess) {
    cudaEventRecord(w->e0, w->stream);
    gbgc_chain_kernel<<<grid, kCBD, smem, w->stream>>>(L);
    err = cudaGetLastError();
    cudaEventRecord(w->e1, w->stream);
  }
  if (err != cudaSuccess) {
    gbgc_chains_release(w);
    set_error(std::string("gbgc_chains: ") + cudaGetErrorString(err));
    return PA_ERR_CUDA;
  }
  return PA_OK;
}

int gbgc_chains_wait(GbgcWave* w, double* tr_out, uint8_t* zb_out, double* head_out, unsigned long long* counters_out) {
  if (w->n_chains <= 0) return PA_OK;
  cudaError_t err = cudaStreamSynchronize(w->stream);
  // plain copies into the caller's (pageable) arrays: they run on the copy engine while the next wave's kernel computes
  if (err == cudaSuccess) err = cudaMemcpy(tr_out, w->dev_tr, w->b_tr, cudaMemcpyDeviceToHost);
  if (err == cudaSuccess) err = cudaMemcpy(head_out, w->dev_head, w->b_head, cudaMemcpyDeviceToHost);
  if (err == cudaSuccess) err = cudaMemcpy(zb_out, w->dev_zb, w->b_zb, cudaMemcpyDeviceToHost);
  if (err == cudaSuccess && counters_out) {
    unsigned long long c[16];
    err = cudaMemcpy(c, w->dev_ctr, sizeof(c), cudaMemcpyDeviceToHost);
    if (err == cudaSuccess)
      for (int k = 0; k < 16; k++) counters_out[k] += c[k];
  }
  if (err != cudaSuccess) {
    set_error(std::string("gbgc_chains: ") + cudaGetErrorString(err));
    return PA_ERR_CUDA;
  }
  return PA_OK;
}

void gbgc_chains_release(GbgcWave* w) {
  if (w->e0) cudaEventDestroy(w->e0);
  if (w->e1) cudaEventDestroy(w->e1);
  if (w->d_in) cudaFree(w->d_in);
  if (w->d_scr) cudaFree(w->d_scr);
  if (w->d_out) cudaFree(w->d_out);
  if (w->stream) cudaStreamDestroy(w->stream);
  *w = GbgcWave();
}

}  // namespace pa
