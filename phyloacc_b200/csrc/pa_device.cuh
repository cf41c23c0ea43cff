// pa_device.cuh -- device-side building blocks shared by the replay and chain kernels (sm_100a).
#ifndef PA_DEVICE_CUH
#define PA_DEVICE_CUH

#include <cuda_runtime.h>
#include <math_constants.h>

#include "pa_internal.h"

namespace pa {

// ---------------------------------------------------------------- Philox4x32-10 draw schedule (same as oracle/oracle.c)
__device__ __forceinline__ uint4 philox4(uint32_t seed, uint32_t key, uint32_t idx, uint32_t sub, uint32_t iter, uint32_t stream) {
  uint32_t c0 = idx, c1 = sub, c2 = iter, c3 = stream, k0 = seed, k1 = key;
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}

// (w + 0.5) * 2^-32, exactly: the 32 random bits become the top of the mantissa of a double in [1, 2), and 1 - 2^-33 is subtracted (both
// operands and the difference have at most 33 significant bits, so the one DADD is exact and the value equals the oracle's
// ((double)w + 0.5) * 2^-32).  Replaces an integer-to-double conversion, an add and a multiply per draw.
__device__ __forceinline__ double u01(uint32_t w) {
  return __hiloint2double((int)(0x3FF00000u | (w >> 12)), (int)(w << 20)) - 0.99999999988358467817306518554688;
}

__device__ __forceinline__ uint32_t pick(const uint4& w, int k) { return k == 0 ? w.x : (k == 1 ? w.y : (k == 2 ? w.z : w.w)); }

// first index k with u*total < cumsum_k (weights summed in index order), else the last positive weight
__device__ __forceinline__ int categorical4(double p0, double p1, double p2, double p3, double u) {
  const double tot = ((p0 + p1) + p2) + p3;
  const double x = u * tot;
  double acc = p0;
  if (x < acc) return 0;
  acc += p1;
  if (x < acc) return 1;
  acc += p2;
  if (x < acc) return 2;
  acc += p3;
  if (x < acc) return 3;
  return p3 > 0 ? 3 : (p2 > 0 ? 2 : (p1 > 0 ? 1 : (p0 > 0 ? 0 : -1)));
}
__device__ __forceinline__ int categorical3(double p0, double p1, double p2, double u) {
  const double tot = (p0 + p1) + p2;
  const double x = u * tot;
  double acc = p0;
  if (x < acc) return 0;
  acc += p1;
  if (x < acc) return 1;
  acc += p2;
  if (x < acc) return 2;
  return p2 > 0 ? 2 : (p1 > 0 ? 1 : (p0 > 0 ? 0 : -1));
}

// Marsaglia-Tsang gamma(shape a, scale b) on the draw schedule (see pa_or_gamma in oracle/oracle.c)
__device__ inline double sched_gamma(uint32_t seed, uint32_t key, uint32_t idx, uint32_t sub0, uint32_t iter, uint32_t stream, double a, double b) {
  uint4 w = philox4(seed, key, idx, sub0, iter, stream);
  double boost = 1.0;
  if (a < 1.0) {
    boost = pow(u01(w.w), 1.0 / a);
    a += 1.0;
  }
  const double d = a - 1.0 / 3.0;
  const double c = (1.0 / 3.0) / sqrt(d);
  for (uint32_t attempt = 0;; attempt++) {
    if (attempt > 0) w = philox4(seed, key, idx, sub0 + attempt, iter, stream);
    if (attempt > 1000) return CUDART_NAN;  // non-finite parameters: never spin
    const double x = sqrt(-2.0 * log(u01(w.x))) * cos(6.283185307179586 * u01(w.y));
    double v = 1.0 + c * x;
    if (v <= 0) continue;
    v = v * v * v;
    const double u = u01(w.z);
    if (u < 1 - 0.0331 * x * x * x * x) return b * d * v * boost;
    if (log(u) < 0.5 * x * x + d * (1 - v + log(v))) return b * d * v * boost;
  }
}

__device__ inline double sched_beta(uint32_t seed, uint32_t key, uint32_t which, uint32_t iter, double a, double b) {
  const double x1 = sched_gamma(seed, key, which, 0, iter, PA_ST_TRANS, a, 1.0);
  const double x2 = sched_gamma(seed, key, which, 4096, iter, PA_ST_TRANS, b, 1.0);
  return x1 / (x1 + x2);
}

// log(gsl_ran_gamma_pdf(x, a, b)), evaluated the way the reference does (density first, then log)
__device__ inline double log_gamma_pdf(double x, double a, double b) {
  double dens;
  if (x < 0) dens = 0;
  else if (x == 0) dens = (a == 1) ? 1 / b : 0;
  else if (a == 1) dens = exp(-x / b) / b;
  else dens = exp((a - 1) * log(x / b) - x / b - lgamma(a)) / b;
  return log(dens);
}

// ---------------------------------------------------------------- P(t) from the shared eigensystem
// P[j][i] = P(j -> i | t * rate) = sum_m exp((eval[m] * t) * rate) * A[m][j][i]   (bpp.hpp:375-380, transposed)
// t <= 1e-10 is the identity (bpp.hpp:373,388-391; bpp_c.cpp:222-226).
__device__ __forceinline__ void build_P(const SubstDev& sd, double t, double rate, bool identity_if_tiny, double* __restrict__ P16) {
  if (identity_if_tiny && !(t > 1e-10)) {
#pragma unroll
    for (int k = 0; k < 16; k++) P16[k] = ((k >> 2) == (k & 3)) ? 1.0 : 0.0;
    return;
  }
  const double e0 = exp(sd.eval[0] * t * rate), e1 = exp(sd.eval[1] * t * rate), e2 = exp(sd.eval[2] * t * rate), e3 = exp(sd.eval[3] * t * rate);
#pragma unroll
  for (int k = 0; k < 16; k++) P16[k] = ((e0 * sd.A[0][k] + e1 * sd.A[1][k]) + e2 * sd.A[2][k]) + e3 * sd.A[3][k];
}

// one row j of P(t * rate): r_i = P(j -> i); same arithmetic as build_P
__device__ __forceinline__ void build_P_row(const SubstDev& sd, double t, double rate, int j, double& r0, double& r1, double& r2, double& r3) {
  if (!(t > 1e-10)) {
    r0 = (j == 0) ? 1.0 : 0.0; r1 = (j == 1) ? 1.0 : 0.0; r2 = (j == 2) ? 1.0 : 0.0; r3 = (j == 3) ? 1.0 : 0.0;
    return;
  }
  const double e0 = exp(sd.eval[0] * t * rate), e1 = exp(sd.eval[1] * t * rate), e2 = exp(sd.eval[2] * t * rate), e3 = exp(sd.eval[3] * t * rate);
  const int k = j * 4;
  r0 = ((e0 * sd.A[0][k] + e1 * sd.A[1][k]) + e2 * sd.A[2][k]) + e3 * sd.A[3][k];
  r1 = ((e0 * sd.A[0][k + 1] + e1 * sd.A[1][k + 1]) + e2 * sd.A[2][k + 1]) + e3 * sd.A[3][k + 1];
  r2 = ((e0 * sd.A[0][k + 2] + e1 * sd.A[1][k + 2]) + e2 * sd.A[2][k + 2]) + e3 * sd.A[3][k + 2];
  r3 = ((e0 * sd.A[0][k + 3] + e1 * sd.A[1][k + 3]) + e2 * sd.A[2][k + 3]) + e3 * sd.A[3][k + 3];
}

// build_P_row for four adjacent lanes that hold rows 0..3 of the same branch (lane & 3 == j, all four converged): each lane
// evaluates one of the four exponentials and the quad exchanges them -- same expressions, same values as build_P_row
__device__ __forceinline__ void build_P_row_quad(const SubstDev& sd, double t, double rate, int j, double& r0, double& r1, double& r2, double& r3) {
  if (!(t > 1e-10)) {
    r0 = (j == 0) ? 1.0 : 0.0; r1 = (j == 1) ? 1.0 : 0.0; r2 = (j == 2) ? 1.0 : 0.0; r3 = (j == 3) ? 1.0 : 0.0;
    return;
  }
  // explicit quad mask: the four lanes of a quad take the same branch at every call site (same branch of the tree, rows 0..3), but
  // they need not be CONVERGED there under independent thread scheduling -- __activemask() could miss a source lane
  const unsigned mask = 0xFu << (threadIdx.x & 28u);
  const double ej = exp(sd.eval[j] * t * rate);
  const int base = (int)(threadIdx.x & 28u);
  const double e0 = __shfl_sync(mask, ej, base), e1 = __shfl_sync(mask, ej, base + 1), e2 = __shfl_sync(mask, ej, base + 2), e3 = __shfl_sync(mask, ej, base + 3);
  const int k = j * 4;
  r0 = ((e0 * sd.A[0][k] + e1 * sd.A[1][k]) + e2 * sd.A[2][k]) + e3 * sd.A[3][k];
  r1 = ((e0 * sd.A[0][k + 1] + e1 * sd.A[1][k + 1]) + e2 * sd.A[2][k + 1]) + e3 * sd.A[3][k + 1];
  r2 = ((e0 * sd.A[0][k + 2] + e1 * sd.A[1][k + 2]) + e2 * sd.A[2][k + 2]) + e3 * sd.A[3][k + 2];
  r3 = ((e0 * sd.A[0][k + 3] + e1 * sd.A[1][k + 3]) + e2 * sd.A[2][k + 3]) + e3 * sd.A[3][k + 3];
}

// 256-bit global accesses (sm_100: LDG.E.256 / STG.E.256): one partial = four doubles = one 32-byte sector, moved by one instruction
__device__ __forceinline__ void ldg256(const double* p, double2& u, double2& v) {
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(u.x), "=d"(u.y), "=d"(v.x), "=d"(v.y) : "l"(p));
}
__device__ __forceinline__ void stg256(double* p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

// ---------------------------------------------------------------- element view (pointers into the blob)
struct ElemView {
  const ElemHeader* h;
  const int16_t* parent;
  const int16_t* child;
  const uint8_t* flags;
  const int16_t* tiprow;
  const int16_t* orig;
  const double* dist;
  const int16_t* intlist;
  const int16_t* intidx;
  const int16_t* su;
  const int8_t* slot;
  const int16_t* up_nodes;
  const int16_t* up_off;
  const int16_t* down_nodes;
  const int16_t* down_off;
  const uint8_t* codes;
  const uint16_t* weights;
  const uint32_t* gap;
  const uint16_t* sitepat;
  const uint4* tipplanes;
  const uint32_t* tipamb;
};

__device__ __forceinline__ ElemView make_view(const uint8_t* blob) {
  ElemView v;
  const ElemHeader* h = reinterpret_cast<const ElemHeader*>(blob);
  v.h = h;
  v.parent = reinterpret_cast<const int16_t*>(blob + h->off_parent);
  v.child = reinterpret_cast<const int16_t*>(blob + h->off_child);
  v.flags = blob + h->off_flags;
  v.tiprow = reinterpret_cast<const int16_t*>(blob + h->off_tiprow);
  v.orig = reinterpret_cast<const int16_t*>(blob + h->off_orig);
  v.dist = reinterpret_cast<const double*>(blob + h->off_dist);
  v.intlist = reinterpret_cast<const int16_t*>(blob + h->off_int);
  v.intidx = reinterpret_cast<const int16_t*>(blob + h->off_intidx);
  v.su = reinterpret_cast<const int16_t*>(blob + h->off_su);
  v.slot = reinterpret_cast<const int8_t*>(blob + h->off_slot);
  v.up_nodes = reinterpret_cast<const int16_t*>(blob + h->off_up_nodes);
  v.up_off = reinterpret_cast<const int16_t*>(blob + h->off_up_off);
  v.down_nodes = reinterpret_cast<const int16_t*>(blob + h->off_down_nodes);
  v.down_off = reinterpret_cast<const int16_t*>(blob + h->off_down_off);
  v.codes = blob + h->off_codes;
  v.weights = reinterpret_cast<const uint16_t*>(blob + h->off_weights);
  v.gap = reinterpret_cast<const uint32_t*>(blob + h->off_gap);
  v.sitepat = reinterpret_cast<const uint16_t*>(blob + h->off_sitepat);
  v.tipplanes = reinterpret_cast<const uint4*>(blob + h->off_tipplanes);
  v.tipamb = reinterpret_cast<const uint32_t*>(blob + h->off_tipamb);
  return v;
}

// ---------------------------------------------------------------- Felsenstein pruning of one pattern (linear space, fp64)
// Tree arrays live in shared memory (copied once per chain); partials of recomputed ("dirty") nodes go to a
// per-thread temp stack in shared memory laid out [slot][base][thread]; partials of clean nodes are read from the
// chain's global arena [internal index q][pattern][4].  Rescaling by 2^256 keeps deep trees away from underflow;
// the cumulative exponent count of a subtree travels with its partial.
#define PA_SCALE_UP 1.157920892373162e77     // 2^256
#define PA_SCALE_THRESH 8.636168555094445e-78  // 2^-256
#define PA_LOG_SCALE 177.445678223345999      // 256 * ln 2

struct TreeSmem {   // per-chain tree arrays in shared memory
  int16_t* child;   // [n][2]
  uint8_t* flags;   // [n]
  int16_t* tiprow;  // [n]
  int16_t* su;      // [ni]
  int8_t* slot;     // [n]
  int16_t* intidx;  // [n]
};

// LDG_CODES = false: `codes` may point to shared memory (a staged copy), so it is read with a plain generic load
template <bool READ_G, bool WRITE_G, bool LDG_CODES = true>
__device__ __forceinline__ double prune_pattern(const ElemHeader* h, const TreeSmem& ts, const uint8_t* __restrict__ codes, const double* __restrict__ Pm,
                                                const uint8_t* __restrict__ dirty, double* __restrict__ stack, int* __restrict__ stack_sc,
                                                double* __restrict__ gpart, int16_t* __restrict__ gsc, const double* __restrict__ pi, int p, int BD,
                                                int tid) {
  const int ni = h->ni, Ppad = h->Ppad, half = Ppad >> 1;
  double root_l[4] = {1.0, 1.0, 1.0, 1.0};
  int root_sc = 0;
  bool root_done = false;
  for (int k = 0; k < ni; k++) {
    const int s = ts.su[k];
    if (s < 0) break;
    if (dirty != nullptr && !dirty[s]) continue;
    double out[4] = {1.0, 1.0, 1.0, 1.0};
    int sc = 0;
#pragma unroll
    for (int cc = 0; cc < 2; cc++) {
      const int ch = ts.child[2 * s + cc];
      const uint8_t f = ts.flags[ch];
      if (f & PA_F_MISSING) continue;  // bpp_c.cpp:660
      const double* Pc = Pm + ch * 16;
      if (f & PA_F_TIP) {
        const uint8_t* cp = codes + (size_t)ts.tiprow[ch] * half + (p >> 1);
        const uint32_t byte = LDG_CODES ? __ldg(cp) : *cp;
        const uint32_t code = (byte >> ((p & 1) * 4)) & 15u;
        if (code == 15u) continue;  // gap / N: Tg >= num_base (bpp_c.cpp:660)
        if ((code & (code - 1)) == 0) {
          const int b = __ffs(code) - 1;
#pragma unroll
          for (int j = 0; j < 4; j++) out[j] *= Pc[j * 4 + b];
        } else {
#pragma unroll
          for (int j = 0; j < 4; j++) {
            double m = 0;
#pragma unroll
            for (int b = 0; b < 4; b++) m += ((code >> b) & 1u) ? Pc[j * 4 + b] : 0.0;
            out[j] *= m;
          }
        }
      } else {
        double L[4];
        if (!READ_G || dirty == nullptr || dirty[ch]) {
          const int sl = ts.slot[ch];
#pragma unroll
          for (int i = 0; i < 4; i++) L[i] = stack[(sl * 4 + i) * BD + tid];
          sc += stack_sc[sl * BD + tid];
        } else {
          const int q = ts.intidx[ch];
          const double2* g2 = reinterpret_cast<const double2*>(gpart + ((size_t)q * Ppad + p) * 4);
          const double2 a = g2[0], b2 = g2[1];
          L[0] = a.x; L[1] = a.y; L[2] = b2.x; L[3] = b2.y;
          sc += gsc[(size_t)q * Ppad + p];
        }
#pragma unroll
        for (int j = 0; j < 4; j++) out[j] *= ((Pc[j * 4 + 0] * L[0] + Pc[j * 4 + 1] * L[1]) + Pc[j * 4 + 2] * L[2]) + Pc[j * 4 + 3] * L[3];
      }
    }
    const double mx = fmax(fmax(out[0], out[1]), fmax(out[2], out[3]));
    if (mx < PA_SCALE_THRESH && mx > 0.0) {
#pragma unroll
      for (int j = 0; j < 4; j++) out[j] *= PA_SCALE_UP;
      sc += 1;
    }
    const int sl = ts.slot[s];
#pragma unroll
    for (int j = 0; j < 4; j++) stack[(sl * 4 + j) * BD + tid] = out[j];
    stack_sc[sl * BD + tid] = sc;
    if (WRITE_G) {
      const int q = ts.intidx[s];
      double2* g2 = reinterpret_cast<double2*>(gpart + ((size_t)q * Ppad + p) * 4);
      g2[0] = make_double2(out[0], out[1]);
      g2[1] = make_double2(out[2], out[3]);
      gsc[(size_t)q * Ppad + p] = (int16_t)sc;
    }
    if (s == h->n - 1) {
#pragma unroll
      for (int j = 0; j < 4; j++) root_l[j] = out[j];
      root_sc = sc;
      root_done = true;
    }
  }
  if (!root_done) {
    const int root = h->n - 1;
    if (!(ts.flags[root] & PA_F_MISSING) && READ_G) {  // clean root: read its stored partial
      const int q = ts.intidx[root];
      const double2* g2 = reinterpret_cast<const double2*>(gpart + ((size_t)q * Ppad + p) * 4);
      const double2 a = g2[0], b2 = g2[1];
      root_l[0] = a.x; root_l[1] = a.y; root_l[2] = b2.x; root_l[3] = b2.y;
      root_sc = gsc[(size_t)q * Ppad + p];
    }
  }
  const double lik = ((pi[0] * root_l[0] + pi[1] * root_l[1]) + pi[2] * root_l[2]) + pi[3] * root_l[3];
  return log(lik) - root_sc * PA_LOG_SCALE;
}

// deterministic block-wide sum (warp shuffles, then the warp partials added in order; the leading barrier protects `red` from
// the readers of a previous call)
__device__ __forceinline__ double block_sum(double v, double* red /* >= 33 doubles */, int tid, int BD) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  double s = 0;  // every thread adds the warp partials in the same order: one value, no third barrier
  for (int w = 0; w < (BD >> 5); w++) s += red[w];
  return s;
}

}  // namespace pa
#endif
