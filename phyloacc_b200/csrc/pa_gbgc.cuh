// pa_gbgc.cuh -- device building blocks of the GC-biased model (PhyloAcc-ST-GBGC): the rate matrix Q o f(s + b B) and P(t) = exp(Q t)
// by uniformisation + squaring (see pa_gbgc.cu).  Shared by the batched transition kernel and the GBGC replay kernel.
#ifndef PA_GBGC_CUH
#define PA_GBGC_CUH
#include <cuda_runtime.h>
#include <math_constants.h>

namespace pa {

__device__ __forceinline__ void matmul4(const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ c) {
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) c[i * 4 + j] = ((a[i * 4] * b[j] + a[i * 4 + 1] * b[4 + j]) + a[i * 4 + 2] * b[8 + j]) + a[i * 4 + 3] * b[12 + j];
}

// Q' of BPP_C::getRate, same expression order: off-diagonals scaled, then the row sums of the WHOLE matrix subtracted from the diagonal
__device__ __forceinline__ void gbgc_rate(const double* __restrict__ Q0, double s, double b, double* __restrict__ Q) {
  if (s == 0.0 && b == 0.0) {
#pragma unroll
    for (int k = 0; k < 16; k++) Q[k] = Q0[k];
    return;
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const bool si = (i == 1 || i == 2);  // strong bases C, G
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const bool sj = (j == 1 || j == 2);
      const double B = s + ((!si && sj) ? b : ((si && !sj) ? -b : 0.0));
      const double f = (B == 0.0) ? 1.0 : B / (1.0 - exp(-B));
      Q[i * 4 + j] = Q0[i * 4 + j] * f;
    }
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const double c = ((Q[i * 4] + Q[i * 4 + 1]) + Q[i * 4 + 2]) + Q[i * 4 + 3];
    Q[i * 4 + i] -= c;
  }
}


// P(t) = exp(Q t) for a rate matrix Q (rows sum to zero, off-diagonals >= 0), row-major, P[i][j] = P(i -> j): uniformisation with
// mu = max_i |Q_ii|, A = I + Q / mu, P(tau) = sum_j Poisson(j; mu tau) A^j for tau = t / 2^k with mu tau <= 1/2, then k squarings.
// Every term is non-negative: no cancellation, small entries keep their relative accuracy.
__device__ __forceinline__ void expm_uniformised(const double* __restrict__ Q, double tt, double* __restrict__ P) {
  const double mu = fmax(fmax(-Q[0], -Q[5]), fmax(-Q[10], -Q[15]));
  if (!(mu > 0.0) || !(tt > 0.0)) {
#pragma unroll
    for (int k = 0; k < 16; k++) P[k] = ((k >> 2) == (k & 3)) ? 1.0 : 0.0;
    return;
  }
  int ksq = 0;
  double x = mu * tt;
  while (x > 0.5 && ksq < 60) { x *= 0.5; ksq++; }
  double A[16], term[16], nxt[16];
#pragma unroll
  for (int k = 0; k < 16; k++) A[k] = Q[k] / mu + (((k >> 2) == (k & 3)) ? 1.0 : 0.0);
  const double e = exp(-x);
#pragma unroll
  for (int k = 0; k < 16; k++) { term[k] = ((k >> 2) == (k & 3)) ? e : 0.0; P[k] = term[k]; }
  for (int j = 1; j < 20; j++) {  // x <= 1/2: the Poisson tail beyond 19 terms is below 1e-24
    matmul4(term, A, nxt);
    const double w = x / j;
#pragma unroll
    for (int k = 0; k < 16; k++) { term[k] = nxt[k] * w; P[k] += term[k]; }
  }
  for (int q = 0; q < ksq; q++) {
    matmul4(P, P, nxt);
#pragma unroll
    for (int k = 0; k < 16; k++) P[k] = nxt[k];
  }
}

// a / b for a normal positive b and any finite a: hardware reciprocal seed (rcp.approx.ftz.f64, ~2^-23), two Newton steps and one residual
// correction -- about 9 instructions with no slow-path branch, within an ulp of the IEEE quotient.  Callers guarantee the range (partials
// and transition probabilities are rescaled into [2^-256, 1] before they get here).
__device__ __forceinline__ double div_fast(double a, double b) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  r = fma(fma(-b, r, 1.0), r, r);
  r = fma(fma(-b, r, 1.0), r, r);
  const double q = a * r;
  return fma(fma(-b, q, a), r, q);
}

// log(x) for a normal x > 0 (x == 0 gives -inf): x = 2^e m with m in [sqrt(1/2), sqrt(2)), log m = 2 atanh(s), s = (m - 1) / (m + 1),
// |s| <= 0.1716, odd series to s^23 (truncation < 1e-19), log 2 split in two words.  ~2.5e-16 relative against a 50-digit reference over
// (1e-26, 1]; a third of the instructions of the library routine, which also serves subnormals, negatives and NaN.  Used for the 16 entries
// of every log P(t) of the GBGC chain (the dominant transcendental of that kernel).
__device__ __forceinline__ double log_fast(double x) {
  int hi = __double2hiint(x);
  const int lo = __double2loint(x);
  int e = (hi >> 20) - 1023;
  hi = (hi & 0x000fffff) | 0x3ff00000;
  if (hi >= 0x3ff6a09f) { hi -= 0x00100000; e += 1; }
  const double f = __hiloint2double(hi, lo) - 1.0;
  const double sq = div_fast(f, 2.0 + f), s2 = sq * sq;
  double p = 2.0 / 23.0;
  p = fma(p, s2, 2.0 / 21.0); p = fma(p, s2, 2.0 / 19.0); p = fma(p, s2, 2.0 / 17.0); p = fma(p, s2, 2.0 / 15.0); p = fma(p, s2, 2.0 / 13.0);
  p = fma(p, s2, 2.0 / 11.0); p = fma(p, s2, 2.0 / 9.0); p = fma(p, s2, 2.0 / 7.0); p = fma(p, s2, 2.0 / 5.0); p = fma(p, s2, 2.0 / 3.0);
  const double lm = fma(s2 * sq, p, 2.0 * sq), ed = (double)e;
  const double r = fma(ed, 0x1.62e42fefa38p-1, fma(ed, 0x1.ef35793c7673p-45, lm));
  return x == 0.0 ? -CUDART_INF : r;
}

}  // namespace pa
#endif
