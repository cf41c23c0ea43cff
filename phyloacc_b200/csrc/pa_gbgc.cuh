// pa_gbgc.cuh -- device building blocks of the GC-biased model (PhyloAcc-ST-GBGC): the rate matrix Q o f(s + b B) and P(t) = exp(Q t)
// by uniformisation + squaring (see pa_gbgc.cu).  Shared by the batched transition kernel and the GBGC replay kernel.
#ifndef PA_GBGC_CUH
#define PA_GBGC_CUH
#include <cuda_runtime.h>

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

}  // namespace pa
#endif
