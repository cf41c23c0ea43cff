// pa_gbgc.cu -- PhyloAcc-ST-GBGC groundwork: the per-branch GC-biased transition matrices on the device.
//
// The GC-biased model has no shared eigendecomposition: every (element, branch, proposal) has its own rate matrix
//   Q' = Q o f(s + B(b)),  f(x) = x / (1 - exp(-x)),  B = +b for weak -> strong, -b for strong -> weak   (BPP_C::getRate, GBGC/SRC/bpp_c.cpp:13-30)
// and the reference solves a general 4 x 4 eigenproblem per matrix (BPP_C::getTransition, GBGC/SRC/bpp_c.cpp:1405-1427).
// Here one thread per matrix computes P(t) = exp(Q' t) by uniformisation + squaring: with mu = max_i |Q'_ii| and A = I + Q' / mu
// (a stochastic matrix), P(tau) = sum_j Poisson(j; mu tau) A^j for tau = t / 2^k small enough that mu tau <= 1/2, then k squarings.
// Every term is non-negative, so there is no cancellation: small entries keep their relative accuracy, log P is finite wherever the
// reference's is, and nothing depends on the spectrum being real.  fp64 throughout; about 1.5 k FMA per matrix.
#include <cuda_runtime.h>

#include "pa_gbgc.cuh"
#include "pa_internal.h"

namespace pa {

namespace {

__global__ void __launch_bounds__(128) gbgc_transition_kernel(const double* __restrict__ Q0g, int n, const double* __restrict__ sel, const double* __restrict__ bias,
                                                             const double* __restrict__ t, double* __restrict__ rate_out, double* __restrict__ P_out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  double Q0[16], Q[16];
#pragma unroll
  for (int k = 0; k < 16; k++) Q0[k] = Q0g[k];
  gbgc_rate(Q0, sel[idx], bias[idx], Q);
  if (rate_out)
#pragma unroll
    for (int k = 0; k < 16; k++) rate_out[(size_t)idx * 16 + k] = Q[k];
  double P[16];
  expm_uniformised(Q, t[idx], P);
#pragma unroll
  for (int k = 0; k < 16; k++) P_out[(size_t)idx * 16 + k] = P[k];
}

}  // namespace

// host entry: Q0 [16] row-major (rate i -> j), n matrices; sel / bias / t are host arrays of length n; outputs [n][16] row-major
// (P[i][j] = P(i -> j | t)); rate_out may be null
int gbgc_transition(int device, const double* Q0, int n, const double* sel, const double* bias, const double* t, double* rate_out, double* P_out) {
  if (n <= 0) return PA_OK;
  cudaError_t err = cudaSetDevice(device);
  double *d_in = nullptr, *d_out = nullptr;
  const size_t in_bytes = sizeof(double) * (16 + (size_t)3 * n), out_bytes = sizeof(double) * (size_t)32 * n;
  if (err == cudaSuccess) err = cudaMalloc(&d_in, in_bytes);
  if (err == cudaSuccess) err = cudaMalloc(&d_out, out_bytes);
  if (err == cudaSuccess) err = cudaMemcpy(d_in, Q0, sizeof(double) * 16, cudaMemcpyHostToDevice);
  if (err == cudaSuccess) err = cudaMemcpy(d_in + 16, sel, sizeof(double) * n, cudaMemcpyHostToDevice);
  if (err == cudaSuccess) err = cudaMemcpy(d_in + 16 + n, bias, sizeof(double) * n, cudaMemcpyHostToDevice);
  if (err == cudaSuccess) err = cudaMemcpy(d_in + 16 + 2 * (size_t)n, t, sizeof(double) * n, cudaMemcpyHostToDevice);
  if (err == cudaSuccess) {
    gbgc_transition_kernel<<<(n + 127) / 128, 128>>>(d_in, n, d_in + 16, d_in + 16 + n, d_in + 16 + 2 * (size_t)n, d_out, d_out + (size_t)16 * n);
    err = cudaGetLastError();
  }
  if (err == cudaSuccess) err = cudaDeviceSynchronize();
  if (err == cudaSuccess && rate_out) err = cudaMemcpy(rate_out, d_out, sizeof(double) * 16 * (size_t)n, cudaMemcpyDeviceToHost);
  if (err == cudaSuccess) err = cudaMemcpy(P_out, d_out + (size_t)16 * n, sizeof(double) * 16 * (size_t)n, cudaMemcpyDeviceToHost);
  if (d_in) cudaFree(d_in);
  if (d_out) cudaFree(d_out);
  if (err != cudaSuccess) {
    set_error(std::string("gbgc_transition: ") + cudaGetErrorString(err));
    return PA_ERR_CUDA;
  }
  return PA_OK;
}

}  // namespace pa
