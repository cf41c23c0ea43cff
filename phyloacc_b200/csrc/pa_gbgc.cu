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

// ------------------------------------------------------------------------------------------------------------------------
// The random-number-free half of a PhyloAcc-ST-GBGC iteration at given states (groundwork): one CTA per state
//   * upward pass over compressed site patterns with every branch's exp(t Q o f(s + b B))      Update_Tg(g,.,.,false)  GBGC/SRC/bpp_c.cpp:477-528
//   * downward pass: expected (child base, parent base) pairs per branch, pattern-weighted      Update_Tg(g,.,.,true)   :530-583
//   * the six (Z, B) emissions of every branch, the 6-state upward messages through the 3 x 3
//     Z transition matrices, log_prob_back[root], the ELBO term                                 initial_prob_back       bpp_c2.cpp:354-407
// Linear space, fp64.  With L_s the upward partial of node s, m_s = P_s L_s and q_p the posterior of the parent's base, the joint of
// (child i, parent j) is q_p(j) P_s(j -> i) L_s(i) / m_s(j); the 2^256 rescaling of the partials cancels in that ratio.
#include "pa_device.cuh"

namespace pa {

namespace {

struct GbgcIntegLaunch {
  const uint8_t* blobs; const int64_t* blob_off;
  int n_states, N, n_max, slots_max, P_max;
  const int32_t* elem; const int8_t* Z; const int8_t* B; const int8_t* fixZ; const int8_t* upper;  // [n_states][N] each, original numbering
  const double* sel_n; const double* sel_c; const double* gbc;
  double grate, lrate, prior_gb, cons_to_mis, ncons_to_mis;
  double pi[4];
  double Q0[16];
  unsigned char* scratch; size_t stride;  // per state: logP6 [n_max][6][16] f64, partials [n_max][P_max][4] f64, scale counts [n_max][P_max] i16
  double* out_post;   // [n_states][N][16]  (child base, parent base), original numbering
  double* out_emis;   // [n_states][N][6]
  double* out_root;   // [n_states][8]: log_prob_back[root][0..5], ELBO term, upward log-likelihood at the current classes
};

constexpr int kIBD = 128;

__global__ void __launch_bounds__(kIBD) gbgc_integrate_kernel(GbgcIntegLaunch L) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int BD = kIBD, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int st = blockIdx.x;
  if (st >= L.n_states) return;
  const int64_t off = L.blob_off[L.elem[st]];
  if (off < 0) return;  // filtered element: outputs stay zero
  const ElemView ev = make_view(L.blobs + off);
  const ElemHeader* h = ev.h;
  const int n = h->n, ni = h->ni, N = L.N, n_max = L.n_max, slots_max = L.slots_max, root = n - 1;
  // shared-memory carve-up (cf. replay_kernel)
  double* Pm = reinterpret_cast<double*>(smem_raw);                     // [n_max][16] P(t) of every branch at its current class
  double* stack = Pm + (size_t)n_max * 16;                              // [slots_max][4][BD]
  double* red = stack + (size_t)slots_max * 4 * BD;                     // [34]
  double* post = red + 34;                                              // [n_max][16] expected pairs (child, parent)
  double* wpart = post + (size_t)n_max * 16;                            // [4 warps][16]
  double* emis = wpart + 4 * 16;                                        // [n_max][6]
  double* raw = emis + (size_t)n_max * 6;                               // [n_max] emission of the current class without the constants
  double* back = raw + n_max;                                           // [n_max][6]
  int* stack_sc = reinterpret_cast<int*>(back + (size_t)n_max * 6);     // [slots_max][BD]
  int16_t* s_child = reinterpret_cast<int16_t*>(stack_sc + (size_t)slots_max * BD);  // [n_max][2]
  int16_t* s_tiprow = s_child + 2 * n_max;
  int16_t* s_su = s_tiprow + n_max;
  int16_t* s_intidx = s_su + n_max;
  int16_t* s_parent = s_intidx + n_max;
  uint8_t* s_flags = reinterpret_cast<uint8_t*>(s_parent + n_max);
  int8_t* s_slot = reinterpret_cast<int8_t*>(s_flags + n_max);
  int8_t* s_cls = s_slot + n_max;                                       // current class k = Z + 3 B of every branch
  __shared__ double s_pi[4];
  unsigned char* sc = L.scratch + (size_t)st * L.stride;
  double* logP6 = reinterpret_cast<double*>(sc);                        // [n_max][6][16], row-major P(i -> j)
  double* gpart = logP6 + (size_t)n_max * 96;
  int16_t* gsc = reinterpret_cast<int16_t*>(gpart + (size_t)n_max * L.P_max * 4);

  for (int i = tid; i < n; i += BD) {
    s_child[2 * i] = ev.child[2 * i];
    s_child[2 * i + 1] = ev.child[2 * i + 1];
    s_tiprow[i] = ev.tiprow[i];
    s_intidx[i] = ev.intidx[i];
    s_parent[i] = ev.parent[i];
    s_flags[i] = ev.flags[i];
    s_slot[i] = ev.slot[i];
    const int o = ev.orig[i];
    const int z = L.Z[(size_t)st * N + o], b = L.B[(size_t)st * N + o] ? 1 : 0;
    s_cls[i] = (int8_t)(z + 3 * b);
  }
  for (int i = tid; i < ni; i += BD) s_su[i] = ev.su[i];
  for (int i = tid; i < n * 16; i += BD) post[i] = 0.0;
  if (tid < 4) s_pi[tid] = L.pi[tid];
  __syncthreads();
  // P(t) of the six classes of every branch: class k -> selection 0 / sel_c / sel_n by k % 3, bias gbc if k >= 3
  for (int task = tid; task < 6 * n; task += BD) {
    const int i = task / 6, k = task - 6 * i;
    const double sel = (k % 3 == 2) ? L.sel_n[st] : (k % 3 == 0 ? 0.0 : L.sel_c[st]);
    double Q[16], P16[16];
    gbgc_rate(L.Q0, sel, k >= 3 ? L.gbc[st] : 0.0, Q);
    expm_uniformised(Q, ev.dist[i], P16);
#pragma unroll
    for (int e = 0; e < 16; e++) logP6[((size_t)i * 6 + k) * 16 + e] = log(P16[e]);
    if (k == s_cls[i])
#pragma unroll
      for (int e = 0; e < 16; e++) Pm[i * 16 + e] = P16[e];
  }
  __syncthreads();
  // ---- upward pass, partials of every internal node kept in the scratch arena
  TreeSmem ts{s_child, s_flags, s_tiprow, s_su, s_slot, s_intidx};
  const int P = h->P, Ppad = h->Ppad;
  double acc = 0.0;
  for (int p = tid; p < Ppad; p += BD)
    if (p < P) acc += (double)ev.weights[p] * prune_pattern<false, true>(h, ts, ev.codes, Pm, nullptr, stack, stack_sc, gpart, gsc, s_pi, p, BD, tid);
  const double up_ll = block_sum(acc, red, tid, BD);
  __syncthreads();
  // ---- downward pass: parents before children = decreasing local index; an internal node's partial slot is overwritten by its posterior
  const int half = Ppad >> 1;
  for (int p0 = 0; p0 < Ppad; p0 += BD) {
    const int p = p0 + tid;
    const bool act = p < P;
    const double w = act ? (double)ev.weights[p] : 0.0;
    if (act && !(s_flags[root] & PA_F_MISSING)) {  // root: posterior of its base
      double* g = gpart + ((size_t)s_intidx[root] * Ppad + p) * 4;
      const double r0 = s_pi[0] * g[0], r1 = s_pi[1] * g[1], r2 = s_pi[2] * g[2], r3 = s_pi[3] * g[3];
      const double tot = ((r0 + r1) + r2) + r3;
      g[0] = r0 / tot; g[1] = r1 / tot; g[2] = r2 / tot; g[3] = r3 / tot;
    }
    for (int i = n - 2; i >= 0; i--) {
      if (s_flags[i] & PA_F_MISSING) continue;  // uniform over the CTA
      double J[16];
#pragma unroll
      for (int e = 0; e < 16; e++) J[e] = 0.0;
      if (act) {
        double Lv[4];
        double* gi = nullptr;
        if (s_flags[i] & PA_F_TIP) {
          const uint32_t byte = __ldg(ev.codes + (size_t)s_tiprow[i] * half + (p >> 1));
          const uint32_t code = (byte >> ((p & 1) * 4)) & 15u;
#pragma unroll
          for (int a = 0; a < 4; a++) Lv[a] = ((code >> a) & 1u) ? 1.0 : 0.0;
        } else {
          gi = gpart + ((size_t)s_intidx[i] * Ppad + p) * 4;
#pragma unroll
          for (int a = 0; a < 4; a++) Lv[a] = gi[a];
        }
        const double* Pi = Pm + i * 16;
        const double* qp = gpart + ((size_t)s_intidx[s_parent[i]] * Ppad + p) * 4;
        double qi[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int b = 0; b < 4; b++) {
          const double m = ((Pi[b * 4] * Lv[0] + Pi[b * 4 + 1] * Lv[1]) + Pi[b * 4 + 2] * Lv[2]) + Pi[b * 4 + 3] * Lv[3];
          const double ratio = (m > 0.0) ? qp[b] / m : 0.0;
#pragma unroll
          for (int a = 0; a < 4; a++) {
            const double j = ratio * Pi[b * 4 + a] * Lv[a];
            J[a * 4 + b] = j;  // [child base a][parent base b]
            qi[a] += j;
          }
        }
        if (gi) {
#pragma unroll
          for (int a = 0; a < 4; a++) gi[a] = qi[a];
        }
      }
      // pattern-weighted sum over the CTA, deterministic: warp shuffles, then the warps' partials in order
#pragma unroll
      for (int e = 0; e < 16; e++) {
        double v = J[e] * w;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) wpart[warp * 16 + e] = v;
      }
      __syncthreads();
      if (tid < 16) post[i * 16 + tid] += ((wpart[tid] + wpart[16 + tid]) + wpart[32 + tid]) + wpart[48 + tid];
      __syncthreads();
    }
  }
  // ---- emissions of the six (Z, B) classes (initial_prob_back, bpp_c2.cpp:354-389; missing branches: initMCMC, bpp_c.cpp:219-229)
  for (int task = tid; task < 6 * (n - 1); task += BD) {
    const int i = task / 6, k = task - 6 * i;
    const int o = ev.orig[i];
    const bool fix1 = L.fixZ[(size_t)st * N + o] == 1;
    double e = 0.0;
    if (s_flags[i] & PA_F_MISSING) {
      if (!(s_flags[s_parent[i]] & PA_F_MISSING)) e = (k % 3 == 1) ? log(L.cons_to_mis) : log(L.ncons_to_mis);
      if (fix1 && k % 3 == 2) e = -CUDART_INF;
    } else if (fix1 && k % 3 == 2) {
      e = -CUDART_INF;
    } else {
      const double* lp = logP6 + ((size_t)i * 6 + k) * 16;
      double s = 0.0;
      for (int a = 0; a < 4; a++)
        for (int b = 0; b < 4; b++) {
          const double c = post[i * 16 + a * 4 + b];
          if (c != 0.0) s += lp[b * 4 + a] * c;  // log P(b -> a) x expected pairs (child a, parent b)
        }
      if (k == s_cls[i]) raw[i] = s;
      e = s + (k < 3 ? log(1.0 - L.prior_gb) : log(L.prior_gb)) + (k % 3 == 1 ? log(1.0 - L.cons_to_mis) : log(1.0 - L.ncons_to_mis));
    }
    emis[i * 6 + k] = e;
    back[i * 6 + k] = e;
    L.out_emis[((size_t)st * N + o) * 6 + k] = e;
  }
  for (int i = tid; i < n - 1; i += BD) {
    const int o = ev.orig[i];
    const int k = s_cls[i];
    if ((s_flags[i] & PA_F_MISSING) || (L.fixZ[(size_t)st * N + o] == 1 && k % 3 == 2)) raw[i] = 0.0;
    for (int e = 0; e < 16; e++) L.out_post[((size_t)st * N + o) * 16 + e] = post[i * 16 + e];
  }
  if (tid < 6) back[root * 6 + tid] = 0.0;
  __syncthreads();
  // ---- 6-state upward messages (bpp_c2.cpp:391-405): serial over the post-order, one thread (84 steps of a 6 x 3 log-sum-exp)
  if (tid == 0) {
    double elbo = 0.0;
    for (int i = 0; i < n - 1; i++) {
      elbo += raw[i];
      const int o = ev.orig[i], par = s_parent[i];
      const bool up = L.upper[(size_t)st * N + o] != 0;
      // log TM_Int (child Z, parent Z): (0,0) = 1-g, (1,0) = g, (1,1) = 1-l (1 on the outgroup's ancestors), (2,1) = l (0 there), (2,2) = 1
      double lt[9];
      for (int e = 0; e < 9; e++) lt[e] = -CUDART_INF;
      lt[0] = log(1.0 - L.grate); lt[3] = log(L.grate);
      lt[4] = up ? 0.0 : log(1.0 - L.lrate); lt[7] = up ? -CUDART_INF : log(L.lrate);
      lt[8] = 0.0;
      double x[18], mx = -CUDART_INF;
      for (int c = 0; c < 6; c++)
        for (int j = 0; j < 3; j++) {
          x[c * 3 + j] = lt[(c % 3) * 3 + j] + back[i * 6 + c];
          if (x[c * 3 + j] > mx) mx = x[c * 3 + j];
        }
      for (int j = 0; j < 3; j++) {
        double msg = -CUDART_INF;
        if (mx > -CUDART_INF) {
          double s = 0.0;
          for (int c = 0; c < 6; c++) s += exp(x[c * 3 + j] - mx);
          msg = log(s) + mx;
        }
        back[par * 6 + j] += msg;
        back[par * 6 + 3 + j] += msg;
      }
    }
    for (int k = 0; k < 6; k++) L.out_root[(size_t)st * 8 + k] = back[root * 6 + k];
    L.out_root[(size_t)st * 8 + 6] = elbo;
    L.out_root[(size_t)st * 8 + 7] = up_ll;
  }
}

size_t gbgc_integ_smem(int n_max, int slots_max) {
  size_t b = (size_t)n_max * 16 * 8 + (size_t)slots_max * 4 * kIBD * 8 + 34 * 8 + (size_t)n_max * 16 * 8 + 64 * 8 + (size_t)n_max * 6 * 8 + (size_t)n_max * 8 +
             (size_t)n_max * 6 * 8 + (size_t)slots_max * kIBD * 4;
  b += (size_t)n_max * (2 + 1 + 1 + 1 + 1) * 2 + (size_t)n_max * 3;
  return (b + 63) & ~size_t(63);
}

}  // namespace

int gbgc_integrate(int device, cudaStream_t stream, const uint8_t* d_blobs, const int64_t* d_blob_off, int N, int n_max, int slots_max, int P_max, const double* pi,
                   const double* Q0, int n_states, const int32_t* elem, const int8_t* Z, const int8_t* B, const int8_t* fixZ, const int8_t* upper,
                   const double* sel_n, const double* sel_c, const double* gbc, double grate, double lrate, double prior_gb, double cons_to_mis,
                   double ncons_to_mis, double* post_out, double* emis_out, double* root_out) {
  if (n_states <= 0) return PA_OK;
  cudaError_t err = cudaSetDevice(device);
  const size_t ns = (size_t)n_states;
  const size_t stride = (((size_t)n_max * 96 * 8 + (size_t)n_max * P_max * 4 * 8 + (size_t)n_max * P_max * 2) + 255) & ~size_t(255);
  const size_t b_in = ns * 4 + 4 * ns * N + 3 * ns * 8 + 64, b_out = ns * N * 16 * 8 + ns * N * 6 * 8 + ns * 8 * 8;
  unsigned char *d_in = nullptr, *d_out = nullptr, *d_scr = nullptr;
  if (err == cudaSuccess) err = cudaMalloc(&d_in, b_in);
  if (err == cudaSuccess) err = cudaMalloc(&d_out, b_out);
  if (err == cudaSuccess) err = cudaMalloc(&d_scr, stride * ns);
  GbgcIntegLaunch L;
  L.blobs = d_blobs; L.blob_off = d_blob_off; L.n_states = n_states; L.N = N; L.n_max = n_max; L.slots_max = slots_max; L.P_max = P_max;
  double* dd = reinterpret_cast<double*>(d_in);
  L.sel_n = dd; L.sel_c = dd + ns; L.gbc = dd + 2 * ns;
  int32_t* de = reinterpret_cast<int32_t*>(dd + 3 * ns);
  L.elem = de;
  int8_t* d8 = reinterpret_cast<int8_t*>(de + ns);
  L.Z = d8; L.B = d8 + ns * N; L.fixZ = d8 + 2 * ns * N; L.upper = d8 + 3 * ns * N;
  L.grate = grate; L.lrate = lrate; L.prior_gb = prior_gb; L.cons_to_mis = cons_to_mis; L.ncons_to_mis = ncons_to_mis;
  for (int k = 0; k < 4; k++) L.pi[k] = pi[k];
  for (int k = 0; k < 16; k++) L.Q0[k] = Q0[k];
  L.scratch = d_scr; L.stride = stride;
  L.out_post = reinterpret_cast<double*>(d_out);
  L.out_emis = L.out_post + ns * N * 16;
  L.out_root = L.out_emis + ns * N * 6;
  auto up = [&](const void* dst, const void* src, size_t bytes) {
    if (err == cudaSuccess) err = cudaMemcpyAsync(const_cast<void*>(dst), src, bytes, cudaMemcpyHostToDevice, stream);
  };
  up(L.sel_n, sel_n, ns * 8); up(L.sel_c, sel_c, ns * 8); up(L.gbc, gbc, ns * 8); up(L.elem, elem, ns * 4);
  up(L.Z, Z, ns * N); up(L.B, B, ns * N); up(L.fixZ, fixZ, ns * N); up(L.upper, upper, ns * N);
  if (err == cudaSuccess) err = cudaMemsetAsync(d_out, 0, b_out, stream);
  const size_t smem = gbgc_integ_smem(n_max, slots_max);
  if (err == cudaSuccess) err = cudaFuncSetAttribute(gbgc_integrate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err == cudaSuccess) {
    gbgc_integrate_kernel<<<n_states, kIBD, smem, stream>>>(L);
    err = cudaGetLastError();
  }
  if (err == cudaSuccess) err = cudaMemcpyAsync(post_out, L.out_post, ns * N * 16 * 8, cudaMemcpyDeviceToHost, stream);
  if (err == cudaSuccess) err = cudaMemcpyAsync(emis_out, L.out_emis, ns * N * 6 * 8, cudaMemcpyDeviceToHost, stream);
  if (err == cudaSuccess) err = cudaMemcpyAsync(root_out, L.out_root, ns * 8 * 8, cudaMemcpyDeviceToHost, stream);
  if (err == cudaSuccess) err = cudaStreamSynchronize(stream);
  if (d_in) cudaFree(d_in);
  if (d_out) cudaFree(d_out);
  if (d_scr) cudaFree(d_scr);
  if (err != cudaSuccess) {
    set_error(std::string("gbgc_integrate: ") + cudaGetErrorString(err));
    return PA_ERR_CUDA;
  }
  return PA_OK;
}

}  // namespace pa
