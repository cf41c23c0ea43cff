// pa_summary.cu -- per-chain summaries on the device: what the reference computes from the traces after Gibbs
// never leaves HBM as a trace; one CTA per chain reduces its 6 x T doubles + T packed Z rows to one record of a few hundred bytes:
//
//   MonitorChain   trace_full_loglik = loglik + log Gamma-pdf(c) [+ log Gamma-pdf(n)], arg-max after burn-in     bpp_c.cpp:768-814
//   Output_init    medians of the five FULL rate traces (element T/2 of the sorted trace), frequencies of
//                  Z = 0 / 1 / 2 per node over the post-burn-in rows                                          bpp_c2.cpp:742-781
//   Eval2          uncertain nodes, clustering of the post-burn-in rows into Z configurations (walked from the
//                  last row down, configuration 0 = Max_Z), per-configuration moments, prior of Z, log-det,
//                  the weighted estimate                                                                      bpp_c2.cpp:14-197,318-349
//
// Rows are clustered serially by one warp in the reference's order, so the per-configuration sums are accumulated in the same
// order as on the host; the medians are exact order statistics (radix select on the 64-bit patterns, any T).  What stays on the
// host: the rare Eval2 fallback that needs pruning at the posterior-mean rates (bpp_c2.cpp:199-316, through replay_kernel), chains
// with more distinct configurations than the table holds, and VERBOSE trace output -- for those the chain's trace is downloaded.
#include <cuda_runtime.h>
#include <math_constants.h>

#include "pa_device.cuh"
#include "pa_summary.cuh"

namespace pa {

namespace {

constexpr int kBD = 256;
constexpr int kCap = SUMMARY_MAX_CONFIGS;  // configurations the shared-memory table holds
constexpr int kRowTile = 128;              // trace rows staged per step of the clustering
constexpr int kTab = 2 * SUMMARY_MAX_CONFIGS;  // slots of the configuration hash table (power of two)

__device__ __forceinline__ double lnbeta_dev(double a, double b) { return lgamma(a) + lgamma(b) - lgamma(a + b); }

// k-th smallest (0-based) of T non-negative doubles in global memory: radix select over the 64-bit patterns (non-negative IEEE
// doubles order like unsigned integers), eight 8-bit passes, all threads of the CTA; `hist` is 256 ints of shared memory
__device__ double radix_select(const double* __restrict__ x, int T, int k, int* hist, unsigned long long* s_prefix, int tid) {
  unsigned long long prefix = 0ull, mask = 0ull;
  for (int shift = 56; shift >= 0; shift -= 8) {
    for (int i = tid; i < 256; i += kBD) hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < T; i += kBD) {
      unsigned long long key = (unsigned long long)__double_as_longlong(x[i]);
      if (key >> 63) key = 0ull;  // -0.0 / negative (never produced by the chain): order as zero
      if ((key & mask) == prefix) atomicAdd(&hist[(int)((key >> shift) & 0xffull)], 1);
    }
    __syncthreads();
    if (tid < 32) {
      // the bin where the running count passes k: every lane owns 8 consecutive bins, a warp prefix sum finds the lane, the lane its bin
      int local[8], sum = 0;
#pragma unroll
      for (int j = 0; j < 8; j++) { local[j] = hist[tid * 8 + j]; sum += local[j]; }
      int incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (tid >= o) incl += v;
      }
      const int excl = incl - sum;
      const unsigned hit = __ballot_sync(0xffffffffu, incl > k);
      const int owner = __ffs(hit) - 1;  // T > k, so some lane's inclusive count exceeds k
      if (tid == owner) {
        int acc = excl, b = 0;
        for (; b < 7; b++) {
          if (acc + local[b] > k) break;
          acc += local[b];
        }
        s_prefix[0] = prefix | ((unsigned long long)(tid * 8 + b) << shift);
        s_prefix[1] = (unsigned long long)(k - acc);
      }
    }
    __syncthreads();
    prefix = s_prefix[0];
    k = (int)s_prefix[1];
    mask |= 0xffull << shift;
    __syncthreads();
  }
  return __longlong_as_double((long long)prefix);
}

__global__ void __launch_bounds__(kBD) summary_kernel(SummaryLaunch L) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ch = blockIdx.x;
  if (ch >= L.n_chains) return;
  const int T = L.T, ZW = L.ZW, burn = L.num_burn, num_mcmc = L.num_mcmc;
  const int elem = L.chain_list[2 * ch], mdl = L.chain_list[2 * ch + 1];
  const ElemView ev = make_view(L.blobs + L.blob_off[elem]);
  const int n = ev.h->n, root = n - 1;
  const bool null_model = (mdl == PA_M0);
  const double* __restrict__ td = L.trace_d + (size_t)ch * 6 * T;  // loglik, n, c, g, l, l2
  const uint32_t* __restrict__ tz = L.trace_z + (size_t)ch * T * ZW;
  unsigned char* rec = L.out + (size_t)ch * L.rec_bytes;
  SummaryHead* head = reinterpret_cast<SummaryHead*>(rec);
  uint32_t* out_maxz = reinterpret_cast<uint32_t*>(rec + sizeof(SummaryHead));
  uint32_t* out_lastz = out_maxz + ZW;
  uint16_t* out_cnt = reinterpret_cast<uint16_t*>(out_lastz + ZW);  // [n_max][3]

  // shared-memory carve-up
  double* s_red = reinterpret_cast<double*>(smem_raw);                  // [kBD] arg-max values
  int* s_redi = reinterpret_cast<int*>(s_red + kBD);                    // [kBD] arg-max rows
  int* s_hist = s_redi + kBD;                                           // [256]
  unsigned long long* s_pref = reinterpret_cast<unsigned long long*>(s_hist + 256);  // [2]
  double* s_acc = reinterpret_cast<double*>(s_pref + 2);                // [kCap][6]: sum c^2, c, n^2, n, c n, full / num_mcmc
  int* s_num = reinterpret_cast<int*>(s_acc + kCap * 6);                // [kCap]
  int* s_first = s_num + kCap;                                          // [kCap] trace row of the configuration's representative
  int* s_iv = s_first + kCap;                                           // [8] scalars
  uint32_t* s_mask = reinterpret_cast<uint32_t*>(s_iv + 8);             // [ZW] 2-bit fields of the uncertain nodes
  uint32_t* s_maxz = s_mask + ZW;                                       // [ZW]
  uint32_t* s_cfg = s_maxz + ZW;                                        // [kCap][ZW] representative rows
  uint32_t* s_rows = s_cfg + (size_t)kCap * ZW;                         // [kRowTile][ZW] staged trace rows
  double* s_full = reinterpret_cast<double*>(s_rows + (size_t)kRowTile * ZW + ((kRowTile * ZW) & 1));  // [kRowTile]
  double* s_rc = s_full + kRowTile;                                     // [kRowTile] c
  double* s_rn = s_rc + kRowTile;                                       // [kRowTile] n
  int* s_cnt = reinterpret_cast<int*>(s_rn + kRowTile);                 // [n][3] Z frequencies
  uint16_t* s_tab = reinterpret_cast<uint16_t*>(s_cnt + (size_t)L.n_max_pad * 3);  // [kTab] configuration index + 1 by hash slot, 0 = empty

  const double lg_ca = lgamma(L.cprior_a), lg_na = lgamma(L.nprior_a);
  auto full_of = [&](int m) {  // trace_full_loglik (bpp_c.cpp:775-806); density first, then log, as the reference evaluates it
    const double c = td[2 * (size_t)T + m];
    double f = td[m] + 0.0;
    {
      double dens;
      if (c < 0) dens = 0;
      else if (c == 0) dens = (L.cprior_a == 1) ? 1 / L.cprior_b : 0;
      else if (L.cprior_a == 1) dens = exp(-c / L.cprior_b) / L.cprior_b;
      else dens = exp((L.cprior_a - 1) * log(c / L.cprior_b) - c / L.cprior_b - lg_ca) / L.cprior_b;
      f += log(dens);
    }
    if (!null_model) {
      const double x = td[(size_t)T + m];
      double dens;
      if (x < 0) dens = 0;
      else if (x == 0) dens = (L.nprior_a == 1) ? 1 / L.nprior_b : 0;
      else if (L.nprior_a == 1) dens = exp(-x / L.nprior_b) / L.nprior_b;
      else dens = exp((L.nprior_a - 1) * log(x / L.nprior_b) - x / L.nprior_b - lg_na) / L.nprior_b;
      f += log(dens);
    }
    return f;
  };

  // ---------------- MonitorChain: arg-max of the full log-likelihood after burn-in (first maximum wins, bpp_c.cpp:808-814)
  {
    double best = -CUDART_INF;
    int best_m = -1;
    for (int m = burn + tid; m < T; m += kBD) {
      const double f = full_of(m);
      if (f > best) { best = f; best_m = m; }
    }
    s_red[tid] = best;
    s_redi[tid] = best_m;
    __syncthreads();
    for (int o = kBD / 2; o > 0; o >>= 1) {
      if (tid < o) {
        const double a = s_red[tid], b = s_red[tid + o];
        const int ia = s_redi[tid], ib = s_redi[tid + o];
        // larger value wins; among equal values the EARLIER row (the reference keeps the first maximum it meets)
        if (ib >= 0 && (ia < 0 || b > a || (b == a && ib < ia))) { s_red[tid] = b; s_redi[tid] = ib; }
      }
      __syncthreads();
    }
  }
  const int max_m = s_redi[0];
  const double max_ll = s_red[0];
  __syncthreads();
  for (int w = tid; w < ZW; w += kBD) {
    const uint32_t word = (max_m >= 0) ? tz[(size_t)max_m * ZW + w] : 0u;
    s_maxz[w] = word;
    s_mask[w] = 0u;
    out_maxz[w] = word;
    out_lastz[w] = tz[(size_t)(T - 1) * ZW + w];
  }
  for (int i = tid; i < n * 3; i += kBD) s_cnt[i] = 0;
  if (tid < 8) s_iv[tid] = 0;
  __syncthreads();

  // ---------------- Output_init: medians of the sorted FULL traces (bpp_c2.cpp:742-760)
  double med[5];
  for (int k = 0; k < 5; k++) med[k] = radix_select(td + (size_t)(k + 1) * T, T, T / 2, s_hist, s_pref, tid);

  // ---------------- uncertain nodes (a post-burn-in row differs from Max_Z there) and Z frequencies per node (bpp_c2.cpp:757-781)
  {
    // thread = (row stripe, word): OR of the differing 2-bit fields
    for (int w = 0; w < ZW; w++) {
      uint32_t diff = 0u;
      const uint32_t mz = s_maxz[w];
      for (int m = burn + tid; m < T; m += kBD) {
        const uint32_t x = tz[(size_t)m * ZW + w] ^ mz;
        diff |= (x | (x >> 1)) & 0x55555555u;
      }
      diff |= diff << 1;
      for (int o = 16; o > 0; o >>= 1) diff |= __shfl_xor_sync(0xffffffffu, diff, o);
      if (lane == 0 && diff) atomicOr(&s_mask[w], diff);
    }
    // frequencies: warp = row stripe, lane = node (+32, +64, ...); private counters, one shared atomic per (lane, node, state)
    for (int i0 = 0; i0 < n; i0 += 32) {
      const int i = i0 + lane;
      int c0 = 0, c1 = 0, c2 = 0;
      if (i < n) {
        const int w = i >> 4, sh = 2 * (i & 15);
        for (int m = burn + warp; m < T; m += kBD / 32) {
          const uint32_t z = (tz[(size_t)m * ZW + w] >> sh) & 3u;
          c0 += (z == 0u); c1 += (z == 1u); c2 += (z == 2u);
        }
        atomicAdd(&s_cnt[i * 3], c0);
        atomicAdd(&s_cnt[i * 3 + 1], c1);
        atomicAdd(&s_cnt[i * 3 + 2], c2);
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < n * 3; i += kBD) out_cnt[i] = (uint16_t)min(s_cnt[i], 65535);

  // ---------------- Eval2 (bpp_c2.cpp:14-197): clustering of the post-burn-in rows, last row first
  double logml = CUDART_NAN;
  int need_host = 0;
  if (max_m < 0 || num_mcmc > 65535) need_host = (max_m >= 0);
  else {
    if (tid == 0) {
      s_num[0] = 0;
      s_first[0] = max_m;
      for (int k = 0; k < 6; k++) s_acc[k] = 0.0;
      s_iv[0] = 1;  // configurations so far
    }
    for (int w = tid; w < ZW; w += kBD) s_cfg[w] = s_maxz[w];
    for (int k = tid; k < kTab; k += kBD) s_tab[k] = 0;
    __syncthreads();
    if (warp == 0) {  // configuration 0 = Max_Z goes into the table under its own hash
      uint32_t hv = 0u;
      for (int w = lane; w < ZW; w += 32) hv ^= (s_maxz[w] & s_mask[w]) * (0x9E3779B1u * (2u * (uint32_t)w + 1u));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) hv ^= __shfl_xor_sync(0xffffffffu, hv, o);
      hv ^= hv >> 15;
      if (lane == 0) s_tab[hv & (uint32_t)(kTab - 1)] = 1;
    }
    __syncthreads();
    for (int hi = T; hi > burn && !s_iv[1]; hi -= kRowTile) {
      const int lo = max(burn, hi - kRowTile), cnt = hi - lo;
      // stage rows lo .. hi-1 (and their full log-likelihood / rates) cooperatively
      for (int idx = tid; idx < cnt * ZW; idx += kBD) s_rows[idx] = tz[(size_t)lo * ZW + idx];
      for (int r = tid; r < cnt; r += kBD) {
        s_full[r] = full_of(lo + r);
        s_rc[r] = td[2 * (size_t)T + lo + r];
        s_rn[r] = td[(size_t)T + lo + r];
      }
      __syncthreads();
      if (warp == 0) {
        int ncfg = s_iv[0];
        for (int r = cnt - 1; r >= 0; r--) {
          const uint32_t* row = s_rows + (size_t)r * ZW;
          // look the row up by the hash of its uncertain nodes' states (open addressing, <= 50 % load): the cost of a row does not
          // grow with the number of configurations; indices are still handed out in order of first appearance
          uint32_t hv = 0u;
          for (int w = lane; w < ZW; w += 32) hv ^= (row[w] & s_mask[w]) * (0x9E3779B1u * (2u * (uint32_t)w + 1u));
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) hv ^= __shfl_xor_sync(0xffffffffu, hv, o);
          hv ^= hv >> 15;
          uint32_t slot = hv & (uint32_t)(kTab - 1);
          int j = -1;
          for (;;) {
            const int idx = s_tab[slot];
            if (idx == 0) break;
            const uint32_t* cf = s_cfg + (size_t)(idx - 1) * ZW;
            bool eq = true;
            for (int w = lane; w < ZW; w += 32) eq = eq && (((row[w] ^ cf[w]) & s_mask[w]) == 0u);
            if (__all_sync(0xffffffffu, eq)) { j = idx - 1; break; }
            slot = (slot + 1u) & (uint32_t)(kTab - 1);
          }
          if (j < 0) {
            if (ncfg >= kCap) {  // more configurations than the table holds: the host summarises this chain
              if (lane == 0) s_iv[1] = 1;
              break;
            }
            j = ncfg++;
            for (int w = lane; w < ZW; w += 32) s_cfg[(size_t)j * ZW + w] = row[w];
            if (lane == 0) {
              s_tab[slot] = (uint16_t)(j + 1);
              s_num[j] = 0;
              s_first[j] = lo + r;
              for (int k = 0; k < 6; k++) s_acc[j * 6 + k] = 0.0;
            }
            __syncwarp();
          }
          if (lane == 0) {  // the reference's accumulation order (bpp_c2.cpp:96-109)
            const double c = s_rc[r], nr = s_rn[r];
            // separately rounded products and sums (no FMA contraction): the sample covariance below cancels catastrophically for
            // near-constant rates, so the host's / the reference's rounding sequence is part of the result
            s_num[j] += 1;
            s_acc[j * 6 + 0] = __dadd_rn(s_acc[j * 6 + 0], __dmul_rn(c, c));  // pow(x, 2) == x * x exactly
            s_acc[j * 6 + 1] = __dadd_rn(s_acc[j * 6 + 1], c);
            if (!null_model) {
              s_acc[j * 6 + 2] = __dadd_rn(s_acc[j * 6 + 2], __dmul_rn(nr, nr));
              s_acc[j * 6 + 3] = __dadd_rn(s_acc[j * 6 + 3], nr);
              s_acc[j * 6 + 4] = __dadd_rn(s_acc[j * 6 + 4], __dmul_rn(c, nr));
            }
            s_acc[j * 6 + 5] = __dadd_rn(s_acc[j * 6 + 5], __ddiv_rn(s_full[r], (double)num_mcmc));
          }
          __syncwarp();
        }
        if (lane == 0) s_iv[0] = ncfg;
      }
      __syncthreads();
    }
    if (s_iv[1]) need_host = 1;
    else {
      // per configuration with weight >= 2 %: prior of Z (bpp_c2.cpp:131-163), log-det, its term of the estimate; thread = configuration
      const int ncfg = s_iv[0];
      double* s_term = s_full;  // reuse: [kRowTile] >= terms of one pass over <= kRowTile configurations
      double VB = 0.0;
      int effective_sum = 0;
      for (int c0 = 0; c0 < ncfg; c0 += kRowTile) {
        const int c = c0 + tid;
        if (tid < kRowTile) {
          double term = 0.0;
          int eff = 0;
          if (c < ncfg) {
            const int num = s_num[c];
            const double w = (double)num / num_mcmc;
            if (w >= 0.02) {
              const uint32_t* cf = s_cfg + (size_t)c * ZW;
              double nZ[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, nF[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
              for (int i = 0; i < n - 1; i++) {
                const int p = ev.parent[i];
                const int zi = (cf[i >> 4] >> (2 * (i & 15))) & 3u, zp = (cf[p >> 4] >> (2 * (p & 15))) & 3u;
                const int k = zi * 3 + zp;
                if (zi < 3 && zp < 3) {
                  nZ[k] += 1;
                  if (!((ev.flags[i] >> (2 + mdl)) & 1)) nF[k] += 1;
                }
              }
              double pz = lnbeta_dev(L.prior_g_a + nZ[3] + nZ[6], L.prior_g_b + nZ[0]) - lnbeta_dev(L.prior_g_a, L.prior_g_b);
              pz += lnbeta_dev(L.prior_l_a + nF[7], L.prior_l_b + nF[4]) - lnbeta_dev(L.prior_l_a, L.prior_l_b);
              if (L.prior_l2_a > 0) pz += lnbeta_dev(L.prior_l2_a + nF[5], L.prior_l2_b + nF[8]) - lnbeta_dev(L.prior_l2_a, L.prior_l2_b);
              const int zr = (cf[root >> 4] >> (2 * (root & 15))) & 3u;
              pz += (zr < 2) ? log(0.5) : -CUDART_INF;
              const double dnum = (double)num;
              const double sum_full = __dadd_rn(s_acc[c * 6 + 5], __ddiv_rn(__dmul_rn(dnum, pz), (double)num_mcmc));
              const double vc = __dsub_rn(s_acc[c * 6 + 0], __ddiv_rn(__dmul_rn(s_acc[c * 6 + 1], s_acc[c * 6 + 1]), dnum));
              double logdet;
              if (null_model) logdet = __dsub_rn(log(vc), log(dnum));
              else {
                const double vn = __dsub_rn(s_acc[c * 6 + 2], __ddiv_rn(__dmul_rn(s_acc[c * 6 + 3], s_acc[c * 6 + 3]), dnum));
                const double cv = __dsub_rn(s_acc[c * 6 + 4], __ddiv_rn(__dmul_rn(s_acc[c * 6 + 3], s_acc[c * 6 + 1]), dnum));
                logdet = __dsub_rn(log(__dsub_rn(__dmul_rn(vc, vn), __dmul_rn(cv, cv))), __dmul_rn(2.0, log(dnum)));
              }
              if (logdet == logdet) {  // NaN log-dets are skipped (bpp_c2.cpp:177)
                eff = num;
                // VB -= w log(num); VB += sum_full; VB += 0.5 logdet num / num_mcmc  -- summed by thread 0 in configuration order
                term = 0.0;
                s_acc[c * 6 + 0] = -__dmul_rn(w, log(dnum));
                s_acc[c * 6 + 1] = sum_full;
                s_acc[c * 6 + 2] = __ddiv_rn(__dmul_rn(__dmul_rn(0.5, logdet), dnum), (double)num_mcmc);
              }
            }
          }
          s_term[tid] = term;
          s_redi[tid] = eff;
        }
        __syncthreads();
        if (tid == 0) {
          for (int k = 0; k < kRowTile && c0 + k < ncfg; k++) {
            if (!s_redi[k]) continue;
            effective_sum += s_redi[k];
            VB = __dadd_rn(VB, s_acc[(c0 + k) * 6 + 0]);
            VB = __dadd_rn(VB, s_acc[(c0 + k) * 6 + 1]);
            VB = __dadd_rn(VB, s_acc[(c0 + k) * 6 + 2]);
          }
        }
        __syncthreads();
      }
      if (tid == 0) {
        if (effective_sum == 0) s_iv[2] = 1;  // Eval2's pruning fallback (bpp_c2.cpp:199-316): host + replay kernel
        else {
          const double sc = __dmul_rn(VB, __ddiv_rn((double)num_mcmc, (double)effective_sum));
          const double fin = null_model ? __dadd_rn(__dadd_rn(sc, log((double)num_mcmc)), __dmul_rn(0.5, __dadd_rn(log(2 * CUDART_PI), 1.0)))
                                        : __dadd_rn(__dadd_rn(__dadd_rn(sc, log((double)num_mcmc)), log(2 * CUDART_PI)), 1.0);
          s_red[0] = fin;
        }
      }
      __syncthreads();
      if (s_iv[2]) need_host = 1;
      else logml = s_red[0];
    }
  }
  if (tid == 0) {
    head->max_loglik = max_ll;
    head->logml = logml;
    for (int k = 0; k < 5; k++) head->med[k] = med[k];
    for (int k = 0; k < 5; k++) head->last[k] = td[(size_t)(k + 1) * T + (T - 1)];
    head->max_m = max_m;
    head->flags = need_host ? SUMMARY_NEEDS_HOST : 0;
    head->n_configs = s_iv[0];
    head->pad_ = 0;
  }
}

}  // namespace

size_t summary_record_bytes(int n_max, int ZW) {
  size_t b = sizeof(SummaryHead) + (size_t)2 * ZW * 4 + (size_t)n_max * 3 * 2;
  return (b + 15) & ~size_t(15);
}

static size_t summary_smem_bytes(int n_max, int ZW) {
  size_t b = (size_t)kBD * 8 + (size_t)kBD * 4 + 256 * 4 + 16;                       // red, redi, hist, prefix
  b += (size_t)kCap * 6 * 8 + (size_t)kCap * 4 * 2 + 8 * 4;                            // acc, num, first, iv
  b += (size_t)2 * ZW * 4 + (size_t)kCap * ZW * 4 + (size_t)kRowTile * ZW * 4 + 8;     // mask, maxz, cfg, rows
  b += (size_t)3 * kRowTile * 8 + (size_t)n_max * 3 * 4 + (size_t)kTab * 2 + 16;       // full, c, n, cnt, hash table
  return (b + 63) & ~size_t(63);
}

int launch_summaries(const SummaryLaunch& L, int n_max, cudaStream_t stream) {
  if (L.n_chains == 0) return PA_OK;
  const size_t smem = summary_smem_bytes(n_max, L.ZW);
  cudaError_t err = cudaFuncSetAttribute(summary_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) {
    set_error(std::string("summary kernel shared memory: ") + cudaGetErrorString(err));
    return PA_ERR_LIMIT;
  }
  summary_kernel<<<L.n_chains, kBD, smem, stream>>>(L);
  err = cudaGetLastError();
  if (err != cudaSuccess) {
    set_error(std::string("summary kernel launch: ") + cudaGetErrorString(err));
    return PA_ERR_CUDA;
  }
  return PA_OK;
}

}  // namespace pa
