// pa_gbgc_host.cpp -- host side of the PhyloAcc-ST-GBGC chain: per-model constraints (BPP_C::initMCMC,
// GBGC/SRC/bpp_c.cpp:143-180), the launch (pa_gbgc_chain.cu), and what the reference computes from a finished chain: MonitorChain's arg-max
// (bpp_c.cpp:655-664), Output_init's summaries (bpp_c2.cpp:603-650) and the variant's Eval2 (variational bound over the configurations of Z
// and, within each, of B; bpp_c2.cpp:29-285).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <thread>
#include <vector>

#include "pa_engine.h"
#include "pa_internal.h"
#include "pa_nvtx.h"

extern "C" void pa_gbgc_params_default(pa_gbgc_params* p) {  // GBGC/SRC/main.cpp:50-86, bpp.hpp:201, bpp.cpp:217-219, bpp_c.hpp:358-361
  memset(p, 0, sizeof(*p));
  p->sel_c0 = -2; p->sel_n0 = 0; p->gbc0 = 3;
  p->cprior_a = 5; p->cprior_b = 0.4;
  p->nprior_a = 5; p->nprior_b = 0.4; p->nprior_c = 2;
  p->gbprior_a = 4; p->gbprior_b = 0.8; p->gbprior_c = 0.2;
  p->cons_to_mis = 0.01; p->ncons_to_mis = 0.1;
  p->grate = 0.8; p->lrate = 0.1;
  p->prop_c = 1.2; p->prop_n = 1.2; p->prop_gb = 0.2;
  p->num_burn = 200; p->num_mcmc = 800;
  p->seed = 7728;
}

namespace pa {

namespace {

// log|det| and sign of a d x d matrix (d <= 3) by LU with partial pivoting (arma::log_det)
void log_det_small(const double* A, int d, double& logdet, double& sign) {
  double M[9];
  for (int k = 0; k < d * d; k++) M[k] = A[k];
  sign = 1;
  logdet = 0;
  for (int c = 0; c < d; c++) {
    int piv = c;
    for (int r = c + 1; r < d; r++)
      if (std::fabs(M[r * d + c]) > std::fabs(M[piv * d + c])) piv = r;
    if (piv != c) {
      for (int k = 0; k < d; k++) std::swap(M[c * d + k], M[piv * d + k]);
      sign = -sign;
    }
    const double u = M[c * d + c];
    if (u == 0) { logdet = -INFINITY; sign = 0; return; }
    if (u < 0) sign = -sign;
    logdet += std::log(std::fabs(u));
    for (int r = c + 1; r < d; r++) {
      const double f = M[r * d + c] / u;
      for (int k = c; k < d; k++) M[r * d + k] -= f * M[c * d + k];
    }
  }
}

struct GbgcChainView {
  int T, N, burn, mcmc, resz;
  const double *loglik, *full, *c, *n, *gb;  // [T]
  const int8_t* Z;                            // [T][N] original numbering, -1 pruned
  const int8_t* B;                            // [T][N]
  const ElemPlan* plan;
};

// BPP_C::Output_init of the GBGC variant (bpp_c2.cpp:603-650): the tail of each rate trace is sorted and element T / 2 of the WHOLE trace is
// reported as the "median"; P(B = 1) and P(Z = -1, 0, 1, 2) per node from the post-burn-in samples, with P(Z = -1) set to 1 for a missing node
// on top of whatever its samples counted.  med = {n_rate, c_rate, gBC}.
void gbgc_output_init(int T, int N, int burn, const double* n_rate, const double* c_rate, const double* gbc, const int8_t* Z, const int8_t* B,
                      const uint8_t* missing, double* med3, double* post_b, double* post_z) {
  const int mcmc = T - burn;
  auto med = [&](const double* x) {
    std::vector<double> v(x, x + T);
    std::sort(v.begin() + burn, v.end());
    return v[T / 2];
  };
  med3[0] = med(n_rate);
  med3[1] = med(c_rate);
  med3[2] = med(gbc);
  std::vector<int> cz((size_t)N * 4, 0), cb(N, 0);  // row by row: the traces are [T][N]
  for (int m = burn; m < T; m++) {
    const int8_t* zr = Z + (size_t)m * N;
    const int8_t* br = B + (size_t)m * N;
    for (int s = 0; s < N; s++) {
      cz[(size_t)s * 4 + zr[s] + 1]++;
      cb[s] += br[s];
    }
  }
  for (int s = 0; s < N; s++) {
    if (missing[s]) cz[(size_t)s * 4] = mcmc;
    for (int q = 0; q < 4; q++) post_z[(size_t)s * 4 + q] = (double)cz[(size_t)s * 4 + q] / mcmc;
    post_b[s] = (double)cb[s] / mcmc;
  }
}

// BPP_C::Eval2 of the GBGC variant (bpp_c2.cpp:29-285)
double gbgc_eval2(const GbgcChainView& v, const int8_t* maxZ, double max_loglik) {
  const int T = v.T, N = v.N, burn = v.burn, mcmc = v.mcmc;
  std::vector<int> uncertain;
  {
    std::vector<int> same_as_max(N, 0);  // row by row: the trace is [T][N]
    for (int i = burn; i < T; i++) {
      const int8_t* zi = v.Z + (size_t)i * N;
      for (int s = 0; s < N; s++) same_as_max[s] += zi[s] == maxZ[s];
    }
    for (int s = 0; s < N; s++)
      if (same_as_max[s] < mcmc) uncertain.push_back(s);
  }
  const int d = 2 + (v.resz != 0 ? 1 : 0);
  struct Conf {
    const int8_t* z;
    int cnt = 0;
    double mean[3] = {0, 0, 0}, var[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, sum_full = 0;
    // the distinct sets of gBGC branches seen with this configuration (getShifts, bpp_c2.cpp:14-27), as bit masks over the nodes, and how
    // often each was seen, in order of first appearance (a linear search over vectors of node ids made a chain's Eval2 quadratic in MCMC)
    std::unordered_map<std::string, int> bindex;
    std::vector<int> bcnt;
  };
  std::vector<Conf> conf(1);
  conf[0].z = maxZ;
  std::string shifts((size_t)(N + 7) / 8, '\0');
  for (int i = T - 1; i >= burn; i--) {
    const int8_t* zi = v.Z + (size_t)i * N;
    size_t j = 0;
    for (; j < conf.size(); j++) {
      bool same = true;
      for (int s : uncertain)
        if (zi[s] != conf[j].z[s]) { same = false; break; }
      if (same) break;
    }
    if (j == conf.size()) {
      conf.emplace_back();
      conf[j].z = zi;
    }
    Conf& c = conf[j];
    c.cnt++;
    std::fill(shifts.begin(), shifts.end(), '\0');  // the non-missing nodes of the pruned tree with B = 1, as a set
    for (int s : v.plan->nodes)
      if (!v.plan->missing[s] && v.B[(size_t)i * N + s]) shifts[s >> 3] |= (char)(1 << (s & 7));
    auto ins = c.bindex.emplace(shifts, (int)c.bcnt.size());
    if (ins.second) c.bcnt.push_back(0);
    c.bcnt[ins.first->second]++;
    const double cr = v.c[i], gb = v.gb[i], nr = v.n[i];
    c.var[1 * 3 + 1] += cr * cr; c.mean[1] += cr;
    c.var[0] += gb * gb; c.mean[0] += gb;
    c.var[0 * 3 + 1] += cr * gb;
    if (v.resz != 0) {
      c.var[2 * 3 + 2] += nr * nr; c.mean[2] += nr;
      c.var[1 * 3 + 2] += cr * nr;
      c.var[0 * 3 + 2] += gb * nr;
    }
    c.sum_full += v.full[i];
  }
  auto cov_logdet = [&](const double* var, const double* mean, int cnt, bool allow_full, double& logdet) {
    double S[9];
    for (int a = 0; a < d; a++)
      for (int b = 0; b < d; b++) {
        const double u = a <= b ? var[a * 3 + b] : var[b * 3 + a];  // symmatu
        S[a * d + b] = (u - mean[a] * mean[b] / cnt) / (cnt - 1);
      }
    double sign = 1;
    if (allow_full) log_det_small(S, d, logdet, sign);
    if (!allow_full || sign < 0) {
      logdet = 0;
      for (int a = 0; a < d; a++) logdet += std::log(S[a * d + a]);
    }
  };
  double VB = 0;
  int effective = 0;
  for (const Conf& c : conf) {
    const double w = (double)c.cnt / mcmc;
    if (w < 0.05) continue;
    double logdet;
    cov_logdet(c.var, c.mean, c.cnt, c.cnt > 10, logdet);
    if (logdet == -INFINITY) continue;
    effective += c.cnt;
    double tempB = 0;
    for (int b : c.bcnt) tempB += (double)b * std::log((double)b);
    VB -= tempB;
    VB += (double)(c.cnt - 1) / 2 * d;
    VB += c.sum_full;
    VB += 0.5 * logdet * c.cnt;
  }
  if (effective == 0) {  // no major configuration: q(r, Z) = q(r) q(Z) over the whole chain
    double mean0[3] = {0, 0, 0}, var0[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, tempB = 0;
    VB = 0;
    for (const Conf& c : conf) {
      for (int a = 0; a < 3; a++) mean0[a] += c.mean[a];
      for (int a = 0; a < 9; a++) var0[a] += c.var[a];
      VB += c.sum_full;
      for (int b : c.bcnt) tempB += (double)b * std::log((double)b);
    }
    VB -= tempB;
    double logdet;
    cov_logdet(var0, mean0, mcmc, true, logdet);
    if (logdet != -INFINITY) {
      VB += 0.5 * logdet * mcmc;
      VB += (double)(mcmc - 1) / 2 * d;
      effective = mcmc;
    }
  }
  if (v.resz == 1 && effective == 0) return max_loglik;
  return VB / effective + std::log((double)effective) + (double)d / 2 * std::log(2 * M_PI);
}

}  // namespace

}  // namespace pa

// The variant's Eval2 on a given chain (host-only; what pa_engine_gbgc_run applies to every finished chain): exposed for parity tests against
// the reference's own Eval2 (tests/test_gbgc_eval2.py)
extern "C" int pa_gbgc_eval2(int32_t N, int32_t burn, int32_t mcmc, int32_t resz, const double* full, const double* c_rate, const double* n_rate,
                             const double* gbc, const int8_t* Z, const int8_t* B, const int32_t* nodes, int32_t n_nodes, const uint8_t* missing,
                             double* loglik_out, double* max_loglik_out) {
  if (N <= 0 || burn < 0 || mcmc < 2 || !full || !c_rate || !n_rate || !gbc || !Z || !B || !nodes || !missing || !loglik_out) return PA_ERR_ARG;
  const int T = burn + mcmc;
  pa::ElemPlan pl;
  pl.nodes.assign(nodes, nodes + n_nodes);
  pl.missing.assign(missing, missing + N);
  double best = -INFINITY;
  int best_m = -1;
  for (int m = burn; m < T; m++)
    if (full[m] > best) { best = full[m]; best_m = m; }
  std::vector<int8_t> mz(N, -1);
  if (best_m >= 0) mz.assign(Z + (size_t)best_m * N, Z + (size_t)(best_m + 1) * N);
  pa::GbgcChainView v{T, N, burn, mcmc, resz, nullptr, full, c_rate, n_rate, gbc, Z, B, &pl};
  *loglik_out = pa::gbgc_eval2(v, mz.data(), best);
  if (max_loglik_out) *max_loglik_out = best;
  return PA_OK;
}

// The variant's Output_init summaries on a given chain (host-only; parity tests against the reference's own writer)
extern "C" int pa_gbgc_output_init(int32_t N, int32_t burn, int32_t mcmc, const double* n_rate, const double* c_rate, const double* gbc, const int8_t* Z,
                                   const int8_t* B, const uint8_t* missing, double* med3_out, double* post_b_out, double* post_z_out) {
  if (N <= 0 || burn < 0 || mcmc < 1 || !n_rate || !c_rate || !gbc || !Z || !B || !missing || !med3_out || !post_b_out || !post_z_out) return PA_ERR_ARG;
  pa::gbgc_output_init(burn + mcmc, N, burn, n_rate, c_rate, gbc, Z, B, missing, med3_out, post_b_out, post_z_out);
  return PA_OK;
}

namespace pa {

int Engine::gbgc_run(const int32_t* ids, int32_t n_ids, const int32_t* targets, int32_t n_targets, const pa_gbgc_params* prm, int model_mask,
                     pa_gbgc_results* out) {
  if (!loaded) {
    set_error("pa_engine_gbgc_run called before pa_engine_load_alignment");
    return PA_ERR_STATE;
  }
  if (device < 0) {
    set_error("no CUDA device bound to this engine (there is no CPU fallback)");
    return PA_ERR_CUDA;
  }
  if (!prm || !out || n_ids < 0 || (n_ids > 0 && !ids) || prm->num_burn < 0 || prm->num_mcmc < 2) {
    set_error("pa_engine_gbgc_run: bad arguments");
    return PA_ERR_ARG;
  }
  const int N = model.tree.N, S = model.tree.S, T = prm->num_burn + prm->num_mcmc;
  if (out->n_ids != n_ids || out->N != N || out->T != T) {
    set_error("pa_engine_gbgc_run: results struct sized for another run (n_ids, N, T must match)");
    return PA_ERR_ARG;
  }
  for (int t = 0; t < n_targets; t++)
    if (targets[t] < 0 || targets[t] >= S) {
      set_error("pa_engine_gbgc_run: target species out of range");
      return PA_ERR_ARG;
    }
  static const int kResz[3] = {0, 2, 1};  // M0, M1, M2 -> the reference's resZ
  std::vector<int32_t> chain_elem, chain_resz, chain_pos, chain_mdl;
  for (int i = 0; i < n_ids; i++) {
    if (ids[i] < 0 || ids[i] >= (int)plans.size()) {
      set_error("element id out of range");
      return PA_ERR_ARG;
    }
    out->filtered[i] = plans[ids[i]].filtered ? 1 : 0;
    if (plans[ids[i]].filtered) continue;
    for (int mdl = 0; mdl < 3; mdl++)
      if (model_mask & (1 << mdl)) {
        chain_elem.push_back(ids[i]);
        chain_resz.push_back(kResz[mdl]);
        chain_pos.push_back(i);
        chain_mdl.push_back(mdl);
      }
  }
  const int nc = (int)chain_elem.size();
  // longest chains first: cost ~ patterns
  std::vector<int> order(nc);
  for (int k = 0; k < nc; k++) order[k] = k;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return plans[chain_elem[a]].P > plans[chain_elem[b]].P; });
  std::vector<int32_t> l_elem(nc), l_resz(nc);
  std::vector<int8_t> fixZ((size_t)nc * N, -1), upper((size_t)nc * N, 0);
  for (int k = 0; k < nc; k++) {
    const int c = order[k];
    const ElemPlan& pl = plans[chain_elem[c]];
    l_elem[k] = chain_elem[c];
    l_resz[k] = chain_resz[c];
    int8_t* fz = fixZ.data() + (size_t)k * N;
    for (int32_t v : pl.upper_c) { fz[v] = 1; upper[(size_t)k * N + v] = 1; }
    fz[pl.root] = 0;
    if (chain_resz[c] == 2) {
      for (int s = 0; s < S; s++) fz[s] = 1;
      for (int t = 0; t < n_targets; t++) fz[targets[t]] = -1;
    } else if (chain_resz[c] == 0) {
      for (size_t q = 0; q + 1 < pl.nodes.size(); q++) fz[pl.nodes[q]] = 1;
    }
  }
  // ---- waves: the chains go through the device in W launches (every W-th chain of the longest-first order, so the waves are alike), two
  // in flight; the host-side summaries of wave w (arg-max, medians, frequencies, the variant's Eval2) run under the kernel of wave w + 1
  const int slots = std::max(1, num_sms) * 3;
  const int W = nc < 4 * slots ? 1 : std::min(8, nc / (2 * slots));
  const int burn = prm->num_burn, mcmc = prm->num_mcmc;
  out->kernel_ms = 0;
  out->d2h_bytes = 0;
  unsigned long long ctr[16] = {0};
  struct WaveHost {
    std::vector<int> ks;  // positions in the launch order
    std::vector<int32_t> elem, resz;
    std::vector<int8_t> fixZ, upper;
    std::vector<double> tr, head;
    std::vector<uint8_t> zb;
  };
  std::vector<WaveHost> wh(W);
  std::vector<GbgcWave> wd(std::min(W, 3));  // device slots, reused round-robin: wave w runs in slot w % 3
  for (int k = 0; k < nc; k++) wh[k % W].ks.push_back(k);
  auto launch = [&](int w) {
    WaveHost& h = wh[w];
    const size_t n = h.ks.size();
    h.elem.resize(n); h.resz.resize(n); h.fixZ.resize(n * N); h.upper.resize(n * N);
    for (size_t q = 0; q < n; q++) {
      const int k = h.ks[q];
      h.elem[q] = l_elem[k];
      h.resz[q] = l_resz[k];
      memcpy(h.fixZ.data() + q * N, fixZ.data() + (size_t)k * N, N);
      memcpy(h.upper.data() + q * N, upper.data() + (size_t)k * N, N);
    }
    PA_NVTX_RANGE("gbgc_launch");
    return gbgc_chains_launch(device, num_sms, d_blobs, d_blob_off, N, n_max, slots_max, P_max, model.tree.pi, model.tree.Q, (int)n, h.elem.data(), h.resz.data(),
                              h.fixZ.data(), h.upper.data(), prm, &wd[w % 3]);
  };
  auto release_all = [&]() {
    for (GbgcWave& g : wd) gbgc_chains_release(&g);
  };
  auto summarise = [&](int c, const double* d, const uint8_t* zbk, const double* hd) {
    const int i = chain_pos[c], mdl = chain_mdl[c];
    const ElemPlan& pl = plans[chain_elem[c]];
    const size_t im = (size_t)i * 3 + mdl;
    std::vector<int8_t> Z((size_t)T * N, -1), B((size_t)T * N, 0);
    const int n = (int)pl.nodes.size();
    for (int m = 0; m < T; m++)
      for (int q = 0; q < n; q++) {
        const uint8_t b = zbk[(size_t)m * n_max + q];
        Z[(size_t)m * N + pl.nodes[q]] = (int8_t)(b & 3);
        B[(size_t)m * N + pl.nodes[q]] = (int8_t)(b >> 2);
      }
    // MonitorChain's arg-max (first maximum wins)
    double best = -INFINITY;
    int best_m = -1;
    for (int m = burn; m < T; m++)
      if (d[T + m] > best) { best = d[T + m]; best_m = m; }
    out->loglik_max[im] = best;
    int8_t* mz = out->max_z + im * N;
    int8_t* mb = out->max_b + im * N;
    for (int s = 0; s < N; s++) {
      mz[s] = best_m >= 0 ? Z[(size_t)best_m * N + s] : (int8_t)-1;
      mb[s] = best_m >= 0 ? B[(size_t)best_m * N + s] : (int8_t)0;
    }
    gbgc_output_init(T, N, burn, d + 3 * (size_t)T, d + 2 * (size_t)T, d + 4 * (size_t)T, Z.data(), B.data(), pl.missing.data(), out->med + im * 3,
                     out->post_b + im * N, out->post_z + im * N * 4);
    GbgcChainView v{T, N, burn, mcmc, chain_resz[c], d, d + T, d + 2 * (size_t)T, d + 3 * (size_t)T, d + 4 * (size_t)T, Z.data(), B.data(), &pl};
    out->loglik[im] = gbgc_eval2(v, mz, best);
    out->accept[im * 2] = (int32_t)hd[1];
    out->accept[im * 2 + 1] = (int32_t)hd[2];
    if (out->trace) memcpy(out->trace + im * 5 * T, d, sizeof(double) * 5 * T);
    if (out->trace_z) memcpy(out->trace_z + im * (size_t)T * N, Z.data(), (size_t)T * N);
    if (out->trace_b) memcpy(out->trace_b + im * (size_t)T * N, B.data(), (size_t)T * N);
  };
  int rc = launch(0);
  if (rc == PA_OK && W > 1) rc = launch(1);
  if (rc != PA_OK) { release_all(); return rc; }
  cudaEvent_t e_first = wd[0].e0;  // start of the first kernel: the reported kernel time is the span to the end of the last one
  wd[0].e0 = nullptr;            // (slot 0 gets a new start event when it is launched again)
  const int nthr = std::max(1, std::min<int>(16, (int)std::thread::hardware_concurrency()));
  for (int w = 0; w < W; w++) {
    WaveHost& h = wh[w];
    const size_t n = h.ks.size();
    h.tr.resize(n * 5 * T); h.head.resize(n * 8); h.zb.resize(n * (size_t)T * n_max);
    {
      PA_NVTX_RANGE("gbgc_wait");
      rc = gbgc_chains_wait(&wd[w % 3], h.tr.data(), h.zb.data(), h.head.data(), ctr);
    }
    if (rc == PA_OK && e_first && wd[w % 3].e1) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, e_first, wd[w % 3].e1) == cudaSuccess) out->kernel_ms = std::max(out->kernel_ms, (double)ms);
    }
    if (rc == PA_OK && w + 2 < W) rc = launch(w + 2);
    if (rc != PA_OK) break;
    out->d2h_bytes += (uint64_t)(h.tr.size() * 8 + h.head.size() * 8 + h.zb.size());
    PA_NVTX_RANGE("gbgc_summaries");
    std::vector<std::thread> pool;
    std::atomic<size_t> next{0};
    for (int t = 0; t < nthr; t++)
      pool.emplace_back([&]() {
        for (;;) {
          const size_t q = next.fetch_add(1);
          if (q >= n) break;
          summarise(order[h.ks[q]], h.tr.data() + q * 5 * T, h.zb.data() + q * (size_t)T * n_max, h.head.data() + q * 8);
        }
      });
    for (auto& th : pool) th.join();
    std::vector<double>().swap(h.tr);
    std::vector<uint8_t>().swap(h.zb);
  }
  if (e_first) cudaEventDestroy(e_first);
  release_all();
  if (rc != PA_OK) return rc;
  out->kernel_launches = (uint64_t)(nc > 0 ? W : 0);
  out->node_pattern_updates = ctr[0];
  out->chain_iterations = ctr[2];
  out->pt_builds = ctr[1];
  out->branch_pattern_updates = ctr[3];
  for (int k = 0; k < 12; k++) out->phase_cycles[k] = ctr[4 + k];
  return PA_OK;
}

}  // namespace pa
