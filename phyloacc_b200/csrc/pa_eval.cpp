// pa_eval.cpp -- host side of a sampling run: chain scheduling, trace collection, and the per-chain summaries the
// reference computes after Gibbs: MaxLoglik / Max_Z (MonitorChain, bpp_c.cpp:768-814), Eval2 (bpp_c2.cpp:14-353) and
// Output_init's medians / posterior Z frequencies (bpp_c2.cpp:739-785).
#include <cuda_runtime.h>
#include <time.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "pa_chain.cuh"
#include "pa_engine.h"
#include "pa_nvtx.h"
#include "pa_summary.cuh"

namespace pa {


#define PA_CUDA(call)                                                                         \
  do {                                                                                        \
    cudaError_t err__ = (call);                                                               \
    if (err__ != cudaSuccess) {                                                               \
      set_error(std::string("CUDA error: ") + cudaGetErrorString(err__) + " at " #call);     \
      return PA_ERR_CUDA;                                                                     \
    }                                                                                         \
  } while (0)

static double gamma_pdf(double x, double a, double b) {  // gsl_ran_gamma_pdf
  if (x < 0) return 0;
  if (x == 0) return (a == 1) ? 1 / b : 0;
  if (a == 1) return std::exp(-x / b) / b;
  return std::exp((a - 1) * std::log(x / b) - x / b - std::lgamma(a)) / b;
}
// the same density with lgamma(a) supplied by the caller (a is a constant of the run: one lgamma instead of one per iteration)
static double gamma_pdf_lg(double x, double a, double b, double lga) {
  if (x < 0) return 0;
  if (x == 0) return (a == 1) ? 1 / b : 0;
  if (a == 1) return std::exp(-x / b) / b;
  return std::exp((a - 1) * std::log(x / b) - x / b - lga) / b;
}
static double lnbeta(double a, double b) { return std::lgamma(a) + std::lgamma(b) - std::lgamma(a + b); }

// ------------------------------------------------------------------------------------------ hyper-parameters
namespace {

// host copy of the device draw schedule (pa_device.cuh: philox4 / u01 / sched_gamma)
struct U4 { uint32_t x, y, z, w; };
U4 philox4_host(uint32_t seed, uint32_t key, uint32_t idx, uint32_t sub, uint32_t iter, uint32_t stream) {
  uint32_t c0 = idx, c1 = sub, c2 = iter, c3 = stream, k0 = seed, k1 = key;
  for (int r = 0; r < 10; r++) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return U4{c0, c1, c2, c3};
}
double u01_host(uint32_t w) { return ((double)w + 0.5) * 2.3283064365386963e-10; }
double gamma_host(uint32_t seed, uint32_t key, uint32_t idx, uint32_t sub0, uint32_t iter, uint32_t stream, double a, double b) {
  U4 w = philox4_host(seed, key, idx, sub0, iter, stream);
  double boost = 1.0;
  if (a < 1.0) {
    boost = std::pow(u01_host(w.w), 1.0 / a);
    a += 1.0;
  }
  const double d = a - 1.0 / 3.0, c = (1.0 / 3.0) / std::sqrt(d);
  for (uint32_t attempt = 0;; attempt++) {
    if (attempt > 0) w = philox4_host(seed, key, idx, sub0 + attempt, iter, stream);
    if (attempt > 1000) return NAN;
    const double x = std::sqrt(-2.0 * std::log(u01_host(w.x))) * std::cos(6.283185307179586 * u01_host(w.y));
    double v = 1.0 + c * x;
    if (v <= 0) continue;
    v = v * v * v;
    const double u = u01_host(w.z);
    if (u < 1 - 0.0331 * x * x * x * x) return b * d * v * boost;
    if (std::log(u) < 0.5 * x * x + d * (1 - v + std::log(v))) return b * d * v * boost;
  }
}

}  // namespace

#define PA_ST_HYPER 4u
#define PA_KEY_HYPER 3u  // element keys are 4 * element + model, model < 3

// BPP::sample_hyperparam (bpp.cpp:324-477): independence MH on the Gamma priors of the two rates and multiplicative
// uniform MH on the Beta priors of l, l2, g, from the CURRENT states of the full-model chains (bpp.cur_*; elements
// whose chain never ran keep the INIT_* values, as in the reference).  Draws come from the Philox schedule (stream
// PA_ST_HYPER) in the reference's order of consumption.
// The sweep depends on the elements only through ten sums and their count (PA_HYPER_NSUMS): sum n, sum log n, sum c, sum log c,
// sum log l, sum log(1-l), sum log l2, sum log(1-l2), sum log g, sum log(1-g), count.  With the elements sharded over ranks each
// rank contributes the sums of its own ids, one all-reduce adds them, and every rank runs the same sweep (same Philox draws).
int Engine::hyper_sums(const int32_t* ids, int32_t n_ids, double* sums) {
  const pa_params& P = model.prm;
  const size_t C = plans.size();
  for (int k = 0; k < PA_HYPER_NSUMS; k++) sums[k] = 0.0;
  for (int i = 0; i < n_ids; i++) {
    const int c = ids[i];
    if (c < 0 || (size_t)c >= C) {
      set_error("element id out of range");
      return PA_ERR_ARG;
    }
    const size_t slot = (size_t)c * 3 + PA_M2;
    const bool have = cur_set.size() == C * 3 && cur_set[slot];
    const double rn = have ? cur_rates[slot * 5 + 0] : P.init_acce_rate, rc = have ? cur_rates[slot * 5 + 1] : P.init_conserve_rate;
    const double rg = have ? cur_rates[slot * 5 + 2] : P.init_grate, rl = have ? cur_rates[slot * 5 + 3] : P.init_lrate;
    const double rl2 = have ? cur_rates[slot * 5 + 4] : P.init_lrate2;
    sums[0] += rn; sums[1] += std::log(rn);
    sums[2] += rc; sums[3] += std::log(rc);
    sums[4] += std::log(rl); sums[5] += std::log(1 - rl);
    sums[6] += std::log(rl2); sums[7] += std::log(1 - rl2);
    sums[8] += std::log(rg); sums[9] += std::log(1 - rg);
  }
  sums[10] = (double)n_ids;
  return PA_OK;
}

int Engine::sample_hyper(int iter, const int32_t* ids, int32_t n_ids, pa_hyper* out) {
  double sums[PA_HYPER_NSUMS];
  const int rc = hyper_sums(ids, n_ids, sums);
  return rc != PA_OK ? rc : sample_hyper_from_sums(iter, sums, out);
}

int Engine::sample_hyper_from_sums(int iter, const double* sums, pa_hyper* out) {
  pa_params& P = model.prm;
  const uint32_t seed = (uint32_t)P.seed, it = (uint32_t)iter;
  uint32_t n_draw = 0;
  auto gam = [&](double a, double b) { return gamma_host(seed, PA_KEY_HYPER, n_draw++, 0u, it, PA_ST_HYPER, a, b); };
  auto uni = [&]() { return u01_host(philox4_host(seed, PA_KEY_HYPER, n_draw++, 0u, it, PA_ST_HYPER).x); };
  const double p = 1, r = 1, q = 0.1, sc = 0.1;  // hyper-priors of the Gamma shapes / scales (bpp.cpp:328-329)
  const double vna = 100, vnb = 100, vca = 100, vcb = 100;
  const double n_ids_d = sums[10];
  double& nprior_a = P.acce_prior_a; double& nprior_b = P.acce_prior_b;
  double& cprior_a = P.conserve_prior_a; double& cprior_b = P.conserve_prior_b;
  const double nprior_a_prop = gam(vna, nprior_a / vna);
  const double cprior_a_prop = gam(vca, cprior_a / vca);
  const double nprior_b_prop = gam(vnb, nprior_b / vnb);
  const double cprior_b_prop = gam(vcb, cprior_b / vcb);
  auto rate_step = [&](double sum_r, double log_prod_r, double& a, double& b, double a_prop, double b_prop, double va, double vb) {
    double M = (a_prop - 1) * (std::log(p) + log_prod_r) - (q + sum_r) / b_prop - (n_ids_d + r) * std::lgamma(a_prop) - std::log(b_prop) * a_prop * (sc + n_ids_d);
    M -= (a - 1) * (std::log(p) + log_prod_r) - (q + sum_r) / b - (n_ids_d + r) * std::lgamma(a) - std::log(b) * a * (sc + n_ids_d);
    const double H = std::log(gamma_pdf(a, va, a_prop / va)) - std::log(gamma_pdf(a_prop, va, a / va)) + std::log(gamma_pdf(b, vb, b_prop / vb)) -
                     std::log(gamma_pdf(b_prop, vb, b / vb));
    if (std::log(uni()) < M + H) {
      a = a_prop;
      b = b_prop;
    }
  };
  rate_step(sums[0], sums[1], nprior_a, nprior_b, nprior_a_prop, nprior_b_prop, vna, vnb);
  rate_step(sums[2], sums[3], cprior_a, cprior_b, cprior_a_prop, cprior_b_prop, vca, vcb);
  auto beta_step = [&](double log_p, double log_pc, double& a, double& b) {
    const double a_prop = a * (uni() * (1.3 - 0.7) + 0.7);
    const double b_prop = b * (uni() * (1.3 - 0.7) + 0.7);
    double M = (a_prop - a) * (log_p - 1) + (b_prop - b) * (log_pc - 1);
    M += n_ids_d * (lnbeta(a, b) - lnbeta(a_prop, b_prop));
    const double H = std::log(a) - std::log(a_prop) + std::log(b) - std::log(b_prop);
    if (std::log(uni()) < M + H) {
      a = a_prop;
      b = b_prop;
    }
  };
  beta_step(sums[4], sums[5], P.hyper_lrate_a, P.hyper_lrate_b);
  beta_step(sums[6], sums[7], P.hyper_lrate2_a, P.hyper_lrate2_b);
  beta_step(sums[8], sums[9], P.hyper_grate_a, P.hyper_grate_b);
  if (out) {
    out->nprior_a = nprior_a; out->nprior_b = nprior_b; out->cprior_a = cprior_a; out->cprior_b = cprior_b;
    out->prior_l_a = P.hyper_lrate_a; out->prior_l_b = P.hyper_lrate_b; out->prior_g_a = P.hyper_grate_a; out->prior_g_b = P.hyper_grate_b;
    out->prior_l2_a = P.hyper_lrate2_a; out->prior_l2_b = P.hyper_lrate2_b;
  }
  return PA_OK;
}

// ------------------------------------------------------------------------------------------ launch
int Engine::sample(const int32_t* ids, int32_t n_ids, int model_mask, bool warm_start, int chain_iter) {
  PA_NVTX_RANGE("pa_engine_sample");
  if (!loaded) {
    set_error("pa_engine_sample called before pa_engine_load_alignment");
    return PA_ERR_STATE;
  }
  if (device < 0) {
    set_error("no CUDA device bound to this engine (there is no CPU fallback)");
    return PA_ERR_CUDA;
  }
  const pa_params& P = model.prm;
  T = P.burnin + P.mcmc;
  if (T < 2 || P.adapt_freq < 1) {
    set_error("BURNIN + MCMC must be >= 2 and ADAPT_FREQ >= 1");
    return PA_ERR_ARG;
  }
  run_ids.assign(ids, ids + n_ids);
  chain_of.assign((size_t)n_ids * 3, -1);
  chain_elem_pos.clear();
  chain_model.clear();
  std::vector<double> cost;
  for (int i = 0; i < n_ids; i++) {
    const int c = ids[i];
    if (c < 0 || c >= (int)plans.size()) {
      set_error("element id out of range");
      return PA_ERR_ARG;
    }
    if (plans[c].filtered) continue;
    for (int mdl = 0; mdl < 3; mdl++) {
      if (!((model_mask >> mdl) & 1)) continue;
      chain_of[(size_t)i * 3 + mdl] = (int)chain_elem_pos.size();
      chain_elem_pos.push_back(i);
      chain_model.push_back(mdl);
      cost.push_back((double)plans[c].GG * plans[c].nodes.size() * (mdl == 0 ? 0.8 : 1.0));
    }
  }
  n_chains = (int)chain_elem_pos.size();
  if (sampled && !collected && device >= 0) {  // a launch that was never collected may still be writing its completion flags
    cudaSetDevice(device);
    cudaStreamSynchronize(stream);
  }
  sampled = true;
  collected = false;
  traces_on_host = false;
  d2h_bytes = 0;
  device_summaries = !(getenv("PA_HOST_SUMMARIES") && atoi(getenv("PA_HOST_SUMMARIES")) != 0);
  done_prefix = fetched_prefix = 0;
  counters = pa_counters{};
  if (n_chains == 0) return PA_OK;
  ZW = (n_max + 15) / 16;
  // longest chains first so that the tail of the persistent-CTA queue is short
  std::vector<int> order(n_chains);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost[a] > cost[b]; });
  // chain index on the device = position in `order`; keep the mapping
  std::vector<int32_t> list((size_t)n_chains * 2);
  std::vector<int> dev_index(n_chains);
  for (int k = 0; k < n_chains; k++) {
    const int ch = order[k];
    list[2 * k] = run_ids[chain_elem_pos[ch]];
    list[2 * k + 1] = chain_model[ch];
    dev_index[ch] = k;
  }
  for (auto& v : chain_of)
    if (v >= 0) v = dev_index[v];
  {
    std::vector<int32_t> pos2(n_chains), mdl2(n_chains);
    for (int ch = 0; ch < n_chains; ch++) {
      pos2[dev_index[ch]] = chain_elem_pos[ch];
      mdl2[dev_index[ch]] = chain_model[ch];
    }
    chain_elem_pos.swap(pos2);
    chain_model.swap(mdl2);
  }

  PA_CUDA(cudaSetDevice(device));
  // device buffers of a run are kept between runs and only grow (allocation is synchronous and slow next to a launch)
  auto reserve = [&](void*& p, size_t& cap, size_t bytes) -> cudaError_t {
    if (bytes <= cap && p) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  };

  // one persistent launch: CTAs pull chains (longest first) from an atomic queue
  const int BD = 256;  // threads per CTA of the chain kernel's layouts
  const size_t smem = chain_smem_bytes(n_max, model.tree.S, BD);
  if ((int)smem > smem_optin) {
    set_error("tree too large for the chain kernel's shared-memory layouts (" + std::to_string(n_max) + " nodes)");
    return PA_ERR_LIMIT;
  }
  // resident CTAs per SM: bounded by shared memory (228 KB per SM, 1 KB reserved per CTA) and by the layout's launch bounds
  int per_sm = std::max(1, std::min(chain_ctas_per_sm(n_max, model.tree.S, BD), (int)(233472 / (smem + 1024))));
  if (const char* env = getenv("PA_CTAS_PER_SM_RUNTIME")) per_sm = std::max(1, std::min(per_sm, atoi(env)));  // tuning knob
  int grid = std::min(n_chains, num_sms * per_sm);
  const size_t stride = chain_arena_bytes(n_max, P_max);
  // the arena is sized by the longest element of the batch: one very long element inflates every CTA's share.  Run with fewer
  // resident CTAs rather than fail when the full grid's arena does not fit beside the traces
  if (stride * (size_t)grid > cap_arena) {
    size_t free_b = 0, total_b = 0;
    PA_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const size_t traces = (sizeof(double) * 6 + sizeof(uint32_t) * (size_t)((n_max + 15) / 16)) * (size_t)T * n_chains;
    const size_t need_other = (traces > cap_trace_d + cap_trace_z ? traces - (cap_trace_d + cap_trace_z) : 0) + ((size_t)64 << 20);
    const size_t avail = free_b + cap_arena > need_other ? free_b + cap_arena - need_other : 0;
    if (stride * (size_t)grid > avail) {
      const int fit = (int)(avail / stride);
      if (fit < 1) {
        set_error("not enough device memory for one chain arena (" + std::to_string(stride >> 20) + " MiB per resident chain)");
        return PA_ERR_LIMIT;
      }
      grid = std::min(grid, fit);
    }
  }
  PA_CUDA(reserve((void*&)d_chain_list, cap_chain_list, sizeof(int32_t) * 2 * n_chains));
  PA_CUDA(reserve((void*&)d_trace_d, cap_trace_d, sizeof(double) * 6 * (size_t)T * n_chains));
  PA_CUDA(reserve((void*&)d_trace_z, cap_trace_z, sizeof(uint32_t) * (size_t)ZW * T * n_chains));
  PA_CUDA(reserve((void*&)d_arena, cap_arena, stride * grid));
  PA_CUDA(reserve((void*&)d_counters, cap_counters, sizeof(unsigned long long) * 24));
  PA_CUDA(reserve((void*&)d_queue, cap_queue, sizeof(int)));
  PA_CUDA(cudaMemcpyAsync(d_chain_list, list.data(), sizeof(int32_t) * 2 * n_chains, cudaMemcpyHostToDevice, stream));
  PA_CUDA(cudaMemsetAsync(d_counters, 0, sizeof(unsigned long long) * 24, stream));
  PA_CUDA(cudaMemsetAsync(d_queue, 0, sizeof(int), stream));
  rec_bytes = summary_record_bytes(n_max, ZW);
  if (device_summaries) {
    PA_CUDA(reserve((void*&)d_summary, cap_d_summary, rec_bytes * (size_t)n_chains));
    if (rec_bytes * (size_t)n_chains > cap_h_summary || !h_summary) {
      if (h_summary) cudaFreeHost(h_summary);
      h_summary = nullptr;
      cap_h_summary = 0;
      PA_CUDA(cudaHostAlloc((void**)&h_summary, rec_bytes * (size_t)n_chains, cudaHostAllocDefault));
      cap_h_summary = rec_bytes * (size_t)n_chains;
    }
  }
  // host-summary path only: completion flags in mapped pinned memory (grow-only), the trace buffers on the host and the download stream
  int* d_done = nullptr;
  if (!device_summaries) {
  if (sizeof(int) * (size_t)n_chains > cap_h_done || !h_done) {
    if (h_done) cudaFreeHost(h_done);
    h_done = nullptr;
    cap_h_done = 0;
    PA_CUDA(cudaHostAlloc((void**)&h_done, sizeof(int) * (size_t)n_chains, cudaHostAllocMapped));
    cap_h_done = sizeof(int) * (size_t)n_chains;
  }
  memset(h_done, 0, sizeof(int) * (size_t)n_chains);
  PA_CUDA(cudaHostGetDevicePointer((void**)&d_done, h_done, 0));
  if (!copy_stream) PA_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
  {
    auto reserve_host = [&](void*& p, size_t& cap, size_t bytes) -> cudaError_t {
      if (bytes <= cap && p) return cudaSuccess;
      if (p) cudaFreeHost(p);
      p = nullptr;
      cap = 0;
      cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocDefault);
      if (e == cudaSuccess) cap = bytes;
      return e;
    };
    PA_CUDA(reserve_host((void*&)h_trace_d, cap_h_trace_d, sizeof(double) * 6 * (size_t)T * n_chains));
    PA_CUDA(reserve_host((void*&)h_trace_z, cap_h_trace_z, sizeof(uint32_t) * (size_t)ZW * T * n_chains));
  }
  }  // !device_summaries
  bool any_warm = false;
  if (warm_start && !cur_set.empty()) {
    std::vector<double> ir((size_t)n_chains * 8, 0.0);
    std::vector<uint32_t> iz((size_t)n_chains * ZW, 0u);
    for (int k = 0; k < n_chains; k++) {
      const size_t slot = (size_t)list[2 * k] * 3 + list[2 * k + 1];
      if (!cur_set[slot]) continue;
      any_warm = true;
      for (int j = 0; j < 5; j++) ir[(size_t)k * 8 + j] = cur_rates[slot * 5 + j];
      ir[(size_t)k * 8 + 5] = 1.0;
      for (int w = 0; w < ZW; w++) iz[(size_t)k * ZW + w] = cur_z[slot * ZW + w];
    }
    if (any_warm) {
      PA_CUDA(reserve((void*&)d_init_rates, cap_init_rates, sizeof(double) * ir.size()));
      PA_CUDA(reserve((void*&)d_init_z, cap_init_z, sizeof(uint32_t) * iz.size()));
      PA_CUDA(cudaMemcpyAsync(d_init_rates, ir.data(), sizeof(double) * ir.size(), cudaMemcpyHostToDevice, stream));
      PA_CUDA(cudaMemcpyAsync(d_init_z, iz.data(), sizeof(uint32_t) * iz.size(), cudaMemcpyHostToDevice, stream));
      PA_CUDA(cudaStreamSynchronize(stream));  // the staging vectors go out of scope
    }
  }

  ChainLaunch L;
  L.blobs = d_blobs;
  L.blob_off = d_blob_off;
  L.chain_list = d_chain_list;
  L.n_chains = n_chains;
  L.sd = model.subst_dev;
  ChainParams& cp = L.cp;
  cp.cprior_a = P.conserve_prior_a; cp.cprior_b = P.conserve_prior_b; cp.nprior_a = P.acce_prior_a; cp.nprior_b = P.acce_prior_b;
  cp.prior_l_a = P.hyper_lrate_a; cp.prior_l_b = P.hyper_lrate_b; cp.prior_l2_a = P.hyper_lrate2_a; cp.prior_l2_b = P.hyper_lrate2_b;
  cp.prior_g_a = P.hyper_grate_a; cp.prior_g_b = P.hyper_grate_b;
  cp.init_crate = P.init_conserve_rate; cp.init_nrate = P.init_acce_rate;
  cp.init_lrate = P.init_lrate; cp.init_lrate2 = (P.hyper_lrate2_a == 0) ? 0.0 : P.init_lrate2; cp.init_grate = P.init_grate;
  cp.cub = P.cub; cp.nlb = P.nlb; cp.log_constomis = std::log(P.constomis);
  cp.ropt = P.ropt; cp.num_burn = P.burnin; cp.num_mcmc = P.mcmc; cp.adaptive_freq = P.adapt_freq;
  cp.seed = (uint32_t)P.seed + 0x9E3779B9u * (uint32_t)chain_iter;  // outer CHAIN iterations draw from distinct streams
  cp.pad_ = 0;
  L.n_max = n_max; L.slots_max = slots_max; L.T = T; L.ZW = ZW; L.S_rows = model.tree.S;
  L.trace_d = d_trace_d; L.trace_z = d_trace_z;
  L.init_rates = any_warm ? d_init_rates : nullptr;
  L.init_z = any_warm ? d_init_z : nullptr;
  L.counters = d_counters;
  L.sel = nullptr; L.n_sel = n_chains; L.P_max = P_max;
  L.arena = d_arena; L.arena_stride = stride; L.queue = d_queue;
  L.done = d_done;
  int rc = launch_chains(L, grid, BD, stream);
  if (rc != PA_OK) return rc;
  counters.kernel_launches = 1;
  if (device_summaries) {  // one CTA per chain reduces its trace to a summary record; same stream, right behind the chains
    SummaryLaunch SL;
    SL.blobs = d_blobs; SL.blob_off = d_blob_off; SL.chain_list = d_chain_list;
    SL.n_chains = n_chains; SL.T = T; SL.ZW = ZW; SL.num_burn = P.burnin; SL.num_mcmc = P.mcmc;
    SL.n_max_pad = (n_max + 1) & ~1;
    SL.trace_d = d_trace_d; SL.trace_z = d_trace_z;
    SL.cprior_a = cp.cprior_a; SL.cprior_b = cp.cprior_b; SL.nprior_a = cp.nprior_a; SL.nprior_b = cp.nprior_b;
    SL.prior_l_a = cp.prior_l_a; SL.prior_l_b = cp.prior_l_b; SL.prior_l2_a = cp.prior_l2_a; SL.prior_l2_b = cp.prior_l2_b;
    SL.prior_g_a = cp.prior_g_a; SL.prior_g_b = cp.prior_g_b;
    SL.out = d_summary; SL.rec_bytes = rec_bytes;
    rc = launch_summaries(SL, n_max, stream);
    if (rc != PA_OK) return rc;
    counters.kernel_launches = 2;
  }
  return PA_OK;
}

// Download the traces of chains [fetched_prefix, hi) (device order = queue order: the kernel finishes them roughly in that
// order and flags each one in mapped pinned memory).  Called while the kernel is still running: the copies go over a second
// stream, so the host summaries of finished chains overlap the chains that are still being sampled.
int Engine::fetch_range(int lo, int hi) {
  (void)lo;
  if (!sampled) {
    set_error("nothing sampled yet");
    return PA_ERR_STATE;
  }
  hi = std::min(hi, n_chains);
  if (device_summaries) {
    set_error("fetch_range is part of the host-summary path");
    return PA_ERR_STATE;
  }
  if (collected || hi <= fetched_prefix) return PA_OK;
  PA_CUDA(cudaSetDevice(device));
  volatile int* done = h_done;
  bool finished = false;
  while (done_prefix < hi) {
    while (done_prefix < n_chains && done[done_prefix]) done_prefix++;
    if (done_prefix >= hi) break;
    if (finished) {  // the launch is over and a flag is missing: the kernel did not run through
      set_error("chain kernel ended without completing every chain");
      return PA_ERR_CUDA;
    }
    const cudaError_t q = cudaStreamQuery(stream);
    if (q == cudaSuccess) finished = true;  // everything is flagged by now; scan once more
    else if (q != cudaErrorNotReady) {
      set_error(std::string("chain kernel: ") + cudaGetErrorString(q));
      return PA_ERR_CUDA;
    } else {
      struct timespec ts = {0, 100000};  // 0.1 ms
      nanosleep(&ts, nullptr);
    }
  }
  const int a = fetched_prefix, b = std::min(n_chains, std::max(hi, done_prefix));
  const size_t per_d = (size_t)6 * T, per_z = (size_t)ZW * T;
  PA_CUDA(cudaMemcpyAsync(h_trace_d + a * per_d, d_trace_d + a * per_d, sizeof(double) * per_d * (b - a), cudaMemcpyDeviceToHost, copy_stream));
  PA_CUDA(cudaMemcpyAsync(h_trace_z + a * per_z, d_trace_z + a * per_z, sizeof(uint32_t) * per_z * (b - a), cudaMemcpyDeviceToHost, copy_stream));
  PA_CUDA(cudaStreamSynchronize(copy_stream));
  d2h_bytes += (sizeof(double) * per_d + sizeof(uint32_t) * per_z) * (size_t)(b - a);
  fetched_prefix = b;
  return PA_OK;
}

int Engine::fetch_finish() {
  if (collected || n_chains == 0) {
    collected = true;
    return PA_OK;
  }
  PA_CUDA(cudaSetDevice(device));
  unsigned long long hc[24];
  PA_CUDA(cudaMemcpyAsync(hc, d_counters, sizeof(hc), cudaMemcpyDeviceToHost, stream));
  PA_CUDA(cudaStreamSynchronize(stream));
  counters.node_pattern_updates = hc[0];
  counters.root_evals = hc[1];
  counters.pt_builds = hc[2];
  counters.history_draws = hc[3];
  counters.chain_iterations = hc[4];
  counters.node_site_updates = hc[5];
  for (int k = 0; k < 10; k++) phase_cycles[k] = hc[8 + k];
  // the last state of every chain becomes the warm start of its next run (Gibbs with UpHyper, bpp_c.cpp:474-485)
  {
    const size_t C = plans.size();
    if (cur_set.size() != C * 3 || cur_z.size() != C * 3 * (size_t)ZW) {
      cur_set.assign(C * 3, 0);
      cur_rates.assign(C * 15, 0.0);
      cur_z.assign(C * 3 * (size_t)ZW, 0u);
    }
    for (int ch = 0; ch < n_chains; ch++) {
      const size_t slot = (size_t)run_ids[chain_elem_pos[ch]] * 3 + chain_model[ch];
      const double* d = &h_trace_d[(size_t)ch * 6 * T];
      for (int j = 0; j < 5; j++) cur_rates[slot * 5 + j] = d[(size_t)(j + 1) * T + (T - 1)];
      for (int w = 0; w < ZW; w++) cur_z[slot * ZW + w] = h_trace_z[((size_t)ch * T + (T - 1)) * ZW + w];
      cur_set[slot] = 1;
    }
  }
  collected = true;
  return PA_OK;
}

int Engine::fetch_traces() {
  if (!sampled) {
    set_error("nothing sampled yet");
    return PA_ERR_STATE;
  }
  if (device_summaries) {
    // traces stay in HBM; they come to the host only on demand (pa_engine_get_trace, VERBOSE output), all chains in one copy
    if (traces_on_host || n_chains == 0) return PA_OK;
    PA_CUDA(cudaSetDevice(device));
    auto reserve_host = [&](void*& p, size_t& cap, size_t bytes) -> cudaError_t {
      if (bytes <= cap && p) return cudaSuccess;
      if (p) cudaFreeHost(p);
      p = nullptr;
      cap = 0;
      cudaError_t e = cudaHostAlloc(&p, bytes, cudaHostAllocDefault);
      if (e == cudaSuccess) cap = bytes;
      return e;
    };
    const size_t bd = sizeof(double) * 6 * (size_t)T * n_chains, bz = sizeof(uint32_t) * (size_t)ZW * T * n_chains;
    PA_CUDA(reserve_host((void*&)h_trace_d, cap_h_trace_d, bd));
    PA_CUDA(reserve_host((void*&)h_trace_z, cap_h_trace_z, bz));
    PA_CUDA(cudaMemcpyAsync(h_trace_d, d_trace_d, bd, cudaMemcpyDeviceToHost, stream));
    PA_CUDA(cudaMemcpyAsync(h_trace_z, d_trace_z, bz, cudaMemcpyDeviceToHost, stream));
    PA_CUDA(cudaStreamSynchronize(stream));
    d2h_bytes += bd + bz;
    traces_on_host = true;
    return PA_OK;
  }
  if (collected || n_chains == 0) {
    collected = true;
    return PA_OK;
  }
  const int rc = fetch_range(0, n_chains);
  return rc != PA_OK ? rc : fetch_finish();
}

int Engine::get_trace(int32_t i, int32_t mdl, ChainTraceHost& tr) {
  int rc = fetch_traces();
  if (rc != PA_OK) return rc;
  if (i < 0 || i >= (int)run_ids.size() || mdl < 0 || mdl > 2) {
    set_error("bad chain index");
    return PA_ERR_ARG;
  }
  const int ch = chain_of[(size_t)i * 3 + mdl];
  if (ch < 0) {
    set_error("element was filtered; it has no chain");
    return PA_ERR_ARG;
  }
  return unpack_trace(ch, tr);
}

// host copy of chain `ch` (device order) from the downloaded trace buffers
int Engine::unpack_trace(int ch, ChainTraceHost& tr) const {
  return unpack_trace_from(ch, &h_trace_d[(size_t)ch * 6 * T], &h_trace_z[(size_t)ch * T * ZW], tr);
}

int Engine::unpack_trace_from(int ch, const double* d, const uint32_t* z, ChainTraceHost& tr) const {
  const ElemPlan& pl = plans[run_ids[chain_elem_pos[ch]]];
  const int N = model.tree.N, n = (int)pl.nodes.size();
  tr.loglik.assign(d, d + T);
  tr.n_rate.assign(d + T, d + 2 * T);
  tr.c_rate.assign(d + 2 * T, d + 3 * T);
  tr.g_rate.assign(d + 3 * T, d + 4 * T);
  tr.l_rate.assign(d + 4 * T, d + 5 * T);
  tr.l2_rate.assign(d + 5 * T, d + 6 * T);
  tr.Z.assign((size_t)T * N, -1);
  const int32_t* nodes = pl.nodes.data();
  for (int m = 0; m < T; m++) {
    int8_t* row = &tr.Z[(size_t)m * N];
    const uint32_t* zw = z + (size_t)m * ZW;
    for (int k0 = 0; k0 < n; k0 += 16) {  // one packed word = 16 nodes of the pruned tree, 2 bits each
      uint32_t word = zw[k0 >> 4];
      const int kn = std::min(16, n - k0);
      for (int k = 0; k < kn; k++, word >>= 2) row[nodes[k0 + k]] = (int8_t)(word & 3u);
    }
  }
  return PA_OK;
}

// ------------------------------------------------------------------------------------------ summaries
namespace {

struct Eval2Work {  // state carried to the (rare) fallback that needs pruning on the device (bpp_c2.cpp:199-316)
  bool need_fallback = false;
  std::vector<std::vector<int8_t>> configZ;
  std::vector<double> priorZ;
  double rn = 0, rc = 0, varc = 0, varn = 0, cov = 0, meann = 0, meanc = 0;
};

double prior_z_config(const Engine& eng, const ElemPlan& pl, int mdl, const int8_t* cfgZ) {  // bpp_c2.cpp:131-163
  const pa_params& P = eng.model.prm;
  const int N = eng.model.tree.N;
  const uint8_t* fixZ = &pl.fixZ[(size_t)mdl * N];
  double nZ[9] = {0}, nF[9] = {0};
  const int n = (int)pl.nodes.size();
  for (int i = 0; i < n - 1; i++) {
    const int s = pl.nodes[i];
    const int k = cfgZ[s] * 3 + cfgZ[pl.parent2[s]];
    nZ[k] += 1;
    if (!fixZ[s]) nF[k] += 1;
  }
  double pz = lnbeta(P.hyper_grate_a + nZ[3] + nZ[6], P.hyper_grate_b + nZ[0]) - lnbeta(P.hyper_grate_a, P.hyper_grate_b);
  pz += lnbeta(P.hyper_lrate_a + nF[7], P.hyper_lrate_b + nF[4]) - lnbeta(P.hyper_lrate_a, P.hyper_lrate_b);
  if (P.hyper_lrate2_a > 0) pz += lnbeta(P.hyper_lrate2_a + nF[5], P.hyper_lrate2_b + nF[8]) - lnbeta(P.hyper_lrate2_a, P.hyper_lrate2_b);
  const double prior_z[3] = {std::log(0.5), std::log(0.5), -INFINITY};
  pz += prior_z[cfgZ[pl.root]];
  return pz;
}

// everything except the pruning fallback; returns VB (possibly to be replaced by the fallback)
double eval2_main(const Engine& eng, const ElemPlan& pl, int mdl, const ChainTraceHost& tr, const std::vector<double>& full,
                  const int8_t* Max_Z, Eval2Work& work) {
  const pa_params& P = eng.model.prm;
  const int N = eng.model.tree.N, num_burn = P.burnin, num_mcmc = P.mcmc, T = num_burn + num_mcmc;
  const bool null_model = (mdl == PA_M0);  // resZ == 0
  std::vector<int> uncertain;
  {
    std::vector<uint8_t> differs(N, 0);  // a node is uncertain iff some post-burn-in row differs from Max_Z there
    for (int i = num_burn; i < T; i++) {
      const int8_t* row = &tr.Z[(size_t)i * N];
      for (int s = 0; s < N; s++) differs[s] |= (uint8_t)(row[s] != Max_Z[s]);
    }
    for (int s = 0; s < N; s++)
      if (differs[s]) uncertain.push_back(s);
  }
  std::vector<const int8_t*> configZ{Max_Z};
  std::vector<int> num{0};
  std::vector<double> var_c{0}, mean_c{0}, var_n{0}, mean_n{0}, cov{0}, sum_full{0};
  for (int i = T - 1; i >= num_burn; i--) {
    const int8_t* zi = &tr.Z[(size_t)i * N];
    size_t j = 0;
    for (; j < configZ.size(); j++) {
      size_t k = 0;
      for (; k < uncertain.size(); k++)
        if (zi[uncertain[k]] != configZ[j][uncertain[k]]) break;
      if (k == uncertain.size()) break;
    }
    if (j == configZ.size()) {
      configZ.push_back(zi);
      num.push_back(0); var_c.push_back(0); mean_c.push_back(0); var_n.push_back(0); mean_n.push_back(0); cov.push_back(0); sum_full.push_back(0);
    }
    num[j]++;
    var_c[j] += std::pow(tr.c_rate[i], 2);
    mean_c[j] += tr.c_rate[i];
    if (!null_model) {
      var_n[j] += std::pow(tr.n_rate[i], 2);
      mean_n[j] += tr.n_rate[i];
      cov[j] += tr.c_rate[i] * tr.n_rate[i];
    }
    sum_full[j] += full[i] / num_mcmc;
  }
  double VB = 0, logdet;
  int effective_sum = 0;
  const int numConf = (int)configZ.size();
  for (int i = 0; i < numConf; i++) {
    const double w = (double)num[i] / num_mcmc;
    if (w < 0.02) continue;
    const double priorZ = prior_z_config(eng, pl, mdl, configZ[i]);
    sum_full[i] += num[i] * priorZ / num_mcmc;
    if (null_model) {
      logdet = std::log(var_c[i] - std::pow(mean_c[i], 2) / num[i]) - std::log(num[i]);
    } else {
      const double vc = var_c[i] - std::pow(mean_c[i], 2) / num[i];
      const double vn = var_n[i] - std::pow(mean_n[i], 2) / num[i];
      const double cv = cov[i] - mean_n[i] * mean_c[i] / num[i];
      logdet = std::log(vc * vn - cv * cv) - 2 * std::log(num[i]);
    }
    if (std::isnan(logdet)) continue;
    effective_sum += num[i];
    VB -= w * std::log(num[i]);
    VB += sum_full[i];
    VB += 0.5 * logdet * num[i] / num_mcmc;
  }
  if (effective_sum == 0) {
    work.need_fallback = true;
    for (int i = 0; i < numConf; i++) {
      work.varc += var_c[i]; work.meanc += mean_c[i];
      if (!null_model) { work.meann += mean_n[i]; work.varn += var_n[i]; work.cov += cov[i]; }
      work.configZ.emplace_back(configZ[i], configZ[i] + N);
      work.priorZ.push_back(prior_z_config(eng, pl, mdl, configZ[i]));
    }
    work.rn = work.meann / num_mcmc;
    work.rc = work.meanc / num_mcmc;
    return NAN;
  }
  if (!null_model) return VB * ((double)num_mcmc / effective_sum) + std::log(num_mcmc) + std::log(2 * M_PI) + 1;
  return VB * ((double)num_mcmc / effective_sum) + std::log(num_mcmc) + 0.5 * (std::log(2 * M_PI) + 1);
}

double eval2_fallback_finish(const Engine& eng, int mdl, const Eval2Work& w, const double* logliks) {  // bpp_c2.cpp:289-314
  const pa_params& P = eng.model.prm;
  const int num_mcmc = P.mcmc;
  const bool null_model = (mdl == PA_M0);
  std::vector<double> s(w.configZ.size());
  double mx = -INFINITY;
  for (size_t i = 0; i < s.size(); i++) {
    s[i] = logliks[i] + w.priorZ[i];
    if (s[i] > mx) mx = s[i];
  }
  double VB;
  if (mx == -INFINITY) VB = -INFINITY;
  else {
    double acc = 0;
    for (double v : s) acc += std::exp(v - mx);
    VB = std::log(acc) + mx;
  }
  VB += std::log(gamma_pdf(w.rc, P.conserve_prior_a, P.conserve_prior_b));
  if (!null_model) VB += std::log(gamma_pdf(w.rn, P.acce_prior_a, P.acce_prior_b));
  double varc = w.varc, varn = w.varn, cv = w.cov;
  if (null_model) {
    varc -= std::pow(w.meanc, 2) / num_mcmc;
    const double logdet = std::log(varc) - std::log(num_mcmc);
    VB += 0.5 * logdet + 0.5 * std::log(2 * M_PI);
  } else {
    varn -= std::pow(w.meann, 2) / num_mcmc;
    varc -= std::pow(w.meanc, 2) / num_mcmc;
    cv -= w.meann * w.meanc / num_mcmc;
    const double logdet = std::log(varc * varn - cv * cv) - 2 * std::log(num_mcmc);
    VB += 0.5 * logdet + std::log(2 * M_PI);
  }
  return VB;
}

// MonitorChain, Eval2 (all but its pruning fallback) and Output_init's medians / Z frequencies of one chain from its host trace
static void host_chain_summary(const Engine& eng, const ElemPlan& pl, int i, int mdl, ChainTraceHost& tr, pa_results* out, Eval2Work& work, double& logml_out,
                           double lg_ca, double lg_na) {
  const pa_params& P = eng.model.prm;
  const int N = eng.model.tree.N, num_burn = P.burnin, num_mcmc = P.mcmc, T = num_burn + num_mcmc;
  // MonitorChain: full log-likelihood and its arg-max after burn-in (bpp_c.cpp:768-814)
  std::vector<double> full(T);
  double max_ll = -INFINITY;
  int max_m = -1;
  for (int m = num_burn; m < T; m++) {  // only the post-burn-in rows are used (arg-max here, Eval2's averages below)
    full[m] = tr.loglik[m] + 0.0;
    full[m] += std::log(gamma_pdf_lg(tr.c_rate[m], P.conserve_prior_a, P.conserve_prior_b, lg_ca));
    if (mdl != PA_M0) full[m] += std::log(gamma_pdf_lg(tr.n_rate[m], P.acce_prior_a, P.acce_prior_b, lg_na));
    if (full[m] > max_ll) {
      max_ll = full[m];
      max_m = m;
    }
  }
  std::vector<int8_t> Max_Z(N, 0);
  if (max_m >= 0) Max_Z.assign(&tr.Z[(size_t)max_m * N], &tr.Z[(size_t)max_m * N] + N);
  if (out->loglik_max) out->loglik_max[i * 3 + mdl] = max_ll;
  if (out->max_Z) memcpy(out->max_Z + ((size_t)i * 3 + mdl) * N, Max_Z.data(), N);
  if (max_m >= 0) logml_out = eval2_main(eng, pl, mdl, tr, full, Max_Z.data(), work);
  else logml_out = NAN;
  // Output_init: medians of the sorted FULL traces, Z frequencies after burn-in (bpp_c2.cpp:742-781)
  if (out->median_rates) {
    double* md = &out->median_rates[((size_t)i * 3 + mdl) * 5];
    std::vector<double>* cols[5] = {&tr.n_rate, &tr.c_rate, &tr.g_rate, &tr.l_rate, &tr.l2_rate};
    for (int k = 0; k < 5; k++) {  // the element the reference reads after sorting the whole trace (bpp_c2.cpp:742-760)
      std::nth_element(cols[k]->begin(), cols[k]->begin() + cols[k]->size() / 2, cols[k]->end());
      md[k] = (*cols[k])[cols[k]->size() / 2];
    }
  }
  if (out->post_Z) {
    double* pz = &out->post_Z[((size_t)i * 3 + mdl) * N * 4];
    std::vector<int> cnt((size_t)N * 4, 0);
    {
      // per-value passes over the row-major trace (vectorisable compare-and-add) instead of a scattered histogram
      std::vector<int> c0(N, 0), c1(N, 0), c2(N, 0);
      for (int m = num_burn; m < T; m++) {
        const int8_t* row = &tr.Z[(size_t)m * N];
        for (int s = 0; s < N; s++) {
          c0[s] += row[s] == 0;
          c1[s] += row[s] == 1;
          c2[s] += row[s] == 2;
        }
      }
      for (int s = 0; s < N; s++) {
        cnt[s * 4 + 1] = c0[s];
        cnt[s * 4 + 2] = c1[s];
        cnt[s * 4 + 3] = c2[s];
        cnt[s * 4] = num_mcmc - c0[s] - c1[s] - c2[s];  // Z == -1: the node is outside the element's pruned tree
      }
    }
    for (int s = 0; s < N; s++) {
      if (pl.missing[s]) cnt[s * 4] = num_mcmc;
      for (int k = 0; k < 4; k++) pz[s * 4 + k] = (double)cnt[s * 4 + k] / num_mcmc;
    }
  }
}

// Eval2 fallback (no configuration carries >= 2% weight with a finite log-det): exact log-likelihood of every
// configuration at the posterior-mean rates, by the replay kernel, in one batch (bpp_c2.cpp:199-316)
static int eval2_fallback_batch(Engine& eng, const std::vector<Eval2Work>& works, std::vector<double>& logml) {
  std::vector<int32_t> elem;
  std::vector<int8_t> Z;
  std::vector<double> nr, cr;
  const int n_chains = (int)works.size();
  for (int ch = 0; ch < n_chains; ch++) {
    const Eval2Work& w = works[ch];
    if (!w.need_fallback) continue;
    for (size_t k = 0; k < w.configZ.size(); k++) {
      elem.push_back(eng.run_ids[eng.chain_elem_pos[ch]]);
      Z.insert(Z.end(), w.configZ[k].begin(), w.configZ[k].end());
      nr.push_back(eng.chain_model[ch] == PA_M0 ? 1.0 : w.rn);
      cr.push_back(w.rc);
    }
  }
  if (elem.empty()) return PA_OK;
  std::vector<double> ll(elem.size());
  const int rc = eng.loglik_replay((int32_t)elem.size(), elem.data(), Z.data(), nr.data(), cr.data(), ll.data());
  if (rc != PA_OK) return rc;
  size_t off = 0;
  for (int ch = 0; ch < n_chains; ch++) {
    const Eval2Work& w = works[ch];
    if (!w.need_fallback) continue;
    logml[ch] = eval2_fallback_finish(eng, eng.chain_model[ch], w, &ll[off]);
    off += w.configZ.size();
  }
  return PA_OK;
}

}  // namespace

int Engine::collect(pa_results* out) {
  PA_NVTX_RANGE("pa_engine_collect");
  const bool timing = getenv("PA_TIMING") != nullptr;
  const double t_start = omp_get_wtime();
  if (!sampled) {
    set_error("nothing sampled yet");
    return PA_ERR_STATE;
  }
  int rc = PA_OK;
  double t_wait = 0.0;
  const int n_ids = (int)run_ids.size();
  if (out->n < n_ids || out->N != model.tree.N) {
    set_error("pa_results too small for this run");
    return PA_ERR_ARG;
  }
  if (device_summaries) return collect_device(out);
  const pa_params& P = model.prm;
  const int N = model.tree.N, num_burn = P.burnin, num_mcmc = P.mcmc;
  for (int i = 0; i < n_ids; i++) {
    const bool filt = plans[run_ids[i]].filtered;
    if (out->filtered) out->filtered[i] = filt;
    if (out->loglik) for (int k = 0; k < 3; k++) out->loglik[i * 3 + k] = 0;
    if (out->loglik_max) for (int k = 0; k < 3; k++) out->loglik_max[i * 3 + k] = 0;
    if (out->max_Z) memset(out->max_Z + (size_t)i * 3 * N, 0, (size_t)3 * N);
    if (out->median_rates) for (int k = 0; k < 15; k++) out->median_rates[(size_t)i * 15 + k] = 0;
    if (out->post_Z) for (size_t k = 0; k < (size_t)3 * N * 4; k++) out->post_Z[(size_t)i * 3 * N * 4 + k] = 0;
  }
  const double lg_ca = std::lgamma(P.conserve_prior_a), lg_na = std::lgamma(P.acce_prior_a);
  std::vector<Eval2Work> works(n_chains);
  std::vector<double> logml(n_chains, 0.0);
  int err = PA_OK;
  const int nt = std::max(1, P.num_thread);
  // Chains are summarised in the order the kernel finishes them (device order), chunk by chunk, while later chains are still being
  // sampled: wait for a chunk's completion flags, download its traces on the copy stream, summarise it on the host threads.
  const bool streaming = !collected && n_chains > 0;
  const int chunk = streaming ? 192 : std::max(1, n_chains);
  for (int lo = 0, hi = 0; lo < n_chains; lo = hi) {
    hi = std::min(n_chains, lo + chunk);
    if (streaming) {
      const double t0 = omp_get_wtime();
      rc = fetch_range(lo, hi);
      if (rc != PA_OK) return rc;
      hi = std::max(hi, fetched_prefix);
      t_wait += omp_get_wtime() - t0;
    }
#pragma omp parallel for schedule(dynamic, 4) num_threads(nt)
    for (int ch = lo; ch < hi; ch++) {
      const int i = chain_elem_pos[ch], mdl = chain_model[ch];
      const ElemPlan& pl = plans[run_ids[i]];
      ChainTraceHost tr;
      if (unpack_trace(ch, tr) != PA_OK) {
#pragma omp critical
        err = PA_ERR_STATE;
        continue;
      }
      host_chain_summary(*this, pl, i, mdl, tr, out, works[ch], logml[ch], lg_ca, lg_na);
    }
  }  // chunks
  if (err != PA_OK) return err;
  rc = fetch_finish();
  if (rc != PA_OK) return rc;
  const double t_sum = omp_get_wtime();
  rc = eval2_fallback_batch(*this, works, logml);
  if (rc != PA_OK) return rc;
  if (out->loglik)
    for (int ch = 0; ch < n_chains; ch++) out->loglik[chain_elem_pos[ch] * 3 + chain_model[ch]] = logml[ch];
  if (timing)
    fprintf(stderr, "[pa timing] collect: waiting for chains + trace download %.3f s, MonitorChain/Eval2/summaries %.3f s (%d threads, overlapped with the kernel), Eval2 fallback %.3f s\n",
            t_wait, t_sum - t_start - t_wait, nt, omp_get_wtime() - t_sum);
  return PA_OK;
}


// Summaries from the device records (pa_summary.cu); traces stay in HBM.  Chains whose Eval2 needs the host (pruning fallback at the
// posterior-mean rates, more Z configurations than the device table holds) have their own trace downloaded and go through the
// host path above.
int Engine::collect_device(pa_results* out) {
  const bool timing = getenv("PA_TIMING") != nullptr;
  const double t_start = omp_get_wtime();
  const pa_params& P = model.prm;
  const int N = model.tree.N, num_mcmc = P.mcmc, n_ids = (int)run_ids.size();
  for (int i = 0; i < n_ids; i++) {
    const bool filt = plans[run_ids[i]].filtered;
    if (out->filtered) out->filtered[i] = filt;
    if (out->loglik) for (int k = 0; k < 3; k++) out->loglik[i * 3 + k] = 0;
    if (out->loglik_max) for (int k = 0; k < 3; k++) out->loglik_max[i * 3 + k] = 0;
    if (out->max_Z) memset(out->max_Z + (size_t)i * 3 * N, 0, (size_t)3 * N);
    if (out->median_rates) for (int k = 0; k < 15; k++) out->median_rates[(size_t)i * 15 + k] = 0;
    if (out->post_Z) for (size_t k = 0; k < (size_t)3 * N * 4; k++) out->post_Z[(size_t)i * 3 * N * 4 + k] = 0;
  }
  if (n_chains == 0) {
    collected = true;
    return PA_OK;
  }
  PA_CUDA(cudaSetDevice(device));
  unsigned long long hc[24];
  if (!collected) {
    PA_CUDA(cudaMemcpyAsync(h_summary, d_summary, rec_bytes * (size_t)n_chains, cudaMemcpyDeviceToHost, stream));
    PA_CUDA(cudaMemcpyAsync(hc, d_counters, sizeof(hc), cudaMemcpyDeviceToHost, stream));
    PA_CUDA(cudaStreamSynchronize(stream));
    d2h_bytes += rec_bytes * (size_t)n_chains + sizeof(hc);
    counters.node_pattern_updates = hc[0];
    counters.root_evals = hc[1];
    counters.pt_builds = hc[2];
    counters.history_draws = hc[3];
    counters.chain_iterations = hc[4];
    counters.node_site_updates = hc[5];
    for (int k = 0; k < 10; k++) phase_cycles[k] = hc[8 + k];
  }
  const double t_copy = omp_get_wtime();
  const int nt = std::max(1, P.num_thread);
  std::vector<int> host_chains;
  {
    const size_t C = plans.size();
    if (cur_set.size() != C * 3 || cur_z.size() != C * 3 * (size_t)ZW) {
      cur_set.assign(C * 3, 0);
      cur_rates.assign(C * 15, 0.0);
      cur_z.assign(C * 3 * (size_t)ZW, 0u);
    }
  }
#pragma omp parallel for schedule(static) num_threads(nt)
  for (int ch = 0; ch < n_chains; ch++) {
    const int i = chain_elem_pos[ch], mdl = chain_model[ch];
    const ElemPlan& pl = plans[run_ids[i]];
    const int n = (int)pl.nodes.size();
    const unsigned char* rec = h_summary + (size_t)ch * rec_bytes;
    const SummaryHead* hd = reinterpret_cast<const SummaryHead*>(rec);
    const uint32_t* maxz = reinterpret_cast<const uint32_t*>(rec + sizeof(SummaryHead));
    const uint32_t* lastz = maxz + ZW;
    const uint16_t* cnt = reinterpret_cast<const uint16_t*>(lastz + ZW);
    if (out->loglik_max) out->loglik_max[i * 3 + mdl] = hd->max_loglik;
    if (out->loglik) out->loglik[i * 3 + mdl] = hd->logml;
    if (out->max_Z) {
      int8_t* mz = out->max_Z + ((size_t)i * 3 + mdl) * N;
      if (hd->max_m >= 0) {
        memset(mz, -1, N);
        for (int k = 0; k < n; k++) mz[pl.nodes[k]] = (int8_t)((maxz[k >> 4] >> (2 * (k & 15))) & 3u);
      }
    }
    if (out->median_rates)
      for (int k = 0; k < 5; k++) out->median_rates[((size_t)i * 3 + mdl) * 5 + k] = hd->med[k];
    if (out->post_Z) {  // bpp_c2.cpp:757-781: frequencies of Z = -1, 0, 1, 2; missing nodes force the first column to 1
      double* pz = &out->post_Z[((size_t)i * 3 + mdl) * N * 4];
      for (int s = 0; s < N; s++) pz[s * 4] = 1.0;  // nodes outside the pruned tree: Z == -1 in every row
      for (int k = 0; k < n; k++) {
        const int s = pl.nodes[k];
        const int c0 = cnt[k * 3], c1 = cnt[k * 3 + 1], c2 = cnt[k * 3 + 2];
        pz[s * 4 + 0] = pl.missing[s] ? 1.0 : (double)(num_mcmc - c0 - c1 - c2) / num_mcmc;
        pz[s * 4 + 1] = (double)c0 / num_mcmc;
        pz[s * 4 + 2] = (double)c1 / num_mcmc;
        pz[s * 4 + 3] = (double)c2 / num_mcmc;
      }
    }
    // the last state of every chain becomes the warm start of its next run (Gibbs with UpHyper, bpp_c.cpp:474-485)
    const size_t slot = (size_t)run_ids[i] * 3 + mdl;
    for (int j = 0; j < 5; j++) cur_rates[slot * 5 + j] = hd->last[j];
    for (int w = 0; w < ZW; w++) cur_z[slot * ZW + w] = lastz[w];
    cur_set[slot] = 1;
    if (hd->flags & SUMMARY_NEEDS_HOST) {
#pragma omp critical
      host_chains.push_back(ch);
    }
  }
  const double t_fin = omp_get_wtime();
  counters.host_eval_chains = host_chains.size();
  // chains that need the host: download their traces (one chain at a time: they are rare) and run the host summary + fallback
  if (!host_chains.empty()) {
    std::sort(host_chains.begin(), host_chains.end());
    const double lg_ca = std::lgamma(P.conserve_prior_a), lg_na = std::lgamma(P.acce_prior_a);
    std::vector<Eval2Work> works(n_chains);
    std::vector<double> logml(n_chains, 0.0);
    const size_t nh = host_chains.size(), per_d = (size_t)6 * T, per_z = (size_t)ZW * T;
    std::vector<double> bd;
    std::vector<uint32_t> bz;
    if (!traces_on_host) {  // their traces, queued back to back, one synchronisation
      bd.resize(nh * per_d);
      bz.resize(nh * per_z);
      for (size_t k = 0; k < nh; k++) {
        const int ch = host_chains[k];
        PA_CUDA(cudaMemcpyAsync(bd.data() + k * per_d, d_trace_d + (size_t)ch * per_d, sizeof(double) * per_d, cudaMemcpyDeviceToHost, stream));
        PA_CUDA(cudaMemcpyAsync(bz.data() + k * per_z, d_trace_z + (size_t)ch * per_z, sizeof(uint32_t) * per_z, cudaMemcpyDeviceToHost, stream));
      }
      PA_CUDA(cudaStreamSynchronize(stream));
      d2h_bytes += nh * (sizeof(double) * per_d + sizeof(uint32_t) * per_z);
    }
    pa_results scratch = *out;
    scratch.median_rates = nullptr;  // medians / frequencies / arg-max already come from the device record
    scratch.post_Z = nullptr;
    scratch.max_Z = nullptr;
    scratch.loglik_max = nullptr;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt)
    for (int k = 0; k < (int)nh; k++) {
      const int ch = host_chains[k];
      const double* src_d = traces_on_host ? &h_trace_d[(size_t)ch * per_d] : bd.data() + (size_t)k * per_d;
      const uint32_t* src_z = traces_on_host ? &h_trace_z[(size_t)ch * per_z] : bz.data() + (size_t)k * per_z;
      ChainTraceHost tr;
      unpack_trace_from(ch, src_d, src_z, tr);
      const int i = chain_elem_pos[ch], mdl = chain_model[ch];
      host_chain_summary(*this, plans[run_ids[i]], i, mdl, tr, &scratch, works[ch], logml[ch], lg_ca, lg_na);
    }
    const int rc = eval2_fallback_batch(*this, works, logml);
    if (rc != PA_OK) return rc;
    if (out->loglik)
      for (int ch : host_chains) out->loglik[chain_elem_pos[ch] * 3 + chain_model[ch]] = logml[ch];
  }
  collected = true;
  if (timing)
    fprintf(stderr, "[pa timing] collect (device summaries): wait + %zu B of records %.3f s, unpack %.3f s, %zu chains through the host path %.3f s\n",
            rec_bytes * (size_t)n_chains, t_copy - t_start, t_fin - t_copy, host_chains.size(), omp_get_wtime() - t_fin);
  return PA_OK;
}

}  // namespace pa
