// pa_capi.cpp -- extern "C" surface declared in include/phyloacc_b200.h.  No exceptions cross this boundary.
#include <cstring>
#include <new>
#include <string>

#include "pa_engine.h"
#include "pa_internal.h"

thread_local std::string g_last;
namespace pa {
void set_last_error(const std::string& s) { g_last = s; }
}
namespace {
template <class F>
int guarded(F&& f) {
  try {
    const int rc = f();
    if (rc != PA_OK) g_last = pa::get_error();
    return rc;
  } catch (const std::bad_alloc&) {
    g_last = "out of host memory";
    return PA_ERR_LIMIT;
  } catch (const std::exception& e) {
    g_last = std::string("internal error: ") + e.what();
    return PA_ERR_ARG;
  } catch (...) {
    g_last = "unknown internal error";
    return PA_ERR_ARG;
  }
}
}  // namespace

extern "C" {

const char* pa_last_error(void) { return g_last.c_str(); }
const char* pa_version(void) { return "phyloacc_b200 0.1 (sm_100a)"; }

int pa_engine_create(const pa_tree* tree, const pa_subst* subst, const pa_params* params, int device, pa_engine** out) {
  if (!tree || !subst || !params || !out) {
    g_last = "null argument";
    return PA_ERR_ARG;
  }
  *out = nullptr;
  return guarded([&]() {
    pa_engine* e = new pa_engine();
    const int rc = e->eng.init(tree, subst, params, device);
    if (rc != PA_OK) {
      g_last = pa::get_error();
      delete e;
      return rc;
    }
    *out = e;
    return PA_OK;
  });
}

void pa_engine_destroy(pa_engine* e) { delete e; }

int pa_engine_set_stream(pa_engine* e, void* cuda_stream) {
  if (!e) return PA_ERR_ARG;
  e->eng.stream = (cudaStream_t)cuda_stream;
  return PA_OK;
}

int pa_engine_load_alignment(pa_engine* e, const uint8_t* X, int64_t G, int32_t C, const int64_t* elem_start, const int32_t* elem_len) {
  if (!e || (!X && G > 0) || (C > 0 && (!elem_start || !elem_len))) {
    g_last = "null argument";
    return PA_ERR_ARG;
  }
  return guarded([&]() { return e->eng.load_alignment(X, G, C, elem_start, elem_len); });
}

int pa_engine_num_elements(const pa_engine* e) { return e ? (int)e->eng.plans.size() : PA_ERR_ARG; }

int pa_engine_element_info(const pa_engine* e, int32_t elem, pa_element_info* out) {
  if (!e || !out || elem < 0 || elem >= (int)e->eng.plans.size()) {
    g_last = "bad element id";
    return PA_ERR_ARG;
  }
  const pa::ElemPlan& p = e->eng.plans[elem];
  const int N = e->eng.model.tree.N, S = e->eng.model.tree.S;
  out->filtered = p.filtered;
  out->GG = p.GG;
  out->P = p.P;
  out->root = p.root;
  out->n_nodes = (int32_t)p.nodes.size();
  out->n_slots = p.n_slots;
  if (p.filtered) return PA_OK;
  if (((out->pattern_codes || out->pattern_weights) && out->cap_P < p.P) || (out->site_pattern && out->cap_GG < p.GG)) {
    g_last = "pa_element_info: cap_P / cap_GG smaller than the element's pattern / column count";
    return PA_ERR_ARG;
  }
  if (out->missing) memcpy(out->missing, p.missing.data(), N);
  if (out->nodes) memcpy(out->nodes, p.nodes.data(), sizeof(int32_t) * p.nodes.size());
  if (out->children2) memcpy(out->children2, p.children2.data(), sizeof(int32_t) * 2 * N);
  if (out->parent2) memcpy(out->parent2, p.parent2.data(), sizeof(int32_t) * N);
  if (out->distances2) memcpy(out->distances2, p.distances2.data(), sizeof(double) * N);
  if (out->pattern_codes) memcpy(out->pattern_codes, p.pattern_codes.data(), (size_t)p.P * S);
  if (out->pattern_weights) memcpy(out->pattern_weights, p.pattern_weights.data(), sizeof(int32_t) * p.P);
  if (out->site_pattern) memcpy(out->site_pattern, p.site_pattern.data(), sizeof(int32_t) * p.GG);
  if (out->fixZ) memcpy(out->fixZ, p.fixZ.data(), (size_t)3 * N);
  return PA_OK;
}

int pa_engine_loglik_replay(pa_engine* e, int32_t n_states, const int32_t* elem, const int8_t* Z, const double* n_rate, const double* c_rate,
                            double* loglik_out) {
  if (!e || (n_states > 0 && (!elem || !Z || !n_rate || !c_rate || !loglik_out))) {
    g_last = "null argument";
    return PA_ERR_ARG;
  }
  return guarded([&]() { return e->eng.loglik_replay(n_states, elem, Z, n_rate, c_rate, loglik_out); });
}

int pa_engine_sample(pa_engine* e, const int32_t* ids, int32_t n_ids) {
  if (!e || (n_ids > 0 && !ids)) {
    g_last = "null argument";
    return PA_ERR_ARG;
  }
  return guarded([&]() { return e->eng.sample(ids, n_ids); });
}

int pa_engine_sample_ex(pa_engine* e, const int32_t* ids, int32_t n_ids, int32_t model_mask, int32_t warm_start, int32_t chain_iter) {
  if (!e || (n_ids > 0 && !ids) || (model_mask & ~7) || model_mask == 0 || chain_iter < 0) {
    g_last = "bad argument";
    return PA_ERR_ARG;
  }
  return guarded([&]() { return e->eng.sample(ids, n_ids, model_mask, warm_start != 0, chain_iter); });
}

int pa_engine_get_hyper(const pa_engine* e, pa_hyper* out) {
  if (!e || !out) {
    g_last = "null argument";
    return PA_ERR_ARG;
  }
  const pa_params& P = e->eng.model.prm;
  out->nprior_a = P.acce_prior_a; out->nprior_b = P.acce_prior_b; out->cprior_a = P.conserve_prior_a; out->cprior_b = P.conserve_prior_b;
  out->prior_l_a = P.hyper_lrate_a; out->prior_l_b = P.hyper_lrate_b; out->prior_g_a = P.hyper_grate_a; out->prior_g_b = P.hyper_grate_b;
  out->prior_l2_a = P.hyper_lrate2_a; out->prior_l2_b = P.hyper_lrate2_b;
  return PA_OK;
}

int pa_engine_set_hyper(pa_engine* e, const pa_hyper* in) {
  if (!e || !in) {
    g_last = "null argument";
    return PA_ERR_ARG;
  }
  pa_params& P = e->eng.model.prm;
  P.acce_prior_a = in->nprior_a; P.acce_prior_b = in->nprior_b; P.conserve_prior_a = in->cprior_a; P.conserve_prior_b = in->cprior_b;
  P.hyper_lrate_a = in->prior_l_a; P.hyper_lrate_b = in->prior_l_b; P.hyper_grate_a = in->prior_g_a; P.hyper_grate_b = in->prior_g_b;
  P.hyper_lrate2_a = in->prior_l2_a; P.hyper_lrate2_b = in->prior_l2_b;
  return PA_OK;
}

int pa_engine_sample_hyper(pa_engine* e, int32_t iter, const int32_t* ids, int32_t n_ids, pa_hyper* out) {
  if (!e || (n_ids > 0 && !ids)) {
    g_last = "null argument";
    return PA_ERR_ARG;
  }
  return guarded([&]() { return e->eng.sample_hyper(iter, ids, n_ids, out); });
}

int pa_engine_element_blob(pa_engine* e, int32_t elem, uint8_t* buf, int64_t cap, int64_t* size) {
  if (!e || !size) {
    g_last = "null argument";
    return PA_ERR_ARG;
  }
  return guarded([&]() { return e->eng.element_blob(elem, buf, cap, size); });
}

int pa_engine_hyper_sums(pa_engine* e, const int32_t* ids, int32_t n_ids, double* sums) {
  if (!e || !sums || (n_ids > 0 && !ids)) {
    g_last = "null argument";
    return PA_ERR_ARG;
  }
  return guarded([&]() { return e->eng.hyper_sums(ids, n_ids, sums); });
}

int pa_engine_sample_hyper_sums(pa_engine* e, int32_t iter, const double* sums, pa_hyper* out) {
  if (!e || !sums) {
    g_last = "null argument";
    return PA_ERR_ARG;
  }
  return guarded([&]() { return e->eng.sample_hyper_from_sums(iter, sums, out); });
}

int pa_engine_collect(pa_engine* e, pa_results* out) {
  if (!e || !out) {
    g_last = "null argument";
    return PA_ERR_ARG;
  }
  return guarded([&]() { return e->eng.collect(out); });
}

int pa_engine_run(pa_engine* e, const int32_t* ids, int32_t n_ids, pa_results* out) {
  const int rc = pa_engine_sample(e, ids, n_ids);
  if (rc != PA_OK) return rc;
  return pa_engine_collect(e, out);
}

int pa_engine_get_trace(pa_engine* e, int32_t i, int32_t model, pa_trace* out) {
  if (!e || !out) {
    g_last = "null argument";
    return PA_ERR_ARG;
  }
  return guarded([&]() {
    pa::ChainTraceHost tr;
    const int rc = e->eng.get_trace(i, model, tr);
    if (rc != PA_OK) return rc;
    const int T = (int)tr.loglik.size(), N = e->eng.model.tree.N;
    if (out->T < T || out->N < N) {
      pa::set_error("pa_trace buffers too small");
      return PA_ERR_ARG;
    }
    out->T = T;
    out->N = N;
    memcpy(out->loglik, tr.loglik.data(), sizeof(double) * T);
    memcpy(out->n_rate, tr.n_rate.data(), sizeof(double) * T);
    memcpy(out->c_rate, tr.c_rate.data(), sizeof(double) * T);
    memcpy(out->g_rate, tr.g_rate.data(), sizeof(double) * T);
    memcpy(out->l_rate, tr.l_rate.data(), sizeof(double) * T);
    memcpy(out->l2_rate, tr.l2_rate.data(), sizeof(double) * T);
    memcpy(out->Z, tr.Z.data(), (size_t)T * N);
    return PA_OK;
  });
}

int pa_gbgc_transition(int device, const pa_subst* subst, int32_t n, const double* sel, const double* bias, const double* t, double* rate_out,
                       double* P_out) {
  if (!subst || n < 0 || (n > 0 && (!sel || !bias || !t || !P_out))) {
    g_last = "null argument";
    return PA_ERR_ARG;
  }
  return guarded([&]() { return pa::gbgc_transition(device, subst->Q, n, sel, bias, t, rate_out, P_out); });
}

int pa_engine_gbgc_replay(pa_engine* e, int32_t n_states, const int32_t* elem, const int8_t* Z, const int8_t* B, const double* sel_n,
                          const double* sel_c, const double* gbc, double* loglik_out) {
  if (!e || (n_states > 0 && (!elem || !Z || !B || !sel_n || !sel_c || !gbc || !loglik_out))) {
    g_last = "null argument";
    return PA_ERR_ARG;
  }
  return guarded([&]() { return e->eng.gbgc_replay(n_states, elem, Z, B, sel_n, sel_c, gbc, loglik_out); });
}

int pa_engine_gbgc_integrate(pa_engine* e, int32_t n_states, const int32_t* elem, const int8_t* Z, const int8_t* B, const int8_t* fixZ,
                             const double* sel_n, const double* sel_c, const double* gbc, const pa_gbgc_consts* consts, double* post_out,
                             double* emis_out, double* root_out) {
  if (!e || !consts || (n_states > 0 && (!elem || !Z || !B || !fixZ || !sel_n || !sel_c || !gbc || !post_out || !emis_out || !root_out))) {
    g_last = "null argument";
    return PA_ERR_ARG;
  }
  return guarded([&]() { return e->eng.gbgc_integrate(n_states, elem, Z, B, fixZ, sel_n, sel_c, gbc, consts, post_out, emis_out, root_out); });
}

int pa_engine_gt_replay(pa_engine* e, const pa_gt_states* st, double* loglik_out, double* prior_out, double* logpz_out) {
  if (!e || !st || (st->n_states > 0 && (!st->elem || !st->gchild || !st->groot || !st->gmissing || !st->gheights || !st->Z || !st->n_rate || !st->c_rate ||
                                         !st->lin_off || !st->lin_species || !st->lin_node || !st->lin_next || !st->heights || !st->thetas))) {
    pa::set_error("pa_engine_gt_replay: null argument");
    return PA_ERR_ARG;
  }
  return guarded([&]() { return e->eng.gt_replay(st, loglik_out, prior_out, logpz_out); });
}

int pa_engine_gbgc_run(pa_engine* e, const int32_t* ids, int32_t n_ids, const int32_t* targets, int32_t n_targets, const pa_gbgc_params* prm,
                       int model_mask, pa_gbgc_results* out) {
  if (!e || !prm || !out || (n_targets > 0 && !targets)) {
    pa::set_error("pa_engine_gbgc_run: null argument");
    return PA_ERR_ARG;
  }
  return guarded([&]() { return e->eng.gbgc_run(ids, n_ids, targets, n_targets, prm, model_mask, out); });
}

int pa_engine_phase_cycles(pa_engine* e, uint64_t* out10) {
  if (!e || !out10) return PA_ERR_ARG;
  for (int k = 0; k < 10; k++) out10[k] = e->eng.phase_cycles[k];
  return PA_OK;
}

int pa_engine_counters(pa_engine* e, pa_counters* out) {
  if (!e || !out) return PA_ERR_ARG;
  *out = e->eng.counters;
  out->d2h_bytes = e->eng.d2h_bytes;
  return PA_OK;
}

}  // extern "C"
