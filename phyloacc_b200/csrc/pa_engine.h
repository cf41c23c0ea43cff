// pa_engine.h -- the engine object behind the C ABI (one per pa_engine handle).
#ifndef PA_ENGINE_H
#define PA_ENGINE_H

#include <cuda_runtime.h>

#include <vector>

#include "pa_internal.h"

namespace pa {

const std::string& get_error();

struct ChainTraceHost {  // host copy of one chain's trace (Output_sampling columns)
  std::vector<double> loglik, n_rate, c_rate, g_rate, l_rate, l2_rate;
  std::vector<int8_t> Z;  // [T][N] original node numbering, -1 for pruned nodes
};

struct Engine {
  Model model;
  int device = -1;
  int num_sms = 0, smem_optin = 0;
  cudaStream_t stream = nullptr;
  bool loaded = false;
  std::vector<ElemPlan> plans;
  std::vector<int64_t> blob_off;
  int n_max = 0, slots_max = 1, P_max = 0, GG_max = 0;
  size_t h2d_bytes = 0;

  // device-resident batch
  uint8_t* d_blobs = nullptr;
  int64_t* d_blob_off = nullptr;
  size_t cap_blobs = 0, cap_blob_off = 0;
  uint8_t* h_blobs = nullptr;  // pinned staging copy of the blobs (grow-only)
  size_t cap_h_blobs = 0;

  // last sampling run
  std::vector<int32_t> run_ids;          // element ids passed to sample()
  std::vector<int32_t> chain_elem_pos;   // per chain: position in run_ids
  std::vector<int32_t> chain_model;      // per chain: PA_M0..PA_M2
  std::vector<int32_t> chain_of;         // [n_ids][3] -> chain index or -1 (filtered)
  int n_chains = 0, T = 0, ZW = 0;
  int32_t* d_chain_list = nullptr;       // [n_chains][2] (element id, model) sorted by decreasing cost
  double* d_trace_d = nullptr;           // [n_chains][6][T]
  uint32_t* d_trace_z = nullptr;         // [n_chains][T][ZW] 2-bit packed Z by local node index
  uint8_t* d_arena = nullptr;            // per-CTA scratch (partials, scale counts, log P tables)
  size_t cap_chain_list = 0, cap_trace_d = 0, cap_trace_z = 0, cap_arena = 0, cap_counters = 0, cap_queue = 0;
  unsigned long long* d_counters = nullptr;
  int* d_queue = nullptr;
  // last state of every (element, model) chain that has run: the warm start of the next outer CHAIN iteration
  // (bpp.cur_Z / cur_nrate / cur_crate / cur_grate / cur_lrate / cur_lrate2, bpp_c.cpp:474-485)
  std::vector<double> cur_rates;         // [C][3][5] n, c, g, l, l2
  std::vector<uint32_t> cur_z;           // [C][3][ZW] packed local Z
  std::vector<uint8_t> cur_set;          // [C][3]
  double* d_init_rates = nullptr;
  uint32_t* d_init_z = nullptr;
  size_t cap_init_rates = 0, cap_init_z = 0;
  uint8_t* d_replay = nullptr;           // scratch of loglik_replay (grow-only)
  size_t cap_replay = 0;
  double* h_trace_d = nullptr;           // pinned host copies of the traces of the last run (grow-only)
  uint32_t* h_trace_z = nullptr;
  size_t cap_h_trace_d = 0, cap_h_trace_z = 0;
  // per-chain summaries computed on the device (pa_summary.cu): MonitorChain, medians, Z frequencies, Eval2; traces stay in HBM and
  // are downloaded only for chains that need the host (Eval2's pruning fallback), for pa_engine_get_trace and for VERBOSE output
  bool device_summaries = true;          // PA_HOST_SUMMARIES=1 selects round 1's host path (A/B, tests)
  unsigned char* d_summary = nullptr;
  unsigned char* h_summary = nullptr;    // pinned
  size_t cap_d_summary = 0, cap_h_summary = 0, rec_bytes = 0;
  bool traces_on_host = false;           // h_trace_* hold the last run's traces
  size_t d2h_bytes = 0;                  // bytes copied device -> host by the last collect (measured, for bench.py's e2e line)
  bool sampled = false, collected = false;
  // completion flags of the running launch, written by the kernel into mapped pinned memory (one per chain, device order), and the
  // stream the finished chains' traces are downloaded on while the kernel is still running
  int* h_done = nullptr;
  size_t cap_h_done = 0;
  int done_prefix = 0, fetched_prefix = 0;
  cudaStream_t copy_stream = nullptr;
  pa_counters counters{};
  unsigned long long phase_cycles[10] = {0};  // SM cycles spent per chain-kernel phase (thread 0 of every CTA, summed)

  ~Engine();
  void release_device();
  int init(const pa_tree* tree, const pa_subst* subst, const pa_params* params, int dev);
  // global_ids (optional): the batch-wide number of every element (the reference's CC) when this engine holds only a shard of the
  // batch -- it keys the element's random streams, so a chain draws the same numbers whichever GPU runs it
  int load_alignment(const uint8_t* X, int64_t G, int32_t C, const int64_t* elem_start, const int32_t* elem_len, const int32_t* global_ids = nullptr);
  int loglik_replay(int32_t n_states, const int32_t* elem, const int8_t* Z, const double* n_rate, const double* c_rate, double* out);
  int gbgc_replay(int32_t n_states, const int32_t* elem, const int8_t* Z, const int8_t* B, const double* sel_n, const double* sel_c, const double* gbc,
                  double* out);
  int gbgc_integrate(int32_t n_states, const int32_t* elem, const int8_t* Z, const int8_t* B, const int8_t* fixZ, const double* sel_n, const double* sel_c,
                     const double* gbc, const pa_gbgc_consts* k, double* post_out, double* emis_out, double* root_out);
  int gt_replay(const pa_gt_states* gs, double* loglik_out, double* prior_out, double* logpz_out);
  int gbgc_run(const int32_t* ids, int32_t n_ids, const int32_t* targets, int32_t n_targets, const pa_gbgc_params* prm, int model_mask, pa_gbgc_results* out);
  int sample(const int32_t* ids, int32_t n_ids, int model_mask = 7, bool warm_start = false, int chain_iter = 0);
  int sample_hyper(int iter, const int32_t* ids, int32_t n_ids, pa_hyper* out);
  int hyper_sums(const int32_t* ids, int32_t n_ids, double* sums /* [PA_HYPER_NSUMS] */);
  int sample_hyper_from_sums(int iter, const double* sums, pa_hyper* out);
  int element_blob(int32_t elem, uint8_t* buf, int64_t cap, int64_t* size);
  int fetch_traces();
  int fetch_range(int lo, int hi);  // wait for chains [fetched_prefix, hi) of the running launch and download their traces
  int fetch_finish();               // end of the launch: counters, warm-start states
  int unpack_trace(int ch, ChainTraceHost& tr) const;
  int unpack_trace_from(int ch, const double* d, const uint32_t* z, ChainTraceHost& tr) const;
  int collect_device(pa_results* out);
  int collect(pa_results* out);
  int get_trace(int32_t i, int32_t mdl, ChainTraceHost& tr);
};

}  // namespace pa

// the opaque handle of the C ABI
struct pa_engine {
  pa::Engine eng;
  pa::Params params;                       // set by pa_engine_create_from_config
  std::vector<std::string> element_names;  // BED column 0 (profile.cpp:85)
  std::vector<int32_t> ids;                // element ids selected by ID_FILE / BATCH (main.cpp:406-439)
  bool from_config = false;
};
#endif
