// pa_engine.cu -- the engine object behind the C ABI: batch upload, replay kernel, chain kernel launch,
// trace collection.  Kernels are written for sm_100a (B200): fp64 pruning over compressed site patterns,
// P(t) from the shared eigensystem, one CTA per chain with partials resident in L2.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "pa_chain.cuh"
#include "pa_device.cuh"
#include "pa_engine.h"
#include "pa_gbgc.cuh"
#include "pa_internal.h"
#include "pa_nvtx.h"

namespace pa {

#define PA_CUDA(call)                                                                              \
  do {                                                                                             \
    cudaError_t err__ = (call);                                                                    \
    if (err__ != cudaSuccess) {                                                                    \
      set_error(std::string("CUDA error: ") + cudaGetErrorString(err__) + " at " #call);          \
      return PA_ERR_CUDA;                                                                          \
    }                                                                                              \
  } while (0)

// ------------------------------------------------------------------------------------------ replay kernel
// One CTA per (element, Z, n_rate, c_rate) state: builds P(t) for every branch from its class, prunes every
// pattern with the shared-memory temp stack only (no global partials) and reduces sum_p w_p log L_p.
// GBGC = true: the GC-biased model of PhyloAcc-ST-GBGC (groundwork).  n_rate / c_rate are then the accelerated / conserved
// SELECTION COEFFICIENTS, `Bflag` says per branch whether gBGC acts on it, `gbc` is the element's bias, and every branch's P(t) is
// exp(t * Q o f(s + b B)) (BPP_C::initMCMC's six rate matrices + init_cache_TM, GBGC/SRC/bpp_c.cpp:196-204,1563-1621) by uniformisation.
struct Q16 { double q[16]; };
template <bool GBGC>
__global__ void __launch_bounds__(128) replay_kernel(const uint8_t* __restrict__ blobs, const int64_t* __restrict__ blob_off, SubstDev sd,
                                                      int n_states, const int32_t* __restrict__ elem, const int8_t* __restrict__ Z,
                                                      const double* __restrict__ n_rate, const double* __restrict__ c_rate, int N, int n_max,
                                                      int slots_max, double* __restrict__ out, const int8_t* __restrict__ Bflag,
                                                      const double* __restrict__ gbc, Q16 Q0) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int BD = blockDim.x, tid = threadIdx.x;
  const int st = blockIdx.x;
  if (st >= n_states) return;
  const int64_t off = blob_off[elem[st]];
  if (off < 0) {  // filtered element
    if (tid == 0) out[st] = CUDART_NAN;
    return;
  }
  const uint8_t* blob = blobs + off;
  const ElemView ev = make_view(blob);
  const ElemHeader* h = ev.h;
  const int n = h->n, ni = h->ni;
  // shared memory carve-up
  double* Pm = reinterpret_cast<double*>(smem_raw);                     // [n_max][16]
  double* stack = Pm + (size_t)n_max * 16;                              // [slots_max][4][BD]
  double* red = stack + (size_t)slots_max * 4 * BD;                     // [33]
  int* stack_sc = reinterpret_cast<int*>(red + 34);                     // [slots_max][BD]
  int16_t* s_child = reinterpret_cast<int16_t*>(stack_sc + (size_t)slots_max * BD);  // [n_max][2]
  int16_t* s_tiprow = s_child + 2 * n_max;
  int16_t* s_su = s_tiprow + n_max;
  int16_t* s_intidx = s_su + n_max;
  uint8_t* s_flags = reinterpret_cast<uint8_t*>(s_intidx + n_max);
  int8_t* s_slot = reinterpret_cast<int8_t*>(s_flags + n_max);
  __shared__ double s_pi[4];
  for (int i = tid; i < n; i += BD) {
    s_child[2 * i] = ev.child[2 * i];
    s_child[2 * i + 1] = ev.child[2 * i + 1];
    s_tiprow[i] = ev.tiprow[i];
    s_intidx[i] = ev.intidx[i];
    s_flags[i] = ev.flags[i];
    s_slot[i] = ev.slot[i];
    // P(t) of the branch above node i for its class (initMCMC, bpp_c.cpp:187-228; Update_Tg bpp_c.cpp:661-663)
    const int z = Z[(size_t)st * N + ev.orig[i]];
    double P16[16];
    if (GBGC) {
      const double sel = (z == 2) ? n_rate[st] : (z == 0 ? 0.0 : c_rate[st]);
      const double bias = Bflag[(size_t)st * N + ev.orig[i]] ? gbc[st] : 0.0;
      double Q[16];
      gbgc_rate(Q0.q, sel, bias, Q);
      expm_uniformised(Q, ev.dist[i], P16);
    } else {
      const double rate = (z == 2) ? n_rate[st] : (z == 0 ? 1.0 : c_rate[st]);
      build_P(sd, ev.dist[i], rate, true, P16);
    }
#pragma unroll
    for (int k = 0; k < 16; k++) Pm[i * 16 + k] = P16[k];
  }
  for (int i = tid; i < ni; i += BD) s_su[i] = ev.su[i];
  if (tid < 4) s_pi[tid] = sd.pi[tid];
  __syncthreads();
  TreeSmem ts{s_child, s_flags, s_tiprow, s_su, s_slot, s_intidx};
  double acc = 0.0;
  for (int p = tid; p < h->Ppad; p += BD) {
    if (p < h->P) {
      const double ll = prune_pattern<false, false>(h, ts, ev.codes, Pm, nullptr, stack, stack_sc, nullptr, nullptr, s_pi, p, BD, tid);
      acc += (double)ev.weights[p] * ll;
    }
  }
  const double tot = block_sum(acc, red, tid, BD);
  if (tid == 0) out[st] = tot;
}

static size_t replay_smem_bytes(int n_max, int slots_max, int BD) {
  size_t b = (size_t)n_max * 16 * 8 + (size_t)slots_max * 4 * BD * 8 + 34 * 8 + (size_t)slots_max * BD * 4;
  b += (size_t)n_max * (2 + 1 + 1 + 1) * 2 + (size_t)n_max * 2;
  return (b + 15) & ~size_t(15);
}

// ------------------------------------------------------------------------------------------ Engine
Engine::~Engine() { release_device(); }

void Engine::release_device() {
  if (device < 0) return;
  cudaSetDevice(device);
  auto fr = [](void*& p) { if (p) cudaFree(p); p = nullptr; };
  fr((void*&)d_blobs); fr((void*&)d_blob_off);
  cap_blobs = cap_blob_off = 0;
  fr((void*&)d_chain_list); fr((void*&)d_trace_d); fr((void*&)d_trace_z); fr((void*&)d_arena); fr((void*&)d_counters); fr((void*&)d_queue);
  fr((void*&)d_init_rates); fr((void*&)d_init_z); fr((void*&)d_replay); fr((void*&)d_summary);
  cap_d_summary = 0;
  if (h_summary) cudaFreeHost(h_summary);
  h_summary = nullptr;
  cap_h_summary = 0;
  cap_replay = 0;
  cap_init_rates = cap_init_z = 0;
  if (h_trace_d) cudaFreeHost(h_trace_d);
  if (h_trace_z) cudaFreeHost(h_trace_z);
  if (h_done) cudaFreeHost(h_done);
  if (h_blobs) cudaFreeHost(h_blobs);
  h_blobs = nullptr;
  cap_h_blobs = 0;
  h_done = nullptr;
  cap_h_done = 0;
  if (copy_stream) cudaStreamDestroy(copy_stream);
  copy_stream = nullptr;
  h_trace_d = nullptr;
  h_trace_z = nullptr;
  cap_h_trace_d = cap_h_trace_z = 0;
  cap_chain_list = cap_trace_d = cap_trace_z = cap_arena = cap_counters = cap_queue = 0;
}

int Engine::init(const pa_tree* tree, const pa_subst* subst, const pa_params* params, int dev) {
  Tree t;
  t.S = tree->S;
  t.N = tree->N;
  if (t.N != 2 * t.S - 1 || t.S < 2) {
    set_error("tree must be strictly binary with N = 2S-1");
    return PA_ERR_ARG;
  }
  t.parent.assign(tree->parent, tree->parent + t.N);
  t.children.assign(tree->children, tree->children + 2 * t.N);
  t.dist.assign(tree->distances, tree->distances + t.N);
  t.names.assign(t.N, "");
  memcpy(t.pi, subst->pi, sizeof(t.pi));
  memcpy(t.Q, subst->Q, sizeof(t.Q));
  if (!build_model(t, *params, model)) return PA_ERR_ARG;
  device = dev;
  if (device >= 0) {
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device >= count) {
      set_error("no usable CUDA device " + std::to_string(device) + " (this engine has no CPU fallback)");
      device = -1;
      return PA_ERR_CUDA;
    }
    PA_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    PA_CUDA(cudaGetDeviceProperties(&prop, device));
    num_sms = prop.multiProcessorCount;
    smem_optin = (int)prop.sharedMemPerBlockOptin;
  }
  return PA_OK;
}

int Engine::load_alignment(const uint8_t* X, int64_t G, int32_t C, const int64_t* elem_start, const int32_t* elem_len, const int32_t* global_ids) {
  PA_NVTX_RANGE("pa_engine_load_alignment");
  if (C < 0 || G < 0) {
    set_error("bad alignment dimensions");
    return PA_ERR_ARG;
  }
  for (int c = 0; c < C; c++)
    if (elem_start[c] < 0 || elem_len[c] < 0 || elem_start[c] + elem_len[c] > G) {
      set_error("element " + std::to_string(c) + " lies outside the alignment");
      return PA_ERR_ARG;
    }
  plans.assign(C, ElemPlan());
  cur_set.clear();  // a new batch: no chain has a previous state

  const int nt = std::max(1, model.prm.num_thread);
#pragma omp parallel for schedule(dynamic, 16) num_threads(nt)
  for (int c = 0; c < C; c++) plan_element(model, X, G, elem_start[c], elem_len[c], global_ids ? global_ids[c] : c, plans[c]);
  blob_off.assign(C, -1);
  size_t total = 0;
  n_max = 0; slots_max = 1; P_max = 0; GG_max = 0;
  for (int c = 0; c < C; c++) {
    if (plans[c].filtered) continue;
    if (plans[c].GG > 65535) {  // pattern / site indices and pattern weights are 16-bit on the device
      set_error("element " + std::to_string(c) + " has " + std::to_string(plans[c].GG) + " columns; the engine handles at most 65535 per element");
      loaded = false;
      return PA_ERR_LIMIT;
    }
    blob_off[c] = (int64_t)total;
    total += plans[c].blob.size();
    const ElemHeader* h = reinterpret_cast<const ElemHeader*>(plans[c].blob.data());
    n_max = std::max(n_max, h->n);
    slots_max = std::max(slots_max, h->n_slots);
    P_max = std::max(P_max, h->Ppad);
    GG_max = std::max(GG_max, h->GG);
  }
  loaded = true;
  if (device < 0) return PA_OK;
  PA_CUDA(cudaSetDevice(device));
  // only the batch itself is replaced; the buffers of a run (arenas, traces, pinned host copies) persist and grow on demand
  sampled = collected = false;
  // the blobs are gathered into a pinned staging buffer (grow-only, never cleared) by all host threads and go up in one copy
  if (total > cap_h_blobs || !h_blobs) {
    if (h_blobs) cudaFreeHost(h_blobs);
    h_blobs = nullptr;
    cap_h_blobs = 0;
    PA_CUDA(cudaHostAlloc((void**)&h_blobs, std::max<size_t>(total, 16), cudaHostAllocDefault));
    cap_h_blobs = std::max<size_t>(total, 16);
  }
#pragma omp parallel for schedule(static) num_threads(nt)
  for (int c = 0; c < C; c++)
    if (!plans[c].filtered) {
      memcpy(h_blobs + blob_off[c], plans[c].blob.data(), plans[c].blob.size());
      std::vector<uint8_t>().swap(plans[c].blob);  // the packed copy now lives in HBM
    }
  {
    const size_t need_b = std::max<size_t>(total, 16), need_o = sizeof(int64_t) * (size_t)std::max(C, 1);
    if (need_b > cap_blobs || !d_blobs) {
      if (d_blobs) cudaFree(d_blobs);
      d_blobs = nullptr;
      cap_blobs = 0;
      PA_CUDA(cudaMalloc(&d_blobs, need_b));
      cap_blobs = need_b;
    }
    if (need_o > cap_blob_off || !d_blob_off) {
      if (d_blob_off) cudaFree(d_blob_off);
      d_blob_off = nullptr;
      cap_blob_off = 0;
      PA_CUDA(cudaMalloc(&d_blob_off, need_o));
      cap_blob_off = need_o;
    }
  }
  PA_CUDA(cudaMemcpyAsync(d_blobs, h_blobs, total, cudaMemcpyHostToDevice, stream));
  PA_CUDA(cudaMemcpyAsync(d_blob_off, blob_off.data(), sizeof(int64_t) * C, cudaMemcpyHostToDevice, stream));
  PA_CUDA(cudaStreamSynchronize(stream));
  h2d_bytes = total + sizeof(int64_t) * C;
  return PA_OK;
}

int Engine::element_blob(int32_t elem, uint8_t* buf, int64_t cap, int64_t* size) {
  if (!loaded || elem < 0 || elem >= (int)plans.size() || !size) {
    set_error("bad element id");
    return PA_ERR_ARG;
  }
  *size = 0;
  const ElemPlan& pl = plans[elem];
  if (pl.filtered) return PA_OK;
  if (!pl.blob.empty()) {  // no device bound: the image is still on the host
    *size = (int64_t)pl.blob.size();
    if (buf && cap >= *size) memcpy(buf, pl.blob.data(), pl.blob.size());
    return PA_OK;
  }
  if (device < 0 || !d_blobs) {
    set_error("the element image is not available");
    return PA_ERR_STATE;
  }
  PA_CUDA(cudaSetDevice(device));
  ElemHeader h;
  PA_CUDA(cudaMemcpy(&h, d_blobs + blob_off[elem], sizeof(h), cudaMemcpyDeviceToHost));
  *size = h.blob_bytes;
  if (buf && cap >= *size) PA_CUDA(cudaMemcpy(buf, d_blobs + blob_off[elem], (size_t)h.blob_bytes, cudaMemcpyDeviceToHost));
  return PA_OK;
}

int Engine::loglik_replay(int32_t n_states, const int32_t* elem, const int8_t* Z, const double* n_rate, const double* c_rate, double* out) {
  if (!loaded) {
    set_error("pa_engine_loglik_replay called before pa_engine_load_alignment");
    return PA_ERR_STATE;
  }
  if (device < 0) {
    set_error("no CUDA device bound to this engine (there is no CPU fallback)");
    return PA_ERR_CUDA;
  }
  if (n_states <= 0) return PA_OK;
  const int N = model.tree.N;
  for (int i = 0; i < n_states; i++)
    if (elem[i] < 0 || elem[i] >= (int)plans.size()) {
      set_error("element id out of range");
      return PA_ERR_ARG;
    }
  PA_CUDA(cudaSetDevice(device));
  // one persistent scratch buffer (grow-only), carved into the five arrays of a replay batch
  const size_t ns = (size_t)n_states;
  const size_t off_n = 0, off_c = off_n + 8 * ns, off_out = off_c + 8 * ns, off_elem = off_out + 8 * ns, off_Z = off_elem + ((4 * ns + 15) & ~size_t(15));
  const size_t need = off_Z + ns * N;
  if (need > cap_replay || !d_replay) {
    if (d_replay) cudaFree(d_replay);
    d_replay = nullptr;
    cap_replay = 0;
    PA_CUDA(cudaMalloc(&d_replay, need));
    cap_replay = need;
  }
  double* d_n = reinterpret_cast<double*>(d_replay + off_n);
  double* d_c = reinterpret_cast<double*>(d_replay + off_c);
  double* d_out = reinterpret_cast<double*>(d_replay + off_out);
  int32_t* d_elem = reinterpret_cast<int32_t*>(d_replay + off_elem);
  int8_t* d_Z = reinterpret_cast<int8_t*>(d_replay + off_Z);
  PA_CUDA(cudaMemcpyAsync(d_elem, elem, sizeof(int32_t) * n_states, cudaMemcpyHostToDevice, stream));
  PA_CUDA(cudaMemcpyAsync(d_Z, Z, (size_t)n_states * N, cudaMemcpyHostToDevice, stream));
  PA_CUDA(cudaMemcpyAsync(d_n, n_rate, sizeof(double) * n_states, cudaMemcpyHostToDevice, stream));
  PA_CUDA(cudaMemcpyAsync(d_c, c_rate, sizeof(double) * n_states, cudaMemcpyHostToDevice, stream));
  const int BD = 128;
  const size_t smem = replay_smem_bytes(n_max, slots_max, BD);
  if ((int)smem > smem_optin) {
    set_error("tree too large for the replay kernel's shared memory");
    return PA_ERR_LIMIT;
  }
  PA_CUDA(cudaFuncSetAttribute(replay_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  replay_kernel<false><<<n_states, BD, smem, stream>>>(d_blobs, d_blob_off, model.subst_dev, n_states, d_elem, d_Z, d_n, d_c, N, n_max, slots_max, d_out,
                                                       nullptr, nullptr, Q16{});
  PA_CUDA(cudaGetLastError());
  PA_CUDA(cudaMemcpyAsync(out, d_out, sizeof(double) * n_states, cudaMemcpyDeviceToHost, stream));
  PA_CUDA(cudaStreamSynchronize(stream));
  return PA_OK;
}

// PhyloAcc-ST-GBGC groundwork: replayed log-likelihoods under the GC-biased model for states (Z, B per branch, selection coefficients of
// the conserved / accelerated classes, gBGC bias); same pruning as loglik_replay, per-branch P(t) from exp(t * Q o f(s + b B))
int Engine::gbgc_replay(int32_t n_states, const int32_t* elem, const int8_t* Z, const int8_t* B, const double* sel_n, const double* sel_c,
                        const double* gbc, double* out) {
  if (!loaded) {
    set_error("pa_engine_gbgc_replay called before pa_engine_load_alignment");
    return PA_ERR_STATE;
  }
  if (device < 0) {
    set_error("no CUDA device bound to this engine (there is no CPU fallback)");
    return PA_ERR_CUDA;
  }
  if (n_states <= 0) return PA_OK;
  const int N = model.tree.N;
  for (int i = 0; i < n_states; i++)
    if (elem[i] < 0 || elem[i] >= (int)plans.size()) {
      set_error("element id out of range");
      return PA_ERR_ARG;
    }
  PA_CUDA(cudaSetDevice(device));
  const size_t ns = (size_t)n_states;
  const size_t off_n = 0, off_c = 8 * ns, off_g = 16 * ns, off_out = 24 * ns, off_elem = 32 * ns, off_Z = off_elem + ((4 * ns + 15) & ~size_t(15)),
               off_B = off_Z + ((ns * N + 15) & ~size_t(15));
  const size_t need = off_B + ns * N;
  if (need > cap_replay || !d_replay) {
    if (d_replay) cudaFree(d_replay);
    d_replay = nullptr;
    cap_replay = 0;
    PA_CUDA(cudaMalloc(&d_replay, need));
    cap_replay = need;
  }
  double* d_n = reinterpret_cast<double*>(d_replay + off_n);
  double* d_c = reinterpret_cast<double*>(d_replay + off_c);
  double* d_g = reinterpret_cast<double*>(d_replay + off_g);
  double* d_out = reinterpret_cast<double*>(d_replay + off_out);
  int32_t* d_elem = reinterpret_cast<int32_t*>(d_replay + off_elem);
  int8_t* d_Z = reinterpret_cast<int8_t*>(d_replay + off_Z);
  int8_t* d_B = reinterpret_cast<int8_t*>(d_replay + off_B);
  PA_CUDA(cudaMemcpyAsync(d_elem, elem, sizeof(int32_t) * n_states, cudaMemcpyHostToDevice, stream));
  PA_CUDA(cudaMemcpyAsync(d_Z, Z, ns * N, cudaMemcpyHostToDevice, stream));
  PA_CUDA(cudaMemcpyAsync(d_B, B, ns * N, cudaMemcpyHostToDevice, stream));
  PA_CUDA(cudaMemcpyAsync(d_n, sel_n, sizeof(double) * n_states, cudaMemcpyHostToDevice, stream));
  PA_CUDA(cudaMemcpyAsync(d_c, sel_c, sizeof(double) * n_states, cudaMemcpyHostToDevice, stream));
  PA_CUDA(cudaMemcpyAsync(d_g, gbc, sizeof(double) * n_states, cudaMemcpyHostToDevice, stream));
  const int BD = 128;
  const size_t smem = replay_smem_bytes(n_max, slots_max, BD);
  if ((int)smem > smem_optin) {
    set_error("tree too large for the replay kernel's shared memory");
    return PA_ERR_LIMIT;
  }
  Q16 Q0;
  for (int k = 0; k < 16; k++) Q0.q[k] = model.tree.Q[k];
  PA_CUDA(cudaFuncSetAttribute(replay_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  replay_kernel<true><<<n_states, BD, smem, stream>>>(d_blobs, d_blob_off, model.subst_dev, n_states, d_elem, d_Z, d_n, d_c, N, n_max, slots_max, d_out,
                                                      d_B, d_g, Q0);
  PA_CUDA(cudaGetLastError());
  PA_CUDA(cudaMemcpyAsync(out, d_out, sizeof(double) * n_states, cudaMemcpyDeviceToHost, stream));
  PA_CUDA(cudaStreamSynchronize(stream));
  return PA_OK;
}

int Engine::gbgc_integrate(int32_t n_states, const int32_t* elem, const int8_t* Z, const int8_t* B, const int8_t* fixZ, const double* sel_n, const double* sel_c,
                           const double* gbc, const pa_gbgc_consts* k, double* post_out, double* emis_out, double* root_out) {
  if (!loaded) {
    set_error("pa_engine_gbgc_integrate called before pa_engine_load_alignment");
    return PA_ERR_STATE;
  }
  if (device < 0) {
    set_error("no CUDA device bound to this engine (there is no CPU fallback)");
    return PA_ERR_CUDA;
  }
  const int N = model.tree.N;
  std::vector<int8_t> upper((size_t)std::max(n_states, 0) * N, 0);
  for (int i = 0; i < n_states; i++) {
    if (elem[i] < 0 || elem[i] >= (int)plans.size()) {
      set_error("element id out of range");
      return PA_ERR_ARG;
    }
    // the outgroup's ancestors within the element's pruned tree keep conserved -> conserved with probability 1 (GBGC/SRC/bpp.hpp:312-316)
    for (int32_t v : plans[elem[i]].upper_c) upper[(size_t)i * N + v] = 1;
  }
  return pa::gbgc_integrate(device, stream, d_blobs, d_blob_off, N, n_max, slots_max, P_max, model.tree.pi, model.tree.Q, n_states, elem, Z, B, fixZ,
                            upper.data(), sel_n, sel_c, gbc, k->grate, k->lrate, k->prior_gb, k->cons_to_mis, k->ncons_to_mis, post_out, emis_out, root_out);
}

// ------------------------------------------------------------------------------------------ FP64 FMA peak
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double seed) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 0.999999, c = 1e-9;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
      a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

int fp64_peak(int device, int millis, double* tflops_out) {
  PA_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  PA_CUDA(cudaGetDeviceProperties(&prop, device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256;
  double* d = nullptr;
  PA_CUDA(cudaMalloc(&d, sizeof(double) * blocks * threads));
  cudaEvent_t e0, e1;
  PA_CUDA(cudaEventCreate(&e0));
  PA_CUDA(cudaEventCreate(&e1));
  int iters = 2000;
  double best = 0;
  float total_ms = 0;
  for (int rep = 0; rep < 64 && total_ms < (float)millis; rep++) {
    PA_CUDA(cudaEventRecord(e0));
    fp64_peak_kernel<<<blocks, threads>>>(d, iters, 1.0);
    PA_CUDA(cudaEventRecord(e1));
    PA_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    PA_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    total_ms += ms;
    const double flop = 2.0 * 64.0 * iters * (double)blocks * threads;
    if (rep > 0) best = std::max(best, flop / (ms * 1e-3) / 1e12);
    if (ms < 20.f) iters *= 2;
  }
  cudaFree(d);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tflops_out = best;
  return PA_OK;
}

}  // namespace pa

extern "C" int pa_device_fp64_peak(int device, int millis, double* tflops_out) {
  if (!tflops_out) return PA_ERR_ARG;
  return pa::fp64_peak(device, millis, tflops_out);
}
