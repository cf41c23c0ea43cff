// pa_gt.cu -- PhyloAcc-GT groundwork: what the gene-tree variant evaluates for one element at a given GENE tree embedded in
// the species tree -- the coalescent prior of the tree (GTree::priorTree, src/PhyloAcc-GT/genetree.cpp:560-585), log P(missing | Z, tree)
// (GTree::get_logpZ1, :1156-1190) and the log-likelihood of the alignment on the tree (upward half of GTree::Update_Tg, :1330-1418, summed
// as BPP_C::Gibbs does, bpp_c.cpp:423-432).
//
// The reference walks the embedded tree species by species: a gene-tree branch is cut where it crosses a species boundary and every piece
// evolves at the rate of the species branch it lies in (1, c_rate, n_rate by Z).  All pieces share one rate matrix, so their transition
// matrices commute and compose: the branch as a whole has P(sum_k rate_k len_k).  The host folds the embedding into one EFFECTIVE LENGTH per
// gene-tree branch (O(N) scalar work) and a post-order of the gene tree; the kernel is then an ordinary fp64 pruning pass over the element's
// compressed site patterns, but on a topology that differs per state: one CTA per state, thread = pattern, partials of the internal gene
// nodes in a per-CTA global scratch (a gene tree may be a caterpillar, so no bounded stack), 2^256 rescaling as in the ST kernels.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <string>
#include <vector>

#include "pa_device.cuh"
#include "pa_engine.h"
#include "pa_internal.h"
#include "pa_nvtx.h"

namespace pa {

namespace {

constexpr int kGBD = 128;

struct GtLaunch {
  const uint8_t* blobs; const int64_t* blob_off;
  SubstDev sd;
  int n_states, N, S, P_max;
  const int32_t* elem;     // [n_states]
  const int32_t* gchild;   // [n_states][N][2]
  const int32_t* groot;    // [n_states]
  const uint8_t* gmissing; // [n_states][N]
  const double* eff;       // [n_states][N] effective length of the branch above each gene node
  const int16_t* prog;     // [n_states][N] internal gene nodes, children before parents; -1 terminated
  double* scratch;         // [grid][N][P_max][4] partials
  int16_t* scratch_sc;     // [grid][N][P_max] scale counts
  double* out;             // [n_states]
};

__global__ void __launch_bounds__(kGBD) gt_replay_kernel(GtLaunch L) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, N = L.N, S = L.S;
  double* Pm = reinterpret_cast<double*>(smem_raw);  // [N][16] P(i -> j | effective length)
  double* red = Pm + (size_t)N * 16;                 // [34]
  int16_t* s_child = reinterpret_cast<int16_t*>(red + 34);  // [N][2]
  int16_t* s_prog = s_child + 2 * N;                         // [N]
  int16_t* s_row = s_prog + N;                               // [S] code row of a species, -1: not in this element
  uint8_t* s_mis = reinterpret_cast<uint8_t*>(s_row + N);   // [N]
  __shared__ double s_pi[4];
  if (tid < 4) s_pi[tid] = L.sd.pi[tid];
  double* part = L.scratch + (size_t)blockIdx.x * N * L.P_max * 4;
  int16_t* psc = L.scratch_sc + (size_t)blockIdx.x * N * L.P_max;
  for (int st = blockIdx.x; st < L.n_states; st += gridDim.x) {
    __syncthreads();
    const int64_t off = L.blob_off[L.elem[st]];
    if (off < 0) {  // filtered element
      if (tid == 0) L.out[st] = CUDART_NAN;
      continue;
    }
    const ElemView ev = make_view(L.blobs + off);
    const ElemHeader* h = ev.h;
    const int P = h->P, Ppad = h->Ppad, half = Ppad >> 1;
    for (int i = tid; i < N; i += kGBD) {
      s_child[2 * i] = (int16_t)L.gchild[((size_t)st * N + i) * 2];
      s_child[2 * i + 1] = (int16_t)L.gchild[((size_t)st * N + i) * 2 + 1];
      s_prog[i] = L.prog[(size_t)st * N + i];
      s_mis[i] = L.gmissing[(size_t)st * N + i];
      s_row[i] = -1;
      double P16[16];
      build_P(L.sd, L.eff[(size_t)st * N + i], 1.0, true, P16);
#pragma unroll
      for (int k = 0; k < 16; k++) Pm[i * 16 + k] = P16[k];
    }
    __syncthreads();
    for (int i = tid; i < h->n; i += kGBD)  // local tips of the element -> the species they are
      if ((ev.flags[i] & PA_F_TIP) && !(ev.flags[i] & PA_F_MISSING) && ev.orig[i] < S) s_row[ev.orig[i]] = ev.tiprow[i];
    __syncthreads();
    const int root = L.groot[st];
    double acc = 0.0;
    for (int p = tid; p < Ppad; p += kGBD) {
      if (p >= P) continue;
      double rl[4] = {1.0, 1.0, 1.0, 1.0};
      int rsc = 0;
      for (int k = 0; k < N; k++) {
        const int x = s_prog[k];
        if (x < 0) break;
        double out[4] = {1.0, 1.0, 1.0, 1.0};
        int sc = 0;
#pragma unroll
        for (int cc = 0; cc < 2; cc++) {
          const int c = s_child[2 * x + cc];
          if (c < 0 || s_mis[c]) continue;  // genetree.cpp:1366
          const double* Pc = Pm + c * 16;
          if (c < S) {
            const int row = s_row[c];
            if (row < 0) continue;
            const uint32_t byte = __ldg(ev.codes + (size_t)row * half + (p >> 1));
            const uint32_t code = (byte >> ((p & 1) * 4)) & 15u;
            if (code == 15u) continue;  // gap / N: Tg >= num_base
#pragma unroll
            for (int j = 0; j < 4; j++) {
              double m = 0.0;
#pragma unroll
              for (int b = 0; b < 4; b++) m += ((code >> b) & 1u) ? Pc[j * 4 + b] : 0.0;
              out[j] *= m;
            }
          } else {
            const double* g = part + ((size_t)c * L.P_max + p) * 4;
            const double l0 = g[0], l1 = g[1], l2 = g[2], l3 = g[3];
            sc += psc[(size_t)c * L.P_max + p];
#pragma unroll
            for (int j = 0; j < 4; j++) out[j] *= ((Pc[j * 4] * l0 + Pc[j * 4 + 1] * l1) + Pc[j * 4 + 2] * l2) + Pc[j * 4 + 3] * l3;
          }
        }
        const double mx = fmax(fmax(out[0], out[1]), fmax(out[2], out[3]));
        if (mx < PA_SCALE_THRESH && mx > 0.0) {
#pragma unroll
          for (int j = 0; j < 4; j++) out[j] *= PA_SCALE_UP;
          sc += 1;
        }
        double* g = part + ((size_t)x * L.P_max + p) * 4;
        g[0] = out[0]; g[1] = out[1]; g[2] = out[2]; g[3] = out[3];
        psc[(size_t)x * L.P_max + p] = (int16_t)sc;
        if (x == root) {
#pragma unroll
          for (int j = 0; j < 4; j++) rl[j] = out[j];
          rsc = sc;
        }
      }
      const double lik = ((s_pi[0] * rl[0] + s_pi[1] * rl[1]) + s_pi[2] * rl[2]) + s_pi[3] * rl[3];
      acc += (double)ev.weights[p] * (log(lik) - rsc * PA_LOG_SCALE);
    }
    const double tot = block_sum(acc, red, tid, kGBD);
    if (tid == 0) L.out[st] = tot;
  }
}

}  // namespace

// Host half: the embedding folded into effective lengths, the coalescent prior and log P(missing | Z); then the kernel.
int Engine::gt_replay(const pa_gt_states* gs, double* loglik_out, double* prior_out, double* logpz_out) {
  if (!loaded) {
    set_error("pa_engine_gt_replay called before pa_engine_load_alignment");
    return PA_ERR_STATE;
  }
  if (device < 0) {
    set_error("no CUDA device bound to this engine (there is no CPU fallback)");
    return PA_ERR_CUDA;
  }
  const int N = model.tree.N, S = model.tree.S, ns = gs->n_states;
  if (gs->N != N) {
    set_error("pa_engine_gt_replay: states sized for another tree");
    return PA_ERR_ARG;
  }
  if (ns <= 0) return PA_OK;
  if (N > 32767) {
    set_error("pa_engine_gt_replay: tree too large");
    return PA_ERR_LIMIT;
  }
  const std::vector<int32_t>& spar = model.tree.parent;
  std::vector<double> eff((size_t)ns * N, 0.0);
  std::vector<int16_t> prog((size_t)ns * N, -1);
  for (int st = 0; st < ns; st++) {
    if (gs->elem[st] < 0 || gs->elem[st] >= (int)plans.size()) {
      set_error("element id out of range");
      return PA_ERR_ARG;
    }
    const int32_t* gc = gs->gchild + (size_t)st * N * 2;
    for (int k = 0; k < 2 * N; k++)
      if (gc[k] < -1 || gc[k] >= N) {
        set_error("pa_engine_gt_replay: gene-tree child out of range");
        return PA_ERR_ARG;
      }
    const double* gh = gs->gheights + (size_t)st * N;
    const uint8_t* mis = gs->gmissing + (size_t)st * N;
    const int8_t* Z = gs->Z + (size_t)st * N;
    const double lc = gs->cons_to_mis, ln = gs->ncons_to_mis;
    auto loss = [&](int s, bool m) {
      const double p = (Z[s] == 1) ? lc : ln;
      return m ? std::log(p) : std::log(1 - p);
    };
    double prior = 0, lpz = 0;
    for (int s = 0; s < S; s++) lpz += loss(s, mis[s] != 0);
    double* ef = eff.data() + (size_t)st * N;
    // lineages grouped by species (the reference's parent_gene2): [lo, hi) of every species in the flattened list
    const int64_t l0 = gs->lin_off[st], l1 = gs->lin_off[st + 1];
    std::vector<std::vector<int64_t>> by_sp(N);
    for (int64_t k = l0; k < l1; k++) {
      const int sp = gs->lin_species[k];
      if (sp < 0 || sp >= N || gs->lin_node[k] < 0 || gs->lin_node[k] >= N) {
        set_error("pa_engine_gt_replay: lineage entry out of range");
        return PA_ERR_ARG;
      }
      by_sp[sp].push_back(k);
    }
    for (int ss = 0; ss < N; ss++) {
      if (by_sp[ss].empty()) continue;
      const double rate = (Z[ss] == 2) ? gs->n_rate[st] : (Z[ss] == 1 ? gs->c_rate[st] : 1.0);
      const double h_ss = gs->heights[ss];
      const int sp = spar[ss];
      const bool has_parent = sp >= 0 && sp < N;
      // coalescences inside this species (temp_coal): the distinct targets that are not the lineage itself, by height
      std::vector<int> coal;
      for (int64_t k : by_sp[ss]) {
        const int nx = gs->lin_next[k];
        if (nx != gs->lin_node[k] && nx >= 0 && nx < N && std::find(coal.begin(), coal.end(), nx) == coal.end()) coal.push_back(nx);
      }
      std::sort(coal.begin(), coal.end(), [&](int a, int b) { return gh[a] < gh[b]; });
      for (int g : coal)
        for (int cc = 0; cc < 2; cc++) {
          const int chi = gc[2 * g + cc];
          if (chi >= 0) ef[chi] += rate * (gh[g] - std::max(gh[chi], h_ss));  // genetree.cpp:1352-1358
        }
      for (int64_t k : by_sp[ss]) {
        const int first = gs->lin_node[k];
        if (has_parent && gs->lin_next[k] == first) ef[first] += rate * (gs->heights[sp] - std::max(gh[first], h_ss));  // :1385-1388
        if (ss >= S && gh[first] >= h_ss) lpz += loss(ss, mis[first] != 0);                                             // :1172-1186
      }
      if (ss >= S) {  // priorTree, genetree.cpp:564-581
        int nn = (int)by_sp[ss].size() - (int)coal.size();
        if (nn != 0) {
          double cur = h_ss;
          const double th = gs->thetas[ss];
          for (int g : coal) {
            prior -= (gh[g] - cur) * (nn - 1) * nn / th;
            prior += std::log(2) - std::log(th);
            nn--;
            cur = gh[g];
          }
          if (nn > 1 && has_parent) prior -= (gs->heights[sp] - cur) * (nn - 1) * nn / th;
        }
      }
    }
    if (prior_out) prior_out[st] = prior;
    if (logpz_out) logpz_out[st] = lpz;
    // post-order of the internal gene nodes below the root (iterative)
    const int root = gs->groot[st];
    if (root < 0 || root >= N) {
      set_error("pa_engine_gt_replay: gene-tree root out of range");
      return PA_ERR_ARG;
    }
    int16_t* pg = prog.data() + (size_t)st * N;
    int np = 0;
    std::vector<std::pair<int, int>> stack;
    stack.push_back({root, 0});
    int guard = 0;
    while (!stack.empty()) {
      auto& top = stack.back();
      const int x = top.first;
      if (x < S || (gc[2 * x] < 0 && gc[2 * x + 1] < 0)) { stack.pop_back(); continue; }
      if (top.second < 2) {
        const int c = gc[2 * x + top.second++];
        if (c >= 0 && c < N) stack.push_back({c, 0});
      } else {
        if (np >= N) break;
        pg[np++] = (int16_t)x;
        stack.pop_back();
      }
      if (++guard > 8 * N + 16) {
        set_error("pa_engine_gt_replay: the gene tree has a cycle");
        return PA_ERR_ARG;
      }
    }
  }
  if (!loglik_out) return PA_OK;
  cudaError_t err = cudaSetDevice(device);
  const int grid = std::min(ns, std::max(1, num_sms) * 4);
  const size_t b_part = (size_t)grid * N * P_max * 4 * 8, b_sc = (size_t)grid * N * P_max * 2;
  const size_t b_in = (size_t)ns * (4 + 8 * N + 4 + N + 8 * N + 2 * N) + 64, b_out = (size_t)ns * 8;
  unsigned char *d_in = nullptr, *d_scr = nullptr;
  if (err == cudaSuccess) err = cudaMalloc(&d_in, b_in + b_out + 64);
  if (err == cudaSuccess) err = cudaMalloc(&d_scr, b_part + b_sc);
  GtLaunch L;
  L.blobs = d_blobs; L.blob_off = d_blob_off; L.sd = model.subst_dev; L.n_states = ns; L.N = N; L.S = S; L.P_max = P_max;
  unsigned char* q = d_in;
  auto carve = [&](size_t bytes) { unsigned char* r = q; q += (bytes + 15) & ~size_t(15); return r; };
  L.out = reinterpret_cast<double*>(carve(b_out));
  L.eff = reinterpret_cast<double*>(carve((size_t)ns * N * 8));
  L.gchild = reinterpret_cast<int32_t*>(carve((size_t)ns * N * 8));
  L.elem = reinterpret_cast<int32_t*>(carve((size_t)ns * 4));
  L.groot = reinterpret_cast<int32_t*>(carve((size_t)ns * 4));
  L.prog = reinterpret_cast<int16_t*>(carve((size_t)ns * N * 2));
  L.gmissing = carve((size_t)ns * N);
  L.scratch = reinterpret_cast<double*>(d_scr);
  L.scratch_sc = reinterpret_cast<int16_t*>(d_scr + b_part);
  auto up = [&](const void* dst, const void* src, size_t bytes) {
    if (err == cudaSuccess) err = cudaMemcpyAsync(const_cast<void*>(dst), src, bytes, cudaMemcpyHostToDevice, stream);
  };
  up(L.eff, eff.data(), (size_t)ns * N * 8);
  up(L.gchild, gs->gchild, (size_t)ns * N * 8);
  up(L.elem, gs->elem, (size_t)ns * 4);
  up(L.groot, gs->groot, (size_t)ns * 4);
  up(L.prog, prog.data(), (size_t)ns * N * 2);
  up(L.gmissing, gs->gmissing, (size_t)ns * N);
  const size_t smem = ((size_t)N * 16 + 34) * 8 + (size_t)N * 2 * 4 + (size_t)N + 16;
  if (err == cudaSuccess && (int)smem > smem_optin) {
    cudaFree(d_in); cudaFree(d_scr);
    set_error("tree too large for the gene-tree replay kernel's shared memory");
    return PA_ERR_LIMIT;
  }
  if (err == cudaSuccess) err = cudaFuncSetAttribute(gt_replay_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err == cudaSuccess) {
    PA_NVTX_RANGE("gt_replay");
    gt_replay_kernel<<<grid, kGBD, smem, stream>>>(L);
    err = cudaGetLastError();
  }
  if (err == cudaSuccess) err = cudaMemcpyAsync(loglik_out, L.out, b_out, cudaMemcpyDeviceToHost, stream);
  if (err == cudaSuccess) err = cudaStreamSynchronize(stream);
  if (d_in) cudaFree(d_in);
  if (d_scr) cudaFree(d_scr);
  if (err != cudaSuccess) {
    set_error(std::string("gt_replay: ") + cudaGetErrorString(err));
    return PA_ERR_CUDA;
  }
  return PA_OK;
}

}  // namespace pa
