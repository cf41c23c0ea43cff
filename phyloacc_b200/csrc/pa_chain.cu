// pa_chain.cu -- the on-device PhyloAcc-ST chain: one persistent CTA per (element, model) chain, T Gibbs/MH
// iterations with no host round-trip.  Mirrors BPP_C::initMCMC + BPP_C::Gibbs (bpp_c.cpp:131-519) of the reference:
//
//   MonitorChain  -> trace row (Z packed 2 bit/node, loglik, n, c, g, l, l2)                 bpp_c.cpp:741-834
//   Update_Tg(.,.,true) for every site: ancestral bases, top-down, from per-PATTERN partials  bpp_c.cpp:672-733
//   getEmission: parent->child base counts per branch  x  log P_k                           bpp_c2.cpp:355-390
//   Update_Z_subtree: 3-state upward messages + top-down sampling of Z                       bpp_c.cpp:1182-1293
//   getUpdateNode + Update_Tg(.,.,false): re-prune ancestors of changed Z                    bpp_c.cpp:620-671
//   sample_rate (n, then c): MH with the histories integrated out (log_f_Xz)                 bpp_c.cpp:860-1176
//   sample_transition: Beta draws for g, l, l2                                               bpp_c.cpp:522-613
//
// Data layout: tip codes (4 bit), weights, gap masks and the pruning program of the element come from its blob in
// HBM; per-pattern partials of every internal node live in this CTA's arena (HBM, L2-resident while the chain
// runs); proposal evaluations keep their partials in a per-thread shared-memory stack and never touch the arena.
// Random numbers: Philox4x32-10 indexed by (chain, iteration, purpose, site/node), see oracle/oracle.h.
#include <cuda_runtime.h>

#include <cstdio>

#include "pa_chain.cuh"
#include "pa_device.cuh"

namespace pa {

namespace {

constexpr int kBD = 256;
constexpr int kNW = kBD / 32;

// scalar slots in shared memory
enum { SC_N = 0, SC_C, SC_G, SC_L, SC_L2, SC_LL, SC_PROPN, SC_PROPC, SC_PROPOSAL, SC_LLNEW, SC_R, SC_COUNT = 16 };
enum { IV_CHAIN = 0, IV_ANY_CHANGED, IV_HAS1, IV_HAS2, IV_VALID, IV_ACCEPT, IV_RATE_CHANGED, IV_STALE, IV_ACC_N, IV_ACC_C, IV_ZFALLBACK, IV_NDIRTY, IV_COUNT = 16 };

struct Smem {
  double *Pcur, *Pprop, *stack, *red, *emis, *back, *eamb, *dist, *uz, *scal, *T;  // T: [2][9] linear Z transition (free, fixed)
  int *stack_sc, *ivars, *nz;
  uint32_t *cnt, *xs;
  int16_t *parent, *child, *tiprow, *su, *intidx, *intlist, *up_nodes, *up_off, *down_nodes, *down_off;
  uint8_t *flags, *fix, *dirty, *stale, *cnt8;
  int8_t *slot, *Z, *Znew;
};

__host__ __device__ inline size_t align16(size_t x) { return (x + 15) & ~size_t(15); }

__host__ __device__ inline size_t carve(Smem* s, unsigned char* base, int n, int slots, int BD) {
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align16(o + bytes); return base ? base + r : (unsigned char*)nullptr; };
  const int XW = (n + 15) / 16 + 1;
  unsigned char* p;
  p = take(sizeof(double) * n * 16); if (s) s->Pcur = (double*)p;
  p = take(sizeof(double) * n * 16); if (s) s->Pprop = (double*)p;
  p = take(sizeof(double) * slots * 4 * BD); if (s) s->stack = (double*)p;
  p = take(sizeof(double) * 34); if (s) s->red = (double*)p;
  p = take(sizeof(double) * n * 3); if (s) s->emis = (double*)p;
  p = take(sizeof(double) * n * 3); if (s) s->back = (double*)p;
  p = take(sizeof(double) * n * 3); if (s) s->eamb = (double*)p;
  p = take(sizeof(double) * n); if (s) s->dist = (double*)p;
  p = take(sizeof(double) * n); if (s) s->uz = (double*)p;
  p = take(sizeof(double) * SC_COUNT); if (s) s->scal = (double*)p;
  p = take(sizeof(double) * 18); if (s) s->T = (double*)p;
  p = take(sizeof(int) * slots * BD); if (s) s->stack_sc = (int*)p;
  p = take(sizeof(int) * IV_COUNT); if (s) s->ivars = (int*)p;
  p = take(sizeof(int) * 18); if (s) s->nz = (int*)p;
  p = take(sizeof(uint32_t) * n * 16); if (s) s->cnt = (uint32_t*)p;
  p = take(sizeof(uint32_t) * XW * BD); if (s) s->xs = (uint32_t*)p;
  p = take(sizeof(int16_t) * n); if (s) s->parent = (int16_t*)p;
  p = take(sizeof(int16_t) * 2 * n); if (s) s->child = (int16_t*)p;
  p = take(sizeof(int16_t) * n); if (s) s->tiprow = (int16_t*)p;
  p = take(sizeof(int16_t) * n); if (s) s->su = (int16_t*)p;
  p = take(sizeof(int16_t) * n); if (s) s->intidx = (int16_t*)p;
  p = take(sizeof(int16_t) * n); if (s) s->intlist = (int16_t*)p;
  p = take(sizeof(int16_t) * n); if (s) s->up_nodes = (int16_t*)p;
  p = take(sizeof(int16_t) * (n + 1)); if (s) s->up_off = (int16_t*)p;
  p = take(sizeof(int16_t) * n); if (s) s->down_nodes = (int16_t*)p;
  p = take(sizeof(int16_t) * (n + 1)); if (s) s->down_off = (int16_t*)p;
  p = take(n); if (s) s->flags = p;
  p = take(n); if (s) s->fix = p;
  p = take(n); if (s) s->dirty = p;
  p = take(n); if (s) s->stale = p;
  p = take((size_t)(BD / 32) * n * 16); if (s) s->cnt8 = p;
  p = take(n); if (s) s->slot = (int8_t*)p;
  p = take(n); if (s) s->Z = (int8_t*)p;
  p = take(n); if (s) s->Znew = (int8_t*)p;
  return o;
}

__device__ __forceinline__ double class_rate(int z, double n_rate, double c_rate) { return z == 2 ? n_rate : (z == 0 ? 1.0 : c_rate); }

// mark the ancestors of node i dirty (getUpdateNode, bpp_c.cpp:620-642 / 836-858)
__device__ __forceinline__ void mark_ancestors(const int16_t* parent, uint8_t* dirty, int i) {
  int j = parent[i];
  while (j >= 0) {
    if (dirty[j]) break;
    dirty[j] = 1;
    j = parent[j];
  }
}

// one pruning pass over all patterns of the element; returns sum_p w_p log L_p when REDUCE
template <bool READ_G, bool WRITE_G, bool REDUCE>
__device__ __forceinline__ double prune_all(const ElemView& ev, const Smem& S, const TreeSmem& ts, const double* Pm, const uint8_t* dirty,
                                            double* gpart, int16_t* gsc, const double* pi, int tid) {
  double acc = 0.0;
  const int P = ev.h->P, Ppad = ev.h->Ppad;
  for (int p = tid; p < Ppad; p += kBD) {
    if (p < P) {
      const double ll = prune_pattern<READ_G, WRITE_G>(ev.h, ts, ev.codes, Pm, dirty, S.stack, S.stack_sc, gpart, gsc, pi, p, kBD, tid);
      if (REDUCE) acc += (double)ev.weights[p] * ll;
    }
  }
  if (REDUCE) return block_sum(acc, S.red, tid, kBD);
  return 0.0;
}

// log-space serial Z update, a literal restatement of Update_Z_subtree; used only when the scaled linear pass underflows
__device__ void z_update_logspace(const Smem& S, int n, const double* Tlin, uint32_t* changed_flag) {
  // log transition matrices
  double lt[2][9];
  for (int f = 0; f < 2; f++)
    for (int k = 0; k < 9; k++) lt[f][k] = log(Tlin[f * 9 + k]);
  for (int i = 0; i < n; i++) {
    for (int k = 0; k < 3; k++) S.back[i * 3 + k] = S.emis[i * 3 + k];
    if (S.fix[i]) S.back[i * 3 + 2] = -CUDART_INF;
  }
  for (int i = 0; i < n - 1; i++) {
    const int p = S.parent[i];
    const double* t = lt[S.fix[i] ? 1 : 0];
    for (int j = 0; j < 3; j++) {
      double col[3], mx = -CUDART_INF, s = 0;
      for (int a = 0; a < 3; a++) {
        col[a] = t[a * 3 + j] + S.back[i * 3 + a];
        if (col[a] > mx) mx = col[a];
      }
      for (int a = 0; a < 3; a++) s += exp(col[a] - mx);
      S.back[p * 3 + j] += log(s) + mx;
    }
  }
  const int root = n - 1;
  {
    double v[3] = {S.back[root * 3] + log(0.5), S.back[root * 3 + 1] + log(0.5), -CUDART_INF};
    double mx = fmax(v[0], fmax(v[1], v[2]));
    S.Znew[root] = (int8_t)categorical3(exp(v[0] - mx), exp(v[1] - mx), exp(v[2] - mx), S.uz[root]);
  }
  for (int i = n - 2; i >= 0; i--) {
    const double* t = lt[S.fix[i] ? 1 : 0];
    const int zp = S.Znew[S.parent[i]];
    double v[3];
    for (int a = 0; a < 3; a++) v[a] = t[a * 3 + zp] + S.back[i * 3 + a];
    const double mx = fmax(v[0], fmax(v[1], v[2]));
    S.Znew[i] = (int8_t)categorical3(exp(v[0] - mx), exp(v[1] - mx), exp(v[2] - mx), S.uz[i]);
  }
}

__global__ void __launch_bounds__(kBD, 2) chain_kernel(ChainLaunch L) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem S;
  carve(&S, smem_raw, L.n_max, L.slots_max, kBD);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const ChainParams& cp = L.cp;
  const SubstDev& sd = L.sd;
  __shared__ double s_pi[4];
  if (tid < 4) s_pi[tid] = sd.pi[tid];

  // this CTA's arena: partials [ni][Ppad][4] f64, scale counts [ni][Ppad] i16, log P tables [3][n_max][16] f64
  unsigned char* arena = L.arena + (size_t)blockIdx.x * L.arena_stride;
  double* logP = reinterpret_cast<double*>(arena);
  double* gpart = logP + (size_t)3 * L.n_max * 16;
  int16_t* gsc = reinterpret_cast<int16_t*>(gpart + (size_t)L.n_max * L.P_max * 4);

  unsigned long long w_npu = 0, w_root = 0, w_pt = 0, w_hist = 0, w_iter = 0;  // work counters (thread 0)
  long long ph0 = 0, ph1 = 0, ph2 = 0, ph3 = 0, ph4 = 0, ph5 = 0, ph6 = 0, ph7 = 0, ph8 = 0, ph9 = 0;  // SM cycles per phase (thread 0)
  long long t_last = clock64();
#define PA_PHASE(var)                \
  if (tid == 0) {                    \
    const long long t_now = clock64(); \
    var += t_now - t_last;           \
    t_last = t_now;                  \
  }

  for (;;) {
    __syncthreads();
    if (tid == 0) S.ivars[IV_CHAIN] = atomicAdd(L.queue, 1);
    __syncthreads();
    const int chain = S.ivars[IV_CHAIN];
    if (chain >= L.n_chains) break;
    const int elem = L.chain_list[2 * chain], mdl = L.chain_list[2 * chain + 1];
    const int resZ = (mdl == PA_M0) ? 0 : (mdl == PA_M1 ? 2 : 1);
    const uint32_t key = (uint32_t)elem * 4u + (uint32_t)mdl;
    const ElemView ev = make_view(L.blobs + L.blob_off[elem]);
    const ElemHeader* h = ev.h;
    const int n = h->n, ni = h->ni, GG = h->GG, Ppad = h->Ppad, root = n - 1;
    const int T = L.T;

    // ---------------- initMCMC (bpp_c.cpp:131-343)
    for (int i = tid; i < n; i += kBD) {
      S.parent[i] = ev.parent[i];
      S.child[2 * i] = ev.child[2 * i];
      S.child[2 * i + 1] = ev.child[2 * i + 1];
      S.tiprow[i] = ev.tiprow[i];
      S.intidx[i] = ev.intidx[i];
      S.up_nodes[i] = ev.up_nodes[i];
      S.down_nodes[i] = ev.down_nodes[i];
      const uint8_t f = ev.flags[i];
      S.flags[i] = f;
      S.fix[i] = (f >> (2 + mdl)) & 1;
      S.slot[i] = ev.slot[i];
      S.dist[i] = ev.dist[i];
      S.Z[i] = (i == root) ? 0 : 1;  // bpp_c.cpp:141-146
      S.dirty[i] = 0;
      S.stale[i] = 0;
    }
    for (int i = tid; i < ni; i += kBD) {
      S.su[i] = ev.su[i];
      S.intlist[i] = ev.intlist[i];
    }
    for (int i = tid; i <= h->n_up; i += kBD) S.up_off[i] = ev.up_off[i];
    for (int i = tid; i <= h->n_down; i += kBD) S.down_off[i] = ev.down_off[i];
    if (tid == 0) {
      S.scal[SC_N] = cp.init_nrate;
      S.scal[SC_C] = cp.init_crate;
      S.scal[SC_G] = cp.init_grate;
      S.scal[SC_L] = cp.init_lrate;
      S.scal[SC_L2] = cp.init_lrate2;
      S.scal[SC_PROPC] = 1.2;                      // bpp_c.cpp:135
      S.scal[SC_PROPN] = (double)1 / GG * 10;      // bpp_c.cpp:136
      S.ivars[IV_ACC_N] = 0;
      S.ivars[IV_ACC_C] = 0;
      S.ivars[IV_STALE] = 0;
    }
    __syncthreads();
    const TreeSmem ts{S.child, S.flags, S.tiprow, S.su, S.slot, S.intidx};
    // P(t) of every branch for its class, and the log P tables of the three classes (emission look-ups)
    for (int task = tid; task < 3 * n; task += kBD) {
      const int k = task / n, i = task - k * n;
      double P16[16];
      const double rate = (k == 0) ? 1.0 : (k == 1 ? cp.init_crate : cp.init_nrate);
      build_P(sd, S.dist[i], rate, true, P16);
      const int zi = S.Z[i];
      const int kz = (zi == 0) ? 0 : (zi == 2 ? 2 : 1);
#pragma unroll
      for (int e = 0; e < 16; e++) {
        logP[((size_t)k * L.n_max + i) * 16 + e] = log(P16[e]);
        if (k == kz) S.Pcur[i * 16 + e] = P16[e];
      }
    }
    if (tid == 0) {  // linear Z transition matrices (bpp_c.cpp:301-332), index [child state a][parent state j] -> a*3+j
      const double g = cp.init_grate, l = cp.init_lrate, l2 = cp.init_lrate2;
      double* Tf = S.T;       // free
      double* Tx = S.T + 9;   // fixZ == 1
      const double y = 1 - l;
      Tf[0] = 1 - g; Tf[3] = g; Tf[6] = 0; Tf[1] = 0; Tf[4] = y; Tf[7] = 1 - y; Tf[2] = 0; Tf[5] = l2; Tf[8] = 1 - l2;
      Tx[0] = 1 - g; Tx[3] = g; Tx[6] = 0; Tx[1] = 0; Tx[4] = 1; Tx[7] = 0; Tx[2] = 1; Tx[5] = 1; Tx[8] = 1;
      w_pt += 3ull * n;
    }
    __syncthreads();
    // initial upward pass over every pattern (Gibbs, bpp_c.cpp:355-356), partials stored in the arena
    {
      const double ll = prune_all<false, true, true>(ev, S, ts, S.Pcur, nullptr, gpart, gsc, s_pi, tid);
      if (tid == 0) {
        S.scal[SC_LL] = ll;
        int cnt_nodes = 0;
        for (int k = 0; k < ni && S.su[k] >= 0; k++) cnt_nodes++;
        w_npu += (unsigned long long)cnt_nodes * h->P;
        w_root += h->P;
      }
    }
    __syncthreads();

    PA_PHASE(ph0)
    double* tr_d = L.trace_d + (size_t)chain * 6 * T;
    uint32_t* tr_z = L.trace_z + (size_t)chain * T * L.ZW;

    for (int m = 0; m < T; m++) {
      // ---------------- MonitorChain: trace row m
      for (int w = tid; w < L.ZW; w += kBD) {
        uint32_t word = 0;
        for (int k = 0; k < 16; k++) {
          const int i = w * 16 + k;
          if (i < n) word |= (uint32_t)(S.Z[i] & 3) << (2 * k);
        }
        tr_z[(size_t)m * L.ZW + w] = word;
      }
      if (tid < 6) {
        const int src = (tid == 0) ? SC_LL : (tid == 1 ? SC_N : (tid == 2 ? SC_C : (tid == 3 ? SC_G : (tid == 4 ? SC_L : SC_L2))));
        tr_d[(size_t)tid * T + m] = S.scal[src];
      }
      if (m == T - 1) break;
      if (tid == 0) w_iter++;
      PA_PHASE(ph1)

      // ---------------- sample ancestral histories site by site + count parent->child base pairs per branch
      for (int i = tid; i < n * 16; i += kBD) S.cnt[i] = 0;
      for (int i = tid; i < n * 3; i += kBD) S.eamb[i] = 0.0;
      const int XW = (ni + 15) >> 4;
      const double cur_n = S.scal[SC_N], cur_c = S.scal[SC_C];
      __syncthreads();
      for (int g0 = 0; g0 < GG; g0 += kBD) {
        const int g = g0 + tid;
        const bool valid = g < GG;
        const int pat = valid ? ev.sitepat[g] : 0;
        uint32_t gw0 = 0, gw1 = 0;
        if (h->W >= 1) gw0 = ev.gap[pat];
        if (h->W >= 2) gw1 = ev.gap[Ppad + pat];
        for (int w = 0; w < XW; w++) S.xs[w * kBD + tid] = 0;
        uint4 rnd = make_uint4(0, 0, 0, 0);
        int have = -1;
        for (int q = ni - 1; q >= 0; q--) {
          const int s = S.intlist[q];
          if (S.flags[s] & PA_F_MISSING) continue;
          const bool isroot = (s == root);
          bool gapbit;
          if (q < 32) gapbit = (gw0 >> q) & 1u;
          else if (q < 64) gapbit = (gw1 >> (q - 32)) & 1u;
          else gapbit = (ev.gap[(size_t)(q >> 5) * Ppad + pat] >> (q & 31)) & 1u;
          const bool active = valid && (isroot || !gapbit);
          int x = 0, xp = 0;
          if (!isroot) {
            const int qp = S.intidx[S.parent[s]];
            xp = (S.xs[(qp >> 4) * kBD + tid] >> (2 * (qp & 15))) & 3;
          }
          if (active) {
            if ((q >> 2) != have) {
              rnd = philox4(cp.seed, key, (uint32_t)g, (uint32_t)(q >> 2), (uint32_t)m, PA_ST_HIST);
              have = q >> 2;
            }
            const double u = u01(pick(rnd, q & 3));
            const double2* g2 = reinterpret_cast<const double2*>(gpart + ((size_t)q * Ppad + pat) * 4);
            const double2 a = g2[0], b = g2[1];
            double p0, p1, p2, p3;
            if (isroot) {
              p0 = a.x * s_pi[0]; p1 = a.y * s_pi[1]; p2 = b.x * s_pi[2]; p3 = b.y * s_pi[3];
            } else {
              const double* Pr = S.Pcur + s * 16 + xp * 4;
              p0 = Pr[0] * a.x; p1 = Pr[1] * a.y; p2 = Pr[2] * b.x; p3 = Pr[3] * b.y;
            }
            x = categorical4(p0, p1, p2, p3, u);
            if (x < 0) x = 0;
            S.xs[(q >> 4) * kBD + tid] |= (uint32_t)x << (2 * (q & 15));
          }
          if (!isroot) {  // counts of (parent base, child base) over the 32 sites of this warp
            const uint32_t code = (uint32_t)(xp * 4 + x);
            const uint32_t bv = __ballot_sync(0xffffffffu, active);
            const uint32_t b0 = __ballot_sync(0xffffffffu, code & 1u), b1 = __ballot_sync(0xffffffffu, code & 2u);
            const uint32_t b2 = __ballot_sync(0xffffffffu, code & 4u), b3 = __ballot_sync(0xffffffffu, code & 8u);
            if (lane < 16) {
              uint32_t mm = bv;
              mm &= (lane & 1) ? b0 : ~b0;
              mm &= (lane & 2) ? b1 : ~b1;
              mm &= (lane & 4) ? b2 : ~b2;
              mm &= (lane & 8) ? b3 : ~b3;
              S.cnt8[((size_t)warp * n + s) * 16 + lane] = (uint8_t)__popc(mm);
            }
          }
        }
        // tip branches: the tip's base is data, the parent's base was just sampled
        for (int i = 0; i < n; i++) {
          const uint8_t f = S.flags[i];
          if (!(f & PA_F_TIP) || (f & PA_F_MISSING)) continue;
          uint32_t code4 = 15u;
          if (valid) code4 = (ev.codes[(size_t)S.tiprow[i] * (Ppad >> 1) + (pat >> 1)] >> ((pat & 1) * 4)) & 15u;
          const int qp = S.intidx[S.parent[i]];
          const int xp = (S.xs[(qp >> 4) * kBD + tid] >> (2 * (qp & 15))) & 3;
          const bool onehot = (code4 & (code4 - 1)) == 0;  // code4 is never 0
          const bool act = valid && code4 != 15u && onehot;
          const uint32_t code = (uint32_t)(xp * 4 + (onehot ? (__ffs(code4) - 1) : 0));
          const uint32_t bv = __ballot_sync(0xffffffffu, act);
          const uint32_t b0 = __ballot_sync(0xffffffffu, code & 1u), b1 = __ballot_sync(0xffffffffu, code & 2u);
          const uint32_t b2 = __ballot_sync(0xffffffffu, code & 4u), b3 = __ballot_sync(0xffffffffu, code & 8u);
          if (lane < 16) {
            uint32_t mm = bv;
            mm &= (lane & 1) ? b0 : ~b0;
            mm &= (lane & 2) ? b1 : ~b1;
            mm &= (lane & 4) ? b2 : ~b2;
            mm &= (lane & 8) ? b3 : ~b3;
            S.cnt8[((size_t)warp * n + i) * 16 + lane] = (uint8_t)__popc(mm);
          }
          if (valid && code4 != 15u && !onehot) {  // ambiguity code: log sum_{b in code} P_k(xp -> b) (bpp_c2.cpp:372-378)
            for (int k = 0; k < 3; k++) {
              const double* lp = logP + ((size_t)k * L.n_max + i) * 16 + xp * 4;
              double sum = 0;
              for (int b = 0; b < 4; b++)
                if ((code4 >> b) & 1u) sum += exp(lp[b]);
              atomicAdd(&S.eamb[i * 3 + k], log(sum));
            }
          }
        }
        __syncthreads();
        for (int idx = tid; idx < (n - 1) * 16; idx += kBD) {  // merge the per-warp tile counts
          uint32_t sum = 0;
          const int nd = idx >> 4;
          if (!(S.flags[nd] & PA_F_MISSING)) {
            for (int w = 0; w < kNW; w++) sum += S.cnt8[(size_t)w * n * 16 + idx];
            S.cnt[idx] += sum;
          }
        }
        __syncthreads();
      }
      if (tid == 0) w_hist += (unsigned long long)GG * ni;

      PA_PHASE(ph2)
      // ---------------- getEmission: e_s(k) = sum_bins count * log P_k (bpp_c2.cpp:355-390); missing nodes keep constants
      for (int task = tid; task < n * 3; task += kBD) {
        const int i = task / 3, k = task - 3 * i;
        double e = 0.0;
        if (i != root) {
          if (S.flags[i] & PA_F_MISSING) e = (k == 1) ? cp.log_constomis : 0.0;
          else {
            e = S.eamb[task];
            const double* lp = logP + ((size_t)k * L.n_max + i) * 16;
            for (int b = 0; b < 16; b++) {
              const uint32_t c = S.cnt[i * 16 + b];
              if (c) e += (double)c * lp[b];
            }
          }
        }
        S.emis[task] = e;
      }
      __syncthreads();

      PA_PHASE(ph3)
      // ---------------- Update_Z_subtree (bpp_c.cpp:1182-1293): scaled linear messages by tree level, one warp
      if (warp == 0) {
        if (lane == 0) S.ivars[IV_ZFALLBACK] = 0;
        for (int i = lane; i < n; i += 32) {
          double e0 = S.emis[i * 3], e1 = S.emis[i * 3 + 1], e2 = S.emis[i * 3 + 2];
          if (S.fix[i]) e2 = -CUDART_INF;
          const double mx = fmax(e0, fmax(e1, e2));
          S.back[i * 3] = exp(e0 - mx);
          S.back[i * 3 + 1] = exp(e1 - mx);
          S.back[i * 3 + 2] = exp(e2 - mx);
          S.uz[i] = u01(philox4(cp.seed, key, (uint32_t)i, 0u, (uint32_t)m, PA_ST_Z).x);
          if (!(mx > -CUDART_INF) || mx != mx) S.ivars[IV_ZFALLBACK] = 1;
        }
        __syncwarp();
        for (int lv = 1; lv < h->n_up; lv++) {
          for (int idx = S.up_off[lv] + lane; idx < S.up_off[lv + 1]; idx += 32) {
            const int i = S.up_nodes[idx];
            double b0 = S.back[i * 3], b1 = S.back[i * 3 + 1], b2 = S.back[i * 3 + 2];
#pragma unroll
            for (int cc = 0; cc < 2; cc++) {
              const int ch = S.child[2 * i + cc];
              const double* Tm = S.T + (S.fix[ch] ? 9 : 0);
              const double v0 = S.back[ch * 3], v1 = S.back[ch * 3 + 1], v2 = S.back[ch * 3 + 2];
              b0 *= (Tm[0] * v0 + Tm[3] * v1) + Tm[6] * v2;
              b1 *= (Tm[1] * v0 + Tm[4] * v1) + Tm[7] * v2;
              b2 *= (Tm[2] * v0 + Tm[5] * v1) + Tm[8] * v2;
            }
            const double mx = fmax(b0, fmax(b1, b2));
            if (!(mx > 0.0) || mx != mx) S.ivars[IV_ZFALLBACK] = 1;
            const double inv = 1.0 / mx;
            S.back[i * 3] = b0 * inv;
            S.back[i * 3 + 1] = b1 * inv;
            S.back[i * 3 + 2] = b2 * inv;
          }
          __syncwarp();
        }
        if (S.ivars[IV_ZFALLBACK] == 0) {
          if (lane == 0) {  // root: prior (1/2, 1/2, 0) (bpp_c.hpp:197-200)
            const int z = categorical3(S.back[root * 3] * 0.5, S.back[root * 3 + 1] * 0.5, 0.0, S.uz[root]);
            S.Znew[root] = (int8_t)z;
            if (z < 0) S.ivars[IV_ZFALLBACK] = 1;
          }
          __syncwarp();
          for (int lv = 1; lv < h->n_down; lv++) {
            for (int idx = S.down_off[lv] + lane; idx < S.down_off[lv + 1]; idx += 32) {
              const int i = S.down_nodes[idx];
              const int zp = S.Znew[S.parent[i]];
              const double* Tm = S.T + (S.fix[i] ? 9 : 0);
              const int z = (zp < 0) ? -1 : categorical3(Tm[0 + zp] * S.back[i * 3], Tm[3 + zp] * S.back[i * 3 + 1], Tm[6 + zp] * S.back[i * 3 + 2], S.uz[i]);
              S.Znew[i] = (int8_t)z;
              if (z < 0) S.ivars[IV_ZFALLBACK] = 1;
            }
            __syncwarp();
          }
        }
        __syncwarp();
        if (S.ivars[IV_ZFALLBACK] != 0 && lane == 0) z_update_logspace(S, n, S.T, nullptr);
        __syncwarp();
        if (lane == 0) {
          S.ivars[IV_ANY_CHANGED] = 0;
          S.ivars[IV_HAS1] = 0;
          S.ivars[IV_HAS2] = 0;
        }
      }
      __syncthreads();

      PA_PHASE(ph4)
      // ---------------- changed Z: dirty ancestors, class switch of P(t), class presence
      for (int i = tid; i < n; i += kBD) {
        const int zo = S.Z[i];
        int zn = S.Znew[i];
        if (zn < 0 || zn > 2) zn = zo;  // degenerate weights (NaN): keep the state
        if (zn != zo) {
          if (i != root) {
            mark_ancestors(S.parent, S.dirty, i);
            S.ivars[IV_ANY_CHANGED] = 1;
            double P16[16];
            build_P(sd, S.dist[i], class_rate(zn, cur_n, cur_c), true, P16);
#pragma unroll
            for (int e = 0; e < 16; e++) S.Pcur[i * 16 + e] = P16[e];
          }
          S.Z[i] = (int8_t)zn;
        }
        if (i != root && !(S.flags[i] & PA_F_MISSING)) {
          if (zn == 1) S.ivars[IV_HAS1] = 1;
          if (zn == 2) S.ivars[IV_HAS2] = 1;
        }
      }
      __syncthreads();
      if (S.ivars[IV_ANY_CHANGED]) {  // bpp_c.cpp:422-431
        const double ll = prune_all<true, true, true>(ev, S, ts, S.Pcur, S.dirty, gpart, gsc, s_pi, tid);
        if (tid == 0) {
          S.scal[SC_LL] = ll;
          int cnt_nodes = 0;
          for (int k = 0; k < ni && S.su[k] >= 0; k++) cnt_nodes += S.dirty[S.su[k]];
          w_npu += (unsigned long long)cnt_nodes * h->P;
          w_root += h->P;
        }
        __syncthreads();
      }
      for (int i = tid; i < n; i += kBD) S.dirty[i] = 0;
      __syncthreads();

      PA_PHASE(ph5)
      // ---------------- sample_rate: accelerated rate n (cls 2), then conserved rate c (cls 1)  (bpp_c.cpp:930-1176)
      for (int step = 0; step < 2; step++) {
        const bool neut = (step == 0);
        const int cls = neut ? 2 : 1;
        const uint32_t which = neut ? 0u : 1u;
        const int has = neut ? S.ivars[IV_HAS2] : S.ivars[IV_HAS1];
        const double prior_a = neut ? cp.nprior_a : cp.cprior_a, prior_b = neut ? cp.nprior_b : cp.cprior_b;
        const double old_rate = neut ? S.scal[SC_N] : S.scal[SC_C];
        // dirty set: ancestors of the class' branches, plus nodes whose stored partials are stale
        for (int i = tid; i < n - 1; i += kBD) {
          if (has && !(S.flags[i] & PA_F_MISSING) && S.Z[i] == cls) mark_ancestors(S.parent, S.dirty, i);
          if (S.stale[i]) S.dirty[i] = 1;
        }
        if (tid == 0) {
          if (S.stale[root]) S.dirty[root] = 1;
          int valid = 0;
          double cur_r = old_rate, proposal = 0.0, r = 0.0;
          if (!has) {  // no branch in this class: draw from the prior (bpp_c.cpp:940-964)
            if (resZ != 0) {
              cur_r = sched_gamma(cp.seed, key, which, 1, (uint32_t)m, PA_ST_RATE, prior_a, prior_b);
              if (neut) {
                if (cp.ropt == 1 && cur_r < cp.nlb) cur_r = old_rate;
                else if (cp.ropt == 2 && cur_r < S.scal[SC_C]) cur_r = old_rate;
              } else {
                if (cp.ropt == 1 && cur_r > cp.cub) cur_r = old_rate;
                else if (cp.ropt == 2 && cur_r > S.scal[SC_N]) cur_r = old_rate;
              }
            }
          } else {
            r = u01(philox4(cp.seed, key, which, 0u, (uint32_t)m, PA_ST_RATE).x);
            valid = 1;
            if (neut) {
              proposal = sched_gamma(cp.seed, key, which, 1, (uint32_t)m, PA_ST_RATE, old_rate / S.scal[SC_PROPN], S.scal[SC_PROPN]);
              if (cp.ropt == 1 && proposal < cp.nlb) valid = 0;
              else if (cp.ropt == 2 && proposal < S.scal[SC_C]) valid = 0;
            } else {
              const double pc = S.scal[SC_PROPC];
              const double u = r * (pc - 1 / pc) + 1 / pc;
              proposal = old_rate * u;
              if (cp.ropt == 1 && proposal > cp.cub) valid = 0;
              else if (cp.ropt == 2 && proposal > S.scal[SC_N]) valid = 0;
            }
            if (proposal < 1e-12 || proposal > 1e8) valid = 0;
          }
          S.scal[SC_PROPOSAL] = valid ? proposal : cur_r;
          S.scal[SC_R] = r;
          S.ivars[IV_VALID] = valid;
          S.ivars[IV_ACCEPT] = 0;
          S.ivars[IV_RATE_CHANGED] = (!valid && cur_r != old_rate) ? 1 : 0;
        }
        __syncthreads();
        if (S.ivars[IV_VALID]) {
          const double proposal = S.scal[SC_PROPOSAL];
          for (int i = tid; i < n - 1; i += kBD) {  // P(t * proposal) for the class' branches, current P elsewhere
            if (S.Z[i] == cls) {
              double P16[16];
              build_P(sd, S.dist[i], proposal, true, P16);
#pragma unroll
              for (int e = 0; e < 16; e++) S.Pprop[i * 16 + e] = P16[e];
            } else {
#pragma unroll
              for (int e = 0; e < 16; e++) S.Pprop[i * 16 + e] = S.Pcur[i * 16 + e];
            }
          }
          __syncthreads();
          const double ll_new = prune_all<true, false, true>(ev, S, ts, S.Pprop, S.dirty, gpart, gsc, s_pi, tid);  // log_f_Xz
          if (tid == 0) {
            const double ll_old = S.scal[SC_LL], r = S.scal[SC_R];
            double MH;
            if (neut) {
              const double pn = S.scal[SC_PROPN];
              MH = ll_new - ll_old + log_gamma_pdf(proposal, prior_a, prior_b) - log_gamma_pdf(old_rate, prior_a, prior_b) +
                   log_gamma_pdf(old_rate, proposal / pn, pn) - log_gamma_pdf(proposal, old_rate / pn, pn);
            } else {
              MH = ll_new - ll_old + log_gamma_pdf(proposal, prior_a, prior_b) - log_gamma_pdf(old_rate, prior_a, prior_b) + log(old_rate) - log(proposal);
            }
            if (log(r) < MH) {
              S.ivars[IV_ACCEPT] = 1;
              S.ivars[IV_RATE_CHANGED] = (proposal != old_rate) ? 1 : 0;
              S.scal[SC_LL] = ll_new;
            }
            int cnt_nodes = 0, cnt_cls = 0;
            for (int k = 0; k < ni && S.su[k] >= 0; k++) cnt_nodes += S.dirty[S.su[k]];
            for (int i = 0; i < n - 1; i++) cnt_cls += (S.Z[i] == cls);
            w_npu += (unsigned long long)cnt_nodes * h->P;
            w_root += h->P;
            w_pt += cnt_cls;
          }
          __syncthreads();
        }
        const bool accepted = S.ivars[IV_ACCEPT] != 0;
        const bool rate_changed = S.ivars[IV_RATE_CHANGED] != 0;
        const double new_rate = (accepted || (!S.ivars[IV_VALID])) ? S.scal[SC_PROPOSAL] : old_rate;
        if (rate_changed) {  // rebuild the class' P(t) / log P for every branch (bpp_c.cpp:1115-1131)
          for (int i = tid; i < n - 1; i += kBD) {
            double P16[16];
            build_P(sd, S.dist[i], new_rate, true, P16);
#pragma unroll
            for (int e = 0; e < 16; e++) logP[((size_t)cls * L.n_max + i) * 16 + e] = log(P16[e]);
            if (S.Z[i] == cls) {
#pragma unroll
              for (int e = 0; e < 16; e++) S.Pcur[i * 16 + e] = P16[e];
            }
          }
        }
        for (int i = tid; i < n; i += kBD) {
          if (accepted && S.dirty[i]) S.stale[i] = 1;
          S.dirty[i] = 0;
        }
        __syncthreads();
        if (tid == 0) {
          if (accepted) S.ivars[IV_STALE] = 1;
          if (neut) S.scal[SC_N] = new_rate; else S.scal[SC_C] = new_rate;
          if (rate_changed) {
            if (neut) S.ivars[IV_ACC_N] += 1; else S.ivars[IV_ACC_C] += 1;
            w_pt += n - 1;
          }
          if ((m + 1) % cp.adaptive_freq == 0) {  // adaptive proposal scale (bpp_c.cpp:1145-1171)
            const int kq = (m + 1) / cp.adaptive_freq;
            const double stp = fmin(0.5, 1 / sqrt((double)kq));
            const int acc = neut ? S.ivars[IV_ACC_N] : S.ivars[IV_ACC_C];
            double scale_adj = 1;
            if ((double)acc / cp.adaptive_freq > 0.44) scale_adj = exp(stp);
            else if ((double)acc / cp.adaptive_freq < 0.23) scale_adj = exp(-stp);
            if (neut) {
              double pn = S.scal[SC_PROPN] * scale_adj;
              if (pn > 1) pn = 1;
              S.scal[SC_PROPN] = pn;
              S.ivars[IV_ACC_N] = 0;
            } else {
              double pc = pow(S.scal[SC_PROPC], scale_adj);
              if (pc < 1.0001) pc = 1.0001;
              S.scal[SC_PROPC] = pc;
              S.ivars[IV_ACC_C] = 0;
            }
          }
        }
        __syncthreads();
        if (neut) { PA_PHASE(ph6) } else { PA_PHASE(ph7) }
      }

      // ---------------- commit: bring the stored partials of stale nodes up to date with the accepted rates
      if (S.ivars[IV_STALE]) {
        prune_all<true, true, false>(ev, S, ts, S.Pcur, S.stale, gpart, gsc, s_pi, tid);
        if (tid == 0) {
          int cnt_nodes = 0;
          for (int k = 0; k < ni && S.su[k] >= 0; k++) cnt_nodes += S.stale[S.su[k]];
          w_npu += (unsigned long long)cnt_nodes * h->P;
          S.ivars[IV_STALE] = 0;
        }
        __syncthreads();
        for (int i = tid; i < n; i += kBD) S.stale[i] = 0;
      }

      PA_PHASE(ph8)
      // ---------------- sample_transition (bpp_c.cpp:522-613)
      if (tid < 18) S.nz[tid] = 0;
      __syncthreads();
      for (int i = tid; i < n - 1; i += kBD) {
        const int k = S.Z[i] * 3 + S.Z[S.parent[i]];
        atomicAdd(&S.nz[k], 1);
        if (!S.fix[i]) atomicAdd(&S.nz[9 + k], 1);
      }
      __syncthreads();
      if (tid < 3) {
        const int* nZ = S.nz;
        const int* nF = S.nz + 9;
        if (tid == 0) S.scal[SC_G] = sched_beta(cp.seed, key, 0u, (uint32_t)m, cp.prior_g_a + nZ[1 * 3 + 0] + nZ[2 * 3 + 0], cp.prior_g_b + nZ[0]);
        else if (tid == 1) S.scal[SC_L] = sched_beta(cp.seed, key, 1u, (uint32_t)m, cp.prior_l_a + nF[2 * 3 + 1], cp.prior_l_b + nF[1 * 3 + 1]);
        else S.scal[SC_L2] = (cp.prior_l2_a == 0) ? 0.0 : sched_beta(cp.seed, key, 2u, (uint32_t)m, cp.prior_l2_a + nF[1 * 3 + 2], cp.prior_l2_b + nF[2 * 3 + 2]);
      }
      __syncthreads();
      if (tid == 0) {
        const double g = S.scal[SC_G], l = S.scal[SC_L], l2 = S.scal[SC_L2];
        double* Tf = S.T;
        double* Tx = S.T + 9;
        const double y = 1 - l;
        Tf[0] = 1 - g; Tf[3] = g; Tf[4] = y; Tf[7] = 1 - y; Tf[5] = l2; Tf[8] = 1 - l2;
        Tx[0] = 1 - g; Tx[3] = g;
      }
      __syncthreads();
      PA_PHASE(ph9)
    }  // iterations
  }    // chains

  if (tid == 0 && L.counters) {
    atomicAdd(&L.counters[0], w_npu);
    atomicAdd(&L.counters[1], w_root);
    atomicAdd(&L.counters[2], w_pt);
    atomicAdd(&L.counters[3], w_hist);
    atomicAdd(&L.counters[4], w_iter);
    const long long phs[10] = {ph0, ph1, ph2, ph3, ph4, ph5, ph6, ph7, ph8, ph9};
    for (int k = 0; k < 10; k++) atomicAdd(&L.counters[8 + k], (unsigned long long)phs[k]);
  }
}

}  // namespace

size_t chain_smem_bytes(int n_max, int slots_max, int BD) { return carve(nullptr, nullptr, n_max, slots_max, BD) + 64; }

size_t chain_arena_bytes(int n_max, int P_max) {
  size_t b = (size_t)3 * n_max * 16 * 8 + (size_t)n_max * P_max * 4 * 8 + (size_t)n_max * P_max * 2;
  return (b + 255) & ~size_t(255);
}

int chain_block_dim() { return kBD; }

int launch_chains(const ChainLaunch& L, int grid, int BD, cudaStream_t stream) {
  if (BD != kBD) return PA_ERR_ARG;
  const size_t smem = chain_smem_bytes(L.n_max, L.slots_max, kBD);
  cudaError_t err = cudaFuncSetAttribute(chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (err != cudaSuccess) {
    set_error(std::string("chain kernel shared memory: ") + cudaGetErrorString(err));
    return PA_ERR_LIMIT;
  }
  chain_kernel<<<grid, kBD, smem, stream>>>(L);
  err = cudaGetLastError();
  if (err != cudaSuccess) {
    set_error(std::string("chain kernel launch: ") + cudaGetErrorString(err));
    return PA_ERR_CUDA;
  }
  return PA_OK;
}

}  // namespace pa
