// pa_chain_warp.cu -- the on-device PhyloAcc-ST chain for typical elements: a SMALL CTA (W = 1 or 2 warps) per
// (element, model) chain, 6-7 chains resident per SM.
//
// Same sampler, same draw schedule and same arithmetic as pa_chain.cu (see the phase list there; reference:
// BPP_C::initMCMC + BPP_C::Gibbs, bpp_c.cpp:131-519), laid out for throughput on elements of up to a few hundred site
// patterns:
//   * few threads per chain: the serial parts of one chain (Z messages, MH scalars) overlap with the wide parts of
//     the other chains resident on the SM instead of idling 7 of 8 warps at a barrier;
//   * pruning passes give every lane TWO patterns of a 64-pattern tile (descriptor decode, P(t) loads and address
//     arithmetic amortised, 8 independent FMA chains per lane), run a per-pass FILTERED program whose operands are
//     already resolved (forwarded in registers / arena slab / tip row), and fetch the operands of node k+1 while node k
//     is computed (ping-pong register buffers); there is no shared-memory temp stack: a result that waits for a
//     sibling subtree is re-read from the arena copy it was written to;
//   * history sampling walks the internal nodes top-down over a tile of RS x 32 sites per warp with the partials and
//     tip codes of the next two nodes in flight (three register stages); parent->child base pairs are counted in
//     nibble-packed 64-bit registers and summed over the warp with four REDUX.SUM per branch;
//   * the pruning pass and the history phase are out-of-line functions with their own register allocation; everything
//     they need sits in a context block in shared memory; nothing on the per-iteration path touches local memory.
// Arena (HBM / L2, one per resident CTA): log P tables [3][n_max][16]; partials [2 copies][NIa nodes][2 halves][Pst][2]
// f64 (half 0 = bases A,C; half 1 = G,T; every 16-byte access of a warp is contiguous); scale counts [2][NIa][Pst] i16.
#include <cuda_runtime.h>

#include <cstdio>

#include "pa_chain.cuh"
#include "pa_device.cuh"

namespace pa {

namespace {

constexpr unsigned kFull = 0xffffffffu;

enum { SC_N = 0, SC_C, SC_G, SC_L, SC_L2, SC_LL, SC_PROPN, SC_PROPC, SC_PROPOSAL, SC_MHPART, SC_R, SC_CURR, SC_COUNT = 12 };
enum { IV_ANY_CHANGED = 0, IV_HAS1, IV_HAS2, IV_ZFALLBACK, IV_ACC_N, IV_ACC_C, IV_VALID, IV_NSU, IV_NACT, IV_NH, IV_SLOT, IV_COUNT = 12 };

template <int NMAX_, int W_, int RS_>
struct GSmem {
  static constexpr int NMAX = NMAX_, W = W_, RS = RS_;
  static constexpr int NI = NMAX / 2 + 1;   // internal nodes
  static constexpr int BD = 32 * W;         // threads per chain
  static constexpr int SITE_TILE = 32 * RS;  // sites per warp per history tile
  alignas(16) double Pcur[NMAX * 16];
  union {  // buffers that are live in disjoint phases of an iteration share storage
    struct { alignas(16) double Pprop[NMAX * 16]; };  // rate steps: proposal P(t)
    struct {                                          // history sampling .. Z update
      double emis[NMAX * 3], back[NMAX * 3];
      uint32_t cnt[NMAX * 8];               // 16 u16 bins per branch
      uint8_t xs8[W * NI * SITE_TILE];      // sampled bases [warp][internal node][site of the warp's tile]
    };
  };
  SubstDev sd;
  double dist[NMAX], uz[NMAX];
  double scal[SC_COUNT];
  double T[18];  // linear Z transition matrices [free, fixZ==1][child state a][parent state j]
  double red[W];
  alignas(16) uint4 ops[NI];  // full pruning program (see init)
  alignas(16) uint4 act[NI];  // program of the current pass: only the nodes to recompute, operands resolved
  alignas(8) uint2 hs[NI];    // history descriptor of internal node q: {node | parent's q << 16 (0xffff root), unused}
  alignas(8) uint2 ht[NI];    // its non-missing tip children: {node | tip row << 16} each, 0xffffffff = none
  struct {  // per-chain constants the out-of-line phases read back instead of carrying them in registers
    double* part;             // arena partials of this CTA
    int16_t* sc;              // arena scale counts
    double* logP;             // log P tables [3][n_max][16]
    const uint16_t* weights;  // blob: pattern weights
    const uint16_t* sitepat;  // blob: pattern of each (sorted) site
    const uint32_t* gap;      // blob: gap masks of internal nodes [W][Ppad]
    const uint8_t* codes2;    // blob: tip codes in tile layout
    uint32_t Pst, NIa, crow, seed, key;
    int n, ni, P, Ppad, GG, Wgap, n_max;
  } cx;
  int ivars[IV_COUNT];
  int nz[18];
  unsigned long long work[5];   // node-pattern updates, root evaluations, P(t) builds, history draws, iterations (thread 0)
  long long phase[10], t_last;  // SM cycles per phase (thread 0)
  int16_t parent[NMAX], child[2 * NMAX], up_nodes[NMAX], up_off[NMAX + 1], down_nodes[NMAX], down_off[NMAX + 1];
  uint16_t hq[NI];  // internal indices q of the non-missing internal nodes, root first (history order)
  uint8_t flags[NMAX], fix[NMAX], dirty[NMAX], sel[NMAX];
  int8_t Z[NMAX], Znew[NMAX];
};

template <class SM>
__device__ __forceinline__ SM& smem_of() {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  return *reinterpret_cast<SM*>(smem_raw);
}

template <int W>
__device__ __forceinline__ void cta_sync() {
  if (W == 1) __syncwarp();
  else __syncthreads();
}

struct WArena {
  double* part;       // [2 * NIa slabs][2 halves][Pst][2]
  int16_t* sc;        // [2 * NIa slabs][Pst]
  uint32_t slab_d;    // doubles per slab = 4 * Pst
  uint32_t half_d;    // doubles per half = 2 * Pst
  uint32_t Pst;       // patterns per slab row (multiple of 64)
  uint32_t NIa;       // slabs per copy
};

template <class SM>
__device__ __forceinline__ WArena arena_of(const SM& S) {
  WArena A;
  A.part = S.cx.part; A.sc = S.cx.sc; A.Pst = S.cx.Pst; A.NIa = S.cx.NIa; A.slab_d = A.Pst * 4u; A.half_d = A.Pst * 2u;
  return A;
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

// ---- scalar helpers kept out of line: one copy each, no effect on the register allocation of the hot loops
__device__ __noinline__ double sched_gamma_ol(uint32_t seed, uint32_t key, uint32_t idx, uint32_t sub0, uint32_t iter, uint32_t stream, double a, double b) {
  return sched_gamma(seed, key, idx, sub0, iter, stream, a, b);
}
__device__ __noinline__ double log_gamma_pdf_ol(double x, double a, double b) { return log_gamma_pdf(x, a, b); }

// P(t * rate) of one branch from the shared eigensystem (bpp.hpp:375-380; identity for t <= 1e-10), written to dstP
// (shared memory) and / or as logs to dstLog (arena); same arithmetic as build_P
template <class SM>
__device__ __noinline__ void build_P_to(double t, double rate, double* dstP, double* dstLog) {
  const SubstDev& sd = smem_of<SM>().sd;
  const bool ident = !(t > 1e-10);
  double e0 = 0, e1 = 0, e2 = 0, e3 = 0;
  if (!ident) {
    e0 = exp(sd.eval[0] * t * rate); e1 = exp(sd.eval[1] * t * rate); e2 = exp(sd.eval[2] * t * rate); e3 = exp(sd.eval[3] * t * rate);
  }
#pragma unroll 1
  for (int k = 0; k < 16; k++) {
    const double v = ident ? (((k >> 2) == (k & 3)) ? 1.0 : 0.0) : (((e0 * sd.A[0][k] + e1 * sd.A[1][k]) + e2 * sd.A[2][k]) + e3 * sd.A[3][k]);
    if (dstP) dstP[k] = v;
    if (dstLog) dstLog[k] = log(v);
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------ pruning
// Program of one pass (built by warp 0).  Descriptor of a recomputed node: x = output slab; y, z = children:
//   bits 0..1 kind (0 missing, 1 tip, 2 result of the previous node (registers), 3 arena slab), bits 2..11 child node
//   (its P(t)), bits 12.. payload (tip row | slab); w = node.
template <bool USE_DIRTY, class SM>
__device__ __forceinline__ int build_pass_program(SM& S, int wflip, int n_su, int lane) {
  const uint32_t NIa = S.cx.NIa;
  int base = 0;
  int carry = -1;  // node of the last op included so far
  for (int k0 = 0; k0 < n_su; k0 += 32) {
    const int k = k0 + lane;
    bool inc = false;
    uint4 op = make_uint4(0, 0, 0, 0);
    if (k < n_su) {
      op = S.ops[k];
      inc = !USE_DIRTY || S.dirty[op.x & 0xffff];
    }
    const int s = op.x & 0xffff;
    const unsigned bal = __ballot_sync(kFull, inc);
    // the op that precedes this one IN THE PASS (not in the full program): only its result can be forwarded in registers;
    // everything else is read back from the arena, where it was stored at least one op earlier than the read is issued
    const unsigned before = bal & ((1u << lane) - 1u);
    const int prev_lane = before ? (31 - __clz(before)) : 0;
    const int prev_in_chunk = __shfl_sync(kFull, s, prev_lane);
    const int prev_s = before ? prev_in_chunk : carry;
    if (inc) {
      const uint32_t q = op.x >> 16;
      uint4 d;
      d.x = (uint32_t)(S.sel[s] ^ wflip) * NIa + q;
      d.w = (uint32_t)s;
#pragma unroll
      for (int cc = 0; cc < 2; cc++) {
        const uint32_t kind = (op.z >> (2 * cc)) & 3u;
        const uint32_t ch = cc ? (op.y >> 16) : (op.y & 0xffff);
        const uint32_t xi = cc ? (op.w >> 16) : (op.w & 0xffff);
        uint32_t e = 0;
        if (kind == 1u) e = 1u | (ch << 2) | (xi << 12);
        else if (kind == 2u) {
          const bool dch = !USE_DIRTY || S.dirty[ch];
          if (dch && (int)ch == prev_s) e = 2u | (ch << 2);
          else e = 3u | (ch << 2) | (((uint32_t)(S.sel[ch] ^ (dch ? wflip : 0)) * NIa + xi) << 12);
        }
        if (cc) d.z = e; else d.y = e;
      }
      S.act[base + __popc(before)] = d;
    }
    if (bal) carry = __shfl_sync(kFull, s, 31 - __clz(bal));
    base += __popc(bal);
  }
  return base;
}

template <int NR>
struct ChildIn {
  double v[NR][4];
  int sc[NR];
  uint32_t code;  // tip child: nibble r = allowed-base mask of the lane's pattern in row r
};

template <int NR>
__device__ __forceinline__ void gather_child(uint32_t d, const WArena& A, const uint8_t* __restrict__ cbase, uint32_t crow, uint32_t idx0, ChildIn<NR>& c) {
  const uint32_t kind = d & 3u, pay = d >> 12;
  if (kind == 3u) {
    const double* b = A.part + ((size_t)pay * A.slab_d + idx0 * 2u);
    const int16_t* sp = A.sc + ((size_t)pay * A.Pst + idx0);
#pragma unroll
    for (int r = 0; r < NR; r++) {
      const double2 u = __ldcg(reinterpret_cast<const double2*>(b + r * 64));
      const double2 w = __ldcg(reinterpret_cast<const double2*>(b + A.half_d + r * 64));
      c.v[r][0] = u.x; c.v[r][1] = u.y; c.v[r][2] = w.x; c.v[r][3] = w.y;
      c.sc[r] = __ldcg(sp + r * 32);
    }
  } else if (kind == 1u) {
    c.code = __ldg(cbase + (size_t)pay * crow);
  }
}

// message of one child: m_j = sum_i P_child(j -> i) L_child(i); tips select / add columns; gaps and missing children give 1.
// With FUSE the message is multiplied into m (the other child's message) instead of replacing it.
template <int NR, bool FUSE>
__device__ __forceinline__ void child_message(uint32_t d, const double* __restrict__ Pm, const ChildIn<NR>& c, const double (&o)[NR][4], const int (&osc)[NR],
                                              double (&m)[NR][4], int (&msc)[NR]) {
  const uint32_t kind = d & 3u;
  if (!FUSE) {
#pragma unroll
    for (int r = 0; r < NR; r++) {
      m[r][0] = 1.0; m[r][1] = 1.0; m[r][2] = 1.0; m[r][3] = 1.0;
      msc[r] = 0;
    }
  }
  if (kind >= 2u) {
    const double2* __restrict__ Pc = reinterpret_cast<const double2*>(Pm + ((d >> 2) & 0x3ffu) * 16);
    const double2 r00 = Pc[0], r01 = Pc[1], r10 = Pc[2], r11 = Pc[3], r20 = Pc[4], r21 = Pc[5], r30 = Pc[6], r31 = Pc[7];
#pragma unroll
    for (int r = 0; r < NR; r++) {
      double a0, a1, a2, a3;
      int asc;
      if (kind == 2u) { a0 = o[r][0]; a1 = o[r][1]; a2 = o[r][2]; a3 = o[r][3]; asc = osc[r]; }
      else { a0 = c.v[r][0]; a1 = c.v[r][1]; a2 = c.v[r][2]; a3 = c.v[r][3]; asc = c.sc[r]; }
      const double v0 = ((r00.x * a0 + r00.y * a1) + r01.x * a2) + r01.y * a3;
      const double v1 = ((r10.x * a0 + r10.y * a1) + r11.x * a2) + r11.y * a3;
      const double v2 = ((r20.x * a0 + r20.y * a1) + r21.x * a2) + r21.y * a3;
      const double v3 = ((r30.x * a0 + r30.y * a1) + r31.x * a2) + r31.y * a3;
      if (FUSE) { m[r][0] *= v0; m[r][1] *= v1; m[r][2] *= v2; m[r][3] *= v3; msc[r] += asc; }
      else { m[r][0] = v0; m[r][1] = v1; m[r][2] = v2; m[r][3] = v3; msc[r] = asc; }
    }
  } else if (kind == 1u) {
    const double* Pd = Pm + ((d >> 2) & 0x3ffu) * 16;
#pragma unroll
    for (int r = 0; r < NR; r++) {
      const uint32_t code = (c.code >> (4 * r)) & 15u;
      if (code != 15u) {
        double v0, v1, v2, v3;
        if ((code & (code - 1)) == 0) {
          const int b = __ffs(code) - 1;
          v0 = Pd[b]; v1 = Pd[4 + b]; v2 = Pd[8 + b]; v3 = Pd[12 + b];
        } else {
          const double e0 = (code & 1u) ? 1.0 : 0.0, e1 = (code & 2u) ? 1.0 : 0.0, e2 = (code & 4u) ? 1.0 : 0.0, e3 = (code & 8u) ? 1.0 : 0.0;
          v0 = ((e0 * Pd[0] + e1 * Pd[1]) + e2 * Pd[2]) + e3 * Pd[3];
          v1 = ((e0 * Pd[4] + e1 * Pd[5]) + e2 * Pd[6]) + e3 * Pd[7];
          v2 = ((e0 * Pd[8] + e1 * Pd[9]) + e2 * Pd[10]) + e3 * Pd[11];
          v3 = ((e0 * Pd[12] + e1 * Pd[13]) + e2 * Pd[14]) + e3 * Pd[15];
        }
        if (FUSE) { m[r][0] *= v0; m[r][1] *= v1; m[r][2] *= v2; m[r][3] *= v3; }
        else { m[r][0] = v0; m[r][1] = v1; m[r][2] = v2; m[r][3] = v3; }
      }
    }
  }
}

// one node of the pass program for the lane's NR patterns: both child messages, rescale, store to the arena
template <int NR>
__device__ __forceinline__ void prune_node(const uint4& op, const double* __restrict__ Pm, const WArena& A, uint32_t idx0, const ChildIn<NR>& A0,
                                           const ChildIn<NR>& B0, double (&o)[NR][4], int (&osc)[NR]) {
  // child 0's factor is applied first, then child 1's (same order as the reference's accumulation)
  double m[NR][4];
  int msc[NR];
  child_message<NR, false>(op.y, Pm, A0, o, osc, m, msc);
  child_message<NR, true>(op.z, Pm, B0, o, osc, m, msc);
  double* ob = A.part + ((size_t)op.x * A.slab_d + idx0 * 2u);
  int16_t* sb = A.sc + ((size_t)op.x * A.Pst + idx0);
#pragma unroll
  for (int r = 0; r < NR; r++) {
    double o0 = m[r][0], o1 = m[r][1], o2 = m[r][2], o3 = m[r][3];
    int sc = msc[r];
    // rescale by 2^256 when every entry dropped below 2^-256 (compare high words: all values are >= 0)
    const int hm = max(max(__double2hiint(o0), __double2hiint(o1)), max(__double2hiint(o2), __double2hiint(o3)));
    if (hm < 0x2FF00000 && hm > 0) {
      o0 *= PA_SCALE_UP; o1 *= PA_SCALE_UP; o2 *= PA_SCALE_UP; o3 *= PA_SCALE_UP;
      sc += 1;
    }
    o[r][0] = o0; o[r][1] = o1; o[r][2] = o2; o[r][3] = o3;
    osc[r] = sc;
    __stcg(reinterpret_cast<double2*>(ob + r * 64), make_double2(o0, o1));
    __stcg(reinterpret_cast<double2*>(ob + A.half_d + r * 64), make_double2(o2, o3));
    __stcg(sb + r * 32, (int16_t)sc);
  }
}

// one tile of 32 * NR patterns (lane's patterns: p0 + 32 r + lane) through the pass program; returns the lane's
// sum over its patterns of w_p log L_p
template <int NR, class SM>
__device__ __forceinline__ double prune_tile(SM& S, const double* __restrict__ Pm, const WArena& A, int n_act, const uint8_t* __restrict__ cbase, uint32_t crow,
                                             int P, int p0, int lane) {
  const uint32_t idx0 = (uint32_t)(p0 + lane);
  double o[NR][4];
  int osc[NR];
#pragma unroll
  for (int r = 0; r < NR; r++) {
    o[r][0] = 1.0; o[r][1] = 1.0; o[r][2] = 1.0; o[r][3] = 1.0;
    osc[r] = 0;
  }
  if (n_act > 0) {
    // two operand buffers per child, used alternately: the operands of node k+1 are in flight while node k is computed
    ChildIn<NR> A0, B0, A1, B1;
    A0.code = B0.code = A1.code = B1.code = 0xffu;
    uint4 op = S.act[0];
    gather_child<NR>(op.y, A, cbase, crow, idx0, A0);
    gather_child<NR>(op.z, A, cbase, crow, idx0, B0);
    for (int k = 0; k < n_act; k += 2) {
      uint4 nx = make_uint4(0, 0, 0, 0);
      if (k + 1 < n_act) {
        nx = S.act[k + 1];
        gather_child<NR>(nx.y, A, cbase, crow, idx0, A1);
        gather_child<NR>(nx.z, A, cbase, crow, idx0, B1);
      }
      prune_node<NR>(op, Pm, A, idx0, A0, B0, o, osc);
      if (k + 1 >= n_act) break;
      if (k + 2 < n_act) {
        op = S.act[k + 2];
        gather_child<NR>(op.y, A, cbase, crow, idx0, A0);
        gather_child<NR>(op.z, A, cbase, crow, idx0, B0);
      }
      prune_node<NR>(nx, Pm, A, idx0, A1, B1, o, osc);
    }
  }
  // the last node of the program is the root (or nothing was computed: missing root -> likelihood 1)
  double acc = 0.0;
  const uint16_t* __restrict__ weights = S.cx.weights;
#pragma unroll
  for (int r = 0; r < NR; r++) {
    const int p = p0 + 32 * r + lane;
    if (p < P) {
      const double lik = ((S.sd.pi[0] * o[r][0] + S.sd.pi[1] * o[r][1]) + S.sd.pi[2] * o[r][2]) + S.sd.pi[3] * o[r][3];
      acc += (double)__ldg(weights + p) * (log(lik) - osc[r] * PA_LOG_SCALE);
    }
  }
  return acc;
}

// One Felsenstein pass over every pattern of the element (linear space, fp64) through the program in S.act, with the
// current (prop == 0) or the proposal's (prop == 1) transition matrices.  Returns sum_p w_p log L_p to every thread
// (Update_Tg(.,.,false) / log_f_Xz, bpp_c.cpp:647-671, 860-928).  Out of line: its register allocation is its own.
template <class SM>
__device__ __noinline__ double prune_pass(int prop) {
  SM& S = smem_of<SM>();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const double* Pm = prop ? S.Pprop : S.Pcur;
  const WArena A = arena_of(S);
  const int P = S.cx.P, n_act = S.ivars[IV_NACT];
  const uint32_t crow = S.cx.crow;
  const uint8_t* codes2 = S.cx.codes2;
  double acc = 0.0;
  for (int t = warp; t * 64 < P; t += SM::W) {
    const int p0 = t * 64;
    const uint8_t* cbase = codes2 + t * 32 + lane;
    if (P - p0 > 32) acc += prune_tile<2>(S, Pm, A, n_act, cbase, crow, P, p0, lane);
    else acc += prune_tile<1>(S, Pm, A, n_act, cbase, crow, P, p0, lane);
  }
  __syncwarp();
  acc = warp_sum(acc);
  if (SM::W > 1) {
    if (lane == 0) S.red[warp] = acc;
    __syncthreads();
    acc = 0.0;
#pragma unroll
    for (int w = 0; w < SM::W; w++) acc += S.red[w];
    __syncthreads();
  }
  return acc;
}

// ------------------------------------------------------------------------------------------------ history sampling
// add the warp total of nibble-packed per-lane counts (16 bins x 4 bit) to a branch's 16 u16 bins (8 words)
template <int W>
__device__ __forceinline__ void add_counts(uint32_t* __restrict__ words, unsigned long long c, int lane) {
  const uint32_t lo = (uint32_t)c, hi = (uint32_t)(c >> 32);
  const uint32_t e0 = __reduce_add_sync(kFull, lo & 0x0F0F0F0Fu);         // bins 0 2 4 6
  const uint32_t e1 = __reduce_add_sync(kFull, (lo >> 4) & 0x0F0F0F0Fu);  // bins 1 3 5 7
  const uint32_t e2 = __reduce_add_sync(kFull, hi & 0x0F0F0F0Fu);         // bins 8 10 12 14
  const uint32_t e3 = __reduce_add_sync(kFull, (hi >> 4) & 0x0F0F0F0Fu);  // bins 9 11 13 15
  if (lane < 8) {
    const uint32_t e = (lane & 4) ? ((lane & 2) ? e3 : e2) : ((lane & 2) ? e1 : e0);
    const uint32_t v = __byte_perm(e, 0u, (lane & 1) ? 0x4342u : 0x4140u);
    if (W == 1) words[lane] += v;
    else if (v) atomicAdd(&words[lane], v);
  }
}
// bin b of a branch (pair code = parent base * 4 + child base) -> its count in the 8-word layout written by add_counts
__device__ __forceinline__ uint32_t count_of(const uint32_t* __restrict__ words, int b) {
  const int i = (b & 1) + 2 * (b >> 3), k = (b & 7) >> 1;
  return (words[2 * i + (k >> 1)] >> (16 * (k & 1))) & 0xffffu;
}

template <int RS>
struct HStage {  // what one node of the history walk needs from memory, fetched two nodes ahead
  double2 pa[RS], pc[RS];     // the node's partials at the sites' patterns
  uint32_t tb0[RS], tb1[RS];  // raw code bytes of its two tip children
};

// Ancestral-history sampling of iteration m for every site (Update_Tg(.,.,true), bpp_c.cpp:672-733), top-down over the
// internal nodes, with the parent->child base pairs of every branch counted into S.cnt (getEmission's sufficient
// statistic, bpp_c2.cpp:355-390) and the emissions of ambiguity-coded tips added to S.emis.  Out of line: its register
// allocation is its own.
template <class SM>
__device__ __noinline__ void history_phase(int m) {
  constexpr int RS = SM::RS, SITE_TILE = SM::SITE_TILE;
  SM& S = smem_of<SM>();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const WArena A = arena_of(S);
  const int GG = S.cx.GG, Ppad = S.cx.Ppad, Wgap = S.cx.Wgap, n_max = S.cx.n_max, nh = S.ivars[IV_NH];
  const uint32_t crow = S.cx.crow, seed = S.cx.seed, key = S.cx.key;
  const uint16_t* __restrict__ sitepat = S.cx.sitepat;
  const uint32_t* __restrict__ gap = S.cx.gap;
  const uint8_t* __restrict__ codes2 = S.cx.codes2;
  const double* __restrict__ logP = S.cx.logP;
  uint8_t* __restrict__ xs8 = S.xs8 + warp * (SM::NI * SITE_TILE);
  for (int g0 = warp * SITE_TILE; g0 < GG; g0 += SM::W * SITE_TILE) {
    const int nrows = min(RS, (GG - g0 + 31) >> 5);
    uint32_t gpat[RS], gw0[RS], gw1[RS];  // per site row: pattern (bit 31: no site), gap words of internal nodes 0..63
#pragma unroll
    for (int r = 0; r < RS; r++) {
      const int g = g0 + 32 * r + lane;
      const uint32_t pat = (g < GG) ? (uint32_t)__ldg(sitepat + g) : 0u;
      gpat[r] = pat | ((g < GG) ? 0u : 0x80000000u);
      gw0[r] = __ldg(gap + pat);
      gw1[r] = (Wgap >= 2) ? __ldg(gap + Ppad + pat) : 0u;
    }
    uint4 rnd[RS];
    int grp = -1;

    auto load_stage = [&](HStage<RS>& st, int i) {
      const int q = S.hq[i];
      const int s = S.hs[q].x & 0xffff;
      const uint2 tc = S.ht[q];
      const double* pb = A.part + (size_t)((uint32_t)S.sel[s] * A.NIa + (uint32_t)q) * A.slab_d;
#pragma unroll
      for (int r = 0; r < RS; r++) {
        if (r < nrows) {
          const uint32_t pat = gpat[r] & 0xffffu;
          st.pa[r] = __ldcg(reinterpret_cast<const double2*>(pb + pat * 2u));
          st.pc[r] = __ldcg(reinterpret_cast<const double2*>(pb + A.half_d + pat * 2u));
          const uint32_t cidx = (pat >> 6) * 32u + (pat & 31u);
          if (tc.x != 0xffffffffu) st.tb0[r] = __ldg(codes2 + (size_t)(tc.x >> 16) * crow + cidx);
          if (tc.y != 0xffffffffu) st.tb1[r] = __ldg(codes2 + (size_t)(tc.y >> 16) * crow + cidx);
        }
      }
    };

    auto process = [&](const HStage<RS>& st, int i) {
      const int q = S.hq[i];
      const uint32_t d = S.hs[q].x;
      const int s = d & 0xffff;
      if ((q >> 2) != grp) {  // four draws per (site, group of four nodes)
        grp = q >> 2;
#pragma unroll
        for (int r = 0; r < RS; r++)
          if (r < nrows) rnd[r] = philox4(seed, key, (uint32_t)(g0 + 32 * r + lane), (uint32_t)grp, (uint32_t)m, PA_ST_HIST);
      }
      const uint32_t qp = d >> 16;
      const bool isroot = (qp == 0xffffu);
      const uint2 tc = S.ht[q];
      unsigned long long cN = 0ull, cT0 = 0ull, cT1 = 0ull;
      uint32_t amb = 0u, xpk = 0u;  // rows with an ambiguity-coded tip child (bit 2r + child), sampled bases (2 bit per row)
      const int j = q & 3;
#pragma unroll
      for (int r = 0; r < RS; r++) {
        if (r < nrows) {
          const bool valid = !(gpat[r] >> 31);
          bool gapbit;
          if (q < 32) gapbit = (gw0[r] >> q) & 1u;
          else if (q < 64) gapbit = (gw1[r] >> (q - 32)) & 1u;
          else gapbit = (__ldg(gap + (size_t)(q >> 5) * Ppad + (gpat[r] & 0xffffu)) >> (q & 31)) & 1u;
          const bool active = valid && (isroot || !gapbit);
          const int xp = isroot ? 0 : (int)xs8[qp * SITE_TILE + 32 * r + lane];
          int x = 0;
          if (active) {
            const double u = u01(j == 0 ? rnd[r].x : (j == 1 ? rnd[r].y : (j == 2 ? rnd[r].z : rnd[r].w)));
            double p0, p1, p2, p3;
            if (isroot) {
              p0 = st.pa[r].x * S.sd.pi[0]; p1 = st.pa[r].y * S.sd.pi[1]; p2 = st.pc[r].x * S.sd.pi[2]; p3 = st.pc[r].y * S.sd.pi[3];
            } else {
              const double2* Pr = reinterpret_cast<const double2*>(S.Pcur + s * 16 + xp * 4);
              const double2 r0 = Pr[0], r1 = Pr[1];
              p0 = r0.x * st.pa[r].x; p1 = r0.y * st.pa[r].y; p2 = r1.x * st.pc[r].x; p3 = r1.y * st.pc[r].y;
            }
            // first k with u*tot < cumsum_k == number of cumulative sums not exceeding u*tot
            const double c1 = p0, c2 = c1 + p1, c3 = c2 + p2, xu = u * (c3 + p3);
            x = (int)(xu >= c1) + (int)(xu >= c2) + (int)(xu >= c3);
            xs8[q * SITE_TILE + 32 * r + lane] = (uint8_t)x;
            if (!isroot) cN += 1ull << (4 * (xp * 4 + x));
          }
          xpk |= (uint32_t)x << (2 * r);
          // tip children of this node: the tip's base is data, the parent's base was just drawn
          const uint32_t csh = (gpat[r] >> 3) & 4u;
          const uint32_t t0 = (tc.x != 0xffffffffu && valid) ? ((st.tb0[r] >> csh) & 15u) : 15u;
          const uint32_t t1 = (tc.y != 0xffffffffu && valid) ? ((st.tb1[r] >> csh) & 15u) : 15u;
          if (t0 != 15u) {
            if ((t0 & (t0 - 1)) == 0) cT0 += 1ull << (4 * (x * 4 + (__ffs(t0) - 1)));
            else amb |= 1u << (2 * r);
          }
          if (t1 != 15u) {
            if ((t1 & (t1 - 1)) == 0) cT1 += 1ull << (4 * (x * 4 + (__ffs(t1) - 1)));
            else amb |= 2u << (2 * r);
          }
        }
      }
      if (amb) {  // ambiguity-coded tips (rare): log sum_{b in code} P_k(x -> b) for the three classes (bpp_c2.cpp:372-378)
#pragma unroll 1
        for (int e = 0; e < 2 * RS; e++) {
          if (!((amb >> e) & 1u)) continue;
          const int r = e >> 1;
          const int g = g0 + 32 * r + lane;
          const uint32_t pat = (uint32_t)__ldg(sitepat + g);
          const uint32_t trow = ((e & 1) ? tc.y : tc.x) >> 16;
          const uint32_t code4 = ((uint32_t)__ldg(codes2 + (size_t)trow * crow + (pat >> 6) * 32u + (pat & 31u)) >> ((pat >> 3) & 4u)) & 15u;
          const int ti = ((e & 1) ? tc.y : tc.x) & 0xffff;
          const int x = (xpk >> (2 * r)) & 3;
#pragma unroll 1
          for (int kk = 0; kk < 3; kk++) {
            const double* lp = logP + ((size_t)kk * n_max + ti) * 16 + x * 4;
            double sum = 0;
            for (int b = 0; b < 4; b++)
              if ((code4 >> b) & 1u) sum += exp(lp[b]);
            atomicAdd(&S.emis[ti * 3 + kk], log(sum));
          }
        }
      }
      if (!isroot) add_counts<SM::W>(S.cnt + s * 8, cN, lane);
      if (tc.x != 0xffffffffu) add_counts<SM::W>(S.cnt + (tc.x & 0xffff) * 8, cT0, lane);
      if (tc.y != 0xffffffffu) add_counts<SM::W>(S.cnt + (tc.y & 0xffff) * 8, cT1, lane);
    };

    HStage<RS> s0, s1, s2;
    if (nh > 0) load_stage(s0, 0);
    if (nh > 1) load_stage(s1, 1);
    for (int i = 0; i < nh; i += 3) {
      if (i + 2 < nh) load_stage(s2, i + 2);
      process(s0, i);
      if (i + 1 < nh) {
        if (i + 3 < nh) load_stage(s0, i + 3);
        process(s1, i + 1);
      }
      if (i + 2 < nh) {
        if (i + 4 < nh) load_stage(s1, i + 4);
        process(s2, i + 2);
      }
    }
    __syncwarp();
  }
}

// log-space serial Z update, a literal restatement of Update_Z_subtree; used only when the scaled linear pass degenerates
template <class SM>
__device__ __noinline__ void z_update_logspace_w(int n) {
  SM& S = smem_of<SM>();
  double lt[2][9];
  for (int f = 0; f < 2; f++)
    for (int k = 0; k < 9; k++) lt[f][k] = log(S.T[f * 9 + k]);
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
    if (zp < 0 || zp > 2) { S.Znew[i] = -1; continue; }
    for (int a = 0; a < 3; a++) v[a] = t[a * 3 + zp] + S.back[i * 3 + a];
    const double mx = fmax(v[0], fmax(v[1], v[2]));
    S.Znew[i] = (int8_t)categorical3(exp(v[0] - mx), exp(v[1] - mx), exp(v[2] - mx), S.uz[i]);
  }
}

// build the program of a pass on warp 0 and publish its length
template <bool USE_DIRTY, class SM>
__device__ __forceinline__ int make_pass(SM& S, int wflip, int n_su, int tid) {
  if (tid < 32) {
    const int n_act = build_pass_program<USE_DIRTY>(S, wflip, n_su, tid);
    if (tid == 0) S.ivars[IV_NACT] = n_act;
  }
  cta_sync<SM::W>();
  return S.ivars[IV_NACT];
}

template <int NMAX, int W, int RS, int MINB>
__global__ void __launch_bounds__(32 * W, MINB) chain_warp_kernel(ChainLaunch L) {
  typedef GSmem<NMAX, W, RS> SM;
  constexpr int BD = SM::BD;
  SM& S = smem_of<SM>();
  const int tid = threadIdx.x, lane = tid & 31;
  const ChainParams& cp = L.cp;
  {
    const double* src = reinterpret_cast<const double*>(&L.sd);
    double* dst = reinterpret_cast<double*>(&S.sd);
    for (int i = tid; i < (int)(sizeof(SubstDev) / sizeof(double)); i += BD) dst[i] = src[i];
  }

  unsigned char* arena = L.arena + (size_t)blockIdx.x * L.arena_stride;
  double* logP = reinterpret_cast<double*>(arena);

  if (tid < 5) S.work[tid] = 0ull;
  if (tid < 10) S.phase[tid] = 0ll;
  if (tid == 0) S.t_last = clock64();
#define PA_PHASE(k)                     \
  if (tid == 0) {                       \
    const long long t_now = clock64();  \
    S.phase[k] += t_now - S.t_last;     \
    S.t_last = t_now;                   \
  }
#define w_npu S.work[0]
#define w_root S.work[1]
#define w_pt S.work[2]
#define w_hist S.work[3]
#define w_iter S.work[4]

  for (;;) {
    cta_sync<W>();
    if (tid == 0) S.ivars[IV_SLOT] = atomicAdd(L.queue, 1);
    cta_sync<W>();
    const int slot = S.ivars[IV_SLOT];
    if (slot >= L.n_sel) break;
    const int chain = L.sel ? L.sel[slot] : slot;
    const int elem = L.chain_list[2 * chain], mdl = L.chain_list[2 * chain + 1];
    const int resZ = (mdl == PA_M0) ? 0 : (mdl == PA_M1 ? 2 : 1);
    const uint32_t key = (uint32_t)elem * 4u + (uint32_t)mdl;
    const int T = L.T;
    // warm start (initMCMC with iter > 0, bpp_c.cpp:148-177): Z and the five rates continue from the last state of the
    // chain's previous run; otherwise the INIT_* values
    const double* warm = (L.init_rates && L.init_rates[(size_t)chain * 8 + 5] != 0.0) ? L.init_rates + (size_t)chain * 8 : nullptr;
    const double init_n = warm ? warm[0] : cp.init_nrate, init_c = warm ? warm[1] : cp.init_crate, init_g = warm ? warm[2] : cp.init_grate,
                 init_l = warm ? warm[3] : cp.init_lrate, init_l2 = warm ? warm[4] : cp.init_lrate2;
    int n, ni, GG, root, P;
    {
      // ---------------- initMCMC (bpp_c.cpp:131-343)
      const ElemView ev = make_view(L.blobs + L.blob_off[elem]);
      const ElemHeader* h = ev.h;
      n = h->n; ni = h->ni; GG = h->GG; root = n - 1; P = h->P;
      if (tid == 0) {
        const uint32_t Pst = (uint32_t)L.P_max;  // already a multiple of 64 (chain_warp_pst)
        const uint32_t NIa = (uint32_t)L.n_max / 2 + 1;
        S.cx.logP = logP;
        S.cx.part = logP + (size_t)3 * L.n_max * 16;
        S.cx.sc = reinterpret_cast<int16_t*>(S.cx.part + (size_t)2 * NIa * Pst * 4u);
        S.cx.weights = ev.weights; S.cx.sitepat = ev.sitepat; S.cx.gap = ev.gap; S.cx.codes2 = ev.codes2;
        S.cx.Pst = Pst; S.cx.NIa = NIa; S.cx.crow = (uint32_t)h->n_tile64 * 32u; S.cx.seed = cp.seed; S.cx.key = key;
        S.cx.n = n; S.cx.ni = ni; S.cx.P = P; S.cx.Ppad = h->Ppad; S.cx.GG = GG; S.cx.Wgap = h->W; S.cx.n_max = L.n_max;
        S.ivars[IV_NSU] = 0;
      }
      for (int i = tid; i < n; i += BD) {
        S.parent[i] = ev.parent[i];
        S.child[2 * i] = ev.child[2 * i];
        S.child[2 * i + 1] = ev.child[2 * i + 1];
        S.up_nodes[i] = ev.up_nodes[i];
        S.down_nodes[i] = ev.down_nodes[i];
        const uint8_t f = ev.flags[i];
        S.flags[i] = f;
        S.fix[i] = (f >> (2 + mdl)) & 1;
        S.dist[i] = ev.dist[i];
        S.Z[i] = warm ? (int8_t)((L.init_z[(size_t)chain * L.ZW + (i >> 4)] >> (2 * (i & 15))) & 3u) : (int8_t)((i == root) ? 0 : 1);  // bpp_c.cpp:141-151
        S.dirty[i] = 0;
        S.sel[i] = 0;
      }
      for (int i = tid; i <= h->n_up; i += BD) S.up_off[i] = ev.up_off[i];
      for (int i = tid; i <= h->n_down; i += BD) S.down_off[i] = ev.down_off[i];
      cta_sync<W>();
      // full pruning program: one descriptor per computed internal node, in the element's pruning order:
      //   x = node | q << 16;  y = child0 | child1 << 16;  z = kind0 | kind1 << 2;
      //   w = x0 | x1 << 16 (tip row of a tip child, internal index of an internal child); kind 0 missing, 1 tip, 2 internal
      int my_su = 0;
      for (int k = tid; k < ni; k += BD) {
        const int s = ev.su[k];
        uint4 op = make_uint4(0, 0, 0, 0);
        auto kind_of = [&](int c) { const uint8_t f = ev.flags[c]; return (f & PA_F_MISSING) ? 0u : ((f & PA_F_TIP) ? 1u : 2u); };
        if (s >= 0) {
          my_su++;
          const int c0 = ev.child[2 * s], c1 = ev.child[2 * s + 1];
          auto x_of = [&](int c) { return (uint32_t)((ev.flags[c] & PA_F_TIP) ? ev.tiprow[c] : ev.intidx[c]) & 0xffffu; };
          op.x = (uint32_t)s | ((uint32_t)ev.intidx[s] << 16);
          op.y = (uint32_t)c0 | ((uint32_t)c1 << 16);
          op.z = kind_of(c0) | (kind_of(c1) << 2);
          op.w = x_of(c0) | (x_of(c1) << 16);
        }
        S.ops[k] = op;
        // history descriptors of internal node q = k (post-order): parent's q, and the node's non-missing tip children
        const int sq = ev.intlist[k];
        const int par = ev.parent[sq];
        S.hs[k] = make_uint2((uint32_t)sq | ((par >= 0 ? (uint32_t)ev.intidx[par] : 0xffffu) << 16), 0u);
        uint2 t2 = make_uint2(0xffffffffu, 0xffffffffu);
        const int c0 = ev.child[2 * sq], c1 = ev.child[2 * sq + 1];
        if (kind_of(c0) == 1u) t2.x = (uint32_t)c0 | ((uint32_t)ev.tiprow[c0] << 16);
        if (kind_of(c1) == 1u) t2.y = (uint32_t)c1 | ((uint32_t)ev.tiprow[c1] << 16);
        S.ht[k] = t2;
      }
      if (my_su) atomicAdd(&S.ivars[IV_NSU], my_su);  // the computed nodes are a prefix of su
      if (tid == 0) {
        int nh = 0;
        for (int q = ni - 1; q >= 0; q--)
          if (!(ev.flags[ev.intlist[q]] & PA_F_MISSING)) S.hq[nh++] = (uint16_t)q;
        S.ivars[IV_NH] = nh;
        S.scal[SC_N] = init_n;
        S.scal[SC_C] = init_c;
        S.scal[SC_G] = init_g;
        S.scal[SC_L] = init_l;
        S.scal[SC_L2] = init_l2;
        S.scal[SC_PROPC] = 1.2;                  // bpp_c.cpp:135
        S.scal[SC_PROPN] = (double)1 / GG * 10;  // bpp_c.cpp:136
        S.ivars[IV_ACC_N] = 0;
        S.ivars[IV_ACC_C] = 0;
        // linear Z transition matrices (bpp_c.cpp:301-332), index [child state a][parent state j] -> a*3+j
        const double g = init_g, l = init_l, l2 = init_l2;
        double* Tf = S.T;      // free
        double* Tx = S.T + 9;  // fixZ == 1
        const double y = 1 - l;
        Tf[0] = 1 - g; Tf[3] = g; Tf[6] = 0; Tf[1] = 0; Tf[4] = y; Tf[7] = 1 - y; Tf[2] = 0; Tf[5] = l2; Tf[8] = 1 - l2;
        Tx[0] = 1 - g; Tx[3] = g; Tx[6] = 0; Tx[1] = 0; Tx[4] = 1; Tx[7] = 0; Tx[2] = 1; Tx[5] = 1; Tx[8] = 1;
        w_pt += 3ull * n;
      }
      cta_sync<W>();
    }
    const int n_su = S.ivars[IV_NSU];
    // P(t) of every branch for its class, and the log P tables of the three classes (emission look-ups); task = (class, node)
    for (int task = tid; task < 3 * n; task += BD) {
      const int k = task / n, i = task - k * n;
      const double rate = (k == 0) ? 1.0 : (k == 1 ? init_c : init_n);
      const int zi = S.Z[i];
      const bool mine = (k == ((zi == 0) ? 0 : (zi == 2 ? 2 : 1)));
      build_P_to<SM>(S.dist[i], rate, mine ? (S.Pcur + i * 16) : nullptr, logP + ((size_t)k * L.n_max + i) * 16);
    }
    cta_sync<W>();
    // initial upward pass over every pattern (Gibbs, bpp_c.cpp:355-356), partials stored in arena copy 0
    {
      const int n_act = make_pass<false>(S, 0, n_su, tid);
      const double ll = prune_pass<SM>(0);
      if (tid == 0) {
        S.scal[SC_LL] = ll;
        w_npu += (unsigned long long)n_act * P;
        w_root += P;
      }
    }
    cta_sync<W>();
    PA_PHASE(0)

    for (int m = 0; m < T; m++) {
      // ---------------- MonitorChain: trace row m
      {
        double* tr_d = L.trace_d + (size_t)chain * 6 * T;
        uint32_t* tr_z = L.trace_z + (size_t)chain * T * L.ZW;
        for (int w = tid; w < L.ZW; w += BD) {
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
      }
      if (m == T - 1) break;
      if (tid == 0) w_iter++;

      // ---------------- sample ancestral histories + count parent->child base pairs per branch
      for (int i = tid; i < n * 8; i += BD) S.cnt[i] = 0;
      for (int i = tid; i < n * 3; i += BD) S.emis[i] = 0.0;  // collects the ambiguous-tip emissions during sampling
      const double cur_n = S.scal[SC_N], cur_c = S.scal[SC_C];
      cta_sync<W>();
      PA_PHASE(1)
      history_phase<SM>(m);
      cta_sync<W>();
      if (tid == 0) w_hist += (unsigned long long)GG * ni;
      PA_PHASE(2)

      // ---------------- getEmission (bpp_c2.cpp:355-390) fused with the start of Update_Z_subtree (bpp_c.cpp:1186-1200):
      // e_s(k) = sum_bins count * log P_k, then b_s(k) = exp(e_s(k) - max) with state 2 masked on fixZ nodes, and the Z uniforms
      if (tid == 0) S.ivars[IV_ZFALLBACK] = 0;
      for (int task = tid; task < n * 3; task += BD) {
        const int i = task / 3, k = task - 3 * i;
        double e = 0.0;
        if (i != root) {
          if (S.flags[i] & PA_F_MISSING) e = (k == 1) ? cp.log_constomis : 0.0;
          else {
            e = S.emis[task];
            const double* lp = logP + ((size_t)k * L.n_max + i) * 16;
#pragma unroll
            for (int b = 0; b < 16; b++) {
              const uint32_t c = count_of(S.cnt + i * 8, b);
              if (c) e += (double)c * lp[b];
            }
          }
        }
        S.emis[task] = e;
      }
      cta_sync<W>();
      for (int i = tid; i < n; i += BD) {
        double e0 = S.emis[i * 3], e1 = S.emis[i * 3 + 1], e2 = S.emis[i * 3 + 2];
        if (S.fix[i]) e2 = -CUDART_INF;
        const double mx = fmax(e0, fmax(e1, e2));
        S.back[i * 3] = exp(e0 - mx);
        S.back[i * 3 + 1] = exp(e1 - mx);
        S.back[i * 3 + 2] = exp(e2 - mx);
        S.uz[i] = u01(philox4(cp.seed, key, (uint32_t)i, 0u, (uint32_t)m, PA_ST_Z).x);
        if (!(mx > -CUDART_INF) || mx != mx) S.ivars[IV_ZFALLBACK] = 1;
      }
      cta_sync<W>();
      PA_PHASE(3)

      // ---------------- Update_Z_subtree (bpp_c.cpp:1204-1293): scaled linear messages by tree level, on warp 0
      if (tid < 32) {
        int lv_up = 0, lv_down = 0;  // number of levels = first index whose offset reaches n
        while (S.up_off[lv_up] < n) lv_up++;
        while (S.down_off[lv_down] < n) lv_down++;
        for (int lv = 1; lv < lv_up; lv++) {
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
            if (mx < 1e-100) {  // keep the scaled messages away from underflow; the scale cancels in the draws
              const double inv = 1.0 / mx;
              b0 *= inv; b1 *= inv; b2 *= inv;
            }
            S.back[i * 3] = b0;
            S.back[i * 3 + 1] = b1;
            S.back[i * 3 + 2] = b2;
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
          for (int lv = 1; lv < lv_down; lv++) {
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
        if (S.ivars[IV_ZFALLBACK] != 0 && lane == 0) z_update_logspace_w<SM>(n);
        __syncwarp();
        if (lane == 0) {
          S.ivars[IV_ANY_CHANGED] = 0;
          S.ivars[IV_HAS1] = 0;
          S.ivars[IV_HAS2] = 0;
        }
      }
      cta_sync<W>();
      PA_PHASE(4)

      // ---------------- changed Z: dirty ancestors, class switch of P(t), class presence
      for (int i = tid; i < n; i += BD) {
        const int zo = S.Z[i];
        int zn = S.Znew[i];
        if (zn < 0 || zn > 2) zn = zo;  // degenerate weights (NaN): keep the state
        if (zn != zo) {
          if (i != root) {
            mark_ancestors(S.parent, S.dirty, i);
            S.ivars[IV_ANY_CHANGED] = 1;
            build_P_to<SM>(S.dist[i], class_rate(zn, cur_n, cur_c), S.Pcur + i * 16, nullptr);
          }
          S.Z[i] = (int8_t)zn;
        }
        if (i != root && !(S.flags[i] & PA_F_MISSING)) {
          if (zn == 1) S.ivars[IV_HAS1] = 1;
          if (zn == 2) S.ivars[IV_HAS2] = 1;
        }
      }
      cta_sync<W>();
      if (S.ivars[IV_ANY_CHANGED]) {  // bpp_c.cpp:422-431 (recomputed partials overwrite the current copy in place)
        const int n_act = make_pass<true>(S, 0, n_su, tid);
        const double ll = prune_pass<SM>(0);
        if (tid == 0) {
          S.scal[SC_LL] = ll;
          w_npu += (unsigned long long)n_act * P;
          w_root += P;
        }
      }
      for (int i = tid; i < n; i += BD) S.dirty[i] = 0;
      cta_sync<W>();
      PA_PHASE(5)

      // ---------------- sample_rate: accelerated rate n (cls 2), then conserved rate c (cls 1)  (bpp_c.cpp:930-1176)
      for (int step = 0; step < 2; step++) {
        const bool neut = (step == 0);
        const int cls = neut ? 2 : 1;
        const uint32_t which = neut ? 0u : 1u;
        const int has = neut ? S.ivars[IV_HAS2] : S.ivars[IV_HAS1];
        const double prior_a = neut ? cp.nprior_a : cp.cprior_a, prior_b = neut ? cp.nprior_b : cp.cprior_b;
        const double old_rate = neut ? S.scal[SC_N] : S.scal[SC_C];
        const double ll_cur = S.scal[SC_LL];
        if (has)
          for (int i = tid; i < n - 1; i += BD)
            if (!(S.flags[i] & PA_F_MISSING) && S.Z[i] == cls) mark_ancestors(S.parent, S.dirty, i);
        if (tid < 32) {
          // lane 0: proposal (or prior draw); lanes 0..3: the prior / Hastings terms of the MH ratio in parallel
          int valid = 0;
          double cur_r = old_rate, proposal = 0.0, r = 0.0;
          if (lane == 0) {
            if (!has) {  // no branch in this class: draw from the prior (bpp_c.cpp:940-964)
              if (resZ != 0) {
                cur_r = sched_gamma_ol(cp.seed, key, which, 1, (uint32_t)m, PA_ST_RATE, prior_a, prior_b);
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
                proposal = sched_gamma_ol(cp.seed, key, which, 1, (uint32_t)m, PA_ST_RATE, old_rate / S.scal[SC_PROPN], S.scal[SC_PROPN]);
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
          }
          valid = __shfl_sync(kFull, valid, 0);
          proposal = __shfl_sync(kFull, proposal, 0);
          double term = 0.0;
          if (valid) {
            const double pn = S.scal[SC_PROPN];
            double x = proposal, a = prior_a, b = prior_b, sgn = 1.0;               // lane 0: + log pdf(proposal | prior)
            if (lane == 1) { x = old_rate; sgn = -1.0; }                             // lane 1: - log pdf(current  | prior)
            if (lane == 2) { x = old_rate; a = proposal / pn; b = pn; }               // lane 2: + log q(current | proposal)   (n only)
            if (lane == 3) { x = proposal; a = old_rate / pn; b = pn; sgn = -1.0; }  // lane 3: - log q(proposal | current)  (n only)
            if (lane < 2 || (neut && lane < 4)) term = sgn * log_gamma_pdf_ol(x, a, b);
            else if (!neut && lane == 2) term = log(old_rate) - log(proposal);
          }
          const double t1 = __shfl_sync(kFull, term, 1), t2 = __shfl_sync(kFull, term, 2), t3 = __shfl_sync(kFull, term, 3);
          if (lane == 0) {
            S.scal[SC_MHPART] = ((term + t1) + t2) + t3;
            S.scal[SC_PROPOSAL] = proposal;
            S.scal[SC_CURR] = cur_r;
            S.scal[SC_R] = r;
            S.ivars[IV_VALID] = valid;
          }
        }
        cta_sync<W>();
        const int valid = S.ivars[IV_VALID];
        const double proposal = S.scal[SC_PROPOSAL], cur_r = S.scal[SC_CURR];
        int accepted = 0, rate_changed = (!valid && cur_r != old_rate) ? 1 : 0;
        if (valid) {
          // P(t * proposal) for the class' branches, current P elsewhere; task = branch
          for (int i = tid; i < n - 1; i += BD) {
            if (S.Z[i] == cls) {
              build_P_to<SM>(S.dist[i], proposal, S.Pprop + i * 16, nullptr);
            } else {
              const double2* src = reinterpret_cast<const double2*>(S.Pcur + i * 16);
              double2* dst = reinterpret_cast<double2*>(S.Pprop + i * 16);
#pragma unroll
              for (int e = 0; e < 8; e++) dst[e] = src[e];
            }
          }
          const long long tp0 = clock64();
          const int n_act = make_pass<true>(S, 1, n_su, tid);
          const double ll_new = prune_pass<SM>(1);                            // log_f_Xz, into the non-current copy
          if (tid == 0 && !neut) S.phase[8] += clock64() - tp0;               // the c-rate pruning pass alone (also included in the c_rate phase)
          const double MH = (ll_new - ll_cur) + S.scal[SC_MHPART];
          if (log(S.scal[SC_R]) < MH) {
            accepted = 1;
            rate_changed = (proposal != old_rate) ? 1 : 0;
          }
          cta_sync<W>();
          if (tid == 0) {
            if (accepted) S.scal[SC_LL] = ll_new;
            int cnt_cls = 0;
            for (int i = 0; i < n - 1; i++) cnt_cls += (S.Z[i] == cls);
            w_npu += (unsigned long long)n_act * P;
            w_root += P;
            w_pt += cnt_cls;
          }
        }
        const double new_rate = valid ? (accepted ? proposal : old_rate) : cur_r;
        if (rate_changed) {  // rebuild the class' P(t) / log P for every branch (bpp_c.cpp:1115-1131); task = branch
          for (int i = tid; i < n - 1; i += BD)
            build_P_to<SM>(S.dist[i], new_rate, (S.Z[i] == cls) ? (S.Pcur + i * 16) : nullptr, logP + ((size_t)cls * L.n_max + i) * 16);
        }
        for (int i = tid; i < n; i += BD) {
          if (accepted && S.dirty[i]) S.sel[i] ^= 1;  // the proposal's partials become the current ones
          S.dirty[i] = 0;
        }
        if (tid == 0) {
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
        cta_sync<W>();
        if (neut) { PA_PHASE(6) } else { PA_PHASE(7) }
      }

      // ---------------- sample_transition (bpp_c.cpp:522-613): six gamma draws on six lanes
      if (tid < 18) S.nz[tid] = 0;
      cta_sync<W>();
      for (int i = tid; i < n - 1; i += BD) {
        const int k = S.Z[i] * 3 + S.Z[S.parent[i]];
        atomicAdd(&S.nz[k], 1);
        if (!S.fix[i]) atomicAdd(&S.nz[9 + k], 1);
      }
      cta_sync<W>();
      if (tid < 32) {
        const int* nZ = S.nz;
        const int* nF = S.nz + 9;
        const int whichb = lane >> 1;  // 0: g, 1: l, 2: l2
        double a = 1.0, b = 1.0;
        if (whichb == 0) { a = cp.prior_g_a + nZ[1 * 3 + 0] + nZ[2 * 3 + 0]; b = cp.prior_g_b + nZ[0]; }
        else if (whichb == 1) { a = cp.prior_l_a + nF[2 * 3 + 1]; b = cp.prior_l_b + nF[1 * 3 + 1]; }
        else if (whichb == 2) { a = cp.prior_l2_a + nF[1 * 3 + 2]; b = cp.prior_l2_b + nF[2 * 3 + 2]; }
        double x = 0.0;
        if (lane < 6 && !(whichb == 2 && cp.prior_l2_a == 0))
          x = sched_gamma_ol(cp.seed, key, (uint32_t)whichb, (lane & 1) ? 4096u : 0u, (uint32_t)m, PA_ST_TRANS, (lane & 1) ? b : a, 1.0);
        const double y = __shfl_down_sync(kFull, x, 1);
        if (lane == 0) S.scal[SC_G] = x / (x + y);
        if (lane == 2) S.scal[SC_L] = x / (x + y);
        if (lane == 4) S.scal[SC_L2] = (cp.prior_l2_a == 0) ? 0.0 : x / (x + y);
        __syncwarp();
        if (lane == 0) {
          const double g = S.scal[SC_G], l = S.scal[SC_L], l2 = S.scal[SC_L2];
          double* Tf = S.T;
          double* Tx = S.T + 9;
          const double y1 = 1 - l;
          Tf[0] = 1 - g; Tf[3] = g; Tf[4] = y1; Tf[7] = 1 - y1; Tf[5] = l2; Tf[8] = 1 - l2;
          Tx[0] = 1 - g; Tx[3] = g;
        }
      }
      cta_sync<W>();
      PA_PHASE(9)
    }  // iterations
  }    // chains

  cta_sync<W>();
  if (L.counters) {
    if (tid < 5) atomicAdd(&L.counters[tid], S.work[tid]);
    if (tid < 10) atomicAdd(&L.counters[8 + tid], (unsigned long long)S.phase[tid]);
  }
#undef PA_PHASE
#undef w_npu
#undef w_root
#undef w_pt
#undef w_hist
#undef w_iter
}

struct WVariant {
  int nmax, warps;
  size_t smem;
  void (*kernel)(ChainLaunch);
};

template <int NMAX, int W, int RS, int MINB>
WVariant make_wvariant() {
  return WVariant{NMAX, W, sizeof(GSmem<NMAX, W, RS>) + 16, chain_warp_kernel<NMAX, W, RS, MINB>};
}

// compiled layouts: (max nodes, warps per chain, site rows per warp per history tile)
const WVariant* pick_wvariant(int n_max, int warps) {
  // (max nodes, warps per chain, site rows per warp per history tile, resident CTAs per SM the register budget allows)
  static const WVariant variants[] = {make_wvariant<86, 1, 2, 8>(), make_wvariant<86, 2, 2, 7>(), make_wvariant<86, 4, 2, 4>(), make_wvariant<86, 8, 1, 2>(),
                                      make_wvariant<160, 1, 2, 4>(), make_wvariant<160, 2, 2, 4>(), make_wvariant<160, 4, 2, 4>(), make_wvariant<160, 8, 1, 2>(),
                                      make_wvariant<512, 8, 1, 1>()};  // deep trees (Zoonomia-scale, 2 * 240 - 1 nodes): one CTA per SM
  for (const WVariant& v : variants)
    if (n_max <= v.nmax && warps == v.warps) return &v;
  return nullptr;
}

}  // namespace

size_t chain_warp_smem_bytes(int n_max, int warps) {
  const WVariant* v = pick_wvariant(n_max, warps);
  return v ? v->smem : (size_t)1 << 30;
}

int chain_warp_pst(int P_max) { return (P_max + 63) & ~63; }

size_t chain_warp_arena_bytes(int n_max, int P_max) {
  const size_t Pst = (size_t)chain_warp_pst(P_max), NIa = (size_t)n_max / 2 + 1;
  size_t b = (size_t)3 * n_max * 16 * 8 + 2 * NIa * Pst * (4 * 8 + 2);
  return (b + 255) & ~size_t(255);
}

// L.P_max must be chain_warp_pst(P_max of the selected chains)
int launch_chains_warp(const ChainLaunch& L, int grid, int warps, cudaStream_t stream) {
  const WVariant* v = pick_wvariant(L.n_max, warps);
  if (!v) {
    set_error("tree too large for the compiled small-CTA chain-kernel shared-memory layouts");
    return PA_ERR_LIMIT;
  }
  cudaError_t err = cudaFuncSetAttribute(v->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v->smem);
  if (err == cudaSuccess) err = cudaFuncSetAttribute(v->kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (err != cudaSuccess) {
    set_error(std::string("small-CTA chain kernel shared memory: ") + cudaGetErrorString(err));
    return PA_ERR_LIMIT;
  }
  v->kernel<<<grid, 32 * warps, v->smem, stream>>>(L);
  err = cudaGetLastError();
  if (err != cudaSuccess) {
    set_error(std::string("small-CTA chain kernel launch: ") + cudaGetErrorString(err));
    return PA_ERR_CUDA;
  }
  return PA_OK;
}

}  // namespace pa
