// pa_chain.cu -- the on-device PhyloAcc-ST chain: one persistent CTA per (element, model) chain, T Gibbs/MH
// iterations with no host round-trip.  Mirrors BPP_C::initMCMC + BPP_C::Gibbs (bpp_c.cpp:131-519) of the reference:
//
//   MonitorChain  -> trace row (Z packed 2 bit/node, loglik, n, c, g, l, l2)                 bpp_c.cpp:741-834
//   Update_Tg(.,.,true) for every site: ancestral bases, top-down, from per-PATTERN partials  bpp_c.cpp:672-733
//   getEmission: parent->child base counts per branch (popc over bit planes)  x  log P_k    bpp_c2.cpp:355-390
//   Update_Z_subtree: 3-state upward messages + top-down sampling of Z                       bpp_c.cpp:1182-1293
//   getUpdateNode + Update_Tg(.,.,false): re-prune ancestors of changed Z                    bpp_c.cpp:620-671
//   sample_rate (n, then c): MH with the histories integrated out (log_f_Xz)                 bpp_c.cpp:860-1176
//   sample_transition: Beta draws for g, l, l2                                               bpp_c.cpp:522-613
//
// Data layout.  The element's tree, pruning program, weights and gap masks come from its blob in HBM; the packed 4-bit
// tip codes of the current 256-pattern tile are staged in shared memory.  The per-pattern partials of every internal
// node live in this CTA's arena in HBM (L2-resident while the chain runs) in TWO copies; a per-node selector says
// which copy is current.  A pruning pass keeps the partials it recomputes in a per-thread shared-memory stack (for
// its own parents) and streams them to the non-current copy; accepting a proposal flips the selectors of the
// recomputed nodes, rejecting leaves everything untouched.  The shared-memory layout is a compile-time struct
// (SmemT<NMAX, SLOTS>) so that every address is an immediate.  The sampled ancestral bases of a 256-site tile are kept as
// bit planes (four ballots per warp and node) next to a byte matrix for the parent look-ups; the tips' bases come as
// precomputed bit planes with the element (ElemHeader::off_tipplanes), so the pair counts are AND + POPC per 32-site word.
// Every finished chain is flagged in mapped pinned memory (ChainLaunch::done): the host summarises it while the kernel runs.
// Random numbers: Philox4x32-10 indexed by (chain, iteration, purpose, site/node), see oracle/oracle.h.
#include <cuda_runtime.h>

#include <cstdio>

#include "pa_chain.cuh"
#include "pa_device.cuh"

namespace pa {

namespace {

// Per-layout constants (SmemT): BD threads per CTA; a tile is BD patterns / sites; CODEROW staged bytes per tip row
// (BD + 8 patterns); XS bytes per row of the byte matrix of sampled bases (BD / 4 + 1 words: rows of different nodes fall into
// different banks).
#define PA_LAYOUT_CONSTANTS(SM) \
  constexpr int kBD = SM::BD, kNW = SM::BD / 32, kTile = SM::BD, kCodeRow = SM::CODEROW, kXS = SM::XS; \
  (void)kNW; (void)kTile; (void)kCodeRow; (void)kXS

enum { SC_N = 0, SC_C, SC_G, SC_L, SC_L2, SC_LL, SC_PROPN, SC_PROPC, SC_PROPOSAL, SC_MHPART, SC_R,
       // scalars of the two MH steps, computed ahead of time by the last warp while the others sample histories
       SC_PRE_R_N = 16, SC_PRE_PROP_N, SC_PRE_PRIOR_N, SC_PRE_MH_N, SC_PRE_R_C, SC_PRE_PROP_C, SC_PRE_PRIOR_C, SC_PRE_MH_C, SC_COUNT = 24 };
enum { IV_CHAIN = 0, IV_ANY_CHANGED, IV_HAS1, IV_HAS2, IV_VALID, IV_ACCEPT, IV_RATE_CHANGED, IV_ACC_N, IV_ACC_C, IV_ZFALLBACK, IV_NSU, IV_STAGED, IV_NACT, IV_KSLICE, IV_ANYAMB, IV_COUNT = 16 };

template <int NMAX, int SLOTS_, int BD_, bool STAGE_>
struct SmemT {
  static constexpr bool STAGE = STAGE_;    // tip codes of the current tile staged in shared memory (false: read from the blob; deep trees)
  static constexpr bool XS8 = STAGE_;      // sampled bases also kept as a byte matrix (false: the parent's base is read back from the bit planes)
  static constexpr int NI = NMAX / 2 + 1;  // internal nodes / tip rows
  static constexpr int SLOTS = SLOTS_;     // temp-stack slots kept in shared memory; results parked deeper are re-read from the arena
  // CODEROW: BD / 2 code bytes per staged tip row + 16 (rows start 16-byte aligned for the bulk copies and fall into different banks)
  static constexpr int BD = BD_, CODEROW = BD_ / 2 + 16, XS = BD_ + 4, NWARP = BD_ / 32;
  alignas(16) double Pcur[NMAX * 16];  // current P(t) of every branch; the root's slot holds pi in its first row (history sampling)
  union {  // buffers that are live in disjoint phases of an iteration share storage
    struct { alignas(16) double Pprop[NMAX * 16]; };  // rate steps: proposal P(t)
    struct {                                          // history sampling .. Z update
      double emis[NMAX * 3], back[NMAX * 3], eamb[NMAX * 3];
      uint32_t cnt[NMAX * 16];
    };
  };
  union {
    struct {  // pruning passes: per-thread temp stack [slot][base][thread] + scale counts
      alignas(16) double stack[SLOTS * 4 * BD];
      int16_t stack_sc[SLOTS * BD];
    };
    struct {  // history sampling: bit planes of the sampled bases [internal node][warp = 32 sites]{A, C, G, T} and, on the small
              // layouts, the same bases as bytes [internal node | one zero row][site] (4 = no draw)
      alignas(16) uint4 planes[NI * NWARP];
      alignas(4) uint8_t xs8[XS8 ? (NI + 1) * XS : 4];
    };
  };
  ElemView ev;  // pointers into the element's blob
  double red[34];
  double dist[NMAX], uz[NMAX];
  double scal[SC_COUNT];
  double T[18];   // linear Z transition matrices [free, fixZ==1][child state a][parent state j]
  double pi[4];
  alignas(16) uint4 ops[NI];  // pruning program, see init
  alignas(16) uint4 act[NI];  // program of the current pass: only the nodes to recompute, operands resolved (build_pass_program)
  int16_t actk[NI];           // their positions in ops
  // history descriptor of internal node q: {node | 0x8000 if not drawn (missing, padding) | row of the parent's sampled bases << 16 (NI = the
  // zero row, for the root), offset (doubles) of the node's current partials in the arena}
  alignas(8) uint2 hs[NI + 3];
  uint32_t cb[NMAX];          // pair-count descriptor of the branch above node i: child (tip row | internal q) | parent q << 16 | tip << 31
  int ivars[IV_COUNT];
  int nz[18];
  // work counters and SM cycles per phase (thread 0 only; in shared memory so that they do not occupy registers of every thread)
  unsigned long long work[6];  // node-pattern updates, root evaluations, P(t) builds, history draws, iterations, node-site updates
  long long phase[10], t_last;
  int16_t parent[NMAX], child[2 * NMAX], up_nodes[NMAX], up_off[NMAX + 1], down_nodes[NMAX], down_off[NMAX + 1];
  int16_t qof[NMAX];  // internal index of a node (-1 for tips)
  uint16_t tk[BD_];   // pair-count pass: thread -> branch | word slice << 8 (trees with at most 128 branches)
  alignas(4) uint8_t dirty[NMAX];
  uint8_t flags[NMAX], fix[NMAX], sel[NMAX], amb[NMAX];
  int8_t Z[NMAX], Znew[NMAX];
  unsigned long long mbar;  // mbarrier the bulk copies of the tip codes complete on
  alignas(16) uint8_t codes[STAGE_ ? NI * CODEROW : 16];
};

__device__ __forceinline__ double class_rate(int z, double n_rate, double c_rate) { return z == 2 ? n_rate : (z == 0 ? 1.0 : c_rate); }

// mark the ancestors of node i dirty (getUpdateNode, bpp_c.cpp:620-642 / 836-858).  Several threads walk up at once and stop at the
// first ancestor somebody has already marked; the test-and-set is ONE shared-memory atomic on the word that holds the node's flag
// byte, so no thread reads a flag another thread may be writing (round 1's plain read-then-write was the one hazard compute-sanitizer
// racecheck reported; a write-only walk without the early exit cost ~1 % of the iteration).  `dirty` is 4-byte aligned.
__device__ __forceinline__ void mark_ancestors(const int16_t* parent, uint8_t* dirty, int i) {
  int j = parent[i];
  while (j >= 0) {
    const uint32_t bit = 1u << (8 * (j & 3));
    const uint32_t old = atomicOr(reinterpret_cast<unsigned int*>(dirty + (j & ~3)), bit);
    if (old & (0xffu << (8 * (j & 3)))) break;
    j = parent[j];
  }
}

struct Arena {
  double* part;          // [2][n_max][P_max][4]
  int16_t* sc;           // [2][n_max][P_max]
  uint32_t part_stride;  // doubles between the two copies
  uint32_t sc_stride;
};

// ---- 1-D bulk copies global -> shared (the TMA engine: cp.async.bulk, SASS UBLKCP) completing on an mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
               "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// Stage the packed 4-bit tip codes of patterns [pb, pb + BD) (pb a multiple of BD) into shared memory: one 1-D bulk copy per tip row
// (cp.async.bulk: the rows of a tile are BD / 2 contiguous bytes of the blob, 16-byte aligned on both sides), all completing on one
// mbarrier; `parity` is the barrier's phase, flipped by every staging.  Patterns beyond the element's last one are never read.
template <class SM>
__device__ __forceinline__ void ensure_staged(SM& S, const ElemView& ev, int pb, int n_rows, int tid, uint32_t& parity) {
  PA_LAYOUT_CONSTANTS(SM);
  if (!SM::STAGE) return;
  if (S.ivars[IV_STAGED] == pb) return;
  __syncthreads();  // every reader of the previous tile is done
  const int half = ev.h->Ppad >> 1;
  const int byte0 = pb >> 1;
  const uint32_t row_bytes = (uint32_t)min(kBD / 2, half - byte0);
#ifdef PA_STAGE_LDG
  // A/B variant (profiles/r2_kernel_experiments.txt): round 1's staging, a cooperative LDG -> STS loop
  {
    (void)parity;
    const uint32_t* src = reinterpret_cast<const uint32_t*>(ev.codes);
    uint32_t* dst = reinterpret_cast<uint32_t*>(S.codes);
    const int wpr = (int)row_bytes / 4;
    for (int idx = tid; idx < n_rows * wpr; idx += kBD) {
      const int r = idx / wpr, w = idx - r * wpr;
      dst[r * (kCodeRow / 4) + w] = src[(r * half + byte0) / 4 + w];
    }
    if (tid == 0) S.ivars[IV_STAGED] = pb;
    __syncthreads();
    return;
  }
#endif
  if (tid == 0) {
    // reads of the old tile through the generic proxy are ordered before the async proxy overwrites it
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(&S.mbar, row_bytes * (uint32_t)n_rows);
    for (int r = 0; r < n_rows; r++) bulk_g2s(S.codes + r * kCodeRow, ev.codes + (size_t)r * half + byte0, row_bytes, &S.mbar);
    S.ivars[IV_STAGED] = pb;
  }
  mbar_wait(&S.mbar, parity);
  parity ^= 1u;
  __syncthreads();
}

// Program of one pruning pass, built by warp 0: the nodes to recompute (all, or the dirty ones) in the element's pruning
// order, with every operand resolved once instead of per thread and node.  Descriptor of a node:
//   x = offset (doubles) of its partial in the arena copy it is written to; the scale count sits at x >> 2
//   y = kind0 | kind1 << 3 | child0 << 6 | child1 << 16 | park << 26 (stack slot the result is parked in, 63 = none: the
//       parent is the next node of the pass and takes it from registers)
//   z, w = payload of child 0 / 1 by kind: 0 missing; 1 tip: byte offset of its row in the staged codes; 2 result of the
//       previous node (registers); 3 stack: offset (doubles) of its slot; 4 arena: offset (doubles) of its current partial
template <bool USE_DIRTY, class SM>
__device__ __forceinline__ int build_pass_program(SM& S, const Arena& A, int Ppad, int wflip, int n_su, int lane) {
  PA_LAYOUT_CONSTANTS(SM);
  int n_act = 0;
  for (int k0 = 0; k0 < n_su; k0 += 32) {
    const int k = k0 + lane;
    const bool inc = k < n_su && (!USE_DIRTY || S.dirty[S.ops[k].x & 0xffff]);
    const unsigned bal = __ballot_sync(0xffffffffu, inc);
    if (inc) S.actk[n_act + __popc(bal & ((1u << lane) - 1u))] = (int16_t)k;
    n_act += __popc(bal);
  }
  __syncwarp();
  for (int j = lane; j < n_act; j += 32) {
    const uint4 op = S.ops[S.actk[j]];
    const int s = op.x & 0xffff;
    const uint32_t q = op.x >> 16;
    const int prev_s = (j > 0) ? (int)(S.ops[S.actk[j - 1]].x & 0xffff) : -1;
    const int next_s = (j + 1 < n_act) ? (int)(S.ops[S.actk[j + 1]].x & 0xffff) : -2;
    uint4 d;
    d.x = (uint32_t)(S.sel[s] ^ wflip) * A.part_stride + q * (uint32_t)Ppad * 4u;
    // a result that waits for a sibling subtree is parked in its stack slot if that slot is one of the SM::SLOTS kept in
    // shared memory; deeper slots are not kept: the consumer re-reads the partial from the arena copy it was stored to
    uint32_t park = (j + 1 == n_act || next_s == (int)S.parent[s]) ? 63u : ((op.z >> 8) & 0xffu);
    if (park != 63u && park >= (uint32_t)SM::SLOTS) park = 63u;
    uint32_t kinds[2], pay[2];
#pragma unroll
    for (int cc = 0; cc < 2; cc++) {
      const uint32_t kind = (op.z >> (2 * cc)) & 3u;
      const uint32_t ch = cc ? (op.y >> 16) : (op.y & 0xffff);
      const uint32_t xi = cc ? (op.w >> 16) : (op.w & 0xffff);
      kinds[cc] = 0u;
      pay[cc] = 0u;
      if (kind == 1u) {
        kinds[cc] = 1u;
        pay[cc] = xi * (SM::STAGE ? (uint32_t)kCodeRow : (uint32_t)(Ppad >> 1));  // byte offset of the tip's row of codes
      } else if (kind == 2u) {
        const bool dch = !USE_DIRTY || S.dirty[ch];
        if (dch && (int)ch == prev_s) kinds[cc] = 2u;
        else if (dch && ((op.z >> (16 + 8 * cc)) & 0xffu) < (uint32_t)SM::SLOTS) {
          kinds[cc] = 3u;
          pay[cc] = ((op.z >> (16 + 8 * cc)) & 0xffu) * 4u * kBD;
        } else if (dch) {  // parked beyond the shared-memory slots: read back what this pass stored
          kinds[cc] = 4u;
          pay[cc] = (uint32_t)(S.sel[ch] ^ wflip) * A.part_stride + xi * (uint32_t)Ppad * 4u;
        } else {
          kinds[cc] = 4u;
          pay[cc] = (uint32_t)S.sel[ch] * A.part_stride + xi * (uint32_t)Ppad * 4u;
        }
      }
    }
    d.y = kinds[0] | (kinds[1] << 3) | ((op.y & 0xffffu) << 6) | ((op.y >> 16) << 16) | (park << 26);
    d.z = pay[0];
    d.w = pay[1];
    S.act[j] = d;
  }
  return n_act;
}

// One Felsenstein pass over every pattern of the element (linear space, fp64) through the program in S.act.
// Thread = pattern.  Every recomputed node is written to the arena copy its descriptor names; the result of a node is
// forwarded in registers when the next node of the pass is its parent, so the shared-memory stack only holds results
// that wait for a sibling subtree.  Returns sum_p w_p log L_p.
// message of one child to its parent: m_j = sum_i P_child(j -> i) L_child(i).  kind 0 missing (1), 1 tip (the code selects / adds
// columns of P; a gap gives 1), 2 the previous node of the pass (registers), 3 parked on the shared-memory stack, 4 arena
#define PA_CHILD_MSG(KIND, CH, PAY, M0, M1, M2, M3, SCC)                                                                      \
  {                                                                                                                            \
    const double* __restrict__ Pd = Pm + (CH) * 16;                                                                            \
    if ((KIND) >= 2u) {                                                                                                        \
      double a0, a1, a2, a3;                                                                                                   \
      if ((KIND) == 2u) { a0 = o0; a1 = o1; a2 = o2; a3 = o3; SCC = sc; }                                                      \
      else if ((KIND) == 3u) {                                                                                                 \
        const double* st = S.stack + (PAY) + tid;                                                                              \
        a0 = st[0]; a1 = st[kBD]; a2 = st[2 * kBD]; a3 = st[3 * kBD];                                                          \
        SCC = S.stack_sc[((PAY) >> 2) + tid];                                                                                  \
      } else {                                                                                                                 \
        double2 u, v;                                                                                                          \
        ldg256(gpart + (PAY), u, v);                                                                                           \
        a0 = u.x; a1 = u.y; a2 = v.x; a3 = v.y;                                                                                \
        SCC = gsc[(PAY) >> 2];                                                                                                 \
      }                                                                                                                        \
      const double2* __restrict__ Pc = reinterpret_cast<const double2*>(Pd);                                                   \
      const double2 r00 = Pc[0], r01 = Pc[1], r10 = Pc[2], r11 = Pc[3], r20 = Pc[4], r21 = Pc[5], r30 = Pc[6], r31 = Pc[7];     \
      M0 = ((r00.x * a0 + r00.y * a1) + r01.x * a2) + r01.y * a3;                                                              \
      M1 = ((r10.x * a0 + r10.y * a1) + r11.x * a2) + r11.y * a3;                                                              \
      M2 = ((r20.x * a0 + r20.y * a1) + r21.x * a2) + r21.y * a3;                                                              \
      M3 = ((r30.x * a0 + r30.y * a1) + r31.x * a2) + r31.y * a3;                                                              \
    } else {                                                                                                                   \
      SCC = 0;                                                                                                                 \
      const uint32_t code = ((KIND) == 1u) ? (((uint32_t)cp[(PAY)] >> sh) & 15u) : 15u;                                        \
      if ((code & (code - 1u)) == 0u) {                                                                                        \
        const int bb = __ffs(code) - 1;                                                                                        \
        M0 = Pd[bb]; M1 = Pd[4 + bb]; M2 = Pd[8 + bb]; M3 = Pd[12 + bb];                                                       \
      } else if (code == 15u) {                                                                                                \
        M0 = 1.0; M1 = 1.0; M2 = 1.0; M3 = 1.0;                                                                                \
      } else {                                                                                                                 \
        const double e0 = (code & 1u) ? 1.0 : 0.0, e1 = (code & 2u) ? 1.0 : 0.0, e2 = (code & 4u) ? 1.0 : 0.0, e3 = (code & 8u) ? 1.0 : 0.0; \
        M0 = ((e0 * Pd[0] + e1 * Pd[1]) + e2 * Pd[2]) + e3 * Pd[3];                                                            \
        M1 = ((e0 * Pd[4] + e1 * Pd[5]) + e2 * Pd[6]) + e3 * Pd[7];                                                            \
        M2 = ((e0 * Pd[8] + e1 * Pd[9]) + e2 * Pd[10]) + e3 * Pd[11];                                                          \
        M3 = ((e0 * Pd[12] + e1 * Pd[13]) + e2 * Pd[14]) + e3 * Pd[15];                                                        \
      }                                                                                                                        \
    }                                                                                                                          \
  }

// One Felsenstein pass over every pattern of the element (linear space, fp64) through the program in S.act.
// Thread = pattern.  Every recomputed node is written to the arena copy its descriptor names; the result of a node is
// forwarded in registers when the next node of the pass is its parent, so the shared-memory stack only holds results
// that wait for a sibling subtree.  Returns sum_p w_p log L_p.
template <class SM>
__device__ __forceinline__ double prune_pass(const ElemView& ev, SM& S, const double* __restrict__ Pm, const Arena& A, int n_act, int n_rows, int tid,
                                             uint32_t& stage_parity) {
  PA_LAYOUT_CONSTANTS(SM);
  const int P = ev.h->P, Ppad = ev.h->Ppad;
  double acc = 0.0;
  for (int p0 = 0; p0 < Ppad; p0 += kTile) {
    ensure_staged(S, ev, p0, n_rows, tid, stage_parity);
    const int p = p0 + tid;
    if (p < P) {
      const uint8_t* __restrict__ cp = SM::STAGE ? (S.codes + (tid >> 1)) : (ev.codes + (p >> 1));
      const int sh = (p & 1) * 4;
      double* __restrict__ gpart = A.part + (uint32_t)p * 4u;
      int16_t* __restrict__ gsc = A.sc + p;
      double o0 = 1.0, o1 = 1.0, o2 = 1.0, o3 = 1.0;
      int sc = 0;
      uint4 nxt = S.act[0];
      for (int k = 0; k < n_act; k++) {
        const uint4 op = nxt;
        if (k + 1 < n_act) nxt = S.act[k + 1];  // next descriptor is in flight while this node is computed
        double m0, m1, m2, m3, n0, n1, n2, n3;
        int sc0, sc1;
        PA_CHILD_MSG(op.y & 7u, (op.y >> 6) & 0x3ffu, op.z, m0, m1, m2, m3, sc0)
        PA_CHILD_MSG((op.y >> 3) & 7u, (op.y >> 16) & 0x3ffu, op.w, n0, n1, n2, n3, sc1)
        // child 0's factor is applied first, then child 1's (same order as the reference's accumulation)
        o0 = m0 * n0; o1 = m1 * n1; o2 = m2 * n2; o3 = m3 * n3;
        sc = sc0 + sc1;
        // rescale by 2^256 when every entry dropped below 2^-256 (compare high words: all values are >= 0)
        const int hm = max(max(__double2hiint(o0), __double2hiint(o1)), max(__double2hiint(o2), __double2hiint(o3)));
        if (hm < 0x2FF00000 && hm > 0) {
          o0 *= PA_SCALE_UP; o1 *= PA_SCALE_UP; o2 *= PA_SCALE_UP; o3 *= PA_SCALE_UP;
          sc += 1;
        }
        const uint32_t park = op.y >> 26;
        if (park != 63u) {  // the parent is not the next node of the pass: park the result on the stack
          double* st = S.stack + park * 4 * kBD + tid;
          st[0] = o0; st[kBD] = o1; st[2 * kBD] = o2; st[3 * kBD] = o3;
          S.stack_sc[park * kBD + tid] = (int16_t)sc;
        }
        stg256(gpart + op.x, o0, o1, o2, o3);
        gsc[op.x >> 2] = (int16_t)sc;
      }
      // the last node of the program is the root (or nothing was computed: missing root -> likelihood 1)
      const double lik = ((S.pi[0] * o0 + S.pi[1] * o1) + S.pi[2] * o2) + S.pi[3] * o3;
      acc += (double)ev.weights[p] * (log(lik) - sc * PA_LOG_SCALE);
    }
  }
  return block_sum(acc, S.red, tid, kBD);
}
#undef PA_CHILD_MSG

// build the program of a pass on warp 0 and publish its length (one barrier)
template <bool USE_DIRTY, class SM>
__device__ __forceinline__ int make_pass(SM& S, const Arena& A, int Ppad, int wflip, int n_su, int tid) {
  if (tid < 32) {
    const int n_act = build_pass_program<USE_DIRTY>(S, A, Ppad, wflip, n_su, tid);
    if (tid == 0) S.ivars[IV_NACT] = n_act;
  }
  __syncthreads();
  return S.ivars[IV_NACT];
}

// scalar helpers that run on a few lanes once per iteration: one out-of-line copy each keeps the iteration loop small
__device__ __noinline__ double sched_gamma_ol(uint32_t seed, uint32_t key, uint32_t idx, uint32_t sub0, uint32_t iter, uint32_t stream, double a, double b) {
  return sched_gamma(seed, key, idx, sub0, iter, stream, a, b);
}
__device__ __noinline__ double log_gamma_pdf_ol(double x, double a, double b) { return log_gamma_pdf(x, a, b); }

// log-space serial Z update, a literal restatement of Update_Z_subtree; used only when the scaled linear pass degenerates
template <class SM>
__device__ __noinline__ void z_update_logspace(SM& S, int n) {
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

// pair-count pass, rare path: the sites of the current tile at which tip `row` carries an ambiguity code contribute
// log sum_{b in code} P_k(x_parent -> b) to the branch's emissions (bpp_c2.cpp:372-378).  One thread per flagged branch walks the
// tile's words and sites in order, so the sum -- and with it the chain -- is reproducible bit for bit (round 1 added the terms with
// shared-memory atomics from several word slices: same value up to the order of the additions).
template <class SM>
__device__ __noinline__ void count_ambiguous(SM& S, const ElemView& ev, const double* __restrict__ logP, int n_max, int i, uint32_t row, uint32_t qp,
                                             int g0, int nw, size_t tcell) {
  const int Ppad = ev.h->Ppad;
  double e[3] = {0.0, 0.0, 0.0};
  for (int word = 0; word < nw; word++) {
    uint32_t am = __ldg(ev.tipamb + tcell + word);
    if (!am) continue;
    const uint4 pp = S.planes[qp * SM::NWARP + word];
    while (am) {
      const int bit = __ffs(am) - 1;
      am &= am - 1u;
      if (!(((pp.x | pp.y | pp.z | pp.w) >> bit) & 1u)) continue;  // no base drawn for the parent at this site
      const uint32_t xp = ((pp.y >> bit) & 1u) + 2u * ((pp.z >> bit) & 1u) + 3u * ((pp.w >> bit) & 1u);
      const int pat = ev.sitepat[g0 + 32 * word + bit];
      const uint32_t code4 = ((uint32_t)ev.codes[(size_t)row * (Ppad >> 1) + (pat >> 1)] >> ((pat & 1) * 4)) & 15u;
#pragma unroll 1
      for (int kk = 0; kk < 3; kk++) {
        const double* lp = logP + ((size_t)kk * n_max + i) * 16 + xp * 4;
        double sum = 0;
        for (int b = 0; b < 4; b++)
          if ((code4 >> b) & 1u) sum += exp(lp[b]);
        e[kk] += log(sum);
      }
    }
  }
  for (int kk = 0; kk < 3; kk++) S.eamb[i * 3 + kk] += e[kk];
}

template <int NMAX, int SLOTS, int BD, int MINB, bool STAGE>
__global__ void __launch_bounds__(BD, MINB) chain_kernel(ChainLaunch L) {
  static_assert(BD == 256, "the elements' tip planes are laid out for 256-site tiles");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  typedef SmemT<NMAX, SLOTS, BD, STAGE> SM;
  PA_LAYOUT_CONSTANTS(SM);
  SM& S = *reinterpret_cast<SM*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const ChainParams& cp = L.cp;
  const SubstDev& sd = L.sd;
  if (tid < 4) S.pi[tid] = sd.pi[tid];

  // this CTA's arena: log P tables [3][n_max][16] f64, partials [2][n_max][P_max][4] f64, scale counts [2][n_max][P_max] i16
  unsigned char* arena = L.arena + (size_t)blockIdx.x * L.arena_stride;
  double* logP = reinterpret_cast<double*>(arena);
  Arena A;
  A.part = logP + (size_t)3 * L.n_max * 16;
  A.part_stride = (uint32_t)L.n_max * L.P_max * 4;
  A.sc = reinterpret_cast<int16_t*>(A.part + 2 * (size_t)A.part_stride);
  A.sc_stride = (uint32_t)L.n_max * L.P_max;
  const int n_rows = L.S_rows;
  if (SM::STAGE && tid == 0) {
    mbar_init(&S.mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t stage_parity = 0;  // phase of the staging mbarrier

  if (tid < 6) S.work[tid] = 0ull;
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
#define w_nsu S.work[5]

  if (tid == 0) S.ivars[IV_CHAIN] = -1;
  for (;;) {
    __syncthreads();
    if (tid == 0) {
      // the previous chain of this CTA is complete: every thread's trace writes precede the barrier above; publish it to the host,
      // which downloads and summarises finished chains while the kernel is still running (pa_eval.cpp: fetch_range)
      const int prev = S.ivars[IV_CHAIN];
      if (prev >= 0 && L.done) {
        __threadfence_system();
        *reinterpret_cast<volatile int*>(L.done + prev) = 1;
      }
      S.ivars[IV_CHAIN] = atomicAdd(L.queue, 1);
      S.ivars[IV_STAGED] = -1;
    }
    __syncthreads();
    const int qslot = S.ivars[IV_CHAIN];
    if (qslot >= L.n_sel) break;
    const int chain = L.sel ? L.sel[qslot] : qslot;
    const int elem = L.chain_list[2 * chain], mdl = L.chain_list[2 * chain + 1];
    const int resZ = (mdl == PA_M0) ? 0 : (mdl == PA_M1 ? 2 : 1);
    // random streams are keyed by the element's batch-wide number (= elem unless this engine holds a shard of the batch)
    const uint32_t key = (uint32_t)reinterpret_cast<const ElemHeader*>(L.blobs + L.blob_off[elem])->elem_id * 4u + (uint32_t)mdl;
    // the element's array pointers live in shared memory: in registers they spill, and local memory misses the (small) L1 here
    if (tid == 0) S.ev = make_view(L.blobs + L.blob_off[elem]);
    __syncthreads();
    const ElemView& ev = S.ev;
    const ElemHeader* h = ev.h;
    const int n = h->n, ni = h->ni, GG = h->GG, Ppad = h->Ppad, root = n - 1;
    const int T = L.T;
    // warm start (initMCMC with iter > 0, bpp_c.cpp:148-177): Z and the five rates continue from the last state of the
    // chain's previous run; otherwise the INIT_* values
    const double* warm = (L.init_rates && L.init_rates[(size_t)chain * 8 + 5] != 0.0) ? L.init_rates + (size_t)chain * 8 : nullptr;
    const double init_n = warm ? warm[0] : cp.init_nrate, init_c = warm ? warm[1] : cp.init_crate, init_g = warm ? warm[2] : cp.init_grate,
                 init_l = warm ? warm[3] : cp.init_lrate, init_l2 = warm ? warm[4] : cp.init_lrate2;

    // ---------------- initMCMC (bpp_c.cpp:131-343)
    for (int i = tid; i < n; i += kBD) {
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
      S.amb[i] = 0;
      S.qof[i] = ev.intidx[i];
    }
    // pair-count descriptors: the branch above node i joins the sampled base of its parent with the node's own sampled
    // base (internal node) or its data (tip)
    for (int i = tid; i < n; i += kBD) {
      const uint8_t f = ev.flags[i];
      const int par = ev.parent[i];
      uint32_t d = 0xffffffffu;
      if (par >= 0 && !(f & PA_F_MISSING))
        d = ((f & PA_F_TIP) ? (0x80000000u | (uint32_t)ev.tiprow[i]) : (uint32_t)ev.intidx[i]) | ((uint32_t)ev.intidx[par] << 16);
      S.cb[i] = d;
    }
    for (int i = tid; i <= h->n_up; i += kBD) S.up_off[i] = ev.up_off[i];
    for (int i = tid; i <= h->n_down; i += kBD) S.down_off[i] = ev.down_off[i];
    // pruning program: one packed descriptor per computed internal node, in the element's pruning order:
    //   x = node | q << 16;  y = child0 | child1 << 16;  z = kind0 | kind1 << 2 | fwd bits 4..6 | slot << 8 | slot0 << 16 | slot1 << 24;
    //   w = x0 | x1 << 16 (tip row of a tip child, internal index of an internal child); kind 0 missing, 1 tip, 2 internal
    for (int k = tid; k < ni; k += kBD) {
      const int s = ev.su[k];
      uint4 op = make_uint4(0, 0, 0, 0);
      auto kind_of = [&](int c) { const uint8_t f = ev.flags[c]; return (f & PA_F_MISSING) ? 0u : ((f & PA_F_TIP) ? 1u : 2u); };
      if (s >= 0) {
        const int c0 = ev.child[2 * s], c1 = ev.child[2 * s + 1];
        auto x_of = [&](int c) { return (uint32_t)((ev.flags[c] & PA_F_TIP) ? ev.tiprow[c] : ev.intidx[c]) & 0xffffu; };
        const uint32_t k0 = kind_of(c0), k1 = kind_of(c1);
        op.x = (uint32_t)s | ((uint32_t)ev.intidx[s] << 16);
        op.y = (uint32_t)c0 | ((uint32_t)c1 << 16);
        op.z = k0 | (k1 << 2) | ((uint32_t)(uint8_t)ev.slot[s] << 8) | ((uint32_t)(uint8_t)(k0 == 2 ? ev.slot[c0] : 0) << 16) |
               ((uint32_t)(uint8_t)(k1 == 2 ? ev.slot[c1] : 0) << 24);
        // register forwarding: bit 4 / 5 = child 0 / 1 is the previous op, bit 6 = the parent is the next op
        if (k > 0 && k0 == 2 && ev.su[k - 1] == c0) op.z |= 16u;
        if (k > 0 && k1 == 2 && ev.su[k - 1] == c1) op.z |= 32u;
        if (k + 1 < ni && ev.su[k + 1] >= 0 && ev.su[k + 1] == ev.parent[s]) op.z |= 64u;
        op.w = x_of(c0) | (x_of(c1) << 16);
      }
      S.ops[k] = op;
      // history descriptor of internal node q = k (post-order); the root's "parent" is the zero row (base 0 selects the row of
      // pi kept in the root's slot of Pcur); every partial starts in arena copy 0
      const int sq = ev.intlist[k];
      const int par = ev.parent[sq];
      S.hs[k] = make_uint2((uint32_t)sq | ((ev.flags[sq] & PA_F_MISSING) ? 0x8000u : 0u) | ((par >= 0 ? (uint32_t)ev.intidx[par] : (uint32_t)SM::NI) << 16),
                           (uint32_t)k * (uint32_t)Ppad * 4u);
    }
    if (tid < 3) S.hs[ni + tid] = make_uint2(0x8000u, 0u);  // the sampler walks the nodes in groups of four
    {  // pair-count pass: word slices per branch and the thread -> (branch, slice) map (no divisions in the iteration loop)
      const int nb = n - 1, K = max(1, kBD / nb);
      S.tk[tid] = (uint16_t)((K > 1 && tid < nb * K) ? ((tid % nb) | ((tid / nb) << 8)) : 0xffff);
      if (tid == 0) S.ivars[IV_KSLICE] = K;
    }
    if (tid == 0) {
      int nsu = 0;
      while (nsu < ni && ev.su[nsu] >= 0) nsu++;
      S.ivars[IV_NSU] = nsu;
      S.ivars[IV_ANYAMB] = 0;
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
    __syncthreads();
    const int n_su = S.ivars[IV_NSU];
    // P(t) of every branch for its class, and the log P tables of the three classes (emission look-ups); task = (class, node, row)
    for (int task = tid; task < 12 * n; task += kBD) {
      const int k = task / (4 * n), rem = task - k * 4 * n, i = rem >> 2, j = rem & 3;
      const double rate = (k == 0) ? 1.0 : (k == 1 ? init_c : init_n);
      double r0, r1, r2, r3;
      build_P_row(sd, S.dist[i], rate, j, r0, r1, r2, r3);
      double* lp = logP + ((size_t)k * L.n_max + i) * 16 + j * 4;
      lp[0] = log(r0); lp[1] = log(r1); lp[2] = log(r2); lp[3] = log(r3);
      const int zi = S.Z[i];
      if (k == ((zi == 0) ? 0 : (zi == 2 ? 2 : 1))) {
        double* pc = S.Pcur + i * 16 + j * 4;
        pc[0] = r0; pc[1] = r1; pc[2] = r2; pc[3] = r3;
      }
    }
    __syncthreads();
    if (tid < 4) S.Pcur[root * 16 + tid] = sd.pi[tid];  // the root has no branch: its first row serves the root draw of the history sampler
    // initial upward pass over every pattern (Gibbs, bpp_c.cpp:355-356), partials stored in arena copy 0
    {
      const int n_act = make_pass<false>(S, A, Ppad, 0, n_su, tid);
      const double ll = prune_pass(ev, S, S.Pcur, A, n_act, n_rows, tid, stage_parity);
      if (tid == 0) {
        S.scal[SC_LL] = ll;
        w_npu += (unsigned long long)n_act * h->P;
        w_nsu += (unsigned long long)n_act * GG;
        w_root += h->P;
      }
    }
    __syncthreads();
    PA_PHASE(0)


    for (int m = 0; m < T; m++) {
      // ---------------- MonitorChain: trace row m
      for (int w = tid; w < L.ZW; w += kBD) {
        uint32_t word = 0;
        for (int k = 0; k < 16; k++) {
          const int i = w * 16 + k;
          if (i < n) word |= (uint32_t)(S.Z[i] & 3) << (2 * k);
        }
        L.trace_z[((size_t)chain * T + m) * L.ZW + w] = word;
      }
      if (tid < 6) {
        const int src = (tid == 0) ? SC_LL : (tid == 1 ? SC_N : (tid == 2 ? SC_C : (tid == 3 ? SC_G : (tid == 4 ? SC_L : SC_L2))));
        L.trace_d[((size_t)chain * 6 + tid) * T + m] = S.scal[src];
      }
      if (m == T - 1) break;
      if (tid == 0) w_iter++;

      // ---------------- sample ancestral histories site by site + count parent->child base pairs per branch
      for (int i = tid; i < n * 16; i += kBD) S.cnt[i] = 0;
      for (int i = tid; i < n * 3; i += kBD) S.eamb[i] = 0.0;
      __syncthreads();
      PA_PHASE(1)
      if (warp == kNW - 1) {
        // The random draws and the prior / Hastings terms of this iteration's two MH steps depend only on the rates and
        // proposal scales at the start of the iteration, not on Z: the last warp (idle below 225 sites) evaluates them
        // now -- same draw-schedule slots, same arithmetic as at the steps -- so that the steps only pick values.
        const double old_n = S.scal[SC_N], old_c = S.scal[SC_C], pn = S.scal[SC_PROPN], pc = S.scal[SC_PROPC];
        const double r_n = u01(philox4(cp.seed, key, 0u, 0u, (uint32_t)m, PA_ST_RATE).x);
        const double r_c = u01(philox4(cp.seed, key, 1u, 0u, (uint32_t)m, PA_ST_RATE).x);
        double v = 0.0;  // lane 0: n proposal; lane 1: n drawn from its prior; lane 2: c drawn from its prior (bpp_c.cpp:940-964,998)
        if (resZ != 0 && lane < 3)
          v = sched_gamma_ol(cp.seed, key, lane == 2 ? 1u : 0u, 1, (uint32_t)m, PA_ST_RATE, lane == 0 ? old_n / pn : (lane == 1 ? cp.nprior_a : cp.cprior_a),
                             lane == 0 ? pn : (lane == 1 ? cp.nprior_b : cp.cprior_b));
        const double prop_n = __shfl_sync(0xffffffffu, v, 0), prior_n = __shfl_sync(0xffffffffu, v, 1), prior_c = __shfl_sync(0xffffffffu, v, 2);
        const double prop_c = old_c * (r_c * (pc - 1 / pc) + 1 / pc);  // bpp_c.cpp:1010-1011
        double x = prop_n, a = cp.nprior_a, b = cp.nprior_b, sgn = 1.0;                    // lane 0: + log pdf(n proposal | prior)
        if (lane == 1) { x = old_n; sgn = -1.0; }                                         // lane 1: - log pdf(n current  | prior)
        if (lane == 2) { x = old_n; a = prop_n / pn; b = pn; }                             // lane 2: + log q(current | proposal)
        if (lane == 3) { x = prop_n; a = old_n / pn; b = pn; sgn = -1.0; }                // lane 3: - log q(proposal | current)
        if (lane == 4) { x = prop_c; a = cp.cprior_a; b = cp.cprior_b; }                   // lane 4: + log pdf(c proposal | prior)
        if (lane == 5) { x = old_c; a = cp.cprior_a; b = cp.cprior_b; sgn = -1.0; }       // lane 5: - log pdf(c current  | prior)
        double term = 0.0;
        if (lane >= 4 ? lane < 6 : resZ != 0) term = sgn * log_gamma_pdf_ol(x, a, b);
        if (lane == 6) term = log(old_c) - log(prop_c);
        const double t0 = __shfl_sync(0xffffffffu, term, 0), t1 = __shfl_sync(0xffffffffu, term, 1), t2 = __shfl_sync(0xffffffffu, term, 2),
                     t3 = __shfl_sync(0xffffffffu, term, 3), t4 = __shfl_sync(0xffffffffu, term, 4), t5 = __shfl_sync(0xffffffffu, term, 5),
                     t6 = __shfl_sync(0xffffffffu, term, 6);
        if (lane == 0) {
          S.scal[SC_PRE_R_N] = r_n; S.scal[SC_PRE_PROP_N] = prop_n; S.scal[SC_PRE_PRIOR_N] = prior_n; S.scal[SC_PRE_MH_N] = ((t0 + t1) + t2) + t3;
          S.scal[SC_PRE_R_C] = r_c; S.scal[SC_PRE_PROP_C] = prop_c; S.scal[SC_PRE_PRIOR_C] = prior_c; S.scal[SC_PRE_MH_C] = ((t4 + t5) + t6) + 0.0;
        }
      }
      for (int g0 = 0; g0 < GG; g0 += kTile) {
        const int g = g0 + tid;
        const bool valid = g < GG;
        const int pat = valid ? (int)ev.sitepat[g] : 0;
        const double* __restrict__ gp = A.part + (uint32_t)pat * 4u;
        uint8_t* xs = S.xs8 + tid;
        if (SM::XS8) xs[SM::NI * kXS] = 0;
        const int rootq = ni - 1, Gtop = (ni - 1) >> 2;
        // warps beyond the element's last site have nothing to draw (their words of the planes are never read)
        if (g0 + (tid & ~31) < GG) {
          uint32_t gw = 0;
          // the partials of a group's four nodes do not depend on the draws: they are requested while the previous group's results
          // are being stored and this group's random numbers are generated (one L2 / HBM round trip per 4 nodes)
          double2 pa[4], pc[4];
#pragma unroll
          for (int j = 3; j >= 0; j--) {
            ldg256(gp + S.hs[4 * Gtop + j].y, pa[j], pc[j]);
          }
          for (int G = Gtop; G >= 0; G--) {
            if (G == Gtop || (G & 7) == 7) {  // "all-gap" flags of internal nodes 32 * (G >> 3) .. + 31 at this site's pattern (bpp_c.cpp:686-690)
              const int wi = G >> 3;
              gw = ev.gap[(size_t)wi * Ppad + pat];
              if (wi == (rootq >> 5)) gw &= ~(1u << (rootq & 31));  // the root is always drawn (bpp_c.cpp:676-682)
              if (!valid) gw = 0xffffffffu;                         // lanes beyond the element draw nothing
            }
            // everything the group's four nodes read is fetched before anything is stored: their descriptors and the sampled bases of
            // parents outside the group; a parent inside the group is forwarded in registers
            uint2 d[4];
            uint32_t xpar[4];
#pragma unroll
            for (int j = 3; j >= 0; j--) d[j] = S.hs[4 * G + j];
#pragma unroll
            for (int j = 3; j >= 0; j--) {
              const uint32_t qp = d[j].x >> 16;
              if (SM::XS8) xpar[j] = xs[qp * kXS];
              else if (qp >= (uint32_t)SM::NI || (int)(qp >> 2) == G) xpar[j] = 0u;
              else {
                const uint4 pp = S.planes[qp * kNW + warp];
                xpar[j] = ((pp.y >> lane) & 1u) + 2u * ((pp.z >> lane) & 1u) + 3u * ((pp.w >> lane) & 1u);
              }
            }
            const uint4 rnd = philox4(cp.seed, key, (uint32_t)g, (uint32_t)G, (uint32_t)m, PA_ST_HIST);
            const uint32_t gbits = gw >> ((4 * G) & 31);
            int x[4];
#pragma unroll
            for (int j = 3; j >= 0; j--) {
              const uint32_t s = d[j].x & 0x7fffu, qp = d[j].x >> 16;
              uint32_t xp = xpar[j];
#pragma unroll
              for (int jj = 3; jj > j; jj--)
                if (qp == (uint32_t)(4 * G + jj)) xp = (uint32_t)x[jj];
              // branch-free, so that the four draws overlap: a node that is not drawn (missing / padding, "all-gap" at this site, lane
              // without a site) evaluates the same expressions on whatever it fetched and discards the result
              const double u = u01(j == 0 ? rnd.x : (j == 1 ? rnd.y : (j == 2 ? rnd.z : rnd.w)));
              const double2* Pr = reinterpret_cast<const double2*>(S.Pcur + s * 16 + (xp & 3u) * 4);
              const double2 r0 = Pr[0], r1 = Pr[1];
              const double p0 = r0.x * pa[j].x, p1 = r0.y * pa[j].y, p2 = r1.x * pc[j].x, p3 = r1.y * pc[j].y;
              // first k with u*tot < cumsum_k == number of cumulative sums not exceeding u*tot
              const double c1 = p0, c2 = c1 + p1, c3 = c2 + p2, xu = u * (c3 + p3);
              const int xd = (int)(xu >= c1) + (int)(xu >= c2) + (int)(xu >= c3);
              x[j] = ((d[j].x & 0x8000u) || ((gbits >> j) & 1u)) ? 4 : xd;  // 4 = no draw
            }
            if (G > 0) {
#pragma unroll
              for (int j = 3; j >= 0; j--) {
                ldg256(gp + S.hs[4 * (G - 1) + j].y, pa[j], pc[j]);
              }
            }
#pragma unroll
            for (int j = 3; j >= 0; j--) {
              const int q = 4 * G + j;
              if (d[j].x & 0x8000u) continue;  // uniform over the CTA
              if (SM::XS8) xs[q * kXS] = (uint8_t)x[j];
              const uint32_t b0 = __ballot_sync(0xffffffffu, x[j] == 0), b1 = __ballot_sync(0xffffffffu, x[j] == 1),
                             b2 = __ballot_sync(0xffffffffu, x[j] == 2), b3 = __ballot_sync(0xffffffffu, x[j] == 3);
              if (lane == 0) S.planes[q * kNW + warp] = make_uint4(b0, b1, b2, b3);
            }
            if (!SM::XS8) __syncwarp();  // the nodes' children read the planes back
          }
        }
        __syncthreads();
        // ---- parent->child base pairs of every branch over the tile's sites (getEmission's sufficient statistic,
        // bpp_c2.cpp:355-390): thread = (branch, slice of the tile's 32-site words); a bin is the population count of
        // (parent's plane of base x) & (child's plane of base y); the tips' planes come precomputed with the element
        {
          const int nw = (min(kTile, GG - g0) + 31) >> 5;
          const int nb = n - 1;
          const int K = S.ivars[IV_KSLICE];  // word slices per branch
          const size_t tile_row0 = (size_t)(g0 / kTile) * (size_t)h->S;
          for (int task = tid; task < nb * K; task += kBD) {  // more than kBD / 2 branches (deep trees): one slice, several branches per thread
            int k = 0, i = task;
            if (K > 1) {
              const uint32_t t = S.tk[tid];
              k = t >> 8;
              i = t & 0xffu;
            }
            const uint32_t d = S.cb[i];
            if (d == 0xffffffffu) continue;
            const uint32_t qp = (d >> 16) & 0x7fffu, cx = d & 0xffffu;
            const bool tip = (d >> 31) != 0;
            const uint4* __restrict__ prow = S.planes + qp * kNW;
            const uint4* __restrict__ crow = S.planes + cx * kNW;
            const size_t tcell = (tile_row0 + cx) * 8;
            for (int w0 = k; w0 < nw; w0 += 4 * K) {  // at most 4 words (128 sites) per flush: 8-bit fields
              uint32_t c0 = 0u, c1 = 0u, c2 = 0u, c3 = 0u;
              const int w1 = min(nw, w0 + 4 * K);
              for (int w = w0; w < w1; w += K) {
                const uint4 a = prow[w];
                uint4 c;
                if (tip) {
                  c = __ldg(ev.tipplanes + tcell + w);
                  if (__ldg(ev.tipamb + tcell + w)) {  // rare: handled by one thread per branch below (fixed summation order)
                    S.amb[i] = 1;
                    S.ivars[IV_ANYAMB] = 1;
                  }
                } else c = crow[w];
                c0 += (uint32_t)__popc(a.x & c.x) + ((uint32_t)__popc(a.x & c.y) << 8) + ((uint32_t)__popc(a.x & c.z) << 16) + ((uint32_t)__popc(a.x & c.w) << 24);
                c1 += (uint32_t)__popc(a.y & c.x) + ((uint32_t)__popc(a.y & c.y) << 8) + ((uint32_t)__popc(a.y & c.z) << 16) + ((uint32_t)__popc(a.y & c.w) << 24);
                c2 += (uint32_t)__popc(a.z & c.x) + ((uint32_t)__popc(a.z & c.y) << 8) + ((uint32_t)__popc(a.z & c.z) << 16) + ((uint32_t)__popc(a.z & c.w) << 24);
                c3 += (uint32_t)__popc(a.w & c.x) + ((uint32_t)__popc(a.w & c.y) << 8) + ((uint32_t)__popc(a.w & c.z) << 16) + ((uint32_t)__popc(a.w & c.w) << 24);
              }
#pragma unroll
              for (int b = 0; b < 16; b++) {
                const uint32_t cw4 = (b < 4) ? c0 : ((b < 8) ? c1 : ((b < 12) ? c2 : c3));
                const uint32_t c = (cw4 >> (8 * (b & 3))) & 0xffu;
                if (c) atomicAdd(&S.cnt[i * 16 + b], c);
              }
            }
          }
        }
        __syncthreads();
        if (S.ivars[IV_ANYAMB]) {  // ambiguity-coded tips in this tile (uniform over the CTA)
          const int nw = (min(kTile, GG - g0) + 31) >> 5;
          const size_t tile_row0 = (size_t)(g0 / kTile) * (size_t)h->S;
          for (int i = tid; i < n - 1; i += kBD) {
            if (!S.amb[i]) continue;
            S.amb[i] = 0;
            const uint32_t d = S.cb[i];
            count_ambiguous(S, ev, logP, L.n_max, i, d & 0xffffu, (d >> 16) & 0x7fffu, g0, nw, (tile_row0 + (d & 0xffffu)) * 8);
          }
          __syncthreads();
          if (tid == 0) S.ivars[IV_ANYAMB] = 0;
        }
      }
      if (tid == 0) w_hist += (unsigned long long)GG * ni;
      PA_PHASE(2)

      // ---------------- getEmission (bpp_c2.cpp:355-390) fused with the start of Update_Z_subtree (bpp_c.cpp:1186-1200):
      // e_s(k) = sum_bins count * log P_k, then b_s(k) = exp(e_s(k) - max) with state 2 masked on fixZ nodes, and the Z uniforms
      if (tid == 0) S.ivars[IV_ZFALLBACK] = 0;
      for (int task = tid; task < n * 3; task += kBD) {
        const int i = task / 3, k = task - 3 * i;
        double e = 0.0;
        if (i != root) {
          if (S.flags[i] & PA_F_MISSING) e = (k == 1) ? cp.log_constomis : 0.0;
          else {
            e = S.eamb[task];
            // the 16 log P entries come in one round trip (the table lives in the arena); a bin with no pairs adds nothing (its
            // log P may be -inf: identity P(t) of a zero-length branch)
            const double2* lp2 = reinterpret_cast<const double2*>(logP + ((size_t)k * L.n_max + i) * 16);
            double2 lp[8];
#pragma unroll
            for (int b = 0; b < 8; b += 2) ldg256(reinterpret_cast<const double*>(lp2 + b), lp[b], lp[b + 1]);
            const uint4* c4 = reinterpret_cast<const uint4*>(&S.cnt[i * 16]);
#pragma unroll
            for (int w = 0; w < 4; w++) {
              const uint4 cw = c4[w];
              if (cw.x) e += (double)cw.x * lp[2 * w].x;
              if (cw.y) e += (double)cw.y * lp[2 * w].y;
              if (cw.z) e += (double)cw.z * lp[2 * w + 1].x;
              if (cw.w) e += (double)cw.w * lp[2 * w + 1].y;
            }
          }
        }
        S.emis[task] = e;
      }
      __syncthreads();
      for (int i = tid; i < n; i += kBD) {
        double e0 = S.emis[i * 3], e1 = S.emis[i * 3 + 1], e2 = S.emis[i * 3 + 2];
        if (S.fix[i]) e2 = -CUDART_INF;
        const double mx = fmax(e0, fmax(e1, e2));
        S.back[i * 3] = exp(e0 - mx);
        S.back[i * 3 + 1] = exp(e1 - mx);
        S.back[i * 3 + 2] = exp(e2 - mx);
        S.uz[i] = u01(philox4(cp.seed, key, (uint32_t)i, 0u, (uint32_t)m, PA_ST_Z).x);
        if (!(mx > -CUDART_INF) || mx != mx) S.ivars[IV_ZFALLBACK] = 1;
      }
      __syncthreads();
      PA_PHASE(3)

      // ---------------- Update_Z_subtree (bpp_c.cpp:1204-1293): scaled linear messages by tree level, one warp
      if (warp == 0) {
        // The two transition matrices have six distinct non-trivial entries (bpp_c.cpp:301-332): free [1-g 0 0; g y l2; 0 1-y 1-l2]
        // (rows = child state, columns = parent state), fixZ == 1 [1-g 0 1; g 1 1; 0 0 1]; they stay in registers and the zero / one
        // entries are not multiplied (x + 0 * v and 1 * v are exact, so the messages are the ones of the full 3 x 3 products)
        const double t00 = S.T[0], t10 = S.T[3], t11 = S.T[4], t21 = S.T[7], t12 = S.T[5], t22 = S.T[8];
        for (int lv = 1; lv < h->n_up; lv++) {
          for (int idx = S.up_off[lv] + lane; idx < S.up_off[lv + 1]; idx += 32) {
            const int i = S.up_nodes[idx];
            double b0 = S.back[i * 3], b1 = S.back[i * 3 + 1], b2 = S.back[i * 3 + 2];
#pragma unroll
            for (int cc = 0; cc < 2; cc++) {
              const int ch = S.child[2 * i + cc];
              const bool fx = S.fix[ch] != 0;
              const double v0 = S.back[ch * 3], v1 = S.back[ch * 3 + 1], v2 = S.back[ch * 3 + 2];
              const double f0 = t00 * v0 + t10 * v1;
              const double f1 = fx ? v1 : __fma_rn(t21, v2, __dmul_rn(t11, v1));
              const double f2 = fx ? ((v0 + v1) + v2) : __fma_rn(t22, v2, __dmul_rn(t12, v1));
              b0 *= f0;
              b1 *= f1;
              b2 *= f2;
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
        if (S.ivars[IV_ZFALLBACK] != 0 && lane == 0) z_update_logspace(S, n);
        __syncwarp();
        if (lane == 0) {
          S.ivars[IV_ANY_CHANGED] = 0;
          S.ivars[IV_HAS1] = 0;
          S.ivars[IV_HAS2] = 0;
        }
      }
      __syncthreads();
      PA_PHASE(4)

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
            build_P(sd, S.dist[i], class_rate(zn, S.scal[SC_N], S.scal[SC_C]), true, P16);  // the rates have not moved since the start of the iteration
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
      if (S.ivars[IV_ANY_CHANGED]) {  // bpp_c.cpp:422-431 (recomputed partials overwrite the current copy in place)
        const int n_act = make_pass<true>(S, A, Ppad, 0, n_su, tid);
        const double ll = prune_pass(ev, S, S.Pcur, A, n_act, n_rows, tid, stage_parity);
        if (tid == 0) {
          S.scal[SC_LL] = ll;
          w_npu += (unsigned long long)n_act * h->P;
          w_nsu += (unsigned long long)n_act * GG;
          w_root += h->P;
        }
        __syncthreads();
      }
      for (int i = tid; i < n; i += kBD) S.dirty[i] = 0;
      __syncthreads();
      PA_PHASE(5)

      // ---------------- sample_rate: accelerated rate n (cls 2), then conserved rate c (cls 1)  (bpp_c.cpp:930-1176)
      // (under M0 no branch can be accelerated and n is not drawn from its prior, bpp_c.cpp:940-964: the n step is a no-op)
      for (int step = (resZ == 0) ? 1 : 0; step < 2; step++) {
        const bool neut = (step == 0);
        const int cls = neut ? 2 : 1;
        const uint32_t which = neut ? 0u : 1u;
        const int has = neut ? S.ivars[IV_HAS2] : S.ivars[IV_HAS1];
        const double prior_a = neut ? cp.nprior_a : cp.cprior_a, prior_b = neut ? cp.nprior_b : cp.cprior_b;
        const double old_rate = neut ? S.scal[SC_N] : S.scal[SC_C];
        if (has)
          for (int i = tid; i < n - 1; i += kBD)
            if (!(S.flags[i] & PA_F_MISSING) && S.Z[i] == cls) mark_ancestors(S.parent, S.dirty, i);
        __syncthreads();
        if (warp == 1) {  // the program of the proposal's pass is built next to warp 0's scalar work
          if (has) {
            const int n_act = build_pass_program<true>(S, A, Ppad, 1, n_su, lane);
            if (lane == 0) S.ivars[IV_NACT] = n_act;
          }
        } else if (warp == 0) {
          // the draws and the prior / Hastings terms were evaluated at the start of the iteration (history phase); pick them
          if (lane == 0) {
            int valid = 0;
            double cur_r = old_rate, proposal = 0.0, r = 0.0;
            if (!has) {  // no branch in this class: draw from the prior (bpp_c.cpp:940-964)
              if (resZ != 0) {
                cur_r = neut ? S.scal[SC_PRE_PRIOR_N] : S.scal[SC_PRE_PRIOR_C];
                if (neut) {
                  if (cp.ropt == 1 && cur_r < cp.nlb) cur_r = old_rate;
                  else if (cp.ropt == 2 && cur_r < S.scal[SC_C]) cur_r = old_rate;
                } else {
                  if (cp.ropt == 1 && cur_r > cp.cub) cur_r = old_rate;
                  else if (cp.ropt == 2 && cur_r > S.scal[SC_N]) cur_r = old_rate;
                }
              }
            } else {
              r = neut ? S.scal[SC_PRE_R_N] : S.scal[SC_PRE_R_C];
              valid = 1;
              proposal = neut ? S.scal[SC_PRE_PROP_N] : S.scal[SC_PRE_PROP_C];
              if (neut) {
                if (cp.ropt == 1 && proposal < cp.nlb) valid = 0;
                else if (cp.ropt == 2 && proposal < S.scal[SC_C]) valid = 0;
              } else {
                if (cp.ropt == 1 && proposal > cp.cub) valid = 0;
                else if (cp.ropt == 2 && proposal > S.scal[SC_N]) valid = 0;
              }
              if (proposal < 1e-12 || proposal > 1e8) valid = 0;
            }
            S.scal[SC_MHPART] = valid ? (neut ? S.scal[SC_PRE_MH_N] : S.scal[SC_PRE_MH_C]) : 0.0;
            S.scal[SC_PROPOSAL] = valid ? proposal : cur_r;
            S.scal[SC_R] = r;
            S.ivars[IV_VALID] = valid;
            S.ivars[IV_ACCEPT] = 0;
            S.ivars[IV_RATE_CHANGED] = (!valid && cur_r != old_rate) ? 1 : 0;
          }
        }
        __syncthreads();
        if (S.ivars[IV_VALID]) {
          const double proposal = S.scal[SC_PROPOSAL];
          for (int task = tid; task < 4 * (n - 1); task += kBD) {  // P(t * proposal) for the class' branches, current P elsewhere
            const int i = task >> 2, j = task & 3;
            double* pp = S.Pprop + i * 16 + j * 4;
            if (S.Z[i] == cls) {
              build_P_row_quad(sd, S.dist[i], proposal, j, pp[0], pp[1], pp[2], pp[3]);  // task & 3 == lane & 3
            } else {
              const double* pc = S.Pcur + i * 16 + j * 4;
              pp[0] = pc[0]; pp[1] = pc[1]; pp[2] = pc[2]; pp[3] = pc[3];
            }
          }
          __syncthreads();
        }
        {
          // ---------------- sample_transition (bpp_c.cpp:522-613).  g, l, l2 depend on Z only and nothing reads them before the next
          // iteration: the last warp (no patterns below 225) draws them while the others run the c-rate pruning pass
          if (!neut && warp == kNW - 1) {
            int c10 = 0, c20 = 0, c00 = 0, f21 = 0, f11 = 0, f12 = 0, f22 = 0;  // nZ(child, parent); f*: rows with fixZ != 1 only (bpp_c.cpp:541-552)
            for (int i = lane; i < n - 1; i += 32) {
              const int k = S.Z[i] * 3 + S.Z[S.parent[i]];
              const bool fr = !S.fix[i];
              c10 += (k == 3); c20 += (k == 6); c00 += (k == 0);
              f21 += (fr && k == 7); f11 += (fr && k == 4); f12 += (fr && k == 5); f22 += (fr && k == 8);
            }
            c10 = __reduce_add_sync(0xffffffffu, c10); c20 = __reduce_add_sync(0xffffffffu, c20); c00 = __reduce_add_sync(0xffffffffu, c00);
            f21 = __reduce_add_sync(0xffffffffu, f21); f11 = __reduce_add_sync(0xffffffffu, f11); f12 = __reduce_add_sync(0xffffffffu, f12);
            f22 = __reduce_add_sync(0xffffffffu, f22);
            const int whichb = lane >> 1;  // 0: g, 1: l, 2: l2; six gamma draws on six lanes
            double a = 1.0, b = 1.0;
            if (whichb == 0) { a = cp.prior_g_a + c10 + c20; b = cp.prior_g_b + c00; }
            else if (whichb == 1) { a = cp.prior_l_a + f21; b = cp.prior_l_b + f11; }
            else if (whichb == 2) { a = cp.prior_l2_a + f12; b = cp.prior_l2_b + f22; }
            double x = 0.0;
            if (lane < 6 && !(whichb == 2 && cp.prior_l2_a == 0))
              x = sched_gamma_ol(cp.seed, key, (uint32_t)whichb, (lane & 1) ? 4096u : 0u, (uint32_t)m, PA_ST_TRANS, (lane & 1) ? b : a, 1.0);
            const double y = __shfl_down_sync(0xffffffffu, x, 1);
            const double beta = x / (x + y);
            const double g = __shfl_sync(0xffffffffu, beta, 0), l = __shfl_sync(0xffffffffu, beta, 2);
            const double l2 = (cp.prior_l2_a == 0) ? 0.0 : __shfl_sync(0xffffffffu, beta, 4);
            if (lane == 0) {
              S.scal[SC_G] = g; S.scal[SC_L] = l; S.scal[SC_L2] = l2;
              double* Tf = S.T;
              double* Tx = S.T + 9;
              const double yy = 1 - l;
              Tf[0] = 1 - g; Tf[3] = g; Tf[4] = yy; Tf[7] = 1 - yy; Tf[5] = l2; Tf[8] = 1 - l2;
              Tx[0] = 1 - g; Tx[3] = g;
            }
          }
        }
        if (S.ivars[IV_VALID]) {
          const double proposal = S.scal[SC_PROPOSAL];
          const long long tp0 = clock64();
          const int n_act = S.ivars[IV_NACT];
          const double ll_new = prune_pass(ev, S, S.Pprop, A, n_act, n_rows, tid, stage_parity);  // log_f_Xz, into the non-current copy
          if (tid == 0 && !neut) S.phase[8] += clock64() - tp0;  // the c-rate pruning pass alone (also included in the c_rate phase)
          if (warp == 0) {
            const int cnt_nodes = n_act;
            int cnt_cls = 0;  // work counter: lane-parallel, not a serial loop on one thread
            for (int i = lane; i < n - 1; i += 32) cnt_cls += (S.Z[i] == cls);
            cnt_cls = __reduce_add_sync(0xffffffffu, cnt_cls);
            if (lane == 0) {
              const double MH = (ll_new - S.scal[SC_LL]) + S.scal[SC_MHPART];
              if (log(S.scal[SC_R]) < MH) {
                S.ivars[IV_ACCEPT] = 1;
                S.ivars[IV_RATE_CHANGED] = (proposal != old_rate) ? 1 : 0;
                S.scal[SC_LL] = ll_new;
              }
              w_npu += (unsigned long long)cnt_nodes * h->P;
              w_nsu += (unsigned long long)cnt_nodes * GG;
              w_root += h->P;
              w_pt += cnt_cls;
            }
          }
          __syncthreads();
        }
        const bool accepted = S.ivars[IV_ACCEPT] != 0;
        const bool rate_changed = S.ivars[IV_RATE_CHANGED] != 0;
        const double new_rate = (accepted || (!S.ivars[IV_VALID])) ? S.scal[SC_PROPOSAL] : old_rate;
        if (rate_changed) {  // rebuild the class' P(t) / log P for every branch (bpp_c.cpp:1115-1131); task = (node, row)
          for (int task = tid; task < 4 * (n - 1); task += kBD) {
            const int i = task >> 2, j = task & 3;
            double r0, r1, r2, r3;
            if (accepted && S.Z[i] == cls) {  // the accepted proposal's rows are still in Pprop (same arithmetic, same values)
              const double* pp = S.Pprop + i * 16 + j * 4;
              r0 = pp[0]; r1 = pp[1]; r2 = pp[2]; r3 = pp[3];
            } else build_P_row_quad(sd, S.dist[i], new_rate, j, r0, r1, r2, r3);
            double* lp = logP + ((size_t)cls * L.n_max + i) * 16 + j * 4;
            stg256(lp, log(r0), log(r1), log(r2), log(r3));
            if (S.Z[i] == cls) {
              double* pc = S.Pcur + i * 16 + j * 4;
              pc[0] = r0; pc[1] = r1; pc[2] = r2; pc[3] = r3;
            }
          }
        }
        for (int i = tid; i < n; i += kBD) {
          if (accepted && S.dirty[i]) {  // the proposal's partials become the current ones
            const uint32_t sl = S.sel[i] ^ 1u, q = (uint32_t)S.qof[i];
            S.sel[i] = (uint8_t)sl;
            S.hs[q].y = sl * A.part_stride + q * (uint32_t)Ppad * 4u;
          }
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
        __syncthreads();
        if (neut) { PA_PHASE(6) } else { PA_PHASE(7) }
      }

      // (sample_transition ran on the last warp next to the rate steps, see above)
      PA_PHASE(9)
    }  // iterations
  }    // chains

  __syncthreads();
  if (L.counters) {
    if (tid < 6) atomicAdd(&L.counters[tid], S.work[tid]);
    if (tid < 10) atomicAdd(&L.counters[8 + tid], (unsigned long long)S.phase[tid]);
  }
#undef PA_PHASE
#undef w_npu
#undef w_root
#undef w_pt
#undef w_hist
#undef w_iter
#undef w_nsu
}

struct Variant {
  int nmax, slots, bd, ctas_per_sm;
  size_t smem;
  void (*kernel)(ChainLaunch);
};

template <int NMAX, int SLOTS, int BD, int MINB, bool STAGE>
Variant make_variant() {
  return Variant{NMAX, SLOTS, BD, MINB, sizeof(SmemT<NMAX, SLOTS, BD, STAGE>) + 16, chain_kernel<NMAX, SLOTS, BD, MINB, STAGE>};
}

// compiled layouts: (max nodes of the tree incl. 2 * tip rows - 1, temp-stack slots in shared memory, threads per CTA,
// resident CTAs per SM the register budget is compiled for).  Any tree with at most `nmax` nodes runs on any of them:
// results parked deeper than the shared-memory slots are re-read from the arena.
const Variant* pick_variant(int n_max, int s_rows, int bd) {
  // (a 128-thread layout <96, 2, 128, 5> was measured: 5 CTAs per SM but two tiles for most elements -> 25 % slower)
  // the last layout serves Zoonomia-scale trees (240 taxa, 479 nodes): one CTA per SM, tip codes read from the blob
  static const Variant variants[] = {make_variant<32, 4, 256, 3, true>(), make_variant<96, 4, 256, 3, true>(), make_variant<160, 4, 256, 2, true>(),
                                     make_variant<240, 4, 256, 2, false>(), make_variant<480, 4, 256, 1, false>()};
  const int need = n_max > 2 * s_rows - 1 ? n_max : 2 * s_rows - 1;
  for (const Variant& v : variants)
    if (v.bd == bd && need <= v.nmax && s_rows <= v.nmax / 2 + 1) return &v;
  return nullptr;
}

}  // namespace

size_t chain_smem_bytes(int n_max, int s_rows, int bd) {
  const Variant* v = pick_variant(n_max, s_rows, bd);
  return v ? v->smem : (size_t)1 << 30;
}

size_t chain_arena_bytes(int n_max, int P_max) {
  size_t b = (size_t)3 * n_max * 16 * 8 + 2 * ((size_t)n_max * P_max * 4 * 8 + (size_t)n_max * P_max * 2);
  return (b + 255) & ~size_t(255);
}

int chain_ctas_per_sm(int n_max, int s_rows, int bd) {
  const Variant* v = pick_variant(n_max, s_rows, bd);
  return v ? v->ctas_per_sm : 1;
}

int launch_chains(const ChainLaunch& L, int grid, int BD, cudaStream_t stream) {
  const Variant* v = pick_variant(L.n_max, L.S_rows, BD);
  if (!v) {
    set_error("tree too large for the compiled chain-kernel shared-memory layouts");
    return PA_ERR_LIMIT;
  }
  cudaError_t err = cudaFuncSetAttribute(v->kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v->smem);
  if (err != cudaSuccess) {
    set_error(std::string("chain kernel shared memory: ") + cudaGetErrorString(err));
    return PA_ERR_LIMIT;
  }
  v->kernel<<<grid, BD, v->smem, stream>>>(L);
  err = cudaGetLastError();
  if (err != cudaSuccess) {
    set_error(std::string("chain kernel launch: ") + cudaGetErrorString(err));
    return PA_ERR_CUDA;
  }
  return PA_OK;
}

}  // namespace pa
