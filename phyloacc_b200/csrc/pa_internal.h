// pa_internal.h -- internal structures shared by the host front-end and the CUDA kernels.
//
// Vocabulary follows the reference (src/PhyloAcc-ST): element = one conserved non-coding
// element (a run of alignment columns), pattern = a unique alignment column of an element,
// chain = one (element, model) MCMC run, Z = latent conservation state per branch.
#ifndef PA_INTERNAL_H
#define PA_INTERNAL_H

#include <stdint.h>

#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/phyloacc_b200.h"

#define PA_NB 4  // num_base; the reference's 5-state indel model is dead code (bpp.hpp:247-253)

// node flag bits in ElemHeader::off_flags
#define PA_F_MISSING 1u
#define PA_F_TIP 2u
#define PA_F_FIXZ0 4u  // fixZ == 1 under M0 (bit = PA_F_FIXZ0 << model)

// Philox draw-schedule streams (shared with oracle/oracle.c so that chains are comparable draw for draw)
#define PA_ST_HIST 0u
#define PA_ST_Z 1u
#define PA_ST_RATE 2u
#define PA_ST_TRANS 3u

// Device blob of one element: header + arrays, byte offsets relative to the blob start (16-byte aligned).
// Local node index = position in the reference's post-order `nodes` list of the pruned tree
// (bpp_c.hpp:395), so children always precede parents and the root is n-1.
struct ElemHeader {
  int32_t n;          // nodes in the pruned tree
  int32_t ni;         // internal nodes
  int32_t S;          // species rows in `codes`
  int32_t P;          // unique site patterns
  int32_t Ppad;       // P rounded up to a multiple of 32
  int32_t GG;         // alignment columns after trimming
  int32_t W;          // 32-bit words of the per-pattern gap mask (= ceil(ni/32))
  int32_t n_slots;    // temp-partial slots needed by the pruning traversal
  int32_t n_up;       // number of height levels (Z upward pass)
  int32_t n_down;     // number of depth levels (Z downward pass)
  int32_t elem_id;    // element number in the batch (reference's CC)
  int32_t blob_bytes;
  int32_t off_parent;      // int16[n]   local parent, -1 for the root
  int32_t off_child;       // int16[n][2] local children (children2 order), -1 for tips
  int32_t off_flags;       // uint8[n]
  int32_t off_tiprow;      // int16[n]   species row of a tip, -1 for internal nodes
  int32_t off_orig;        // int16[n]   original node id
  int32_t off_dist;        // double[n]  distances2
  int32_t off_int;         // int16[ni]  internal nodes in post-order (reference's internal_nodes)
  int32_t off_intidx;      // int16[n]   index into off_int, -1 for tips
  int32_t off_su;          // int16[ni]  internal nodes in pruning-traversal order (min. live temporaries)
  int32_t off_slot;        // int8[n]    temp slot of an internal node
  int32_t off_up_nodes;    // int16[n]   nodes sorted by height (tips first)
  int32_t off_up_off;      // int16[n_up+1]
  int32_t off_down_nodes;  // int16[n]   nodes sorted by depth (root first)
  int32_t off_down_off;    // int16[n_down+1]
  int32_t off_codes;       // uint8[S][Ppad/2]  4-bit allowed-base masks, pattern p in byte p>>1, nibble p&1
  int32_t off_weights;     // uint16[Ppad]
  int32_t off_gap;         // uint32[W][Ppad]   bit q set <=> internal node q is "all-gap" at this pattern (bpp_c.hpp:417-428)
  int32_t off_sitepat;     // uint16[GGpad]     pattern of each site, sites sorted by (pattern, column)
  int32_t off_tipplanes;   // uint32[T][S][8][4] bit planes of the tips' bases per 256-site tile (T = ceil(GG/256)): bit t of word w of
                           //                    plane b <=> the species' code at site 256*tile + 32*w + t is exactly base b
  int32_t off_tipamb;      // uint32[T][S][8]    bit set <=> the code there is an ambiguity code (neither one base nor a gap/N)
};

// Substitution model on the device: P(t) = sum_m exp(eval[m] * t * r) * A[m]  with A[m][j][i] = V[j][m] * Vinv[m][i]
struct SubstDev {
  double eval[4];
  double A[4][16];
  double pi[4];
};

// Scalar chain parameters on the device
struct ChainParams {
  double cprior_a, cprior_b, nprior_a, nprior_b;
  double prior_l_a, prior_l_b, prior_l2_a, prior_l2_b, prior_g_a, prior_g_b;
  double init_crate, init_nrate, init_lrate, init_lrate2, init_grate;
  double cub, nlb, log_constomis;
  int32_t ropt, num_burn, num_mcmc, adaptive_freq;
  uint32_t seed;
  int32_t pad_;
};

// ---------------------------------------------------------------------------------------------- host side
namespace pa {

struct Params {  // LoadParams (main.cpp:100-287)
  std::string phytree_path, align_path, segment_path, id_path, output_path, result_prefix = "test";
  std::string outgroup, targetspecies, conservegroup;
  int batch = -1;
  int num_chain = 0;
  bool sample_hyper = false, verbose = false;
  pa_params num;  // numeric part
  Params() { pa_params_default(&num); }
};

struct Tree {  // PhyloTree + BPP tree arrays
  int S = 0, N = 0;
  std::vector<int32_t> parent, children;  // children [N][2]
  std::vector<double> dist;
  std::vector<std::string> names;
  double pi[4] = {0, 0, 0, 0};
  double Q[16] = {0};
};

struct Alignment {  // PhyloProf
  int S = 0;
  int64_t G = 0;
  std::vector<std::string> species_names;
  std::vector<std::string> X;  // lower-cased rows
  std::vector<std::string> element_names;
  std::vector<int64_t> elem_start;
  std::vector<int32_t> elem_len;
};

struct ElemPlan {  // integer part of BPP_C::BPP_C for one element + pattern compression
  bool filtered = true;
  int GG = 0, P = 0, root = -1;
  int n_slots = 0;                              // temp-partial slots the pruning traversal needs (Sethi-Ullman number)
  std::vector<uint8_t> missing;                 // [N]
  std::vector<int32_t> nodes;                   // post-order of the pruned tree (original ids)
  std::vector<int32_t> children2, parent2;      // [N][2], [N]
  std::vector<double> distances2;               // [N]
  std::vector<int32_t> upper_c, upper_conserve_c;
  std::vector<uint8_t> fixZ;                    // [3][N] (M0, M1, M2)
  std::vector<uint8_t> pattern_codes;           // [P][S]
  std::vector<int32_t> pattern_weights;         // [P]
  std::vector<int32_t> site_pattern;            // [GG]
  std::vector<uint8_t> blob;                    // packed device blob (ElemHeader + arrays)
};

struct Model {  // BPP
  Tree tree;
  pa_params prm;
  std::vector<int32_t> outgroup, conservedgroup;
  std::vector<uint8_t> upper, upper_conserve;  // [N] membership
  double eval[4], V[16], Vinv[16];             // Q = V diag(eval) Vinv
  SubstDev subst_dev;
};

// host-side building blocks (pa_host.cpp)
void set_error(const std::string& msg);
bool load_params(const std::string& path, Params& out);
bool load_tree(const std::string& path, Tree& out);
bool load_alignment_files(const std::string& fasta, const std::string& bed, Alignment& out);
bool load_ids(const Params& p, int C, std::vector<int>& ids);
bool match_alignment_to_tree(const Tree& t, Alignment& a);
std::vector<int32_t> species_ids(const Tree& t, const std::string& semicolon_list);
bool build_model(const Tree& t, const pa_params& prm, Model& out);
bool eigen_real4(const double Q[16], double eval[4], double V[16], double Vinv[16]);
int gbgc_transition(int device, const double* Q0, int n, const double* sel, const double* bias, const double* t, double* rate_out, double* P_out);  // pa_gbgc.cu
int gbgc_integrate(int device, cudaStream_t stream, const uint8_t* d_blobs, const int64_t* d_blob_off, int N, int n_max, int slots_max, int P_max, const double* pi,
                   const double* Q0, int n_states, const int32_t* elem, const int8_t* Z, const int8_t* B, const int8_t* fixZ, const int8_t* upper,
                   const double* sel_n, const double* sel_c, const double* gbc, double grate, double lrate, double prior_gb, double cons_to_mis,
                   double ncons_to_mis, double* post_out, double* emis_out, double* root_out);  // pa_gbgc.cu
struct GbgcWave {  // one asynchronous launch of the GBGC chain kernel (pa_gbgc_chain.cu)
  unsigned char *d_in = nullptr, *d_scr = nullptr, *d_out = nullptr;
  void *dev_tr = nullptr, *dev_head = nullptr, *dev_ctr = nullptr, *dev_zb = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;  // around the kernel
  int n_chains = 0;
  size_t b_tr = 0, b_zb = 0, b_head = 0;
  size_t cap_in = 0, cap_scr = 0, cap_out = 0;
};
int gbgc_chains_launch(int device, int num_sms, const uint8_t* d_blobs, const int64_t* d_blob_off, int N, int n_max, int slots_max, int P_max, const double* pi,
                       const double* Q0, int n_chains, const int32_t* chain_elem, const int32_t* chain_resz, const int8_t* fixZ, const int8_t* upper,
                       const pa_gbgc_params* prm, GbgcWave* w);
int gbgc_chains_wait(GbgcWave* w, double* tr_out, uint8_t* zb_out, double* head_out, unsigned long long* counters_out);
void gbgc_chains_release(GbgcWave* w);
void plan_element(const Model& m, const uint8_t* X, int64_t G, int64_t start, int32_t len, int elem_id, ElemPlan& out);

}  // namespace pa

#endif
