// pa_chain.cuh -- on-device Gibbs/MH chain kernel (declared here, defined in pa_chain.cu).
#ifndef PA_CHAIN_CUH
#define PA_CHAIN_CUH
#include "pa_internal.h"
namespace pa {
struct ChainLaunch {
  const uint8_t* blobs; const int64_t* blob_off;
  const int32_t* chain_list;  // [n_chains][2]
  int n_chains;
  const int32_t* sel; int n_sel;  // chains this launch works through (indices into chain_list); sel == nullptr: 0..n_sel-1
  SubstDev sd; ChainParams cp;
  int n_max, slots_max, P_max, T, ZW, S_rows;
  double* trace_d; uint32_t* trace_z;
  const double* init_rates;   // [n_chains][8]: n, c, g, l, l2 rates, flag (0: cold start), 2 unused; nullptr: all cold
  const uint32_t* init_z;     // [n_chains][ZW] packed Z of a warm start
  uint8_t* arena; size_t arena_stride;
  unsigned long long* counters; int* queue;
  int* done;  // [n_sel] set to 1 (system scope) when the chain of queue slot k has written its last trace row; may be nullptr
};
size_t chain_smem_bytes(int n_max, int s_rows, int bd);  // bd: threads per CTA of the layout (256 or 128)
int chain_ctas_per_sm(int n_max, int s_rows, int bd);
size_t chain_arena_bytes(int n_max, int P_max);
int launch_chains(const ChainLaunch& L, int grid, int BD, cudaStream_t stream);
}  // namespace pa
#endif
