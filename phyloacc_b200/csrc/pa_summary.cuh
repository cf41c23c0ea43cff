// pa_summary.cuh -- per-chain summaries computed on the device (declared here, defined in pa_summary.cu).
#ifndef PA_SUMMARY_CUH
#define PA_SUMMARY_CUH
#include <cuda_runtime.h>

#include "pa_internal.h"
namespace pa {

#define SUMMARY_MAX_CONFIGS 1024 // Z configurations per chain the device clusters; chains with more are summarised on the host
#define SUMMARY_NEEDS_HOST 1     // SummaryHead::flags: Eval2 must run on the host for this chain (pruning fallback / table overflow)

// One record per chain: SummaryHead, then Max_Z and the last row's Z (ZW packed words each, local node order, 2 bits per node), then
// the post-burn-in counts of Z = 0 / 1 / 2 per local node (uint16 [n_max][3]).
struct SummaryHead {
  double max_loglik;  // MaxLoglik (bpp_c.cpp:808-814)
  double logml;       // Eval2's estimate (NaN when flags & SUMMARY_NEEDS_HOST)
  double med[5];      // medians of the full n, c, g, l, l2 traces (bpp_c2.cpp:742-760)
  double last[5];     // last row's n, c, g, l, l2: the warm start of the chain's next run (bpp_c.cpp:474-485)
  int32_t max_m;      // trace row of the arg-max (-1: none)
  int32_t flags;
  int32_t n_configs;  // distinct Z configurations met after burn-in
  int32_t pad_;
};

struct SummaryLaunch {
  const uint8_t* blobs; const int64_t* blob_off;
  const int32_t* chain_list;  // [n_chains][2] element, model
  int n_chains, T, ZW, num_burn, num_mcmc;
  int n_max_pad;  // rows of the per-node count table in shared memory (n_max rounded up to an even number)
  const double* trace_d; const uint32_t* trace_z;
  double cprior_a, cprior_b, nprior_a, nprior_b;
  double prior_l_a, prior_l_b, prior_l2_a, prior_l2_b, prior_g_a, prior_g_b;
  unsigned char* out; size_t rec_bytes;
};

size_t summary_record_bytes(int n_max, int ZW);
int launch_summaries(const SummaryLaunch& L, int n_max, cudaStream_t stream);
}  // namespace pa
#endif
