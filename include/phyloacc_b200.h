/* phyloacc_b200.h -- C ABI of the B200-native PhyloAcc-ST engine (libphyloacc_b200.so).
 *
 * The reference has no in-process plugin / FFI: its seam is the OpenMP element loop of
 * src/PhyloAcc-ST/main.cpp:452-500, which owns one BPP_C per element, reads the shared
 * immutable batch model `BPP` and writes per-element slots of it.  This header is the C
 * surface a maintainer would bind at that seam; every entry point
 * names the reference code it replaces (paths relative to src/PhyloAcc-ST/).
 *
 * Conventions: plain pointers and sizes, caller-allocated outputs, no exceptions across
 * the boundary; every function returns PA_OK (0) or a negative error code and leaves a
 * message retrievable with pa_last_error() (thread-local).  A handle is bound to one CUDA
 * device; calls on one handle must be serialised by the caller, different handles are
 * independent.  There is NO CPU fallback: entry points that compute fail with
 * PA_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef PHYLOACC_B200_H
#define PHYLOACC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PA_OK 0
#define PA_ERR_ARG (-1)      /* bad argument / malformed input (reference: exit(1) paths, main.cpp:112,129,162) */
#define PA_ERR_CUDA (-2)     /* CUDA runtime error or no usable device */
#define PA_ERR_STATE (-3)    /* call order violated (e.g. run before load) */
#define PA_ERR_IO (-4)       /* file could not be read / written */
#define PA_ERR_LIMIT (-5)    /* problem exceeds a compiled-in limit (tree too large for shared memory, ...) */

/* model indices follow the reference's resZ (main.cpp:466-494): the ORDER of slots in every
 * per-model array of this header is M0, M1, M2 = resZ 0, 2, 1. */
#define PA_M0 0  /* null:  resZ = 0 */
#define PA_M1 1  /* accelerated only in non-conserved, non-outgroup clades: resZ = 2 */
#define PA_M2 2  /* full:  resZ = 1 */

typedef struct pa_engine pa_engine;

/* Tree in the reference's numbering (newick.cpp:382-400, bpp.cpp:84-117): leaves 0..S-1,
 * internal nodes S..N-1 in post-order, root = N-1, N = 2S-1; children ordered by node id. */
typedef struct pa_tree {
  int32_t S, N;
  const int32_t* parent;    /* [N], -1 for the root */
  const int32_t* children;  /* [N][2], -1 for leaves */
  const double* distances;  /* [N] branch length above each node, already passed through float (newick.cpp:185) */
} pa_tree;

/* PhyloTree.pi / PhyloTree.subs_rate (newick.h); Q rows are "from", columns "to" */
typedef struct pa_subst {
  double pi[4];
  double Q[16];
} pa_subst;

/* Numeric engine parameters, defaults = main.cpp:49-90 (use pa_params_default). */
typedef struct pa_params {
  int32_t seed;                                   /* SEED */
  double init_conserve_rate, init_acce_rate;      /* INIT_CONSERVE_RATE, INIT_ACCE_RATE (ratio0, ratio1) */
  double conserve_prior_a, conserve_prior_b;      /* CONSERVE_PRIOR_A/B */
  double acce_prior_a, acce_prior_b;              /* ACCE_PRIOR_A/B */
  int32_t ropt;                                   /* ROPT */
  double cub, nlb;                                /* CUB, NLB */
  int32_t burnin, mcmc, adapt_freq;               /* BURNIN, MCMC, ADAPT_FREQ */
  double init_lrate, init_lrate2, init_grate;     /* INIT_LRATE, INIT_LRATE2, INIT_GRATE */
  double hyper_lrate_a, hyper_lrate_b;            /* HYPER_LRATE_A/B */
  double hyper_grate_a, hyper_grate_b;            /* HYPER_GRATE_A/B */
  double hyper_lrate2_a, hyper_lrate2_b;          /* HYPER_LRATE2_A/B (A == 0: Dollo, main.cpp:286) */
  double conserve_prop, gap_prop, constomis;      /* CONSERVE_PROP, GAP_PROP, CONSTOMIS */
  char gapchar;                                   /* GAPCHAR */
  int32_t prune_tree;                             /* PRUNE_TREE */
  double trim_gap_percent;                        /* TRIM_GAP_PERCENT */
  int32_t min_len;                                /* MIN_LEN */
  int32_t n_outgroup;  const int32_t* outgroup;   /* OUTGROUP species as leaf ids (bpp.hpp:303-309); id S = "name not found" */
  int32_t n_conserve;  const int32_t* conserve;   /* CONSERVE species as leaf ids (bpp.hpp:237-242) */
  int32_t num_thread;                             /* NUM_THREAD: host threads for the integer front-end and Eval2 */
} pa_params;

void pa_params_default(pa_params* p);

/* Per-element summaries = what the reference stores in bpp.log_liks_* / Max_Z (Gibbs bpp_c.cpp:470-471,
 * Eval2 bpp_c2.cpp:332-349) and prints in Output_init (bpp_c2.cpp:739-785).  All arrays are caller-allocated,
 * indexed [i] over the ids passed to pa_engine_run; model slot order M0, M1, M2. */
typedef struct pa_results {
  int32_t n;                /* capacity (>= number of ids) */
  int32_t N;                /* nodes per tree (row length of the Z arrays) */
  uint8_t* filtered;        /* [n] 1 if the element was filtered by the front-end (bpp_c.hpp:158,288,369) */
  double* loglik;           /* [n][3] marginal log-likelihood estimates (log_liks_null / resZ / sgl) */
  double* loglik_max;       /* [n][3] MaxLoglik (log_liks_Z) */
  int8_t* max_Z;            /* [n][3][N] Max_Z, -1 for pruned nodes */
  double* median_rates;     /* [n][3][5] posterior medians n, c, g, l, l2 over the FULL trace (bpp_c2.cpp:742-755) */
  double* post_Z;           /* [n][3][N][4] frequencies of Z = -1,0,1,2 over post-burn-in samples (bpp_c2.cpp:757-781) */
} pa_results;

/* Raw MCMC traces of one chain (Output_sampling, bpp_c2.cpp:706-736); T = burnin + mcmc rows. */
typedef struct pa_trace {
  int32_t T, N;
  double* loglik; double* n_rate; double* c_rate; double* g_rate; double* l_rate; double* l2_rate;  /* [T] each */
  int8_t* Z;                /* [T][N] */
} pa_trace;

/* Integer front-end of one element (BPP_C::BPP_C, bpp_c.hpp:143-476) for parity tests. Caller-allocated:
 * missing/children2/parent2/distances2/nodes sized N (children2 2N); pattern_codes holds cap_P * S bytes, pattern_weights cap_P
 * entries, site_pattern cap_GG entries.  The scalar fields are always filled (call once with null arrays to size the buffers); when
 * an array is non-null and its capacity is smaller than the element needs, nothing is copied and PA_ERR_ARG is returned. */
typedef struct pa_element_info {
  int32_t cap_P, cap_GG;    /* in: capacities of pattern_codes / pattern_weights (patterns) and site_pattern (columns) */
  int32_t filtered, GG, P, root, n_nodes;
  uint8_t* missing; int32_t* nodes; int32_t* children2; int32_t* parent2; double* distances2;
  uint8_t* pattern_codes;   /* [P][S] nibble masks of the unique columns, sorted lexicographically */
  int32_t* pattern_weights; /* [P] */
  int32_t* site_pattern;    /* [GG] pattern index of every kept column, original column order */
  uint8_t* fixZ;            /* [3][N]: 1 where fixZ == 1 for M0, M1, M2 (initMCMC bpp_c.cpp:155-161,240-289) */
  int32_t n_slots;          /* out: temporaries the pruning traversal of this element's tree needs (no reference counterpart) */
} pa_element_info;

const char* pa_last_error(void);
const char* pa_version(void);

/* BPP::BPP + BPP::InitMCMC (bpp.hpp:173-404, bpp.cpp:186-224): tree arrays, eigensystem of Q, upper sets.
 * device < 0: do not touch CUDA (front-end only; compute entry points will fail with PA_ERR_CUDA). */
int pa_engine_create(const pa_tree* tree, const pa_subst* subst, const pa_params* params, int device, pa_engine** out);
void pa_engine_destroy(pa_engine* e);

/* Launch stream for all device work of this handle (a cudaStream_t; NULL = default stream). */
int pa_engine_set_stream(pa_engine* e, void* cuda_stream);

/* PhyloProf.X + element_start/size (profile.cpp:17-126, bpp.hpp:257-274) followed by the integer part of
 * BPP_C::BPP_C for every element: tip encoding, column trimming, missing species, filters, pruned tree
 * (bpp_c.hpp:143-476), then site-pattern compression (new stage) and upload of the packed batch to HBM.
 * X is [S][G] row-major lower-case characters in TREE leaf order (rows already matched, bpp.cpp:133-184). */
int pa_engine_load_alignment(pa_engine* e, const uint8_t* X, int64_t G, int32_t C, const int64_t* elem_start, const int32_t* elem_len);

int pa_engine_num_elements(const pa_engine* e);
int pa_engine_element_info(const pa_engine* e, int32_t elem, pa_element_info* out);
/* The packed device image of one element (inspection / tests): header, tree arrays, pruning program, 4-bit tip codes per pattern,
 * weights, gap masks, site -> pattern map and the tips' bit planes per 256-site tile (layout: ElemHeader in csrc/pa_internal.h).
 * `*size` receives its length; the bytes are copied when `cap` is large enough.  A filtered element has no image (size 0). */
int pa_engine_element_blob(pa_engine* e, int32_t elem, uint8_t* buf, int64_t cap, int64_t* size);

/* Replayed log-likelihoods: for each state i, log P(X_elem | Z, n_rate, c_rate) exactly as the reference
 * evaluates it after initMCMC (bpp_c.cpp:131-236) + Update_Tg(.,.,false) (bpp_c.cpp:647-671) +
 * sum_g log_exp_sum(lambda[g][root]).  Z is [n_states][N] in original node numbering (-1 for pruned nodes). */
int pa_engine_loglik_replay(pa_engine* e, int32_t n_states, const int32_t* elem, const int8_t* Z, const double* n_rate,
                            const double* c_rate, double* loglik_out);

/* The element loop of main.cpp:452-500: for every id, M0 -> M1 -> M2 chains (initMCMC, Gibbs, Eval2,
 * Output_init).  pa_engine_sample launches the on-device chains and, right behind them on the handle's stream, the
 * summary kernel (MonitorChain's arg-max, the medians and Z frequencies of Output_init, Eval2) -- traces stay in HBM;
 * pa_engine_collect synchronises, copies one small record per chain and fills `out`; only chains whose Eval2 takes the
 * reference's pruning fallback (bpp_c2.cpp:199-316) have their trace downloaded.  pa_engine_run = sample + collect. */
int pa_engine_sample(pa_engine* e, const int32_t* ids, int32_t n_ids);
int pa_engine_collect(pa_engine* e, pa_results* out);

/* Outer CHAIN iterations and SAMPLE_HYPER (main.cpp:445-516).  pa_engine_sample_ex runs the models in `model_mask`
 * (bit PA_M0..PA_M2; the reference runs only the full model M2 under SAMPLE_HYPER 1, main.cpp:463-494); with
 * `warm_start` != 0 every chain continues from the last state of its previous run on this handle -- Z, n, c, g, l,
 * l2 -- as initMCMC does for iter > 0 (bpp_c.cpp:148-177, state saved by Gibbs, bpp_c.cpp:474-485); `chain_iter`
 * selects the random streams of this outer iteration.  The state of a run is captured by pa_engine_collect.
 * pa_engine_sample_hyper is BPP::sample_hyperparam (bpp.cpp:324-477): one independence-MH sweep over the five
 * hyper-parameter pairs given the current rates of the full-model chains of `ids`; the accepted values take effect
 * for the next pa_engine_sample(_ex) and are returned in `out`.
 * With the elements sharded over ranks (one engine per GPU) the sweep needs one exchange: pa_engine_hyper_sums returns this rank's
 * PA_HYPER_NSUMS partial sums over its own `ids` (sum n, sum log n, sum c, sum log c, sum log l, sum log(1-l), sum log l2,
 * sum log(1-l2), sum log g, sum log(1-g), count -- the sufficient statistics of bpp.cpp:340-349,403-452), the caller adds them over
 * ranks (one all-reduce of 11 doubles) and every rank calls pa_engine_sample_hyper_sums with the totals: same draws, same result. */
#define PA_HYPER_NSUMS 11
typedef struct pa_hyper {
  double nprior_a, nprior_b, cprior_a, cprior_b;          /* ACCE_PRIOR_A/B, CONSERVE_PRIOR_A/B */
  double prior_l_a, prior_l_b, prior_g_a, prior_g_b;      /* HYPER_LRATE_A/B, HYPER_GRATE_A/B */
  double prior_l2_a, prior_l2_b;                          /* HYPER_LRATE2_A/B */
} pa_hyper;
int pa_engine_sample_ex(pa_engine* e, const int32_t* ids, int32_t n_ids, int32_t model_mask, int32_t warm_start, int32_t chain_iter);
int pa_engine_get_hyper(const pa_engine* e, pa_hyper* out);
int pa_engine_set_hyper(pa_engine* e, const pa_hyper* in);
int pa_engine_sample_hyper(pa_engine* e, int32_t iter, const int32_t* ids, int32_t n_ids, pa_hyper* out);
int pa_engine_hyper_sums(pa_engine* e, const int32_t* ids, int32_t n_ids, double* sums /* [PA_HYPER_NSUMS] */);
int pa_engine_sample_hyper_sums(pa_engine* e, int32_t iter, const double* sums /* [PA_HYPER_NSUMS] */, pa_hyper* out);
int pa_engine_run(pa_engine* e, const int32_t* ids, int32_t n_ids, pa_results* out);

/* Trace of one chain of the last pa_engine_sample (i = position in ids, model = PA_M0..PA_M2). */
int pa_engine_get_trace(pa_engine* e, int32_t i, int32_t model, pa_trace* out);

/* Work counters of the last pa_engine_sample, for the roofline: node-pattern updates,
 * root evaluations, P(t) builds, history draws, kernel launches. */
typedef struct pa_counters {
  uint64_t node_pattern_updates, root_evals, pt_builds, history_draws, chain_iterations, kernel_launches;
  uint64_t node_site_updates; /* like node_pattern_updates, counted on UNCOMPRESSED columns (BASELINE metric "pruned site-liks/sec") */
  uint64_t d2h_bytes;         /* bytes copied device -> host for the last run so far: summary records, counters, and any traces that were
                               * asked for (pa_engine_get_trace, VERBOSE) or needed by Eval2's host fallback */
  uint64_t host_eval_chains;  /* chains of the last run whose Eval2 ran on the host (pruning fallback, configuration-table overflow) */
} pa_counters;
int pa_engine_counters(pa_engine* e, pa_counters* out);
/* SM cycles per phase of the chain kernel, summed over CTAs (thread 0 clock64 deltas): init, trace, history sampling,
 * emission, Z update, Z-dirty pruning, n-rate step, c-rate step, commit, transitions. */
int pa_engine_phase_cycles(pa_engine* e, uint64_t* out10);

/* Measurement helper for the roofline: sustained FP64 FMA throughput (TFLOP/s, 2 flop per FMA) of `device`, from a
 * register-resident DFMA micro-benchmark run for about `millis` ms. Not part of the reference's interface. */
int pa_device_fp64_peak(int device, int millis, double* tflops_out);

/* PhyloAcc-ST-GBGC building block (the chain itself is pa_engine_gbgc_run below): the per-branch GC-biased matrices of BPP_C::getRate and
 * BPP_C::getTransition (src/PhyloAcc-ST-GBGC/SRC/bpp_c.cpp:13-30,1405-1427) for n (selection, bias, branch length) triples on `device`:
 * rate_out (may be NULL) and P_out are [n][16] row-major, rate i -> j and P(i -> j | t).  The reference returns log P transposed. */
int pa_gbgc_transition(int device, const pa_subst* subst, int32_t n, const double* sel, const double* bias, const double* t, double* rate_out,
                       double* P_out);

/* PhyloAcc-ST-GBGC building block: replayed log-likelihoods under the GC-biased model.  For each state i: element elem[i], latent class Z
 * ([n_states][N], original node numbering) and gBGC indicator B ([n_states][N], 0 / 1) per branch, selection coefficients of the
 * accelerated / conserved classes, the element's gBGC bias.  Every branch gets exp(t * Q o f(s + b B)) with s = 0 / sel_c / sel_n by class
 * and b = gbc on the branches with B = 1 -- the six rate matrices of BPP_C::initMCMC and init_cache_TM (GBGC/SRC/bpp_c.cpp:196-204,
 * 1563-1621) -- and the upward pass of Update_Tg(., ., ., false) (:477-528); result = sum over columns of log sum_b pi_b L_root(b). */
int pa_engine_gbgc_replay(pa_engine* e, int32_t n_states, const int32_t* elem, const int8_t* Z, const int8_t* B, const double* sel_n,
                          const double* sel_c, const double* gbc, double* loglik_out);

/* PhyloAcc-ST-GBGC building block: the random-number-free half of one GBGC iteration at given states -- expected (child base, parent base)
 * pairs per branch (Update_Tg(g,.,.,true), GBGC/SRC/bpp_c.cpp:530-583), the six (Z, B) emissions of every branch, the 6-state upward messages
 * and the ELBO term (initial_prob_back, GBGC/SRC/bpp_c2.cpp:354-407).  fixZ is the model's constraint per node (1: class 2 forbidden), as
 * BPP_C::initMCMC sets it (GBGC/SRC/bpp_c.cpp:145-180).  post_out [n_states][N][16] (row: child base, column: parent base), emis_out
 * [n_states][N][6] (class k = Z + 3 B), root_out [n_states][8]: log_prob_back[root][0..5], the ELBO term, the upward log-likelihood. */
typedef struct pa_gbgc_consts {
  double grate, lrate;              /* PREP_GRATE, PREP_LRATE: the 3 x 3 transition matrix of Z (GBGC/SRC/bpp.hpp:286-293) */
  double prior_gb;                  /* nGBprior_c: prior probability that gBGC acts on a branch */
  double cons_to_mis, ncons_to_mis; /* CONSTOMIS and its non-conserved counterpart */
} pa_gbgc_consts;
int pa_engine_gbgc_integrate(pa_engine* e, int32_t n_states, const int32_t* elem, const int8_t* Z, const int8_t* B, const int8_t* fixZ,
                             const double* sel_n, const double* sel_c, const double* gbc, const pa_gbgc_consts* consts, double* post_out,
                             double* emis_out, double* root_out);

/* PhyloAcc-ST-GBGC: the chain itself.  BPP_C::Gibbs of src/PhyloAcc-ST-GBGC/SRC/bpp_c.cpp:259-345 for the listed elements
 * under M0 / M1 / M2 (the reference's resZ 0 / 2 / 1; main.cpp:395-412), followed by the variant's own Eval2 (variational bound,
 * bpp_c2.cpp:29-285) and Output_init summaries (bpp_c2.cpp:603-650).  Scalars = LoadParams' GBGC keys and constructor constants
 * (GBGC/SRC/main.cpp:50-86,108-193; bpp.hpp:201; bpp.cpp:217-219; bpp_c.hpp:358-361). */
typedef struct pa_gbgc_params {
  double sel_c0, sel_n0, gbc0;                    /* CONSERVE_RATE (-2), ACCE_RATE (0), initial gBGC bias (3) */
  double cprior_a, cprior_b;                      /* CONSERVE_PRIOR_A / B: Gamma prior of -sel_c */
  double nprior_a, nprior_b, nprior_c;            /* ACCE_PRIOR_A / B and the shift (2): Gamma prior of sel_n + shift */
  double gbprior_a, gbprior_b, gbprior_c;         /* Gamma prior of the bias; prior probability that gBGC acts on a branch */
  double cons_to_mis, ncons_to_mis;               /* CONSTOMIS and its non-conserved counterpart (0.01, 0.1) */
  double grate, lrate;                            /* PREP_GRATE, PREP_LRATE (0.8, 0.1) */
  double prop_c, prop_n, prop_gb;                 /* proposal widths (1.2, 1.2, 0.2; not adapted inside Gibbs: update_rates(.., false)) */
  int32_t num_burn, num_mcmc;                     /* BURNIN, MCMC */
  uint32_t seed;
  int32_t pad_;
} pa_gbgc_params;
void pa_gbgc_params_default(pa_gbgc_params* p);

/* Caller-allocated outputs; model index 0 / 1 / 2 = M0 / M1 / M2.  Optional arrays may be NULL. */
typedef struct pa_gbgc_results {
  int32_t n_ids, N, T;
  uint8_t* filtered;   /* [n_ids] */
  double* loglik;      /* [n_ids][3]  Eval2's estimate of log P(X | model): loglik_NUll, loglik_RES, loglik_all */
  double* loglik_max;  /* [n_ids][3]  max over post-burn-in iterations of log P(X, Z, B, r) */
  double* med;         /* [n_ids][3][3]  n_rate, c_rate, gBC as Output_init takes them (element T/2 of the trace with its tail sorted) */
  double* post_b;      /* [n_ids][3][N]  P(B = 1) per node */
  double* post_z;      /* [n_ids][3][N][4]  P(Z = -1, 0, 1, 2) per node */
  int8_t* max_z;       /* [n_ids][3][N]  Z at the arg-max, -1 for pruned nodes */
  int8_t* max_b;       /* [n_ids][3][N] */
  int32_t* accept;     /* [n_ids][3][2]  iterations with an accepted rate move / accepted (Z, B) update */
  double* trace;       /* optional [n_ids][3][5][T]: log P(X | Z, B, r), log P(X, Z, B, r), c_rate, n_rate, gBC */
  int8_t* trace_z;     /* optional [n_ids][3][T][N] */
  int8_t* trace_b;     /* optional [n_ids][3][T][N] */
  double kernel_ms;    /* duration of the chain kernel (CUDA events) */
  uint64_t node_pattern_updates, chain_iterations; /* upward passes: internal nodes x patterns; iterations over all chains */
  uint64_t pt_builds, branch_pattern_updates;      /* (branch, class) P(t) matrices built; downward passes: branches x patterns */
  uint64_t d2h_bytes;                              /* bytes copied device -> host by this call */
  uint64_t phase_cycles[12]; /* SM cycles per kernel phase, thread 0 of every CTA summed (setup, monitor, downward pass, emissions, messages,
                                P(t) builds, proposal emissions, proposal messages, (Z, B) draws, upward pass, ELBO + MH, commit) */
  uint64_t kernel_launches;  /* launches of the chain kernel: the chains go through the device in waves */
} pa_gbgc_results;
/* targets: node ids of TARGETSPECIES (the only tips free to accelerate under M1).  model_mask: bit m set = run model m. */
int pa_engine_gbgc_run(pa_engine* e, const int32_t* ids, int32_t n_ids, const int32_t* targets, int32_t n_targets, const pa_gbgc_params* prm,
                       int model_mask, pa_gbgc_results* out);

/* BPP_C::Eval2 of the GBGC variant (src/PhyloAcc-ST-GBGC/SRC/bpp_c2.cpp:29-285) on a GIVEN chain: the variational estimate of log P(X | model)
 * from the post-burn-in samples -- configurations of Z over the uncertain nodes, within each the sets of gBGC branches, a Gaussian over
 * (gBC, c_rate[, n_rate]).  Host-only (no device needed): it is what pa_engine_gbgc_run applies to every finished chain, exposed so that it
 * can be pinned against the reference's own Eval2 on the same trace.  full / c_rate / n_rate / gbc: [burn + mcmc]; Z, B: [burn + mcmc][N];
 * nodes: the element's pruned tree (post-order, original ids), missing: [N]; resz: the reference's resZ (0 null, 2 restricted, 1 full).
 * max_loglik_out (may be NULL): MonitorChain's MaxLoglik. */
int pa_gbgc_eval2(int32_t N, int32_t burn, int32_t mcmc, int32_t resz, const double* full, const double* c_rate, const double* n_rate,
                  const double* gbc, const int8_t* Z, const int8_t* B, const int32_t* nodes, int32_t n_nodes, const uint8_t* missing,
                  double* loglik_out, double* max_loglik_out);

/* BPP_C::Output_init of the GBGC variant (bpp_c2.cpp:603-650) on a GIVEN chain, host-only: med3_out = {n_rate, c_rate, gBC} as the reference takes
 * them (element T/2 of the trace whose post-burn-in tail is sorted), post_b_out [N] = P(B = 1), post_z_out [N][4] = P(Z = -1, 0, 1, 2) with the
 * reference's rule for missing nodes.  What pa_engine_gbgc_run applies to every finished chain; exposed for parity tests against the reference's
 * own writer. */
int pa_gbgc_output_init(int32_t N, int32_t burn, int32_t mcmc, const double* n_rate, const double* c_rate, const double* gbc, const int8_t* Z,
                        const int8_t* B, const uint8_t* missing, double* med3_out, double* post_b_out, double* post_z_out);

/* PhyloAcc-GT groundwork (no GT sampler yet): what the gene-tree variant evaluates for an element at a given GENE tree
 * embedded in the species tree, for n_states states at once --
 *   loglik_out: log P(X | gene tree, Z, rates), the upward half of GTree::Update_Tg (src/PhyloAcc-GT/genetree.cpp:1330-1418) summed over the
 *               columns as BPP_C::Gibbs does (bpp_c.cpp:423-432); fp64 pruning on the device over the element's compressed site patterns;
 *   prior_out:  the multispecies-coalescent density of the tree, GTree::priorTree (genetree.cpp:560-585);
 *   logpz_out:  log P(missing | Z, tree), GTree::get_logpZ1 (genetree.cpp:1156-1190).
 * The gene tree is passed as GTree holds it (genetree.hpp:28-45): children_gene / heights_gene / missing_gene over the N node slots of the
 * species tree (tips = species), its root, and parent_gene2 flattened -- for every species branch the gene lineages inside it and where each
 * goes (itself: it leaves the branch uncoalesced; another node: it coalesces into that node inside the branch; N: the root's own entry).
 * A gene-tree branch evolves piecewise at the rate of the species branches it crosses (1 / c_rate / n_rate by Z).  Output arrays may be NULL. */
typedef struct pa_gt_states {
  int32_t n_states, N;
  const int32_t* elem;         /* [n_states] element of the loaded batch */
  const int32_t* gchild;       /* [n_states][N][2] children_gene, -1 = none */
  const int32_t* groot;        /* [n_states] */
  const uint8_t* gmissing;     /* [n_states][N] missing_gene */
  const double* gheights;      /* [n_states][N] heights_gene (root of the species tree = 0, tips negative) */
  const int8_t* Z;             /* [n_states][N] latent class of every SPECIES branch */
  const double* n_rate;        /* [n_states] */
  const double* c_rate;        /* [n_states] */
  const int64_t* lin_off;      /* [n_states + 1] range of each state in the three lineage arrays */
  const int32_t* lin_species;  /* parent_gene2: species branch, */
  const int32_t* lin_node;     /*               gene lineage inside it, */
  const int32_t* lin_next;     /*               where the lineage goes */
  const double* heights;       /* [N] heights of the species nodes (BPP::InitPhyloTree, PhyloAcc-GT/bpp.cpp:81-84) */
  const double* thetas;        /* [N] population-size parameter of every species branch (main.cpp:376-414) */
  double cons_to_mis, ncons_to_mis; /* CONSTOMIS and its non-conserved counterpart */
} pa_gt_states;
int pa_engine_gt_replay(pa_engine* e, const pa_gt_states* states, double* loglik_out, double* prior_out, double* logpz_out);

/* File-level front door = the first half of the reference's main() (main.cpp:331-358,406-439): LoadParams,
 * LoadPhyloProfiles, LoadPhyloTree, BPP::BPP, InitMCMC, element-id selection, then pa_engine_load_alignment. */
int pa_engine_create_from_config(const char* cfg_path, int device, pa_engine** out);
/* Borrowed views of what the parsers produced (valid until the handle is destroyed). */
int pa_engine_get_tree(const pa_engine* e, pa_tree* tree_out, pa_subst* subst_out);
const char* pa_engine_node_name(const pa_engine* e, int32_t node);
const char* pa_engine_element_name(const pa_engine* e, int32_t elem);
int pa_engine_get_ids(const pa_engine* e, const int32_t** ids_out, int32_t* n_out);
int pa_engine_get_params(const pa_engine* e, pa_params* out);

/* Whole-program entry = main() of the reference (main.cpp:324-528): parameter file in, result files out
 * (<prefix>_elem_lik.txt, _rate_postZ_M{0,1,2}.txt, _M{0,1,2}_elem_Z.txt, _hyper.txt, _species_names.txt,
 * optional _mcmc_trace_M*_*.txt) and the "time used:" log line. */
int pa_run_from_config(const char* cfg_path);

/* Whole-program entry of the PhyloAcc-ST-GBGC variant = main() of src/PhyloAcc-ST-GBGC/SRC/main.cpp:240-445: the variant's parameter file
 * (LoadParams, main.cpp:89-212: CONSERVE_RATE, ACCE_RATE, PREP_LRATE, PREP_GRATE, GAPCHAR '*', VERBOSE 1 by default ...) in, result files out:
 * <prefix>_hyper.txt, _species_names.txt, _rate_postZ_{0,1,2}.txt (n_rate, c_rate, gBC, P(B) and P(Z = -1..2) per node; file index = resZ),
 * _elem_lik.txt (loglik_NUll, loglik_RES, loglik_all, log_ratio, loglik_Max1..3), _1_elem_Z.txt and, with VERBOSE, _mcmc_trace_{resZ}_{c}.txt
 * (bpp_c2.cpp:567-650, bpp.cpp:536-574).  SAMPLE_HYPER 1 is rejected (exit code 1).  Returns the process exit code. */
int pa_gbgc_run_from_config(const char* cfg_path);

#ifdef __cplusplus
}
#endif
#endif
