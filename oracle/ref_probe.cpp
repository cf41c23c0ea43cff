// oracle/ref_probe.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A driver that links against the UNMODIFIED reference translation units
// (bpp.cpp, bpp_c.cpp, bpp_c2.cpp, newick.cpp, profile.cpp, utils.cpp under
// /root/reference/src/PhyloAcc-ST; see oracle/Makefile) and exposes their
// internal state as hex-float text, so that the oracle restatement and the
// CUDA engine can be pinned against what the reference itself computes:
//   * the tree / eigensystem / null P(t) built by BPP::BPP          (bpp.hpp:173-404)
//   * the per-element front-end of BPP_C::BPP_C                     (bpp_c.hpp:143-476)
//   * replayed log-likelihoods for given (model, Z, n_rate, c_rate) (bpp_c.cpp:131-343,647-671)
//   * emissions + upward Z messages for a given ancestral history   (bpp_c2.cpp:355-390, bpp_c.cpp:1182-1212)
//   * Eval2 for a given trace                                       (bpp_c2.cpp:14-353)
//   * the reference's own WRITERS fed with given traces: Output_sampling / BPP_C::Output_init / BPP::Output_init
//     (bpp_c2.cpp:706-785, bpp.cpp:568-619) -- byte-level format parity of the result files
// The reference's own traces carry only 8 significant digits (bpp_c2.cpp:710),
// which cannot pin 1e-10, hence this probe.
//
// usage: ref_probe <cfg> <script>      (script commands documented in run())
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <set>
#include <map>
#include <fstream>
#include <sstream>
#include <iostream>
#include <algorithm>
#include <cmath>
#include <queue>
#include <cassert>
#include <unistd.h>
#include <armadillo>

#define private public
#define protected public
#include "bpp.hpp"
#include "bpp_c.hpp"
#undef private
#undef protected

using namespace std;
using namespace arma;

struct Cfg {
  map<string, string> kv;
  string get(const string& k, const string& d) const { auto it = kv.find(k); return it == kv.end() ? d : it->second; }
  double getd(const string& k, double d) const { auto it = kv.find(k); return it == kv.end() ? d : atof(it->second.c_str()); }
  int geti(const string& k, int d) const { auto it = kv.find(k); return it == kv.end() ? d : atoi(it->second.c_str()); }
};

static Cfg load_cfg(const char* path) {
  Cfg c;
  ifstream in(path);
  string line;
  while (getline(in, line)) {
    istringstream ss(line);
    string k, v;
    if (!(ss >> k)) continue;
    if (k[0] == '#') continue;
    ss >> v;
    c.kv[k] = v;
  }
  return c;
}

static void dump_vec_hex(const char* key, const double* p, size_t n) {
  printf("%s", key);
  for (size_t i = 0; i < n; i++) printf(" %a", p[i]);
  printf("\n");
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: ref_probe <cfg> <script>\n"); return 2; }
  Cfg cfg = load_cfg(argv[1]);
  // defaults follow main.cpp:49-90
  string outgroup = cfg.get("OUTGROUP", ""), target = cfg.get("TARGETSPECIES", ""), conserve = cfg.get("CONSERVE", "");
  double conserve_prop = cfg.getd("CONSERVE_PROP", 0.8), ratio0 = cfg.getd("INIT_CONSERVE_RATE", 0.5), ratio1 = cfg.getd("INIT_ACCE_RATE", 1);
  int ropt = cfg.geti("ROPT", 1);
  double cub = cfg.getd("CUB", 1), nlb = cfg.getd("NLB", 0.6);
  double npa = cfg.getd("ACCE_PRIOR_A", 10), npb = cfg.getd("ACCE_PRIOR_B", 0.2), cpa = cfg.getd("CONSERVE_PRIOR_A", 5), cpb = cfg.getd("CONSERVE_PRIOR_B", 0.04);
  int seed = cfg.geti("SEED", 5);
  double grate = cfg.getd("INIT_GRATE", 0.5), lrate = cfg.getd("INIT_LRATE", 0.3), lrate2 = cfg.getd("INIT_LRATE2", 0.1);
  double ga = cfg.getd("HYPER_GRATE_A", 3), gb = cfg.getd("HYPER_GRATE_B", 1), la = cfg.getd("HYPER_LRATE_A", 1), lb = cfg.getd("HYPER_LRATE_B", 9);
  double l2a = cfg.getd("HYPER_LRATE2_A", 1), l2b = cfg.getd("HYPER_LRATE2_B", 9);
  if (l2a == 0) lrate2 = 0;
  double missing_thres = cfg.getd("GAP_PROP", 0.8), consToMis = cfg.getd("CONSTOMIS", 0.01), revgap = cfg.getd("TRIM_GAP_PERCENT", 1);
  string gc = cfg.get("GAPCHAR", "-");
  char gapchar = gc[0];
  string ps = cfg.get("PRUNE_TREE", "0");
  bool prune = (ps == "1" || ps == "true");
  int min_length = cfg.geti("MIN_LEN", 50);
  int num_burn = cfg.geti("BURNIN", 200), num_mcmc = cfg.geti("MCMC", 800), num_thin = cfg.geti("ADAPT_FREQ", 500);

  // keep the reference's chatter off stdout while loading
  fflush(stdout);
  streambuf* old_cout = cout.rdbuf();
  ostringstream sink;
  cout.rdbuf(sink.rdbuf());
  FILE* real_stdout = fdopen(dup(fileno(stdout)), "w");
  freopen("/dev/null", "w", stdout);  // printTree uses printf
  PhyloProf profile = LoadPhyloProfiles(cfg.get("ALIGN_FILE", ""), cfg.get("SEG_FILE", ""));
  PhyloTree phytree = LoadPhyloTree(cfg.get("PHYTREE_FILE", ""));
  BPP bpp(0, profile, phytree, string("."), target, outgroup, conserve_prop, conserve, ratio0, ratio1, ropt, cub, nlb, npa, npb, cpa, cpb,
          seed, grate, lrate, lrate2, ga, gb, la, lb, l2a, l2b, 0.0, 0.0, missing_thres, false);
  bpp.InitMCMC(num_burn, num_mcmc, num_thin);
  fflush(stdout);
  // restore stdout for our own dump (cout stays muted: the reference prints from initMCMC / Eval2)
  dup2(fileno(real_stdout), fileno(stdout));

  const int N = bpp.N, S = bpp.S;
  printf("tree S %d N %d C %d G %d\n", S, N, bpp.C, bpp.G);
  printf("names");
  for (int s = 0; s < N; s++) printf(" %s", bpp.nodes_names[s].c_str());
  printf("\n");
  printf("parent");
  for (int s = 0; s < N; s++) printf(" %d", bpp.parent[s]);
  printf("\n");
  printf("children");
  for (int s = 0; s < N; s++) printf(" %d %d", bpp.children[s][0], bpp.children[s][1]);
  printf("\n");
  dump_vec_hex("distances", bpp.distances, N);
  dump_vec_hex("pi", bpp.pi.memptr(), 4);
  {
    double q[16];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) q[i * 4 + j] = bpp.submat(i, j);
    dump_vec_hex("Q", q, 16);
  }
  printf("upper");
  for (set<int>::iterator it = bpp.upper.begin(); it != bpp.upper.end(); ++it) printf(" %d", *it);
  printf("\nupper_conserve");
  for (set<int>::iterator it = bpp.upper_conserve.begin(); it != bpp.upper_conserve.end(); ++it) printf(" %d", *it);
  printf("\nelement_size");
  for (int c = 0; c < bpp.C; c++) printf(" %u", bpp.element_size[c]);
  printf("\n");
  for (int s = 0; s < N - 1; s++) {
    double m[16];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) m[i * 4 + j] = bpp.log_cache_TM_null[s](i, j);  // (child,parent) layout, row-major dump
    printf("logTMnull %d", s);
    for (int i = 0; i < 16; i++) printf(" %a", m[i]);
    printf("\n");
  }

  // ---- script
  ifstream script(argv[2]);
  string line;
  BPP_C* el = NULL;
  int cur_c = -1;
  bool filter = false;
  string w_prefix;
  ofstream w_outZ[3];
  vector<int> w_ids;
  while (getline(script, line)) {
    istringstream ss(line);
    string cmd;
    if (!(ss >> cmd)) continue;
    if (cmd == "element") {
      // element <c> : construct BPP_C (bpp_c.hpp:143-476) and dump the integer front-end
      int c; ss >> c;
      if (el) { delete el; el = NULL; }
      filter = false;
      el = new BPP_C(c, profile, bpp, gapchar, missing_thres, filter, false, consToMis, prune, revgap, min_length);
      cur_c = c;
      printf("element %d filter %d GG %d\n", c, (int)filter, filter ? 0 : el->GG);
      if (filter) continue;
      printf("root %d\n", el->root);
      printf("missing");
      for (int s = 0; s < N; s++) printf(" %d", (int)el->missing[s]);
      printf("\nnodes");
      for (size_t i = 0; i < el->nodes.size(); i++) printf(" %d", el->nodes[i]);
      printf("\nchildren2");
      for (int s = 0; s < N; s++) printf(" %d %d", el->children2[s][0], el->children2[s][1]);
      printf("\nparent2");
      for (int s = 0; s < N; s++) printf(" %d", el->parent2[s]);
      printf("\n");
      dump_vec_hex("distances2", el->distances2, N);
      printf("upper_c");
      for (size_t i = 0; i < el->upper_c.size(); i++) printf(" %d", el->upper_c[i]);
      printf("\nupper_conserve_c");
      for (size_t i = 0; i < el->upper_conserve_c.size(); i++) printf(" %d", el->upper_conserve_c[i]);
      printf("\n");
      // per column: tip masks (bit b set <=> lambda[g][s][b]==0) and Tg for all nodes
      for (int g = 0; g < el->GG; g++) {
        printf("col %d masks", g);
        for (int s = 0; s < S; s++) {
          int m = 0;
          for (int b = 0; b < 4; b++) if (el->lambda[g][s][b] == 0) m |= (1 << b);
          printf(" %d", m);
        }
        printf(" tg");
        for (int s = 0; s < N; s++) printf(" %d", el->Tg[g][s]);
        printf("\n");
      }
    } else if (cmd == "state") {
      // state <resZ> <n_rate> <c_rate> <Z[0..N-1]> : replayed log-likelihood (initMCMC + upward pass as Gibbs does, bpp_c.cpp:355-356)
      if (!el || filter) { printf("state skipped\n"); continue; }
      int resZ; double nr, cr;
      ss >> resZ >> nr >> cr;
      vector<int> Z(N);
      for (int s = 0; s < N; s++) ss >> Z[s];
      bpp.cur_nrate[cur_c] = nr;
      bpp.cur_crate[cur_c] = cr;
      bpp.cur_Z[resZ][cur_c] = Z;
      el->initMCMC(1, bpp, resZ);
      vector<bool> visited(N, false);
      double tot = 0;
      vector<double> persite(el->GG);
      for (int g = 0; g < el->GG; g++) {
        el->Update_Tg(g, visited, bpp, false);
        persite[g] = BPP::log_exp_sum(el->lambda[g][el->root]);
        tot += persite[g];
      }
      printf("loglik %d %a\n", resZ, tot);
      dump_vec_hex("persite", persite.data(), persite.size());
      printf("fixZ");
      for (int s = 0; s < N; s++) printf(" %d", el->fixZ[s]);
      printf("\n");
    } else if (cmd == "history") {
      // history <g_rate> <l_rate> <l2_rate> then GG lines "h <Tg of internal nodes S..N-1>" : set sampled bases, run
      // getEmission + the deterministic upward part of Update_Z_subtree; dumps log_emission and log_prob_back.
      if (!el || filter) { printf("history skipped\n"); continue; }
      double gr, lr, lr2;
      ss >> gr >> lr >> lr2;
      for (int g = 0; g < el->GG; g++) {
        string hl;
        getline(script, hl);
        istringstream hs(hl);
        string tag; hs >> tag;
        for (int s = S; s < N; s++) { int v; hs >> v; if (el->Tg[g][s] < 4 && v >= 0) el->Tg[g][s] = v; }
      }
      // rebuild log_TM_Int with the requested rates (same statements as bpp_c.cpp:573-599)
      for (size_t i = 0; i < el->nodes.size(); i++) {
        int s = el->nodes[i];
        if (el->fixZ[s] == 1) {
          el->log_TM_Int[s](1, 1) = 0; el->log_TM_Int[s](2, 1) = log(0.0);
          el->log_TM_Int[s](0, 0) = log(1 - gr); el->log_TM_Int[s](1, 0) = log(gr); el->log_TM_Int[s](2, 0) = log(0.0);
        } else {
          el->log_TM_Int[s](0, 0) = log(1 - gr); el->log_TM_Int[s](1, 0) = log(gr); el->log_TM_Int[s](2, 0) = log(0.0);
          el->log_TM_Int[s](1, 1) = log(1 - lr); el->log_TM_Int[s](2, 1) = log(lr);
          el->log_TM_Int[s](1, 2) = log(lr2); el->log_TM_Int[s](2, 2) = log(1 - lr2);
        }
      }
      vector<int> changed = el->Update_Z_subtree(4);
      for (int s = 0; s < N; s++) {
        printf("emission %d %a %a %a\n", s, el->log_emission[s][0], el->log_emission[s][1], el->log_emission[s][2]);
      }
      for (int s = 0; s < N; s++) {
        printf("back %d %a %a %a\n", s, el->log_prob_back[s][0], el->log_prob_back[s][1], el->log_prob_back[s][2]);
      }
    } else if (cmd == "hyper") {
      // hyper <iter> <C> then C lines "r <n> <c> <g> <l> <l2>" : set bpp.cur_* and run the reference's own
      // BPP::sample_hyperparam (bpp.cpp:324-477); its cout lines carry the MH ratios and the proposals
      int iter, Cn;
      ss >> iter >> Cn;
      vector<int> ids;
      for (int c = 0; c < Cn; c++) {
        string rl;
        getline(script, rl);
        istringstream rs(rl);
        string tag; double a, b, g, l, l2;
        rs >> tag >> a >> b >> g >> l >> l2;
        if (c < bpp.C) {
          bpp.cur_nrate[c] = a; bpp.cur_crate[c] = b; bpp.cur_grate[c] = g; bpp.cur_lrate[c] = l; bpp.cur_lrate2[c] = l2;
          ids.push_back(c);
        }
      }
      printf("hyper_before %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", bpp.nprior_a, bpp.nprior_b, bpp.cprior_a, bpp.cprior_b,
             bpp.prior_l_a, bpp.prior_l_b, bpp.prior_g_a, bpp.prior_g_b, bpp.prior_l2_a, bpp.prior_l2_b);
      ostringstream cap;
      streambuf* prev = cout.rdbuf(cap.rdbuf());
      const streamsize old_prec = cout.precision(17);
      ofstream devnull("/dev/null");
      bpp.sample_hyperparam(iter, ids, devnull);
      cout.rdbuf(prev);
      cout.precision(old_prec);
      istringstream lines(cap.str());
      string ln;
      while (getline(lines, ln)) printf("hyper_log %s\n", ln.c_str());
      printf("hyper_after %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", bpp.nprior_a, bpp.nprior_b, bpp.cprior_a, bpp.cprior_b,
             bpp.prior_l_a, bpp.prior_l_b, bpp.prior_g_a, bpp.prior_g_b, bpp.prior_l2_a, bpp.prior_l2_b);
    } else if (cmd == "writers") {
      // writers <prefix> : open <prefix>_rate_postZ_M{0,1,2}.txt with main()'s header line (main.cpp:380-399)
      ss >> w_prefix;
      for (int k = 0; k < 3; k++) {
        w_outZ[k].open((w_prefix + "_rate_postZ_M" + to_string(k) + ".txt").c_str());
        w_outZ[k] << "No.\tn_rate\tc_rate\tg_rate\tl_rate\tl2_rate";
        for (int s = 0; s < N; s++)
          for (int q = 0; q < 4; q++) w_outZ[k] << "\t" << bpp.nodes_names[s] << "_" << q;
        w_outZ[k] << endl;
      }
      printf("writers %s\n", w_prefix.c_str());
    } else if (cmd == "out") {
      // out <resZ> <file index> <T> then T lines "t <loglik> <n> <c> <g> <l> <l2> <Z[0..N-1]>": the traces of one chain as Gibbs leaves
      // them; runs the reference's MonitorChain bookkeeping, Eval2, then ITS writers Output_sampling and Output_init (main.cpp:466-494)
      if (!el || filter) { printf("out skipped\n"); continue; }
      int resZ, fidx, T;
      ss >> resZ >> fidx >> T;
      if (T != num_burn + num_mcmc) { printf("out bad T\n"); continue; }
      el->initMCMC(0, bpp, resZ);
      for (int m = 0; m < T; m++) {
        string tl;
        getline(script, tl);
        istringstream ts(tl);
        string tag; double ll, nr, cr, gr, lr, l2r;
        ts >> tag >> ll >> nr >> cr >> gr >> lr >> l2r;
        for (int s = 0; s < N; s++) ts >> el->Z[s];
        el->trace_n_rate[m] = nr; el->trace_c_rate[m] = cr;
        el->trace_g_rate[m] = gr; el->trace_l_rate[m] = lr; el->trace_l2_rate[m] = l2r;
        if (m == 0) {  // MonitorChain(0) recomputes the log-likelihood from lambda: take slot 1's path instead (see eval2 above)
          el->trace_n_rate[1] = nr; el->trace_c_rate[1] = cr;
          el->MonitorChain(1, bpp, ll, 0.0, resZ);
          el->trace_loglik[0] = el->trace_loglik[1];
          el->trace_full_loglik[0] = el->trace_full_loglik[1];
          el->trace_Z[0] = el->trace_Z[1];
        } else {
          el->MonitorChain(m, bpp, ll, 0.0, resZ);
        }
      }
      bpp.log_liks_Z[resZ][cur_c] = el->MaxLoglik;  // Gibbs, bpp_c.cpp:470-471
      bpp.Max_Z[resZ][cur_c] = el->Max_Z;
      el->Eval2(bpp, resZ);
      el->Output_sampling(0, w_prefix, bpp, fidx);
      el->Output_init(w_prefix, w_prefix, bpp, w_outZ[fidx], resZ);
      if (std::find(w_ids.begin(), w_ids.end(), cur_c) == w_ids.end()) w_ids.push_back(cur_c);
      printf("out %d %d done\n", resZ, fidx);
    } else if (cmd == "skip") {
      // skip <c> : an id whose element was filtered -- it still gets its (all-zero) row in _elem_lik.txt / _elem_Z.txt
      int c; ss >> c;
      if (std::find(w_ids.begin(), w_ids.end(), c) == w_ids.end()) w_ids.push_back(c);
    } else if (cmd == "finish") {
      // finish : BPP::Output_init (bpp.cpp:568-619) over the ids seen, in script order
      for (int k = 0; k < 3; k++) w_outZ[k].close();
      bpp.Output_init(profile, w_prefix, w_ids);
      printf("finish %d ids\n", (int)w_ids.size());
    } else if (cmd == "eval2") {
      // eval2 <resZ> <T> then T lines "t <loglik> <n> <c> <Z[0..N-1]>" : fills the traces as Gibbs would, then runs
      // MonitorChain-equivalent bookkeeping (full loglik, Max_Z) through the reference's own MonitorChain and Eval2.
      if (!el || filter) { printf("eval2 skipped\n"); continue; }
      int resZ, T;
      ss >> resZ >> T;
      if (T != num_burn + num_mcmc) { printf("eval2 bad T\n"); continue; }
      // initMCMC for fixZ; Z content is irrelevant here
      el->initMCMC(0, bpp, resZ);
      for (int m = 0; m < T; m++) {
        string tl;
        getline(script, tl);
        istringstream ts(tl);
        string tag; double ll, nr, cr;
        ts >> tag >> ll >> nr >> cr;
        for (int s = 0; s < N; s++) ts >> el->Z[s];
        el->trace_n_rate[m] = nr;
        el->trace_c_rate[m] = cr;
        double l = ll;
        // m==0 recomputes loglik from lambda inside MonitorChain; use m>0 path semantics by temporarily shifting
        if (m == 0) {
          // reproduce bpp_c.cpp:749-758 without touching lambda: emulate by calling with m=1 slot then moving
          el->trace_n_rate[1] = nr; el->trace_c_rate[1] = cr;
          el->MonitorChain(1, bpp, l, 0.0, resZ);
          el->trace_loglik[0] = el->trace_loglik[1];
          el->trace_full_loglik[0] = el->trace_full_loglik[1];
          el->trace_Z[0] = el->trace_Z[1];
          if (num_burn > 1 || true) { /* Max tracking: slot 1 >= num_burn only if num_burn<=1; handled by caller using num_burn>=2 */ }
        } else {
          el->MonitorChain(m, bpp, l, 0.0, resZ);
        }
      }
      bpp.log_liks_null[cur_c] = bpp.log_liks_resZ[cur_c] = bpp.log_liks_sgl[cur_c] = 0;
      el->Eval2(bpp, resZ);
      double v = resZ == 0 ? bpp.log_liks_null[cur_c] : (resZ == 2 ? bpp.log_liks_resZ[cur_c] : bpp.log_liks_sgl[cur_c]);
      printf("eval2 %d %a maxloglik %a failure %d\n", resZ, v, el->MaxLoglik, (int)el->failure);
      printf("maxZ");
      for (int s = 0; s < N; s++) printf(" %d", el->Max_Z[s]);
      printf("\n");
    }
    fflush(stdout);
  }
  if (el) delete el;
  cout.rdbuf(old_cout);
  return 0;
}
