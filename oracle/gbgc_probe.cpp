// oracle/gbgc_probe.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Driver linked against the UNMODIFIED translation units of src/PhyloAcc-ST-GBGC/SRC (oracle/Makefile, target ref-gbgc) that
// dumps, as hex floats, what the GC-biased model computes for a branch:
//   * BPP_C::getRate(Q0, s, b)           Q o f(s + b B), rows renormalised            (GBGC/SRC/bpp_c.cpp:13-30)
//   * BPP_C::getTransition(node, s, b)   log P(t)^T from the eigensystem of that Q    (GBGC/SRC/bpp_c.cpp:1405-1427)
// for the grid of (node, selection, bias) triples in the script, and
//   * replayed log-likelihoods of an element for a given state (Z, B per branch; conserved / accelerated selection coefficients;
//     the element's gBGC bias): the six rate matrices of initMCMC (GBGC/SRC/bpp_c.cpp:196-204), init_cache_TM (:1563-1621), the
//     upward pass Update_Tg(g, ., ., false) (:477-528) and sum_g log_exp_sum(lambda[g][root]).
// It pins oracle/gbgc.py and the CUDA kernels behind pa_gbgc_transition / pa_engine_gbgc_replay (phyloacc_b200/csrc/pa_gbgc.cu).
//
// usage: gbgc_probe <cfg> <script>     script lines: "tm <node> <sel> <bias>", "element <c>", "state <sel_c> <sel_n> <gbc> Z[N] B[N]"
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
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

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: gbgc_probe <cfg> <script>\n"); return 2; }
  map<string, string> kv;
  {
    ifstream in(argv[1]);
    string line;
    while (getline(in, line)) {
      istringstream ss(line);
      string k, v;
      if (!(ss >> k)) continue;
      ss >> v;
      kv[k] = v;
    }
  }
  auto get = [&](const char* k, const char* d) { return kv.count(k) ? kv[k] : string(d); };
  // defaults of GBGC/SRC/main.cpp:40-90
  fflush(stdout);
  streambuf* old_cout = cout.rdbuf();
  ostringstream sink;
  cout.rdbuf(sink.rdbuf());
  FILE* real_stdout = fdopen(dup(fileno(stdout)), "w");
  freopen("/dev/null", "w", stdout);
  PhyloProf profile = LoadPhyloProfiles(get("ALIGN_FILE", ""), get("SEG_FILE", ""));
  PhyloTree phytree = LoadPhyloTree(get("PHYTREE_FILE", ""));
  BPP bpp(0, profile, phytree, string("."), get("TARGETSPECIES", ""), get("OUTGROUP", ""), 0.8, get("CONSERVE", ""), get("REF", ""), -2, 0, 0, 1, 0.6, 5, 0.4, 5,
          0.4, 4, 0.8, 0.2, 1, 1, 7728, 0.1, 0.8, 0.01, 0.1, 0.8, false);
  bpp.InitMCMC(20, 40, 10);
  bool filter = false;
  string gc = get("GAPCHAR", "*");
  BPP_C* elp = new BPP_C(0, profile, bpp, gc[0], 0.8, filter, false);
#define el (*elp)
  fflush(stdout);
  dup2(fileno(real_stdout), fileno(stdout));
  printf("gbgc N %d filter %d\n", bpp.N, (int)filter);
  {
    printf("Q");
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) printf(" %a", bpp.submat(i, j));
    printf("\n");
  }
  ifstream script(argv[2]);
  string line;
  while (getline(script, line)) {
    istringstream ss(line);
    string cmd;
    if (!(ss >> cmd)) continue;
    if (cmd == "tm") {
      int s; double sel, bias;
      ss >> s >> sel >> bias;
      mat Q = el.getRate(bpp.submat, sel, bias);
      mat L = el.getTransition(s, sel, bias, bpp);
      printf("tm %d %a %a t %a\nrate", s, sel, bias, el.distances2[s]);
      for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) printf(" %a", Q(i, j));
      printf("\nlogtm");
      for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) printf(" %a", L(i, j));  // (child, parent) layout like the ST caches
      printf("\n");
    } else if (cmd == "element") {
      int c; ss >> c;
      delete elp;
      filter = false;
      elp = new BPP_C(c, profile, bpp, gc[0], 0.8, filter, false);
      printf("element %d filter %d GG %d root %d nodes %d\n", c, (int)filter, filter ? 0 : el.GG, filter ? -1 : el.root, filter ? 0 : (int)el.nodes.size());
    } else if (cmd == "state") {
      if (filter) { printf("state skipped\n"); continue; }
      double sc, sn, gb;
      ss >> sc >> sn >> gb;
      const int N = bpp.N;
      el.initMCMC(0, bpp, 1);
      for (int s = 0; s < N; s++) { int z; ss >> z; el.Z[s] = z; }
      for (int s = 0; s < N; s++) { int b; ss >> b; el.isGB[s] = (b != 0); }
      for (int s = 0; s < N; s++) el.fixZ[s] = -1;  // init_cache_TM skips the accelerated tables where fixZ == 1
      vector<mat> Qs(6);
      Qs[0] = bpp.submat;
      Qs[1] = el.getRate(bpp.submat, sc, 0);
      Qs[2] = el.getRate(bpp.submat, sn, 0);
      Qs[3] = el.getRate(bpp.submat, 0, gb);
      Qs[4] = el.getRate(bpp.submat, sc, gb);
      Qs[5] = el.getRate(bpp.submat, sn, gb);
      el.init_cache_TM(Qs);
      for (int s = bpp.S; s < N; s++)
        for (int g = 0; g < el.GG; g++) el.lambda[g][s].fill(0);
      vector<bool> visited(N, false);
      double ll = 0;
      for (int g = 0; g < el.GG; g++) {
        el.Update_Tg(g, visited, bpp, false);
        ll += BPP::log_exp_sum(el.lambda[g][el.root]);
      }
      printf("loglik %a\n", ll);
    } else if (cmd == "eval2") {
      // eval2 <resZ> <burn> <mcmc> <trace file>: BPP_C::Eval2 (GBGC/SRC/bpp_c2.cpp:29-285) on a GIVEN chain.  The file holds one row per
      // iteration: log P(X, Z, B, r), c_rate, n_rate, gBC, Z[N], B[N].  MaxLoglik / Max_Z are taken as MonitorChain takes them
      // (bpp_c.cpp:655-664: first maximum of the post-burn-in log P(X, Z, B, r)).
      if (filter) { printf("eval2 skipped\n"); continue; }
      int resZ, burn, mcmc; string path;
      ss >> resZ >> burn >> mcmc >> path;
      const int N = bpp.N, T = burn + mcmc;
      el.num_burn = burn; el.num_mcmc = mcmc;
      el.trace_full_loglik.assign(T, 0.0); el.trace_loglik.assign(T, 0.0);
      el.trace_c_rate.assign(T, 0.0); el.trace_n_rate.assign(T, 0.0); el.trace_n_GB.assign(T, 0.0);
      el.trace_Z.assign(T, vector<int>(N, 1));
      el.trace_GB.assign(T, vector<bool>(N, false));
      ifstream tf(path.c_str());
      for (int m = 0; m < T; m++) {
        tf >> el.trace_full_loglik[m] >> el.trace_c_rate[m] >> el.trace_n_rate[m] >> el.trace_n_GB[m];
        for (int s2 = 0; s2 < N; s2++) { int z; tf >> z; el.trace_Z[m][s2] = z; }
        for (int s2 = 0; s2 < N; s2++) { int b; tf >> b; el.trace_GB[m][s2] = (b != 0); }
      }
      el.MaxLoglik = -INFINITY;
      for (int m = burn; m < T; m++)
        if (el.MaxLoglik < el.trace_full_loglik[m]) { el.MaxLoglik = el.trace_full_loglik[m]; el.Max_Z = el.trace_Z[m]; }
      el.verbose = false;
      bpp.log_liks_null[el.CC] = bpp.log_liks_resZ[el.CC] = bpp.log_liks_sgl[el.CC] = 0;
      el.Eval2(bpp, resZ);
      const double v = resZ == 0 ? bpp.log_liks_null[el.CC] : (resZ == 2 ? bpp.log_liks_resZ[el.CC] : bpp.log_liks_sgl[el.CC]);
      printf("eval2 %d %a\n", resZ, v);
    } else if (cmd == "output_init") {
      // output_init <resZ> <burn> <mcmc> <trace file>: the reference's own row of _rate_postZ_<resZ>.txt (BPP_C::Output_init,
      // GBGC/SRC/bpp_c2.cpp:603-650) for a GIVEN chain (file format as for eval2), echoed as it was written
      if (filter) { printf("output_init skipped\n"); continue; }
      int resZ, burn, mcmc; string path;
      ss >> resZ >> burn >> mcmc >> path;
      const int N = bpp.N, T = burn + mcmc;
      el.num_burn = burn; el.num_mcmc = mcmc;
      el.trace_full_loglik.assign(T, 0.0); el.trace_loglik.assign(T, 0.0);
      el.trace_c_rate.assign(T, 0.0); el.trace_n_rate.assign(T, 0.0); el.trace_n_GB.assign(T, 0.0);
      el.trace_Z.assign(T, vector<int>(N, 1));
      el.trace_GB.assign(T, vector<bool>(N, false));
      ifstream tf(path.c_str());
      for (int m = 0; m < T; m++) {
        tf >> el.trace_full_loglik[m] >> el.trace_c_rate[m] >> el.trace_n_rate[m] >> el.trace_n_GB[m];
        for (int s2 = 0; s2 < N; s2++) { int z; tf >> z; el.trace_Z[m][s2] = z; }
        for (int s2 = 0; s2 < N; s2++) { int b; tf >> b; el.trace_GB[m][s2] = (b != 0); }
      }
      const string rowfile = path + ".row";
      {
        ofstream oz(rowfile.c_str());
        el.Output_init(string("."), string("."), bpp, oz, resZ);
      }
      ifstream rf(rowfile.c_str());
      string row;
      getline(rf, row);
      printf("output_init %d\t%s\n", resZ, row.c_str());
    } else if (cmd == "monitor") {
      // monitor <resZ> <loglik> <sel_c> <sel_n> <gbc> Z[N] B[N]: what Gibbs records for a state -- log P(constraints) of initMCMC
      // (GBGC/SRC/bpp_c.cpp:232-254), logp_Z = prior_Z(Z) - logpZC (:273, bpp_c2.cpp:484-498) and MonitorChain's log P(X, Z, B, r)
      // (bpp_c.cpp:598-652) given the state's log P(X | Z, B, r)
      if (filter) { printf("monitor skipped\n"); continue; }
      int resZ; double ll, sc, sn, gb;
      ss >> resZ >> ll >> sc >> sn >> gb;
      const int N = bpp.N;
      fflush(stdout);
      freopen("/dev/null", "w", stdout);
      el.initMCMC(0, bpp, resZ);
      for (int s2 = 0; s2 < N; s2++) { int z; ss >> z; el.Z[s2] = z; }
      for (int s2 = 0; s2 < N; s2++) { int b; ss >> b; el.isGB[s2] = (b != 0); }
      const double lpz = el.prior_Z(el.Z) - el.logpZC;
      el.trace_loglik[1] = ll; el.trace_c_rate[1] = sc; el.trace_n_rate[1] = sn; el.trace_n_GB[1] = gb;
      int a1 = 0, a2 = 0;
      el.MonitorChain(1, true, a1, a2, bpp, lpz, resZ);
      fflush(stdout);
      dup2(fileno(real_stdout), fileno(stdout));
      printf("monitor %d %a %a %a\n", resZ, el.trace_full_loglik[1], lpz, el.logpZC);
    } else if (cmd == "integ") {
      // integ <resZ> <sel_c> <sel_n> <gbc> Z[N] B[N]: the evaluations of one Gibbs iteration that do not involve random numbers, at a given
      // state -- expected parent / child base pairs per branch (Update_Tg(g, ., ., true), GBGC/SRC/bpp_c.cpp:530-583), the six
      // (Z, B) emissions of every branch, the 6-state upward messages and the ELBO term (initial_prob_back, bpp_c2.cpp:354-407).
      // The model's own fixZ is kept (resZ); where it forbids class 2 the requested Z is lowered to 1 (printed back).
      if (filter) { printf("integ skipped\n"); continue; }
      int resZ; double sc, sn, gb;
      ss >> resZ >> sc >> sn >> gb;
      const int N = bpp.N;
      vector<int> Zin(N), Bin(N);
      for (int s = 0; s < N; s++) ss >> Zin[s];
      for (int s = 0; s < N; s++) ss >> Bin[s];
      el.initMCMC(0, bpp, resZ);
      for (int s = 0; s < N; s++) {
        if (el.fixZ[s] == 1 && Zin[s] == 2) Zin[s] = 1;
        el.Z[s] = Zin[s];
        el.isGB[s] = (Bin[s] != 0);
      }
      vector<mat> Qs(6);
      Qs[0] = bpp.submat;
      Qs[1] = el.getRate(bpp.submat, sc, 0);
      Qs[2] = el.getRate(bpp.submat, sn, 0);
      Qs[3] = el.getRate(bpp.submat, 0, gb);
      Qs[4] = el.getRate(bpp.submat, sc, gb);
      Qs[5] = el.getRate(bpp.submat, sn, gb);
      el.init_cache_TM(Qs);
      for (int s = bpp.S; s < N; s++)
        for (int g = 0; g < el.GG; g++) el.lambda[g][s].fill(0);
      vector<bool> visited(N, false);
      for (int g = 0; g < el.GG; g++) el.Update_Tg(g, visited, bpp, false);
      for (size_t k = 0; k < el.nodes.size(); k++) el.posterior_Tg_next[el.nodes[k]].fill(0);
      for (int g = 0; g < el.GG; g++) el.Update_Tg(g, visited, bpp, true);
      el.posterior_Tg = el.posterior_Tg_next;
      const double elbo = el.initial_prob_back(bpp);
      printf("integ %d elbo %a loglik %a\n", resZ, elbo, el.log_prob_back[el.root][0]);
      printf("Zused");
      for (int s = 0; s < N; s++) printf(" %d", el.Z[s]);
      printf("\nfixZ");
      for (int s = 0; s < N; s++) printf(" %d", el.fixZ[s]);
      printf("\n");
      for (size_t k = 0; k + 1 < el.nodes.size(); k++) {
        const int s = el.nodes[k];
        printf("post %d", s);
        for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) printf(" %a", el.posterior_Tg[s](i, j));  // row: child base, column: parent base
        printf("\nemis %d", s);
        for (int i = 0; i < 6; i++) printf(" %a", el.log_emission[s][i]);
        printf("\ntmint %d", s);
        for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) printf(" %a", el.log_TM_Int[s](i, j));
        printf("\n");
      }
      printf("rootback");
      for (int i = 0; i < 6; i++) printf(" %a", el.log_prob_back[el.root][i]);
      printf("\nconsts %a %a %a\n", bpp.nGBprior_c, bpp.consToMis, bpp.nconsToMis);
    }
  }
  cout.rdbuf(old_cout);
  return 0;
}
