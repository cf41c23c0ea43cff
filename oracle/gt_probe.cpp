// oracle/gt_probe.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Driver built around the UNMODIFIED sources of src/PhyloAcc-GT (oracle/Makefile, target ref-gt): the reference's own main.cpp is
// compiled in (its main renamed) so that LoadParams and every default it holds are the reference's, and the other translation units are
// linked as they are.  It dumps, as hex floats, what PhyloAcc-GT computes for an element at a given state -- SURVEY.md §8 f2 groundwork:
//   * the gene tree embedded in the species tree: GTree::initTree_Sptop (genetree.cpp:468-557, species topology) or GTree::initTree
//     (genetree.cpp:366-466, a draw from the multispecies coalescent) -- children_gene, parent_gene, heights_gene, missing_gene, and per
//     species the lineages crossing it (parent_gene2) and the coalescences inside it (temp_coal)
//   * the coalescent prior of that tree              GTree::priorTree      genetree.cpp:560-585
//   * log P(missing | Z, gene tree)                  GTree::get_logpZ1     genetree.cpp:1156-1190
//   * the upward pass over the embedded gene tree    GTree::Update_Tg(., ., ., false, ...)  genetree.cpp:1330-1418, as BPP_C::Gibbs starts it
//     (bpp_c.cpp:385-440), and the log-likelihood  sum_g log_exp_sum(lambda[N-1][root][g].col(2) + log_pi)
// It pins oracle/gt.py and the CUDA kernel behind pa_engine_gt_replay.
//
//   * with "hist 1": one draw of the substitution history on the embedded tree (Update_Tg(.., true, ..), genetree.cpp:1419-1533, the shim's
//     random stream) and, GIVEN that history, the Z emissions of every species branch under the three classes
//     (BPP_C::getEmission, bpp_c.cpp:1043-1216) -- a deterministic function of the dumped history
// usage: gt_probe <cfg> <script>    script lines: "element <c>", "state <resZ> <coal 0|1> <n_rate> <c_rate> Z[N] [hist]"
#include <dirent.h>
#include <math.h>
#include <stdio.h>
#include <sys/time.h>
#include <unistd.h>

#include <algorithm>
#include <cassert>
#include <cmath>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <numeric>
#include <queue>
#include <random>
#include <set>
#include <sstream>
#include <string>
#include <vector>

#include <armadillo>
#include <gsl/gsl_errno.h>
#include <gsl/gsl_randist.h>
#include <gsl/gsl_rng.h>
#include <gsl/gsl_sf_gamma.h>

// every standard / shim header is in before the access specifiers are opened for the reference's own classes
#define private public
#define protected public
#define main reference_main_unused
#include "main.cpp"  // the reference's, from -I$(GT): globals, LoadParams, DispParams
#undef main
#undef private
#undef protected

static void put_vec(const char* tag, const double* v, int n) {
  printf("%s", tag);
  for (int i = 0; i < n; i++) printf(" %a", v[i]);
  printf("\n");
}

int main(int argc, char** argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: gt_probe <cfg> <script>\n");
    return 2;
  }
  fflush(stdout);
  FILE* real_stdout = fdopen(dup(fileno(stdout)), "w");
  freopen("/dev/null", "w", stdout);
  streambuf* old_cout = cout.rdbuf();
  ostringstream sink;
  cout.rdbuf(sink.rdbuf());
  // ---- the reference's own start-up sequence (PhyloAcc-GT/main.cpp:343-425)
  LoadParams(2, argv);
  PhyloProf profile = LoadPhyloProfiles(align_path, segment_path);
  PhyloTree phytree = LoadPhyloTree(phytree_path);
  PhyloTree_theta tree2 = LoadPhyloTree_theta(tree_coal_unit);
  {
    double theta_cum = 0;
    int count_cum = 0;
    int N = phytree.nodes_names.size();
    vector<int> pos_cum = vector<int>(N, 0);
    for (int i = 0; i < (N - 1); i++) {
      if (phytree.nodes_names[i] != tree2.nodes_names[i]) {
        fprintf(stderr, "two trees do not have the same topology\n");
        return 1;
      }
      if (i < phytree.S) {
        phytree.thetas[i] = 0;
      } else if ((tree2.distances[i] != 1.0) && (tree2.distances[i] != 7.0) && (tree2.distances[i] != 0)) {
        phytree.thetas[i] = 2 * phytree.distances[i] / tree2.distances[i];
        if (phytree.thetas[i] >= theta_cutoff) pos_cum[i] = 1;
        else { theta_cum += phytree.thetas[i]; count_cum += 1; }
      } else {
        pos_cum[i] = 1;
      }
    }
    double theta_aver = theta_cum / count_cum;
    pos_cum[N - 1] = 1;
    for (int i = phytree.S; i < N; i++)
      if (pos_cum[i] == 1) phytree.thetas[i] = theta_aver;
  }
  BPP bpp(0, profile, phytree, output_path, targetspecies, outgroup, conserve_prop, conservegroup, ratio0, ratio1, ropt, cub, nlb, nprior_a, nprior_b, cprior_a,
          cprior_b, seed, seed2, prep_grate, prep_lrate, prep_lrate2, prior_grate_a, prior_grate_b, prior_lrate_a, prior_lrate_b, prior_lrate2_a, prior_lrate2_b,
          indel, indel2, missing_thres, sample_indel, prior_dir_par, br_sample_cutoff, deepcoal_species);
  bpp.InitMCMC(num_burn, num_mcmc, num_thin);
  const int N = bpp.N, S = bpp.S;
  fflush(stdout);
  dup2(fileno(real_stdout), fileno(stdout));
  printf("gt N %d S %d\n", N, S);
  put_vec("heights", bpp.heights.data(), N);
  put_vec("thetas", bpp.thetas, N);
  printf("parent");
  for (int s = 0; s < N; s++) printf(" %d", bpp.parent[s]);
  printf("\n");
  {
    double lp[4];
    for (int b = 0; b < 4; b++) lp[b] = bpp.log_pi[b];
    put_vec("logpi", lp, 4);
    double q[16];
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) q[i * 4 + j] = bpp.submat(i, j);
    put_vec("Q", q, 16);
  }
  BPP_C* el = nullptr;
  ifstream script(argv[2]);
  string line;
  while (getline(script, line)) {
    istringstream ss(line);
    string cmd;
    if (!(ss >> cmd)) continue;
    if (cmd == "element") {
      int c;
      ss >> c;
      delete el;
      bool filter = false;
      freopen("/dev/null", "w", stdout);
      el = new BPP_C(c, profile, bpp, gapchar, missing_thres, filter, false, false, consToMis, block, prune, revgap, min_length);
      fflush(stdout);
      dup2(fileno(real_stdout), fileno(stdout));
      printf("element %d filter %d GG %d\n", c, (int)filter, filter ? 0 : el->GG);
      if (filter) { delete el; el = nullptr; }
    } else if (cmd == "state" && el) {
      int resZ, coal;
      double nr, cr;
      ss >> resZ >> coal >> nr >> cr;
      vector<int> Zin(N);
      for (int s = 0; s < N; s++) ss >> Zin[s];
      string hist_flag;
      ss >> hist_flag;
      const bool hist = hist_flag == "hist";
      freopen("/dev/null", "w", stdout);
      bpp.cur_Z[resZ][el->CC] = Zin;
      bpp.cur_nrate[el->CC] = nr;
      bpp.cur_crate[el->CC] = cr;
      el->initMCMC(0, 5, bpp, resZ, prune, false);
      GTree* gt = el->gtree;
      if (coal) {  // a gene tree drawn from the multispecies coalescent instead of the species topology
        if (prune) {
          set<int> alls;
          for (int s = 0; s < N; s++) alls.insert(s);
          gt->initTree(el->missing, alls, bpp);
        } else {
          gt->initTree(el->missing, bpp.upper, bpp);
        }
        for (int s = S; s < N; s++) { gt->lambda[s].clear(); gt->Tg[s].clear(); }
      }
      // the start of BPP_C::Gibbs (bpp_c.cpp:385-440) with WL off: all columns
      const int GG = el->GG;
      gt->GG = GG;
      for (int s = S; s < N; s++)
        for (map<int, int>::iterator it = gt->parent_gene2[s].begin(); it != gt->parent_gene2[s].end(); it++) {
          gt->lambda[s][it->first].resize(gt->GG, zeros<mat>(bpp.num_base, 3));
          gt->Tg[s][it->first].resize(gt->GG, -2);
        }
      vector<bool> visited = vector<bool>(N, false);
      gt->Update_Tg(GG, visited, bpp, false, el->nodes, el->Z, el->trace_n_rate[0], el->trace_c_rate[0], el->log_pi, el->eigenvec, el->eigenval, el->eigeninv);
      double ll = 0;
      for (int g = 0; g < GG; g++) ll += BPP::log_exp_sum(gt->lambda[N - 1][gt->root][g].col(2) + el->log_pi);
      const double lpz = gt->get_logpZ1(bpp, el->Z, el->consToMis, el->nconsToMis);
      const double prior = gt->priorTree(bpp);
      if (hist) {
        gt->Update_Tg(GG, visited, bpp, true, el->nodes, el->Z, el->trace_n_rate[0], el->trace_c_rate[0], el->log_pi, el->eigenvec, el->eigenval, el->eigeninv);
        el->m = 0;
        el->getEmission(GG, bpp);
      }
      fflush(stdout);
      dup2(fileno(real_stdout), fileno(stdout));
      printf("state resZ %d coal %d root %d\n", resZ, coal, gt->root);
      printf("Z");
      for (int s = 0; s < N; s++) printf(" %d", el->Z[s]);
      printf("\n");
      printf("rates %a %a\n", el->trace_n_rate[0], el->trace_c_rate[0]);
      printf("consmis %a %a\n", el->consToMis, el->nconsToMis);
      printf("gchild");
      for (int s = 0; s < N; s++) printf(" %d %d", gt->children_gene[s][0], gt->children_gene[s][1]);
      printf("\n");
      printf("gparent");
      for (int s = 0; s < N; s++) printf(" %d", gt->parent_gene[s]);
      printf("\n");
      put_vec("gheights", gt->heights_gene, N);
      printf("gmissing");
      for (int s = 0; s < N; s++) printf(" %d", (int)gt->missing_gene[s]);
      printf("\n");
      for (int s = 0; s < N; s++) {
        if (gt->parent_gene2[s].empty()) continue;
        printf("cross %d", s);
        for (map<int, int>::iterator it = gt->parent_gene2[s].begin(); it != gt->parent_gene2[s].end(); it++) printf(" %d %d", it->first, it->second);
        printf("\n");
        if (gt->temp_coal[s].empty()) continue;
        printf("coal %d", s);
        for (size_t k = 0; k < gt->temp_coal[s].size(); k++) printf(" %d", gt->temp_coal[s][k]);
        printf("\n");
      }
      if (hist) {
        for (int s = 0; s < N; s++)
          for (map<int, vector<int> >::iterator it = gt->Tg[s].begin(); it != gt->Tg[s].end(); it++) {
            printf("tg %d %d", s, it->first);
            for (int g = 0; g < GG; g++) printf(" %d", it->second[g]);
            printf("\n");
          }
        // tips with an ambiguity code (Tg = -1): the bases the code allows, in the probe's own column order (the GT front-end moves
        // columns: blocks of identical sites first)
        for (int s = 0; s < S; s++) {
          if (gt->Tg[s].find(s) == gt->Tg[s].end()) continue;
          for (int g = 0; g < GG; g++)
            if (gt->Tg[s][s][g] == -1) {
              printf("amb %d %d", s, g);
              for (int b = 0; b < 4; b++) printf(" %d", gt->lambda[s][s][g](b, 2) == 0 ? 1 : 0);
              printf("\n");
            }
        }
        for (size_t k = 0; k < el->nodes.size(); k++) {
          const int s = el->nodes[k];
          printf("emission %d %a %a %a\n", s, el->log_emission[s][0], el->log_emission[s][1], el->log_emission[s][2]);
        }
      }
      printf("loglik %a\nlogpz %a\nprior %a\nend\n", ll, lpz, prior);
    }
  }
  delete el;
  cout.rdbuf(old_cout);
  return 0;
}
