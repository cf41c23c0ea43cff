// pa_app.cpp -- file-level front door: the reference's main() (main.cpp:324-528) on top of the engine.
#include <dirent.h>

#include <cmath>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <sstream>

#include "pa_engine.h"
#include "pa_internal.h"

namespace pa {
void set_last_error(const std::string& s);

static bool directory_exists(const std::string& p) {  // main.cpp:291-307
  if (p == "") return false;
  DIR* d = opendir(p.c_str());
  if (d) {
    closedir(d);
    return true;
  }
  return false;
}

// LoadParams + LoadPhyloProfiles + LoadPhyloTree + BPP::BPP + id selection (main.cpp:331-358,406-439)
static int create_from_config(const char* cfg_path, int device, pa_engine** out) {
  pa_engine* e = new pa_engine();
  e->from_config = true;
  Params& P = e->params;
  std::cout << "Loading program configurations from " << cfg_path << "......" << std::endl;
  Tree tree;
  Alignment aln;
  if (!load_params(cfg_path, P)) { delete e; return PA_ERR_ARG; }
  if (!load_alignment_files(P.align_path, P.segment_path, aln)) { delete e; return PA_ERR_IO; }
  std::cout << "Loading phylogenetic tree from " << P.phytree_path << "......" << std::endl;
  if (!load_tree(P.phytree_path, tree)) { delete e; return PA_ERR_IO; }
  if (!match_alignment_to_tree(tree, aln)) {
    set_error("alignment has fewer species than the tree");
    delete e;
    return PA_ERR_ARG;
  }
  std::vector<int32_t> og = species_ids(tree, P.outgroup), cg = species_ids(tree, P.conservegroup);
  P.num.n_outgroup = (int32_t)og.size();
  P.num.outgroup = og.data();
  P.num.n_conserve = (int32_t)cg.size();
  P.num.conserve = cg.data();
  pa_tree pt{tree.S, tree.N, tree.parent.data(), tree.children.data(), tree.dist.data()};
  pa_subst ps;
  memcpy(ps.pi, tree.pi, sizeof(ps.pi));
  memcpy(ps.Q, tree.Q, sizeof(ps.Q));
  int rc = e->eng.init(&pt, &ps, &P.num, device);
  P.num.outgroup = nullptr;
  P.num.conserve = nullptr;
  if (rc != PA_OK) { delete e; return rc; }
  e->eng.model.tree.names = tree.names;
  e->element_names = aln.element_names;
  const int C = (int)aln.element_names.size();
  std::vector<int> ids;
  if (!load_ids(P, C, ids)) { delete e; return PA_ERR_IO; }
  e->ids.assign(ids.begin(), ids.end());
  double mean_len = 0;
  for (int c = 0; c < C; c++) mean_len += (double)aln.elem_len[c] / std::max(C, 1);
  std::cout << "# total length = " << aln.G << " (" << C << "). # Species = " << tree.S << ". # elements = " << C
            << ". Mean gene set size = " << mean_len << "." << std::endl;
  std::cout << "# Burn-ins = " << P.num.burnin << ". # MCMC Updates = " << P.num.mcmc << ". # adaptive frequency = " << P.num.adapt_freq
            << ".  RND SEED = " << P.num.seed << "." << std::endl;
  // rows -> one [S][G] byte matrix
  std::vector<uint8_t> X((size_t)tree.S * aln.G);
  for (int s = 0; s < tree.S; s++) memcpy(X.data() + (size_t)s * aln.G, aln.X[s].data(), aln.G);
  std::vector<std::string>().swap(aln.X);
  // elements running past the alignment are clipped like std::string::substr does in the reference (bpp_c.hpp:211)
  for (int c = 0; c < C; c++) {
    if (aln.elem_start[c] > aln.G) aln.elem_start[c] = aln.G;
    if (aln.elem_start[c] + aln.elem_len[c] > aln.G) aln.elem_len[c] = (int32_t)(aln.G - aln.elem_start[c]);
    if (aln.elem_len[c] < 0) aln.elem_len[c] = 0;
  }
  rc = e->eng.load_alignment(X.data(), aln.G, C, aln.elem_start.data(), aln.elem_len.data());
  if (rc != PA_OK) { delete e; return rc; }
  *out = e;
  return PA_OK;
}

// writers: main.cpp:362-401, bpp_c2.cpp:706-736,739-785, bpp.cpp:568-619
static int run_from_config(const char* cfg_path) {
  const time_t start = time(NULL);
  std::cout << std::fixed << std::setprecision(4);
  // peek at RESULT_FOLDER first, as the reference checks it before loading data (main.cpp:334-338)
  {
    Params P0;
    if (!load_params(cfg_path, P0)) {
      std::cerr << get_error() << std::endl;
      return 1;
    }
    if (!directory_exists(P0.output_path)) {
      std::cout << "output path doesn't exist or empty!" << std::endl;
      return 1;
    }
  }
  pa_engine* e = nullptr;
  int rc = create_from_config(cfg_path, 0, &e);
  if (rc != PA_OK) {
    std::cerr << "(Error. " << get_error() << ")" << std::endl;
    return 1;
  }
  const Params& P = e->params;
  const Tree& tree = e->eng.model.tree;
  const int N = tree.N;
  const std::string prefix = P.output_path + "/" + P.result_prefix;
  // model slot -> file suffix: M0, M1, M2 (crossed variable names in the reference)
  std::ofstream out_hyper((prefix + "_hyper.txt").c_str());
  out_hyper << "iter\tnprior_a\tnprior_b\tcprior_a\tcprior_b\tprior_l_a\tprior_l_b\tprior_g_a\tprior_g_b\n";
  out_hyper << 0 << "\t" << P.num.acce_prior_a << "\t" << P.num.acce_prior_b << "\t" << P.num.conserve_prior_a << "\t" << P.num.conserve_prior_b
            << "\t" << P.num.hyper_lrate_a << "\t" << P.num.hyper_lrate_b << "\t" << P.num.hyper_grate_a << "\t" << P.num.hyper_grate_b << std::endl;
  std::ofstream out_lik_hyper;  // SAMPLE_HYPER 1: _elem_lik.txt is opened once and appended to by every outer iteration (main.cpp:373-378)
  if (P.sample_hyper) {
    out_lik_hyper.open((prefix + "_elem_lik.txt").c_str());
    out_lik_hyper.precision(8);
    out_lik_hyper << "No.\tID\tloglik_all\tloglik_Max" << std::endl;
  }
  std::ofstream out_Z[3];
  for (int m = 0; m < 3; m++) {
    out_Z[m].open((prefix + "_rate_postZ_M" + std::to_string(m) + ".txt").c_str());
    out_Z[m] << "No.\tn_rate\tc_rate\tg_rate\tl_rate\tl2_rate";
    for (int s = 0; s < N; s++)
      for (int k = 0; k < 4; k++) out_Z[m] << "\t" << tree.names[s] << "_" << k;
    out_Z[m] << std::endl;
  }
  {
    std::ofstream sp((prefix + "_species_names.txt").c_str());
    for (int s = 0; s < N; s++) sp << tree.names[s] << std::endl;
  }
  std::cout << e->ids.size() << " of elements to be computed" << std::endl;

  const int C = (int)e->eng.plans.size();
  std::vector<int32_t> ids;
  for (int32_t c : e->ids) {
    if (c < 0 || c >= C) {
      std::cout << c << " 1 Standard exception: element id out of range" << std::endl;  // reference: caught per element (main.cpp:497-499)
      continue;
    }
    ids.push_back(c);
  }
  const int n = (int)ids.size();
  std::vector<uint8_t> filtered(std::max(n, 1));
  std::vector<double> loglik((size_t)std::max(n, 1) * 3, 0.0), loglik_max((size_t)std::max(n, 1) * 3, 0.0), med((size_t)std::max(n, 1) * 15, 0.0),
      postZ((size_t)std::max(n, 1) * 3 * N * 4, 0.0);
  std::vector<int8_t> maxZ((size_t)std::max(n, 1) * 3 * N, 0);
  pa_results res{n, N, filtered.data(), loglik.data(), loglik_max.data(), maxZ.data(), med.data(), postZ.data()};
  for (int iter = 0; iter < P.num_chain; iter++) {
    std::cout << "Running MCMC chain " << iter + 1 << " ..." << std::endl;
    // elements go to the device in waves so that the traces of one wave fit comfortably in HBM and host memory
    const int wave = 4096;
    for (int w0 = 0; w0 < n; w0 += wave) {
      const int wn = std::min(wave, n - w0);
      pa_results wres{wn, N, filtered.data() + w0, loglik.data() + (size_t)w0 * 3, loglik_max.data() + (size_t)w0 * 3,
                      maxZ.data() + (size_t)w0 * 3 * N, med.data() + (size_t)w0 * 15, postZ.data() + (size_t)w0 * 3 * N * 4};
      // SAMPLE_HYPER 1 runs only the full model (main.cpp:463-494); outer iterations after the first continue every chain
      // from its last state (initMCMC with iter > 0, bpp_c.cpp:148-177).  The reference saves that state only `if(UpHyper)`
      // (bpp_c.cpp:474-485): with SAMPLE_HYPER 0 and CHAIN > 1 its later iterations restart from bpp.cur_Z (all zeros: every node
      // starts in the background state) and the INIT_* rates -- a fresh chain on a new random stream here (Z = 1 start as in iter 0).
      rc = e->eng.sample(ids.data() + w0, wn, P.sample_hyper ? (1 << PA_M2) : 7, P.sample_hyper && iter > 0, iter);
      if (rc == PA_OK) rc = e->eng.collect(&wres);
      if (rc != PA_OK) {
        std::cerr << "(Error. " << get_error() << ")" << std::endl;
        delete e;
        return 1;
      }
      // VERBOSE: per-chain traces (Output_sampling, bpp_c2.cpp:706-736); file index is the reference's third argument 0,1,2
      if (P.verbose) {
        for (int i = 0; i < wn; i++) {
          if (filtered[w0 + i]) continue;
          for (int m = 0; m < 3; m++) {
            ChainTraceHost tr;
            if (e->eng.get_trace(i, m, tr) != PA_OK) continue;
            const std::string path = prefix + "_mcmc_trace_M" + std::to_string(m) + "_" + std::to_string(ids[w0 + i]) + ".txt";
            std::ofstream ot;
            ot.precision(8);
            if (iter == 0) {
              ot.open(path.c_str());
              ot << "loglik\trate_n\trate_c\tgrate\tlrate\tlrate2\t";
              for (int s = 0; s < N; s++) ot << tree.names[s] << "\t";
              ot << std::endl;
            } else ot.open(path.c_str(), std::ios::app);
            const int T = (int)tr.loglik.size();
            for (int t = 0; t < T; t++) {
              ot << tr.loglik[t] << "\t" << tr.n_rate[t] << "\t" << tr.c_rate[t] << "\t" << tr.g_rate[t] << "\t" << tr.l_rate[t] << "\t" << tr.l2_rate[t] << "\t";
              for (int s = 0; s < N; s++) ot << (int)tr.Z[(size_t)t * N + s] << "\t";
              ot << std::endl;
            }
          }
        }
      }
    }
    (void)res;
    // rows of _rate_postZ_Mk.txt (Output_init, bpp_c2.cpp:776-783; default ofstream precision = 6 significant digits)
    for (int i = 0; i < n; i++) {
      if (filtered[i]) continue;
      for (int m = 0; m < 3; m++) {
        if (P.sample_hyper && m != PA_M2) continue;
        const double* md = &med[((size_t)i * 3 + m) * 5];
        out_Z[m] << ids[i] << "\t" << md[0] << "\t" << md[1] << "\t" << md[2] << "\t" << md[3] << "\t" << md[4];
        const double* pz = &postZ[(((size_t)i * 3 + m) * N) * 4];
        for (int s = 0; s < N; s++)
          for (int k = 0; k < 4; k++) out_Z[m] << "\t" << pz[s * 4 + k];
        out_Z[m] << std::endl;
      }
    }
    if (P.sample_hyper) {
      // BPP::sample_hyperparam + BPP::Output_init0 (main.cpp:504-506, bpp.cpp:324-477, 622-635)
      pa_hyper hy;
      rc = e->eng.sample_hyper(iter, e->ids.data(), (int32_t)e->ids.size(), &hy);
      if (rc != PA_OK) {
        std::cout << " 2 Standard exception: " << get_error() << std::endl;
      } else {
        out_hyper << iter << "\t" << hy.nprior_a << "\t" << hy.nprior_b << "\t" << hy.cprior_a << "\t" << hy.cprior_b << "\t" << hy.prior_l_a << "\t"
                  << hy.prior_l_b << "\t" << hy.prior_g_a << "\t" << hy.prior_g_b << "\t" << hy.prior_l2_a << "\t" << hy.prior_l2_b << std::endl;
      }
      for (int i = 0; i < n; i++)
        out_lik_hyper << ids[i] << "\t" << e->element_names[ids[i]] << "\t" << loglik[i * 3 + PA_M2] << "\t" << loglik_max[i * 3 + PA_M2] << std::endl;
      continue;
    }
    // _elem_lik.txt and _M{k}_elem_Z.txt (BPP::Output_init, bpp.cpp:568-619): one row per id, zeros for filtered ones
    {
      std::ofstream ol((prefix + "_elem_lik.txt").c_str());
      ol.precision(8);
      ol << "No.\tID\tloglik_Null\tloglik_Acc\tloglik_Full\tlogBF1\tlogBF2\tloglik_Max_M0\tloglik_Max_M1\tloglik_Max_M2" << std::endl;
      for (int i = 0; i < n; i++) {
        const double l0 = loglik[i * 3 + 0], l1 = loglik[i * 3 + 1], l2 = loglik[i * 3 + 2];
        ol << ids[i] << "\t" << e->element_names[ids[i]] << "\t" << l0 << "\t" << l1 << "\t" << l2 << "\t";
        ol << l1 - l0 << "\t" << l1 - l2;
        ol << "\t" << loglik_max[i * 3 + 0] << "\t" << loglik_max[i * 3 + 1] << "\t" << loglik_max[i * 3 + 2];
        ol << std::endl;
      }
      for (int m = 0; m < 3; m++) {
        std::ofstream oz((prefix + "_M" + std::to_string(m) + "_elem_Z.txt").c_str());
        oz << "No.";
        for (int s = 0; s < N; s++) oz << "\t" << tree.names[s];
        oz << std::endl;
        for (int i = 0; i < n; i++) {
          oz << ids[i];
          for (int s = 0; s < N; s++) oz << "\t" << (int)maxZ[((size_t)i * 3 + m) * N + s];
          oz << std::endl;
        }
      }
    }
  }
  for (int m = 0; m < 3; m++) out_Z[m].close();
  delete e;
  std::cout << std::endl << std::endl << "time used:  " << (time(NULL) - start) / 60 << " min." << std::endl << std::endl;
  return 0;
}

}  // namespace pa

extern "C" {

int pa_engine_create_from_config(const char* cfg_path, int device, pa_engine** out) {
  if (!cfg_path || !out) return PA_ERR_ARG;
  *out = nullptr;
  try {
    const int rc = pa::create_from_config(cfg_path, device, out);
    if (rc != PA_OK) pa::set_last_error(pa::get_error());
    return rc;
  } catch (const std::exception& ex) {
    pa::set_last_error(std::string("internal error: ") + ex.what());
    return PA_ERR_ARG;
  }
}

int pa_engine_get_tree(const pa_engine* e, pa_tree* t, pa_subst* s) {
  if (!e) return PA_ERR_ARG;
  const pa::Tree& tr = e->eng.model.tree;
  if (t) {
    t->S = tr.S;
    t->N = tr.N;
    t->parent = tr.parent.data();
    t->children = tr.children.data();
    t->distances = tr.dist.data();
  }
  if (s) {
    memcpy(s->pi, tr.pi, sizeof(s->pi));
    memcpy(s->Q, tr.Q, sizeof(s->Q));
  }
  return PA_OK;
}

const char* pa_engine_node_name(const pa_engine* e, int32_t node) {
  if (!e || node < 0 || node >= (int)e->eng.model.tree.names.size()) return "";
  return e->eng.model.tree.names[node].c_str();
}

const char* pa_engine_element_name(const pa_engine* e, int32_t elem) {
  if (!e || elem < 0 || elem >= (int)e->element_names.size()) return "";
  return e->element_names[elem].c_str();
}

int pa_engine_get_ids(const pa_engine* e, const int32_t** ids_out, int32_t* n_out) {
  if (!e || !ids_out || !n_out) return PA_ERR_ARG;
  *ids_out = e->ids.data();
  *n_out = (int32_t)e->ids.size();
  return PA_OK;
}

int pa_engine_get_params(const pa_engine* e, pa_params* out) {
  if (!e || !out) return PA_ERR_ARG;
  *out = e->eng.model.prm;
  out->n_outgroup = (int32_t)e->eng.model.outgroup.size();
  out->outgroup = e->eng.model.outgroup.data();
  out->n_conserve = (int32_t)e->eng.model.conservedgroup.size();
  out->conserve = e->eng.model.conservedgroup.data();
  return PA_OK;
}

int pa_run_from_config(const char* cfg_path) {
  if (!cfg_path) return 1;
  try {
    return pa::run_from_config(cfg_path);
  } catch (const std::exception& ex) {
    std::cerr << "(Error. " << ex.what() << ")" << std::endl;
    return 1;
  }
}

}  // extern "C"
