// pa_gbgc_app.cpp -- file-level front door of the PhyloAcc-ST-GBGC variant: main() of src/PhyloAcc-ST-GBGC/SRC/main.cpp:240-445 on top of
// pa_engine_gbgc_run.  Same inputs (the variant's own parameter keys, .mod, FASTA, BED, id file) and the same result files:
//   <prefix>_hyper.txt, <prefix>_species_names.txt, <prefix>_rate_postZ_{0,1,2}.txt (file index = the reference's resZ: 0 null, 2 restricted,
//   1 full), <prefix>_elem_lik.txt, <prefix>_1_elem_Z.txt, and with VERBOSE (the variant's default) <prefix>_mcmc_trace_{resZ}_{element}.txt.
// Elements are cut into shards over the visible GPUs like the ST program (pa_app.cpp); chains are keyed by the element number.
#include <cmath>
#include <cstring>
#include <ctime>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <sstream>
#include <thread>

#include "pa_app.h"
#include "pa_fmt.h"
#include "pa_nvtx.h"

namespace pa {
void set_last_error(const std::string& s);

namespace {

struct GbgcConfig {
  Params P;            // what the shared loaders and the engine's integer front-end need
  pa_gbgc_params g;
  std::string refspecies;
  int adapt_freq = 100;  // ADAPT_FREQ (num_thin): printed, not used inside Gibbs (update_rates(.., false))
};

std::string trim_chars(std::string s, const char* ts) {
  const size_t a = s.find_first_not_of(ts);
  if (a == std::string::npos) return "";
  const size_t b = s.find_last_not_of(ts);
  return s.substr(a, b - a + 1);
}

// LoadParams of the variant (GBGC/SRC/main.cpp:89-212): first token = key, value read with operator>> (a value that does not parse leaves
// the default); unknown keys are ignored without a message; defaults main.cpp:31-86
bool load_gbgc_params(const char* path, GbgcConfig& c) {
  std::ifstream in(path);
  if (!in) {
    set_error(std::string("Cannot open the parameters file: ") + path);
    return false;
  }
  Params& P = c.P;
  pa_gbgc_params_default(&c.g);
  P.batch = 0;
  P.num_chain = 0;
  P.verbose = true;
  P.sample_hyper = false;
  P.num.gapchar = '*';
  P.num.gap_prop = 0.8;
  P.num.conserve_prop = 0.8;
  P.num.constomis = 0.01;
  P.num.num_thread = 1;
  // the variant's BPP_C constructor neither trims gap columns nor filters short elements nor prunes the tree (GBGC/SRC/bpp_c.hpp)
  P.num.min_len = 0;
  P.num.trim_gap_percent = 1;
  P.num.prune_tree = 0;
  int seed = 7728, ropt = 0;
  double cub = 1, nlb = 0.6, indel = 0, indel2 = 0;
  bool sample_indel = false;
  std::string line;
  while (std::getline(in, line)) {
    std::istringstream ls(line);
    std::string k;
    ls >> k;
    if (k == "PHYTREE_FILE") ls >> P.phytree_path;
    else if (k == "ALIGN_FILE") ls >> P.align_path;
    else if (k == "SEG_FILE") ls >> P.segment_path;
    else if (k == "ID_FILE") ls >> P.id_path;
    else if (k == "BATCH") ls >> P.batch;
    else if (k == "RESULT_FOLDER") ls >> P.output_path;
    else if (k == "PREFIX") ls >> P.result_prefix;
    else if (k == "SEED") ls >> seed;
    else if (k == "CONSERVE_RATE") ls >> c.g.sel_c0;
    else if (k == "ACCE_RATE") ls >> c.g.sel_n0;
    else if (k == "CONSERVE_PRIOR_A") ls >> c.g.cprior_a;
    else if (k == "CONSERVE_PRIOR_B") ls >> c.g.cprior_b;
    else if (k == "ACCE_PRIOR_A") ls >> c.g.nprior_a;
    else if (k == "ACCE_PRIOR_B") ls >> c.g.nprior_b;
    else if (k == "ROPT") ls >> ropt;
    else if (k == "CUB") ls >> cub;
    else if (k == "NLB") ls >> nlb;
    else if (k == "BURNIN") ls >> c.g.num_burn;
    else if (k == "MCMC") ls >> c.g.num_mcmc;
    else if (k == "ADAPT_FREQ") ls >> c.adapt_freq;
    else if (k == "PREP_LRATE") ls >> c.g.lrate;
    else if (k == "PREP_GRATE") ls >> c.g.grate;
    else if (k == "CHAIN") ls >> P.num_chain;
    else if (k == "OUTGROUP") ls >> P.outgroup;
    else if (k == "TARGETSPECIES") ls >> P.targetspecies;
    else if (k == "CONSERVE") ls >> P.conservegroup;
    else if (k == "CONSERVE_PROP") ls >> P.num.conserve_prop;
    else if (k == "GAP_PROP") ls >> P.num.gap_prop;
    else if (k == "REF") ls >> c.refspecies;
    else if (k == "CONSTOMIS") ls >> c.g.cons_to_mis;
    else if (k == "GAPCHAR") ls >> P.num.gapchar;
    else if (k == "INDEL") ls >> indel;
    else if (k == "INDEL2") ls >> indel2;
    else if (k == "SAMPLE_INDEL") ls >> sample_indel;
    else if (k == "SAMPLE_HYPER") ls >> P.sample_hyper;
    else if (k == "VERBOSE") ls >> P.verbose;
    else if (k == "NUM_THREAD") ls >> P.num.num_thread;
  }
  P.phytree_path = trim_chars(P.phytree_path, " \"\t\n");
  P.align_path = trim_chars(P.align_path, " \"\t\n");
  P.output_path = trim_chars(P.output_path, " \"\t\n");
  P.segment_path = trim_chars(P.segment_path, " \"\t\n");
  c.g.seed = (uint32_t)seed;
  P.num.seed = seed;
  P.num.burnin = c.g.num_burn;
  P.num.mcmc = c.g.num_mcmc;
  P.num.adapt_freq = c.adapt_freq;
  P.num.constomis = c.g.cons_to_mis;
  // id selection: the variant takes every element for BATCH 0 (main.cpp:331-347), the shared loader for BATCH -1
  if (P.batch == 0) P.batch = -1;
  return true;
}

int run_gbgc_from_config(const char* cfg_path) {
  const time_t start = time(NULL);
  std::cout << std::fixed << std::setprecision(4);
  std::cout << "Loading input data and running parameters......" << std::endl;
  std::cout << "Loading program configurations from " << cfg_path << "......" << std::endl;
  GbgcConfig cfg;
  if (!load_gbgc_params(cfg_path, cfg)) {
    std::cerr << get_error() << std::endl;
    return 1;
  }
  if (!directory_exists(cfg.P.output_path)) {  // main.cpp:250-254
    std::cout << "output path doesn't exist or empty!" << std::endl;
    return 1;
  }
  if (cfg.P.sample_hyper) {
    std::cerr << "(Error. SAMPLE_HYPER 1 is not supported by the GPU build of PhyloAcc-ST-GBGC: the variant's sweep (bpp.cpp:324-383) takes logs of the negative selection coefficients and is NaN in the reference itself)"
              << std::endl;
    return 1;
  }
  if (cfg.g.num_mcmc < 2 || cfg.g.num_burn < 0) {
    std::cerr << "(Error. BURNIN must be >= 0 and MCMC >= 2)" << std::endl;
    return 1;
  }
  ConfigInputs in;
  in.P = cfg.P;
  int rc = load_inputs(in);
  if (rc != PA_OK) {
    std::cerr << "(Error. " << get_error() << ")" << std::endl;
    return 1;
  }
  std::cout << "# Threads = " << in.P.num.num_thread << std::endl << std::endl;
  const Params& P = in.P;
  const Tree& tree = in.tree;
  const int N = tree.N, S = tree.S, C = in.C, T = cfg.g.num_burn + cfg.g.num_mcmc;
  // TARGETSPECIES -> leaf ids (BPP::BPP, GBGC/SRC/bpp.hpp:222-228); names that match no species are dropped
  std::vector<int32_t> targets;
  for (int32_t t : species_ids(tree, P.targetspecies))
    if (t >= 0 && t < S) targets.push_back(t);
  const std::string prefix = P.output_path + "/" + P.result_prefix;
  static const int kFileOfModel[3] = {0, 2, 1};  // model slot M0 / M1 / M2 -> the reference's resZ, which names the files
  std::ofstream out_hyper((prefix + "_hyper.txt").c_str());
  out_hyper << "iter\tnprior_a\tnprior_b\tcprior_a\tcprior_b\tindel\tindel2\tgrate\tlrate\n";
  std::ofstream out_Z[3];  // indexed by resZ
  for (int r = 0; r < 3; r++) {
    out_Z[r].open((prefix + "_rate_postZ_" + std::to_string(r) + ".txt").c_str());
    out_Z[r] << "No.\tn_rate\tc_rate\tgBC";
    for (int s = 0; s < N; s++) out_Z[r] << "\t" << tree.names[s] << "_B";
    for (int s = 0; s < N; s++)
      for (int k = 0; k < 4; k++) out_Z[r] << "\t" << tree.names[s] << "_" << k;
    out_Z[r] << std::endl;
  }
  {
    std::ofstream sp((prefix + "_species_names.txt").c_str());
    for (int s = 0; s < N; s++) sp << tree.names[s] << std::endl;
  }
  std::vector<int32_t> ids;
  for (int c : in.ids) {
    if (c < 0 || c >= C) {
      std::cout << c << " Standard exception: element id out of range" << std::endl;
      continue;
    }
    ids.push_back(c);
  }
  const int n = (int)ids.size();
  std::cout << in.ids.size() << " of elements to be computed" << std::endl;
  // what BPP keeps per element (bpp.cpp:200-211): zero until an element has been computed
  std::vector<double> lik_null(C, 0.0), lik_res(C, 0.0), lik_all(C, 0.0), lik_max((size_t)3 * C, 0.0);  // lik_max indexed [resZ][c]
  std::vector<int> maxZ_full((size_t)C * N, 0);

  const std::vector<int> devices = pick_devices(n);
  const int n_dev = (int)devices.size();
  std::vector<int> shard_lo(n_dev + 1, 0);
  {
    double total = 0, acc = 0;
    for (int i = 0; i < n; i++) total += std::max(1, in.elem_len[ids[i]]);
    int d = 1;
    for (int i = 0; i < n && d < n_dev; i++) {
      acc += std::max(1, in.elem_len[ids[i]]);
      while (d < n_dev && acc >= total * d / n_dev) shard_lo[d++] = i + 1;
    }
    for (; d <= n_dev; d++) shard_lo[d] = n;
  }
  std::vector<pa_engine*> engines(n_dev, nullptr);
  std::vector<int> dev_rc(n_dev, PA_OK);
  std::vector<std::string> dev_err(n_dev);
  auto for_each_device = [&](const std::function<int(int)>& fn) {
    std::vector<std::thread> th;
    for (int d = 0; d < n_dev; d++)
      th.emplace_back([&, d]() {
        try {
          dev_rc[d] = fn(d);
          if (dev_rc[d] != PA_OK) dev_err[d] = get_error();
        } catch (const std::exception& ex) {
          dev_rc[d] = PA_ERR_ARG;
          dev_err[d] = std::string("internal error: ") + ex.what();
        }
      });
    for (auto& t : th) t.join();
    for (int d = 0; d < n_dev; d++)
      if (dev_rc[d] != PA_OK) {
        std::cerr << "(Error. GPU " << d << ": " << dev_err[d] << ")" << std::endl;
        return false;
      }
    return true;
  };
  auto destroy_all = [&]() {
    for (pa_engine* e : engines) delete e;
  };
  if (n_dev > 1) std::cout << "# element shards = " << n_dev << " (one engine and host thread per shard)." << std::endl;
  const int threads_per_dev = std::max(1, P.num.num_thread / n_dev);
  if (!for_each_device([&](int d) {
        PA_NVTX_RANGE("load shard");
        int r = make_engine(in, devices[d], threads_per_dev, &engines[d]);
        if (r != PA_OK) return r;
        const int lo = shard_lo[d], cnt = shard_lo[d + 1] - lo;
        std::vector<int64_t> st(cnt);
        std::vector<int32_t> ln(cnt), gid(cnt);
        for (int k = 0; k < cnt; k++) {
          st[k] = in.elem_start[ids[lo + k]];
          ln[k] = in.elem_len[ids[lo + k]];
          gid[k] = ids[lo + k];
        }
        return engines[d]->eng.load_alignment(in.X.data(), in.G, cnt, st.data(), ln.data(), gid.data());
      })) {
    destroy_all();
    return 1;
  }
  // rows of the three _rate_postZ files in the order of the id list, whatever the shards: one text block per element and model
  std::vector<std::string> rows((size_t)std::max(n, 1) * 3);
  std::vector<uint8_t> filtered(std::max(n, 1), 0);
  for (int iter = 0; iter < P.num_chain; iter++) {
    std::cout << "Running MCMC chain " << iter + 1 << " ..." << std::endl;
    // BPP::sample_proposal (bpp.cpp): one line per outer iteration; the indel columns belong to the unused 5-state model
    out_hyper << iter << "\t" << cfg.g.nprior_a << "\t" << cfg.g.nprior_b << "\t" << cfg.g.cprior_a << "\t" << cfg.g.cprior_b << "\t" << 0 << "\t" << 0
              << "\t" << cfg.g.grate << "\t" << cfg.g.lrate << std::endl;
    pa_gbgc_params prm = cfg.g;
    prm.seed = cfg.g.seed + (uint32_t)iter * 0x9E3779B9u;  // a new stream per outer iteration (chains restart: the variant's initMCMC ignores iter)
    const bool ok = for_each_device([&](int d) {
      Engine& eng = engines[d]->eng;
      const int lo = shard_lo[d], cnt = shard_lo[d + 1] - lo;
      const int wave = P.verbose ? 512 : 4096;  // traces of a wave are held on the host while its files are written
      for (int w0 = 0; w0 < cnt; w0 += wave) {
        const int wn = std::min(wave, cnt - w0);
        std::vector<int32_t> local(wn);
        for (int k = 0; k < wn; k++) local[k] = w0 + k;
        std::vector<uint8_t> filt(wn);
        std::vector<double> ll((size_t)wn * 3), lmax((size_t)wn * 3), med((size_t)wn * 9), pb((size_t)wn * 3 * N), pz((size_t)wn * 3 * N * 4);
        std::vector<int8_t> mz((size_t)wn * 3 * N), mb((size_t)wn * 3 * N);
        std::vector<int32_t> acc((size_t)wn * 6);
        std::vector<double> tr;
        std::vector<int8_t> tz, tb;
        if (P.verbose) {
          tr.resize((size_t)wn * 3 * 5 * T);
          tz.resize((size_t)wn * 3 * T * N);
          tb.resize((size_t)wn * 3 * T * N);
        }
        pa_gbgc_results res;
        memset(&res, 0, sizeof(res));
        res.n_ids = wn; res.N = N; res.T = T;
        res.filtered = filt.data(); res.loglik = ll.data(); res.loglik_max = lmax.data(); res.med = med.data(); res.post_b = pb.data();
        res.post_z = pz.data(); res.max_z = mz.data(); res.max_b = mb.data(); res.accept = acc.data();
        if (P.verbose) { res.trace = tr.data(); res.trace_z = tz.data(); res.trace_b = tb.data(); }
        int r;
        {
          PA_NVTX_RANGE("gbgc wave");
          r = eng.gbgc_run(local.data(), wn, targets.data(), (int32_t)targets.size(), &prm, 7, &res);
        }
        if (r != PA_OK) return r;
        PA_NVTX_RANGE("gbgc write");
        for (int i = 0; i < wn; i++) {
          const size_t gi = (size_t)lo + w0 + i;
          const int c = ids[gi];
          filtered[gi] = filt[i];
          if (filt[i]) continue;
          lik_null[c] = ll[(size_t)i * 3 + 0]; lik_res[c] = ll[(size_t)i * 3 + 1]; lik_all[c] = ll[(size_t)i * 3 + 2];
          for (int m = 0; m < 3; m++) lik_max[(size_t)kFileOfModel[m] * C + c] = lmax[(size_t)i * 3 + m];
          for (int s = 0; s < N; s++) maxZ_full[(size_t)c * N + s] = mz[((size_t)i * 3 + 2) * N + s];
          for (int m = 0; m < 3; m++) {
            // Output_init (GBGC/SRC/bpp_c2.cpp:603-650): default stream formatting (6 significant digits)
            std::string& row = rows[gi * 3 + m];
            row.clear();
            const double* md = &med[((size_t)i * 3 + m) * 3];
            fmt_i(row, c);
            for (int k = 0; k < 3; k++) { row.push_back('\t'); fmt_g(row, md[k]); }
            for (int s = 0; s < N; s++) { row.push_back('\t'); fmt_g(row, pb[((size_t)i * 3 + m) * N + s]); }
            for (int k = 0; k < 4 * N; k++) { row.push_back('\t'); fmt_g(row, pz[((size_t)i * 3 + m) * N * 4 + k]); }
            row.push_back('\n');
            if (!P.verbose) continue;
            // Output_sampling (bpp_c2.cpp:567-600): log P(X, Z, B, r), n_rate, c_rate, gBC, B per node, Z per node; 8 significant digits
            const std::string path = prefix + "_mcmc_trace_" + std::to_string(kFileOfModel[m]) + "_" + std::to_string(c) + ".txt";
            std::ofstream ot;
            ot.precision(8);
            if (iter == 0) {
              ot.open(path.c_str());
              ot << "loglik\trate_n\trate_c\tGB\t";
              for (int s = 0; s < N; s++) ot << tree.names[s] << "_B\t";
              for (int s = 0; s < N; s++) ot << tree.names[s] << "\t";
              ot << std::endl;
            } else ot.open(path.c_str(), std::ios::app);
            const double* t5 = &tr[((size_t)i * 3 + m) * 5 * T];
            const int8_t* z = &tz[((size_t)i * 3 + m) * T * N];
            const int8_t* b = &tb[((size_t)i * 3 + m) * T * N];
            std::string buf;
            buf.reserve((size_t)T * (4 * 16 + 5 * (size_t)N + 2));
            for (int t = 0; t < T; t++) {
              for (const double v : {t5[(size_t)T + t], t5[(size_t)3 * T + t], t5[(size_t)2 * T + t], t5[(size_t)4 * T + t]}) {
                fmt_g(buf, v, 8);
                buf.push_back('\t');
              }
              for (int s = 0; s < N; s++) { fmt_i(buf, b[(size_t)t * N + s]); buf.push_back('\t'); }
              for (int s = 0; s < N; s++) { fmt_i(buf, z[(size_t)t * N + s]); buf.push_back('\t'); }
              buf.push_back('\n');
            }
            ot.write(buf.data(), (std::streamsize)buf.size());
          }
        }
      }
      return (int)PA_OK;
    });
    if (!ok) {
      destroy_all();
      return 1;
    }
    // the reference appends an element's rows as it finishes it: M0 (file 0), M1 (file 2), M2 (file 1), elements in id order
    for (int i = 0; i < n; i++) {
      if (filtered[i]) continue;
      for (int m = 0; m < 3; m++) out_Z[kFileOfModel[m]] << rows[(size_t)i * 3 + m];
    }
    for (int r = 0; r < 3; r++) out_Z[r].flush();
    // BPP::Output_init (GBGC/SRC/bpp.cpp:536-574): every element of the BED file, zeros where nothing was computed
    {
      std::ofstream ol((prefix + "_elem_lik.txt").c_str());
      ol.precision(8);
      ol << "No.\tID\tloglik_NUll\tloglik_RES\tloglik_all\tlog_ratio\tloglik_Max1\tloglik_Max2\tloglik_Max3" << std::endl;
      for (int c = 0; c < C; c++) {
        ol << c << "\t" << in.element_names[c] << "\t" << lik_null[c] << "\t" << lik_res[c] << "\t" << lik_all[c] << "\t" << lik_res[c] - lik_null[c];
        for (int r = 0; r < 3; r++) ol << "\t" << lik_max[(size_t)r * C + c];
        ol << std::endl;
      }
      std::ofstream oz((prefix + "_1_elem_Z.txt").c_str());
      for (int s = 0; s < N; s++) oz << tree.names[s] << "_B\t";
      for (int s = 0; s < N; s++) oz << tree.names[s] << "\t";
      oz << std::endl;
      for (int c = 0; c < C; c++) {
        for (int s = 0; s < N; s++) oz << maxZ_full[(size_t)c * N + s] << "\t";
        oz << std::endl;
      }
    }
  }
  for (int r = 0; r < 3; r++) out_Z[r].close();
  destroy_all();
  std::cout << std::endl << std::endl << "time used:  " << (time(NULL) - start) / 60 << " min." << std::endl << std::endl;
  return 0;
}

}  // namespace
}  // namespace pa

extern "C" int pa_gbgc_run_from_config(const char* cfg_path) {
  if (!cfg_path) return 1;
  try {
    return pa::run_gbgc_from_config(cfg_path);
  } catch (const std::exception& ex) {
    std::cerr << "(Error. " << ex.what() << ")" << std::endl;
    return 1;
  }
}
