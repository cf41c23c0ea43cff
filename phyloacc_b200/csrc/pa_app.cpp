// pa_app.cpp -- file-level front door: the reference's main() (main.cpp:324-528) on top of the engine.
#include <dirent.h>

#include <cmath>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <functional>
#include <sstream>
#include <thread>

#include <cuda_runtime.h>

#include "pa_app.h"
#include "pa_fmt.h"
#include "pa_engine.h"
#include "pa_nvtx.h"
#include "pa_internal.h"

namespace pa {
void set_last_error(const std::string& s);

bool directory_exists(const std::string& p) {  // main.cpp:291-307
  if (p == "") return false;
  DIR* d = opendir(p.c_str());
  if (d) {
    closedir(d);
    return true;
  }
  return false;
}

static int load_config_inputs(const char* cfg_path, ConfigInputs& in) {
  std::cout << "Loading program configurations from " << cfg_path << "......" << std::endl;
  if (!load_params(cfg_path, in.P)) return PA_ERR_ARG;
  return load_inputs(in);
}

// LoadPhyloProfiles + LoadPhyloTree + id selection for parameters already in in.P (shared by the ST and GBGC front doors)
// PA_TIMING=1: wall time of the host stages on stderr (where the time of a genome-scale batch goes outside the kernels)
struct StageTimer {
  const bool on = getenv("PA_TIMING") != nullptr;
  double t0 = now();
  static double now() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
  }
  void mark(const char* what) {
    if (!on) return;
    const double t = now();
    std::cerr << "[pa timing] " << what << ": " << (t - t0) << " s" << std::endl;
    t0 = t;
  }
};

int load_inputs(ConfigInputs& in) {
  Params& P = in.P;
  Alignment aln;
  StageTimer tm;
  if (!load_alignment_files(P.align_path, P.segment_path, aln)) return PA_ERR_IO;
  tm.mark("FASTA + BED");
  std::cout << "Loading phylogenetic tree from " << P.phytree_path << "......" << std::endl;
  if (!load_tree(P.phytree_path, in.tree)) return PA_ERR_IO;
  if (!match_alignment_to_tree(in.tree, aln)) {
    set_error("alignment has fewer species than the tree");
    return PA_ERR_ARG;
  }
  tm.mark("tree + row matching");
  const Tree& tree = in.tree;
  in.og = species_ids(tree, P.outgroup);
  in.cg = species_ids(tree, P.conservegroup);
  in.element_names = aln.element_names;
  const int C = in.C = (int)aln.element_names.size();
  if (!load_ids(P, C, in.ids)) return PA_ERR_IO;
  double mean_len = 0;
  for (int c = 0; c < C; c++) mean_len += (double)aln.elem_len[c] / std::max(C, 1);
  std::cout << "# total length = " << aln.G << " (" << C << "). # Species = " << tree.S << ". # elements = " << C
            << ". Mean gene set size = " << mean_len << "." << std::endl;
  std::cout << "# Burn-ins = " << P.num.burnin << ". # MCMC Updates = " << P.num.mcmc << ". # adaptive frequency = " << P.num.adapt_freq
            << ".  RND SEED = " << P.num.seed << "." << std::endl;
  // rows -> one [S][G] byte matrix
  in.G = aln.G;
  in.X.resize((size_t)tree.S * aln.G);
  for (int s = 0; s < tree.S; s++) memcpy(in.X.data() + (size_t)s * aln.G, aln.X[s].data(), aln.G);
  std::vector<std::string>().swap(aln.X);
  // elements running past the alignment are clipped like std::string::substr does in the reference (bpp_c.hpp:211)
  for (int c = 0; c < C; c++) {
    if (aln.elem_start[c] > aln.G) aln.elem_start[c] = aln.G;
    if (aln.elem_start[c] + aln.elem_len[c] > aln.G) aln.elem_len[c] = (int32_t)(aln.G - aln.elem_start[c]);
    if (aln.elem_len[c] < 0) aln.elem_len[c] = 0;
  }
  in.elem_start = aln.elem_start;
  in.elem_len = aln.elem_len;
  tm.mark("alignment matrix");
  return PA_OK;
}

// BPP::BPP on one device (bpp.hpp:173-404): an engine handle with the batch model; the alignment is loaded by the caller
int make_engine(const ConfigInputs& in, int device, int num_thread, pa_engine** out) {
  pa_engine* e = new pa_engine();
  e->from_config = true;
  e->params = in.P;
  pa_params num = in.P.num;
  num.n_outgroup = (int32_t)in.og.size();
  num.outgroup = in.og.data();
  num.n_conserve = (int32_t)in.cg.size();
  num.conserve = in.cg.data();
  if (num_thread > 0) num.num_thread = num_thread;
  pa_tree pt{in.tree.S, in.tree.N, in.tree.parent.data(), in.tree.children.data(), in.tree.dist.data()};
  pa_subst ps;
  memcpy(ps.pi, in.tree.pi, sizeof(ps.pi));
  memcpy(ps.Q, in.tree.Q, sizeof(ps.Q));
  const int rc = e->eng.init(&pt, &ps, &num, device);
  if (rc != PA_OK) {
    delete e;
    return rc;
  }
  e->eng.model.tree.names = in.tree.names;
  e->element_names = in.element_names;
  e->ids.assign(in.ids.begin(), in.ids.end());
  *out = e;
  return PA_OK;
}

static int create_from_config(const char* cfg_path, int device, pa_engine** out) {
  ConfigInputs in;
  int rc = load_config_inputs(cfg_path, in);
  if (rc != PA_OK) return rc;
  pa_engine* e = nullptr;
  StageTimer tm;
  rc = make_engine(in, device, 0, &e);
  if (rc != PA_OK) return rc;
  tm.mark("engine");
  rc = e->eng.load_alignment(in.X.data(), in.G, in.C, in.elem_start.data(), in.elem_len.data());
  tm.mark("integer front-end + upload");
  if (rc != PA_OK) {
    delete e;
    return rc;
  }
  *out = e;
  return PA_OK;
}

// GPUs of this process the run is spread over: all visible ones unless PA_DEVICES=<count> says fewer or PA_DEVICE_LIST=<a,b,...> names
// them (an ordinal may repeat: several shards on one GPU, used by the tests), and never more than one shard per
// PA_MIN_ELEMS_PER_DEVICE (default 128) elements -- the driver's 50-locus batches stay on one GPU, genome-scale id files use the box
std::vector<int> pick_devices(int n_elems) {
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count < 1) return {0};  // the engine reports the missing device itself
  std::vector<int> devs;
  if (const char* env = getenv("PA_DEVICE_LIST")) {
    std::stringstream ss(env);
    std::string tok;
    while (std::getline(ss, tok, ','))
      if (!tok.empty()) devs.push_back(std::max(0, std::min(count - 1, atoi(tok.c_str()))));
  }
  if (devs.empty()) {
    if (const char* env = getenv("PA_DEVICES")) count = std::max(1, std::min(count, atoi(env)));
    for (int d = 0; d < count; d++) devs.push_back(d);
  }
  int per = 128;
  if (const char* env = getenv("PA_MIN_ELEMS_PER_DEVICE")) per = std::max(1, atoi(env));
  const int use = std::max(1, std::min((int)devs.size(), n_elems / per));
  devs.resize(use);
  return devs;
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
  ConfigInputs in;
  int rc = load_config_inputs(cfg_path, in);
  if (rc != PA_OK) {
    std::cerr << "(Error. " << get_error() << ")" << std::endl;
    return 1;
  }
  StageTimer tm;
  const Params& P = in.P;
  const Tree& tree = in.tree;
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
  std::cout << in.ids.size() << " of elements to be computed" << std::endl;

  const int C = in.C;
  std::vector<int32_t> ids;
  for (int32_t c : in.ids) {
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
  // ---- shards: the ids are cut into contiguous ranges of equal cost (columns), one engine + host thread per GPU; no traffic between
  // GPUs while sampling, the per-element summaries meet in the host arrays above.  Every engine holds only its shard;
  // chains are keyed by the batch-wide element number, so the files do not depend on the number of GPUs.
  const std::vector<int> devices = pick_devices(n);
  const int n_dev = (int)devices.size();
  std::vector<int> shard_lo(n_dev + 1, 0);
  {
    double total = 0;
    for (int i = 0; i < n; i++) total += std::max(1, in.elem_len[ids[i]]);
    double acc = 0;
    int d = 1;
    for (int i = 0; i < n && d < n_dev; i++) {
      acc += std::max(1, in.elem_len[ids[i]]);
      while (d < n_dev && acc >= total * d / n_dev) shard_lo[d++] = i + 1;
    }
    for (; d <= n_dev; d++) shard_lo[d] = n;
    shard_lo[n_dev] = n;
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
  const int threads_per_dev = std::max(1, P.num.num_thread / n_dev);
  if (n_dev > 1) std::cout << "# element shards = " << n_dev << " (one engine and host thread per shard)." << std::endl;
  if (!for_each_device([&](int d) {
        PA_NVTX_RANGE("load shard");
        StageTimer stt;
        int r = make_engine(in, devices[d], threads_per_dev, &engines[d]);
        if (r != PA_OK) return r;
        stt.mark(("shard " + std::to_string(d) + ": engine + CUDA context").c_str());
        const int lo = shard_lo[d], cnt = shard_lo[d + 1] - lo;
        std::vector<int64_t> st(cnt);
        std::vector<int32_t> ln(cnt), gid(cnt);
        for (int k = 0; k < cnt; k++) {
          st[k] = in.elem_start[ids[lo + k]];
          ln[k] = in.elem_len[ids[lo + k]];
          gid[k] = ids[lo + k];
        }
        r = engines[d]->eng.load_alignment(in.X.data(), in.G, cnt, st.data(), ln.data(), gid.data());
        stt.mark(("shard " + std::to_string(d) + ": integer front-end + upload").c_str());
        return r;
      })) {
    destroy_all();
    return 1;
  }
  tm.mark("file headers, engines (CUDA contexts), integer front-end + upload of every shard");
  for (int iter = 0; iter < P.num_chain; iter++) {
    std::cout << "Running MCMC chain " << iter + 1 << " ..." << std::endl;
    const bool ok = for_each_device([&](int d) {
      Engine& eng = engines[d]->eng;
      const int lo = shard_lo[d], cnt = shard_lo[d + 1] - lo;
      // elements go through the device in waves so that the traces of one wave fit comfortably in HBM
      const int wave = 8192;
      std::vector<int32_t> local(std::min(wave, std::max(cnt, 1)));
      for (int w0 = 0; w0 < cnt; w0 += wave) {
        const int wn = std::min(wave, cnt - w0);
        for (int k = 0; k < wn; k++) local[k] = w0 + k;
        const size_t g0 = (size_t)lo + w0;
        pa_results wres{wn, N, filtered.data() + g0, loglik.data() + g0 * 3, loglik_max.data() + g0 * 3, maxZ.data() + g0 * 3 * N, med.data() + g0 * 15,
                        postZ.data() + g0 * 3 * N * 4};
        // SAMPLE_HYPER 1 runs only the full model (main.cpp:463-494); outer iterations after the first continue every chain
        // from its last state (initMCMC with iter > 0, bpp_c.cpp:148-177).  The reference saves that state only `if(UpHyper)`
        // (bpp_c.cpp:474-485): with SAMPLE_HYPER 0 and CHAIN > 1 its later iterations restart from bpp.cur_Z (all zeros: every node
        // starts in the background state) and the INIT_* rates -- a fresh chain on a new random stream here (Z = 1 start as in iter 0).
        int r;
        {
          PA_NVTX_RANGE("sample wave");
          r = eng.sample(local.data(), wn, P.sample_hyper ? (1 << PA_M2) : 7, P.sample_hyper && iter > 0, iter);
        }
        if (r == PA_OK) {
          PA_NVTX_RANGE("collect wave");
          r = eng.collect(&wres);
        }
        if (r != PA_OK) return r;
        // VERBOSE: per-chain traces (Output_sampling, bpp_c2.cpp:706-736); file index is the reference's third argument 0,1,2
        if (P.verbose) {
          PA_NVTX_RANGE("write traces");
          for (int i = 0; i < wn; i++) {
            if (filtered[g0 + i]) continue;
            for (int m = 0; m < 3; m++) {
              ChainTraceHost tr;
              if (eng.get_trace(i, m, tr) != PA_OK) continue;
              const std::string path = prefix + "_mcmc_trace_M" + std::to_string(m) + "_" + std::to_string(ids[g0 + i]) + ".txt";
              std::ofstream ot;
              ot.precision(8);
              if (iter == 0) {
                ot.open(path.c_str());
                ot << "loglik\trate_n\trate_c\tgrate\tlrate\tlrate2\t";
                for (int s = 0; s < N; s++) ot << tree.names[s] << "\t";
                ot << std::endl;
              } else ot.open(path.c_str(), std::ios::app);
              const int T = (int)tr.loglik.size();
              std::string buf;
              buf.reserve((size_t)T * (6 * 16 + 3 * N + 2));
              for (int t = 0; t < T; t++) {
                for (const double v : {tr.loglik[t], tr.n_rate[t], tr.c_rate[t], tr.g_rate[t], tr.l_rate[t], tr.l2_rate[t]}) {
                  fmt_g(buf, v, 8);
                  buf.push_back('\t');
                }
                for (int s = 0; s < N; s++) {
                  fmt_i(buf, tr.Z[(size_t)t * N + s]);
                  buf.push_back('\t');
                }
                buf.push_back('\n');
              }
              ot.write(buf.data(), (std::streamsize)buf.size());
            }
          }
        }
      }
      return (int)PA_OK;
    });
    if (!ok) {
      destroy_all();
      return 1;
    }
    tm.mark("chains + summaries + collect of every wave");
    // rows of _rate_postZ_Mk.txt (Output_init, bpp_c2.cpp:776-783; default ofstream precision = 6 significant digits)
    {
      PA_NVTX_RANGE("write rate_postZ");
      const int n_chunks = std::max(1, std::min(n, 4 * std::max(1, P.num.num_thread)));
      for (int m = 0; m < 3; m++) {
        if (P.sample_hyper && m != PA_M2) continue;
        std::vector<std::string> chunk(n_chunks);  // rows of contiguous element ranges, formatted by all host threads
#pragma omp parallel for schedule(dynamic) num_threads(std::max(1, P.num.num_thread))
        for (int ch = 0; ch < n_chunks; ch++) {
          std::string& b = chunk[ch];
          const int lo = (int)((long long)n * ch / n_chunks), hi = (int)((long long)n * (ch + 1) / n_chunks);
          b.reserve((size_t)(hi - lo) * (64 + (size_t)N * 4 * 6));
          for (int i = lo; i < hi; i++) {
            if (filtered[i]) continue;
            const double* md = &med[((size_t)i * 3 + m) * 5];
            fmt_i(b, ids[i]);
            for (int k = 0; k < 5; k++) { b.push_back('\t'); fmt_g(b, md[k]); }
            const double* pz = &postZ[(((size_t)i * 3 + m) * N) * 4];
            for (int k = 0; k < 4 * N; k++) { b.push_back('\t'); fmt_g(b, pz[k]); }
            b.push_back('\n');
          }
        }
        for (const std::string& b : chunk) out_Z[m].write(b.data(), (std::streamsize)b.size());
        out_Z[m].flush();
      }
    }
    if (P.sample_hyper) {
      // BPP::sample_hyperparam + BPP::Output_init0 (main.cpp:504-506, bpp.cpp:324-477, 622-635)
      // the sweep needs the elements only through PA_HYPER_NSUMS sums: every shard adds its own, every engine draws the same sweep
      pa_hyper hy;
      double sums[PA_HYPER_NSUMS] = {0};
      rc = PA_OK;
      for (int d = 0; d < n_dev && rc == PA_OK; d++) {
        const int cnt = shard_lo[d + 1] - shard_lo[d];
        std::vector<int32_t> local(cnt);
        for (int k = 0; k < cnt; k++) local[k] = k;
        double part[PA_HYPER_NSUMS];
        rc = engines[d]->eng.hyper_sums(local.data(), cnt, part);
        for (int k = 0; k < PA_HYPER_NSUMS; k++) sums[k] += part[k];
      }
      for (int d = 0; d < n_dev && rc == PA_OK; d++) rc = engines[d]->eng.sample_hyper_from_sums(iter, sums, &hy);
      if (rc != PA_OK) {
        std::cout << " 2 Standard exception: " << get_error() << std::endl;
      } else {
        out_hyper << iter << "\t" << hy.nprior_a << "\t" << hy.nprior_b << "\t" << hy.cprior_a << "\t" << hy.cprior_b << "\t" << hy.prior_l_a << "\t"
                  << hy.prior_l_b << "\t" << hy.prior_g_a << "\t" << hy.prior_g_b << "\t" << hy.prior_l2_a << "\t" << hy.prior_l2_b << std::endl;
      }
      for (int i = 0; i < n; i++)
        out_lik_hyper << ids[i] << "\t" << in.element_names[ids[i]] << "\t" << loglik[i * 3 + PA_M2] << "\t" << loglik_max[i * 3 + PA_M2] << std::endl;
      continue;
    }
    // _elem_lik.txt and _M{k}_elem_Z.txt (BPP::Output_init, bpp.cpp:568-619): one row per id, zeros for filtered ones
    {
      std::ofstream ol((prefix + "_elem_lik.txt").c_str());
      ol.precision(8);
      ol << "No.\tID\tloglik_Null\tloglik_Acc\tloglik_Full\tlogBF1\tlogBF2\tloglik_Max_M0\tloglik_Max_M1\tloglik_Max_M2" << std::endl;
      for (int i = 0; i < n; i++) {
        const double l0 = loglik[i * 3 + 0], l1 = loglik[i * 3 + 1], l2 = loglik[i * 3 + 2];
        ol << ids[i] << "\t" << in.element_names[ids[i]] << "\t" << l0 << "\t" << l1 << "\t" << l2 << "\t";
        ol << l1 - l0 << "\t" << l1 - l2;
        ol << "\t" << loglik_max[i * 3 + 0] << "\t" << loglik_max[i * 3 + 1] << "\t" << loglik_max[i * 3 + 2];
        ol << std::endl;
      }
      for (int m = 0; m < 3; m++) {
        std::ofstream oz((prefix + "_M" + std::to_string(m) + "_elem_Z.txt").c_str());
        oz << "No.";
        for (int s = 0; s < N; s++) oz << "\t" << tree.names[s];
        oz << std::endl;
        std::string b;
        b.reserve((size_t)n * (12 + 3 * (size_t)N));
        for (int i = 0; i < n; i++) {
          fmt_i(b, ids[i]);
          for (int s = 0; s < N; s++) { b.push_back('\t'); fmt_i(b, maxZ[((size_t)i * 3 + m) * N + s]); }
          b.push_back('\n');
        }
        oz.write(b.data(), (std::streamsize)b.size());
      }
    }
  }
  for (int m = 0; m < 3; m++) out_Z[m].close();
  tm.mark("result files");
  destroy_all();
  tm.mark("engines released");
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
