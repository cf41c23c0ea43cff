// pa_host.cpp -- host side of the engine: input parsers with the reference's quirks, the batch model
// (BPP), and the integer front-end of every element (BPP_C::BPP_C) including site-pattern compression.
// Reference citations are relative to /root/reference/src/PhyloAcc-ST/.
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <numeric>
#include <sstream>

#include "pa_internal.h"

namespace pa {

static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }
const std::string& get_error() { return g_error; }

// ------------------------------------------------------------------------------------------ string helpers
// strutils::trim / split (utils.h:335, utils.cpp:85-100): empty tokens are kept.
static std::string trim(std::string s, const char* ts = " \t\n") {
  s.erase(0, s.find_first_not_of(ts));
  s.erase(s.find_last_not_of(ts) + 1);
  return s;
}
static std::vector<std::string> split(const std::string& s, char delim) {
  std::vector<std::string> out;
  std::stringstream ss(s);
  std::string item;
  while (std::getline(ss, item, delim)) out.push_back(item);
  return out;
}

// ------------------------------------------------------------------------------------------ parameters
static bool to_bool(const std::string& s) { return s == "true" || s == "1"; }  // main.cpp:95-97

bool load_params(const std::string& path, Params& p) {  // LoadParams, main.cpp:100-287
  std::ifstream in(path.c_str());
  if (!in) {
    set_error("Cannot open the parameters file: " + path);
    return false;
  }
  std::string line;
  int line_num = 1;
  pa_params& n = p.num;
  while (std::getline(in, line)) {
    std::istringstream ls(line);
    std::vector<std::string> w;
    std::string word;
    while (ls >> word) w.push_back(word);
    if (w.empty() || w[0][0] == '#') continue;  // blank lines and comments
    if (w.size() != 2) {
      std::ostringstream os;
      os << "Line " << line_num << " in the parameter file is not formatted correctly: " << line;
      set_error(os.str());
      return false;
    }
    const std::string &k = w[0], &v = w[1];
    try {
      if (k == "PHYTREE_FILE") p.phytree_path = v;
      else if (k == "ALIGN_FILE") p.align_path = v;
      else if (k == "SEG_FILE") p.segment_path = v;
      else if (k == "ID_FILE") p.id_path = v;
      else if (k == "BATCH") p.batch = std::stoi(v);
      else if (k == "RESULT_FOLDER") p.output_path = v;
      else if (k == "PREFIX") p.result_prefix = v;
      else if (k == "SEED") n.seed = std::stoi(v);
      else if (k == "INIT_CONSERVE_RATE") n.init_conserve_rate = std::stod(v);
      else if (k == "INIT_ACCE_RATE") n.init_acce_rate = std::stod(v);
      else if (k == "CONSERVE_PRIOR_A") n.conserve_prior_a = std::stod(v);
      else if (k == "CONSERVE_PRIOR_B") n.conserve_prior_b = std::stod(v);
      else if (k == "ACCE_PRIOR_A") n.acce_prior_a = std::stod(v);
      else if (k == "ACCE_PRIOR_B") n.acce_prior_b = std::stod(v);
      else if (k == "ROPT") n.ropt = std::stoi(v);
      else if (k == "CUB") n.cub = std::stod(v);
      else if (k == "NLB") n.nlb = std::stod(v);
      else if (k == "BURNIN") n.burnin = std::stoi(v);
      else if (k == "MCMC") n.mcmc = std::stoi(v);
      else if (k == "ADAPT_FREQ") n.adapt_freq = std::stoi(v);
      else if (k == "INIT_LRATE") n.init_lrate = std::stod(v);
      else if (k == "INIT_LRATE2") n.init_lrate2 = std::stod(v);
      else if (k == "INIT_GRATE") n.init_grate = std::stod(v);
      else if (k == "HYPER_LRATE_A") n.hyper_lrate_a = std::stod(v);
      else if (k == "HYPER_LRATE_B") n.hyper_lrate_b = std::stod(v);
      else if (k == "HYPER_GRATE_A") n.hyper_grate_a = std::stod(v);
      else if (k == "HYPER_GRATE_B") n.hyper_grate_b = std::stod(v);
      else if (k == "HYPER_LRATE2_A") n.hyper_lrate2_a = std::stod(v);
      else if (k == "HYPER_LRATE2_B") n.hyper_lrate2_b = std::stod(v);
      else if (k == "CHAIN") p.num_chain = std::stoi(v);
      else if (k == "OUTGROUP") p.outgroup = v;
      else if (k == "TARGETSPECIES") p.targetspecies = v;
      else if (k == "CONSERVE") p.conservegroup = v;
      else if (k == "CONSERVE_PROP") n.conserve_prop = std::stod(v);
      else if (k == "GAP_PROP") n.gap_prop = std::stod(v);
      else if (k == "CONSTOMIS") n.constomis = std::stod(v);
      else if (k == "GAPCHAR") n.gapchar = v[0];
      else if (k == "PRUNE_TREE") n.prune_tree = to_bool(v);
      else if (k == "TRIM_GAP_PERCENT") n.trim_gap_percent = std::stod(v);
      else if (k == "MIN_LEN") n.min_len = std::stoi(v);
      else if (k == "INDEL" || k == "INDEL2" || k == "SAMPLE_INDEL") { /* parsed but unused by ST */ }
      else if (k == "SAMPLE_HYPER") p.sample_hyper = to_bool(v);
      else if (k == "VERBOSE") p.verbose = to_bool(v);
      else if (k == "NUM_THREAD") n.num_thread = std::stoi(v);
      else std::cout << "Unknown parameter: " << k << ", skipping..." << std::endl;
    } catch (const std::exception&) {
      set_error("Bad value for parameter " + k + ": " + v);
      return false;
    }
    line_num++;
  }
  p.phytree_path = trim(p.phytree_path, " \"\t\n");
  p.align_path = trim(p.align_path, " \"\t\n");
  p.output_path = trim(p.output_path, " \"\t\n");
  p.segment_path = trim(p.segment_path, " \"\t\n");
  if (n.hyper_lrate2_a == 0) n.init_lrate2 = 0;  // main.cpp:286
  return true;
}

// ------------------------------------------------------------------------------------------ .mod / Newick
namespace {
struct NwkNode {
  std::vector<int> kids;
  std::string label;
  double dist = 0.0;
  int id = -1;
};

struct NwkParser {
  const std::string& s;
  size_t pos = 0;
  std::vector<NwkNode> nodes;
  explicit NwkParser(const std::string& str) : s(str) {}

  int parse_node() {
    const int me = (int)nodes.size();
    nodes.emplace_back();
    if (pos < s.size() && s[pos] == '(') {
      pos++;
      while (true) {
        const int k = parse_node();
        nodes[me].kids.push_back(k);
        if (pos < s.size() && s[pos] == ',') { pos++; continue; }
        if (pos < s.size() && s[pos] == ')') { pos++; break; }
        throw std::runtime_error("malformed Newick string");
      }
    }
    const size_t st = pos;
    while (pos < s.size() && s[pos] != ',' && s[pos] != ')' && s[pos] != ':' && s[pos] != ';') pos++;
    nodes[me].label = s.substr(st, pos - st);
    if (pos < s.size() && s[pos] == ':') {
      pos++;
      const size_t d0 = pos;
      while (pos < s.size() && s[pos] != ',' && s[pos] != ')' && s[pos] != ';') pos++;
      // the reference stores distances in a float: (float)atof (newick.cpp:185,300,325)
      nodes[me].dist = (double)(float)atof(s.substr(d0, pos - d0).c_str());
    }
    return me;
  }
};
}  // namespace

bool load_tree(const std::string& path, Tree& t) {  // LoadPhyloTree (newick.cpp:431-517) + InitPhyloTree (bpp.cpp:84-117)
  std::ifstream in(path.c_str());
  if (!in) {
    set_error("Cannot open the phylogenetic tree input file: " + path);
    return false;
  }
  std::string line, newick;
  while (std::getline(in, line)) {
    if (line == "") break;  // newick.cpp:457
    line = trim(line);
    std::vector<std::string> sp = split(line, ' ');
    if (sp.empty()) continue;
    if (sp[0] == "BACKGROUND:") {
      int ind = 0;
      for (size_t g = 1; g < sp.size(); g++) {
        if (sp[g] == "") continue;
        if (ind < 4) t.pi[ind] = atof(sp[g].c_str());
        ind++;
      }
    } else if (sp[0] == "RATE_MAT:") {
      for (int i = 0; i < 4; i++) {
        std::getline(in, line);
        std::vector<std::string> tmp = split(trim(line), ' ');
        int ind = 0;
        for (size_t g = 0; g < tmp.size(); g++) {
          if (tmp[g] == "") continue;
          if (ind < 4) t.Q[i * 4 + ind] = atof(tmp[g].c_str());
          ind++;
        }
      }
    } else if (sp[0] == "TREE:") {
      if (sp.size() > 1) newick = trim(sp[1]);
    }
  }
  if (newick.empty()) {
    set_error("No TREE: line in " + path);
    return false;
  }
  NwkParser np(newick);
  int root;
  try {
    root = np.parse_node();
  } catch (const std::exception& e) {
    set_error(std::string("Cannot parse tree: ") + e.what());
    return false;
  }
  int S = 0;
  for (auto& nd : np.nodes)
    if (nd.kids.empty()) S++;
  const int N = 2 * S - 1;
  if ((int)np.nodes.size() != N) {
    set_error("The tree must be strictly binary");
    return false;
  }
  // TravelTree2 (newick.cpp:382-400): leaves in DFS order, internal nodes in post-order
  int leaf = -1, inner = -1;
  std::function<void(int)> number = [&](int v) {
    for (int k : np.nodes[v].kids) number(k);
    if (np.nodes[v].kids.empty()) np.nodes[v].id = ++leaf;
    else np.nodes[v].id = ++inner + S;
  };
  number(root);
  t.S = S;
  t.N = N;
  t.parent.assign(N, -1);
  t.children.assign(2 * N, -1);
  t.dist.assign(N, 0.0);
  t.names.assign(N, "");
  for (auto& nd : np.nodes) {
    if (nd.label.empty()) {  // newick.cpp:416-419
      set_error("Error. Some nodes in the tree aren't labeled. Please label all nodes and try again!");
      return false;
    }
    t.names[nd.id] = nd.label;
    t.dist[nd.id] = nd.dist;
    if (!nd.kids.empty()) {
      if (nd.kids.size() != 2) {
        set_error("The tree must be strictly binary");
        return false;
      }
      int a = np.nodes[nd.kids[0]].id, b = np.nodes[nd.kids[1]].id;
      if (a > b) std::swap(a, b);  // children ordered by node id (dag row scan, bpp.cpp:105-117)
      t.children[2 * nd.id] = a;
      t.children[2 * nd.id + 1] = b;
      t.parent[a] = nd.id;
      t.parent[b] = nd.id;
    }
  }
  return true;
}

// ------------------------------------------------------------------------------------------ FASTA / BED / ids
bool load_alignment_files(const std::string& fasta, const std::string& bed, Alignment& a) {  // profile.cpp:17-126
  // The whole file in one read, one pass over its lines without per-line string objects (the getline / trim / append loop of the reference
  // took 1.6 s for the 146 MB of a 16384-element batch): same rules -- every line trimmed of " \t\n" at both ends, a trimmed line that
  // starts with '>' opens a record (name = the rest, trimmed), any other line is appended to the current record, lower-cased.
  std::string buf;
  {
    FILE* f = fopen(fasta.c_str(), "rb");
    if (!f) {
      set_error("Cannot open the phylogenetic profile input file: " + fasta);
      return false;
    }
    fseek(f, 0, SEEK_END);
    const long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    buf.resize(sz > 0 ? (size_t)sz : 0);
    const size_t got = buf.empty() ? 0 : fread(&buf[0], 1, buf.size(), f);
    fclose(f);
    buf.resize(got);
  }
  unsigned char lower[256];
  for (int c = 0; c < 256; c++) lower[c] = (unsigned char)((c >= 'A' && c <= 'Z') ? c - 'A' + 'a' : c);
  auto is_ws = [](char c) { return c == ' ' || c == '\t' || c == '\n'; };
  std::string whole;
  bool first = true;
  const char* base = buf.data();
  const size_t size = buf.size();
  size_t p = 0;
  while (p < size) {
    const char* nl = (const char*)memchr(base + p, '\n', size - p);
    size_t e = nl ? (size_t)(nl - base) : size;
    const size_t next = e + 1;
    size_t b0 = p;
    while (b0 < e && is_ws(base[b0])) b0++;
    while (e > b0 && is_ws(base[e - 1])) e--;
    if (e > b0 && base[b0] == '>') {
      a.species_names.push_back(trim(std::string(base + b0 + 1, e - b0 - 1)));
      if (!first) {
        const size_t len = whole.size();
        a.X.push_back(std::move(whole));
        whole.clear();
        whole.reserve(len);
      }
      first = false;
    } else if (e > b0) {
      const size_t old = whole.size();
      whole.resize(old + (e - b0));
      unsigned char* dst = reinterpret_cast<unsigned char*>(&whole[old]);
      const unsigned char* src = reinterpret_cast<const unsigned char*>(base + b0);
      for (size_t k = 0; k < e - b0; k++) dst[k] = lower[src[k]];
    }
    p = next;
  }
  a.X.push_back(std::move(whole));
  std::string().swap(buf);
  std::string line;
  a.S = (int)a.species_names.size();
  a.G = a.X.empty() ? 0 : (int64_t)a.X[0].size();
  for (auto& row : a.X)
    if ((int64_t)row.size() != a.G) {
      set_error("alignment rows differ in length");
      return false;
    }
  std::ifstream seg(bed.c_str());
  if (!seg) {
    set_error("Cannot open the segment input file: " + bed);
    return false;
  }
  int64_t start = 0;
  while (std::getline(seg, line)) {
    line = trim(line);
    std::vector<std::string> sp = split(line, '\t');
    if (sp.size() < 3) break;  // profile.cpp:84
    a.element_names.push_back(sp[0]);
    const int64_t s0 = atoi(sp[1].c_str()), s1 = atoi(sp[2].c_str());
    // bpp.hpp:257-274: elements are laid out back to back in file order, only end-start matters
    a.elem_start.push_back(start);
    a.elem_len.push_back((int32_t)(s1 - s0));
    start += (s1 - s0);
  }
  return true;
}

bool load_ids(const Params& p, int C, std::vector<int>& ids) {  // main.cpp:406-439
  ids.clear();
  if (p.id_path == "") {
    if (p.batch == -1) {
      for (int c = 0; c < C; c++) ids.push_back(c);
    } else {
      const int temp = (int)std::ceil(C / 3);
      for (int c = p.batch * temp; c < (p.batch + 1) * temp; c++) {
        if (c >= C) break;
        ids.push_back(c);
      }
    }
    return true;
  }
  std::ifstream in(p.id_path.c_str());
  if (!in) {
    set_error("Cannot open the id file: " + p.id_path);
    return false;
  }
  std::string line;
  while (std::getline(in, line)) {
    std::istringstream ls(line);
    std::string tmp;
    ls >> tmp;
    tmp = trim(tmp);
    if (tmp == "") continue;
    ids.push_back(atoi(tmp.c_str()));
  }
  return true;
}

bool match_alignment_to_tree(const Tree& t, Alignment& a) {  // MatchProfAndTree (bpp.cpp:133-184)
  std::vector<int> reorder(t.S);
  for (int s1 = 0; s1 < t.S; s1++) {
    int found = -1;
    for (int s2 = 0; s2 < a.S; s2++)
      if (t.names[s1] == a.species_names[s2]) { found = s2; break; }
    if (found < 0) {
      std::cout << "The species in phylogenetic profile and tree cannot be matched literally; using file order." << std::endl;
      return a.S >= t.S;
    }
    reorder[s1] = found;
  }
  std::vector<std::string> old = a.X;
  a.X.resize(t.S);
  for (int s = 0; s < t.S; s++) a.X[s] = old[reorder[s]];
  a.species_names.assign(t.names.begin(), t.names.begin() + t.S);
  a.S = t.S;
  return true;
}

std::vector<int32_t> species_ids(const Tree& t, const std::string& list) {  // bpp.hpp:230-242,303-308
  std::vector<int32_t> out;
  const std::string tl = trim(list);
  if (tl.empty()) return out;
  for (const std::string& nm : split(tl, ';')) {
    int pos = t.S;  // std::find(...) == end() -> index S
    for (int s = 0; s < t.S; s++)
      if (t.names[s] == nm) { pos = s; break; }
    out.push_back(pos);
  }
  return out;
}

// ------------------------------------------------------------------------------------------ eigensystem of Q
// Real 4x4 matrix with a real spectrum (reversible substitution models): shifted QR iteration for the
// eigenvalues, null spaces by complete-pivot elimination for the eigenvectors (handles repeated
// eigenvalues such as JC69).  Replaces arma::eig_gen + inv (bpp.hpp:357-362); only products of the form
// V f(Lambda) V^-1 are ever used, which do not depend on eigenvector scaling or order.
namespace {
void null_space4(const double M[16], double tol, std::vector<std::array<double, 4>>& basis) {
  double m[16];
  memcpy(m, M, sizeof(m));
  int colperm[4] = {0, 1, 2, 3};
  int rank = 0;
  for (int k = 0; k < 4; k++) {
    double best = 0;
    int pr = k, pc = k;
    for (int i = k; i < 4; i++)
      for (int j = k; j < 4; j++)
        if (std::fabs(m[i * 4 + j]) > best) { best = std::fabs(m[i * 4 + j]); pr = i; pc = j; }
    if (best <= tol) break;
    if (pr != k) for (int j = 0; j < 4; j++) std::swap(m[k * 4 + j], m[pr * 4 + j]);
    if (pc != k) {
      for (int i = 0; i < 4; i++) std::swap(m[i * 4 + k], m[i * 4 + pc]);
      std::swap(colperm[k], colperm[pc]);
    }
    const double d = m[k * 4 + k];
    for (int j = k; j < 4; j++) m[k * 4 + j] /= d;
    for (int i = 0; i < 4; i++)
      if (i != k) {
        const double f = m[i * 4 + k];
        if (f != 0.0) for (int j = k; j < 4; j++) m[i * 4 + j] -= f * m[k * 4 + j];
      }
    rank++;
  }
  basis.clear();
  for (int f = rank; f < 4; f++) {
    double y[4] = {0, 0, 0, 0};
    y[f] = 1.0;
    for (int i = 0; i < rank; i++) y[i] = -m[i * 4 + f];
    std::array<double, 4> x{};
    for (int i = 0; i < 4; i++) x[colperm[i]] = y[i];
    double nrm = 0;
    for (int i = 0; i < 4; i++) nrm += x[i] * x[i];
    nrm = std::sqrt(nrm);
    for (int i = 0; i < 4; i++) x[i] /= nrm;
    basis.push_back(x);
  }
}

bool invert4(const double A[16], double Ainv[16]) {
  double a[4][8];
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) { a[i][j] = A[i * 4 + j]; a[i][4 + j] = (i == j); }
  for (int c = 0; c < 4; c++) {
    int piv = c;
    for (int r = c + 1; r < 4; r++) if (std::fabs(a[r][c]) > std::fabs(a[piv][c])) piv = r;
    if (a[piv][c] == 0.0) return false;
    if (piv != c) for (int j = 0; j < 8; j++) std::swap(a[c][j], a[piv][j]);
    const double d = a[c][c];
    for (int j = 0; j < 8; j++) a[c][j] /= d;
    for (int r = 0; r < 4; r++)
      if (r != c) {
        const double f = a[r][c];
        if (f != 0.0) for (int j = 0; j < 8; j++) a[r][j] -= f * a[c][j];
      }
  }
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) Ainv[i * 4 + j] = a[i][4 + j];
  return true;
}
}  // namespace

bool eigen_real4(const double Q[16], double eval[4], double V[16], double Vinv[16]) {
  double a[16];
  memcpy(a, Q, sizeof(a));
  double anorm = 0;
  for (int i = 0; i < 16; i++) anorm += a[i] * a[i];
  anorm = std::sqrt(anorm);
  if (anorm == 0) return false;
  // Hessenberg reduction is unnecessary at n = 4: shifted QR on the full active block with Givens-free
  // modified Gram-Schmidt, deflating from the bottom.
  std::vector<double> ev;
  int m = 4, iter = 0;
  while (m > 0) {
    if (m == 1) { ev.push_back(a[0]); break; }
    double off = 0;
    for (int j = 0; j + 1 < m; j++) off += std::fabs(a[(m - 1) * 4 + j]);
    if (off <= 1e-15 * anorm || iter > 5000) {
      ev.push_back(a[(m - 1) * 4 + (m - 1)]);
      m--;
      iter = 0;
      continue;
    }
    const double p = a[(m - 2) * 4 + (m - 2)], q = a[(m - 2) * 4 + (m - 1)], r = a[(m - 1) * 4 + (m - 2)], s = a[(m - 1) * 4 + (m - 1)];
    const double tr = p + s, det = p * s - q * r, disc = tr * tr / 4 - det;
    double mu = s;
    if (disc >= 0) {
      const double sq = std::sqrt(disc), l1 = tr / 2 + sq, l2 = tr / 2 - sq;
      mu = (std::fabs(l1 - s) < std::fabs(l2 - s)) ? l1 : l2;
    }
    if (iter % 11 == 10) mu += 1e-3 * anorm * ((iter / 11) % 2 ? 1 : -1);
    for (int i = 0; i < m; i++) a[i * 4 + i] -= mu;
    double Qm[16] = {0}, R[16] = {0};
    for (int j = 0; j < m; j++) {
      double v[4];
      for (int i = 0; i < m; i++) v[i] = a[i * 4 + j];
      for (int k = 0; k < j; k++) {
        double dot = 0;
        for (int i = 0; i < m; i++) dot += Qm[i * 4 + k] * v[i];
        R[k * 4 + j] = dot;
        for (int i = 0; i < m; i++) v[i] -= dot * Qm[i * 4 + k];
      }
      double nv = 0;
      for (int i = 0; i < m; i++) nv += v[i] * v[i];
      nv = std::sqrt(nv);
      R[j * 4 + j] = nv;
      for (int i = 0; i < m; i++) Qm[i * 4 + j] = nv > 0 ? v[i] / nv : (i == j ? 1.0 : 0.0);
    }
    for (int i = 0; i < m; i++)
      for (int j = 0; j < m; j++) {
        double sum = 0;
        for (int k = 0; k < m; k++) sum += R[i * 4 + k] * Qm[k * 4 + j];
        a[i * 4 + j] = sum;
      }
    for (int i = 0; i < m; i++) a[i * 4 + i] += mu;
    iter++;
  }
  std::sort(ev.begin(), ev.end());
  double scale = 0;
  for (int i = 0; i < 16; i++) scale = std::max(scale, std::fabs(Q[i]));
  int k = 0;
  while (k < 4) {
    int k2 = k + 1;
    while (k2 < 4 && std::fabs(ev[k2] - ev[k]) <= 1e-8 * scale) k2++;
    double lam = 0;
    for (int i = k; i < k2; i++) lam += ev[i];
    lam /= (k2 - k);
    if (std::fabs(lam) < 1e-13 * scale) lam = 0.0;  // rate matrices: rows sum to zero -> exact zero eigenvalue
    double M[16];
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) M[i * 4 + j] = Q[i * 4 + j] - (i == j ? lam : 0.0);
    std::vector<std::array<double, 4>> basis;
    double tol = 1e-13 * scale;
    for (int attempt = 0; attempt < 8; attempt++) {
      null_space4(M, tol, basis);
      if ((int)basis.size() == k2 - k) break;
      tol *= 10;
    }
    if ((int)basis.size() != k2 - k) return false;  // defective or complex spectrum
    for (int b = 0; b < (int)basis.size(); b++) {
      eval[k + b] = lam;
      for (int i = 0; i < 4; i++) V[i * 4 + (k + b)] = basis[b][i];
    }
    k = k2;
  }
  return invert4(V, Vinv);
}

bool build_model(const Tree& t, const pa_params& prm, Model& m) {  // BPP::BPP (bpp.hpp:173-404)
  m.tree = t;
  m.prm = prm;
  m.outgroup.assign(prm.outgroup, prm.outgroup + prm.n_outgroup);
  m.conservedgroup.assign(prm.conserve, prm.conserve + prm.n_conserve);
  m.prm.outgroup = nullptr;
  m.prm.conserve = nullptr;
  const int N = t.N;
  auto upper_of = [&](const std::vector<int32_t>& grp, std::vector<uint8_t>& out) {  // getUppertree (bpp.cpp:486-504)
    out.assign(N, 0);
    for (int32_t p : grp) {
      if (p < 0 || p >= N) continue;  // unknown species name (index S) walks the tree in the reference; ignore here
      while (p != N - 1) {
        out[p] = 1;
        p = t.parent[p];
      }
    }
    out[N - 1] = 1;
  };
  upper_of(m.outgroup, m.upper);
  upper_of(m.conservedgroup, m.upper_conserve);
  if (!eigen_real4(t.Q, m.eval, m.V, m.Vinv)) {
    set_error("rate matrix is not diagonalisable with a real spectrum");
    return false;
  }
  for (int k = 0; k < 4; k++) {
    m.subst_dev.eval[k] = m.eval[k];
    m.subst_dev.pi[k] = t.pi[k];
    for (int j = 0; j < 4; j++)
      for (int i = 0; i < 4; i++) m.subst_dev.A[k][j * 4 + i] = m.V[j * 4 + k] * m.Vinv[k * 4 + i];
  }
  return true;
}

// ------------------------------------------------------------------------------------------ element front-end
static inline uint8_t encode_char(unsigned char ch, char gapchar, int& tg) {  // bpp_c.hpp:220-261
  switch (ch) {
    case 'a': tg = 0; return 1;
    case 'c': tg = 1; return 2;
    case 'g': tg = 2; return 4;
    case 't': tg = 3; return 8;
    case 'r': tg = -1; return 1 | 4;
    case 'y': tg = -1; return 2 | 8;
    case 'k': tg = -1; return 4 | 8;
    case 'm': tg = -1; return 1 | 2;
    case 's': tg = -1; return 2 | 4;
    case 'w': tg = -1; return 1 | 8;
    default: tg = ((char)ch == gapchar) ? 4 : 5; return 15;
  }
}

template <class T>
static int32_t put(std::vector<uint8_t>& blob, const std::vector<T>& v, size_t count) {
  size_t off = (blob.size() + 15) & ~size_t(15);
  blob.resize(off + sizeof(T) * count, 0);
  if (!v.empty()) memcpy(blob.data() + off, v.data(), sizeof(T) * std::min(count, v.size()));
  return (int32_t)off;
}

void plan_element(const Model& m, const uint8_t* X, int64_t G, int64_t start, int32_t len, int elem_id, ElemPlan& out) {
  const Tree& t = m.tree;
  const pa_params& P = m.prm;
  const int S = t.S, N = t.N;
  out = ElemPlan();
  int GG = len;
  if (GG < P.min_len) return;  // bpp_c.hpp:158

  // tip encoding: masks[g][s] + Tg class
  std::vector<uint8_t> masks((size_t)GG * S);
  std::vector<int8_t> tgc((size_t)GG * S);
  for (int s = 0; s < S; s++) {
    const uint8_t* row = X + (size_t)s * G + start;
    for (int g = 0; g < GG; g++) {
      int tg;
      masks[(size_t)g * S + s] = encode_char(row[g], P.gapchar, tg);
      tgc[(size_t)g * S + s] = (int8_t)tg;
    }
  }
  // column trimming (bpp_c.hpp:268-301)
  if (P.trim_gap_percent < 1) {
    std::vector<int> keep;
    for (int g = 0; g < GG; g++) {
      int mis = 0;
      for (int s = 0; s < S; s++) mis += tgc[(size_t)g * S + s] >= 4;
      if (!(mis > S * P.trim_gap_percent)) keep.push_back(g);
    }
    if ((int)keep.size() < P.min_len) return;
    if ((int)keep.size() != GG) {
      for (size_t k = 0; k < keep.size(); k++) {
        memmove(&masks[k * S], &masks[(size_t)keep[k] * S], S);
        memmove(&tgc[k * S], &tgc[(size_t)keep[k] * S], S);
      }
      GG = (int)keep.size();
      masks.resize((size_t)GG * S);
      tgc.resize((size_t)GG * S);
    }
  }
  std::vector<int> num_missing(S, 0);  // bpp_c.hpp:304-314
  for (int g = 0; g < GG; g++)
    for (int s = 0; s < S; s++) num_missing[s] += tgc[(size_t)g * S + s] == 4;

  std::vector<int32_t> children2(t.children), parent2(t.parent);
  std::vector<double> distances2(t.dist);
  std::vector<uint8_t> missing(N, 0);
  for (int s = S; s < N; s++) {  // bpp_c.hpp:333-351
    for (int cc = 0; cc < 2; cc++) {
      const int chi = children2[2 * s + cc];
      if (chi < S && num_missing[chi] > P.gap_prop * GG) missing[chi] = 1;
    }
    if (missing[children2[2 * s]] && missing[children2[2 * s + 1]]) missing[s] = 1;
  }
  int ct = 0;  // bpp_c.hpp:361-373
  for (int32_t c : m.conservedgroup)
    if (c >= 0 && c < N && missing[c]) ct++;
  if (ct > P.conserve_prop * m.conservedgroup.size()) return;

  // pruning / re-rooting along `upper` (getSubtree_missing, bpp_c.cpp:64-113)
  int root = N - 1;
  auto in_upper = [&](int s) { return P.prune_tree ? true : (bool)m.upper[s]; };
  std::function<void(int, int)> sub_missing = [&](int ss, int child) {
    if (!in_upper(ss) || ss < S) return;
    int s = ss;
    while (in_upper(s) && s >= S) {
      const int p0 = children2[2 * s], p1 = children2[2 * s + 1];
      int keepn, lose;
      if (missing[p0] && !missing[p1]) { keepn = p1; lose = p0; }
      else if (!missing[p0] && missing[p1]) { keepn = p0; lose = p1; }
      else {
        sub_missing(p0, 0);
        sub_missing(p1, 1);
        return;
      }
      parent2[lose] = -1;
      distances2[keepn] += distances2[s];
      if (s == root) root = keepn;
      else {
        children2[2 * parent2[ss] + child] = keepn;
        parent2[keepn] = parent2[ss];
      }
      s = keepn;
    }
  };
  sub_missing(N - 1, -1);
  parent2[root] = N;  // bpp_c.hpp:390

  std::vector<int32_t> nodes;
  std::function<void(int)> post = [&](int j) {  // getSubtree (bpp_c.cpp:44-60)
    for (int cc = 0; cc < 2; cc++)
      if (children2[2 * j + cc] != -1) post(children2[2 * j + cc]);
    nodes.push_back(j);
  };
  post(root);
  const int n = (int)nodes.size();
  std::vector<int> local(N, -1);
  for (int i = 0; i < n; i++) local[nodes[i]] = i;

  out.filtered = false;
  out.GG = GG;
  out.root = root;
  out.missing = missing;
  out.nodes = nodes;
  out.children2 = children2;
  out.parent2 = parent2;
  out.distances2 = distances2;
  for (int v : nodes) {
    if (m.upper[v]) out.upper_c.push_back(v);
    if (m.upper_conserve[v]) out.upper_conserve_c.push_back(v);
  }
  // fixZ per model (initMCMC, bpp_c.cpp:155-161,240-289); slot order M0, M1, M2
  out.fixZ.assign((size_t)3 * N, 0);
  for (int mdl = 0; mdl < 3; mdl++) {
    uint8_t* f = &out.fixZ[(size_t)mdl * N];
    for (int v : out.upper_c) f[v] = 1;
    f[root] = 1;
    if (mdl == PA_M1) for (int v : out.upper_conserve_c) f[v] = 1;
    if (mdl == PA_M0) for (int i = 0; i < n - 1; i++) f[nodes[i]] = 1;
  }

  // ---- site-pattern compression (new stage): unique S-nibble columns, sorted lexicographically
  std::vector<int> order(GG);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return memcmp(&masks[(size_t)a * S], &masks[(size_t)b * S], S) < 0; });
  out.site_pattern.assign(GG, 0);
  std::vector<int> sorted_site_pat(GG);
  int Pn = 0;
  for (int k = 0; k < GG; k++) {
    if (k == 0 || memcmp(&masks[(size_t)order[k] * S], &masks[(size_t)order[k - 1] * S], S) != 0) {
      out.pattern_codes.insert(out.pattern_codes.end(), &masks[(size_t)order[k] * S], &masks[(size_t)order[k] * S] + S);
      out.pattern_weights.push_back(0);
      Pn++;
    }
    out.pattern_weights[Pn - 1]++;
    out.site_pattern[order[k]] = Pn - 1;
    sorted_site_pat[k] = Pn - 1;
  }
  out.P = Pn;

  // ---- device blob
  std::vector<int16_t> parent_l(n), child_l(2 * n), tiprow(n), orig(n), intlist, intidx(n, -1);
  std::vector<uint8_t> flags(n, 0);
  std::vector<double> dist(n);
  for (int i = 0; i < n; i++) {
    const int v = nodes[i];
    orig[i] = (int16_t)v;
    parent_l[i] = (v == root) ? (int16_t)-1 : (int16_t)local[parent2[v]];
    const bool tip = children2[2 * v] == -1;
    child_l[2 * i] = tip ? (int16_t)-1 : (int16_t)local[children2[2 * v]];
    child_l[2 * i + 1] = tip ? (int16_t)-1 : (int16_t)local[children2[2 * v + 1]];
    tiprow[i] = tip ? (int16_t)v : (int16_t)-1;
    dist[i] = distances2[v];
    uint8_t f = 0;
    if (missing[v]) f |= PA_F_MISSING;
    if (tip) f |= PA_F_TIP;
    for (int mdl = 0; mdl < 3; mdl++)
      if (out.fixZ[(size_t)mdl * N + v]) f |= (uint8_t)(PA_F_FIXZ0 << mdl);
    flags[i] = f;
    if (!tip) {
      intidx[i] = (int16_t)intlist.size();
      intlist.push_back((int16_t)i);
    }
  }
  const int ni = (int)intlist.size();
  // pruning traversal: visit the child subtree needing more temporaries first (Sethi-Ullman), missing subtrees are skipped
  std::vector<int> need(n, 0);
  auto computed = [&](int i) { return !(flags[i] & PA_F_TIP) && !(flags[i] & PA_F_MISSING); };
  for (int i = 0; i < n; i++) {
    if (!computed(i)) continue;
    int a = child_l[2 * i], b = child_l[2 * i + 1];
    int na = computed(a) ? need[a] : 0, nb2 = computed(b) ? need[b] : 0;
    if (na < nb2) std::swap(na, nb2);
    need[i] = std::max(1, (nb2 > 0) ? std::max(na, nb2 + 1) : na);
  }
  std::vector<int16_t> su;
  std::vector<int8_t> slot(n, -1);
  std::vector<int> free_slots;
  int n_slots = 0;
  std::function<void(int)> visit = [&](int i) {
    int a = child_l[2 * i], b = child_l[2 * i + 1];
    const bool ca = computed(a), cb = computed(b);
    if (ca && cb) {
      if (need[a] < need[b]) std::swap(a, b);
      visit(a);
      visit(b);
    } else if (ca) visit(a);
    else if (cb) visit(b);
    // result slot: reuse a child's slot, free the other
    int sl = -1;
    if (ca) sl = slot[a];
    if (cb) {
      if (sl < 0) sl = slot[b];
      else free_slots.push_back(slot[b]);
    }
    if (sl < 0) {
      if (!free_slots.empty()) { sl = free_slots.back(); free_slots.pop_back(); }
      else sl = n_slots++;
    }
    slot[i] = (int8_t)sl;
    su.push_back((int16_t)i);
  };
  if (computed(n - 1)) visit(n - 1);
  // Z levels: heights (upward) and depths (downward) over ALL nodes of the pruned tree
  std::vector<int> height(n, 0), depth(n, 0);
  for (int i = 0; i < n; i++)
    if (!(flags[i] & PA_F_TIP)) height[i] = 1 + std::max(height[child_l[2 * i]], height[child_l[2 * i + 1]]);
  for (int i = n - 2; i >= 0; i--) depth[i] = depth[parent_l[i]] + 1;
  auto level_sort = [&](const std::vector<int>& key, std::vector<int16_t>& nodes_out, std::vector<int16_t>& off_out) {
    const int L = *std::max_element(key.begin(), key.end()) + 1;
    off_out.assign(L + 1, 0);
    for (int i = 0; i < n; i++) off_out[key[i] + 1]++;
    for (int l = 0; l < L; l++) off_out[l + 1] += off_out[l];
    nodes_out.assign(n, 0);
    std::vector<int> cur(off_out.begin(), off_out.end() - 1);
    for (int i = 0; i < n; i++) nodes_out[cur[key[i]]++] = (int16_t)i;
    return L;
  };
  std::vector<int16_t> up_nodes, up_off, down_nodes, down_off;
  const int n_up = level_sort(height, up_nodes, up_off);
  const int n_down = level_sort(depth, down_nodes, down_off);

  const int Ppad = (Pn + 31) & ~31;
  const int W = (ni + 31) / 32;
  std::vector<uint8_t> codes((size_t)S * (Ppad / 2), 0xFF);
  for (int p = 0; p < Pn; p++)
    for (int s = 0; s < S; s++) {
      uint8_t& b = codes[(size_t)s * (Ppad / 2) + (p >> 1)];
      const uint8_t c = out.pattern_codes[(size_t)p * S + s];
      b = (p & 1) ? (uint8_t)((b & 0x0F) | (c << 4)) : (uint8_t)((b & 0xF0) | c);
    }
  std::vector<uint16_t> weights(Ppad, 0);
  for (int p = 0; p < Pn; p++) weights[p] = (uint16_t)out.pattern_weights[p];
  // gap flags of internal nodes per pattern (bpp_c.hpp:417-428): both children "all-gap" (raw tip codes, regardless of `missing`)
  std::vector<uint32_t> gap((size_t)W * Ppad, 0);
  {
    std::vector<uint8_t> isgap(n);
    for (int p = 0; p < Pn; p++) {
      for (int i = 0; i < n; i++) {
        if (flags[i] & PA_F_TIP) isgap[i] = out.pattern_codes[(size_t)p * S + tiprow[i]] == 15;
        else {
          isgap[i] = isgap[child_l[2 * i]] && isgap[child_l[2 * i + 1]];
          if (isgap[i]) gap[(size_t)(intidx[i] >> 5) * Ppad + p] |= 1u << (intidx[i] & 31);
        }
      }
    }
  }
  const int GGpad = (GG + 31) & ~31;
  std::vector<uint16_t> sitepat(GGpad, 0);
  for (int k = 0; k < GG; k++) sitepat[k] = (uint16_t)sorted_site_pat[k];
  // the tips' data as bit planes over the sites of each 256-site tile (canonical site order): the pair-count pass of the
  // chain kernel intersects them with the ballots of the sampled ancestral bases (getEmission's sufficient statistic)
  const int n_tiles = (GG + 255) / 256;
  std::vector<uint32_t> tipplanes((size_t)n_tiles * S * 32, 0), tipamb((size_t)n_tiles * S * 8, 0);
  for (int k = 0; k < GG; k++) {
    const int tile = k >> 8, w = (k >> 5) & 7, bit = k & 31;
    const uint8_t* col = &out.pattern_codes[(size_t)sorted_site_pat[k] * S];
    for (int s = 0; s < S; s++) {
      const uint8_t c = col[s];
      if (c == 15) continue;
      const size_t cell = ((size_t)tile * S + s) * 8 + w;
      if (c == 1 || c == 2 || c == 4 || c == 8) tipplanes[cell * 4 + (c == 1 ? 0 : (c == 2 ? 1 : (c == 4 ? 2 : 3)))] |= 1u << bit;
      else tipamb[cell] |= 1u << bit;
    }
  }

  std::vector<uint8_t>& blob = out.blob;
  blob.assign(sizeof(ElemHeader), 0);
  ElemHeader h;
  memset(&h, 0, sizeof(h));
  h.n = n; h.ni = ni; h.S = S; h.P = Pn; h.Ppad = Ppad; h.GG = GG; h.W = W; h.n_slots = std::max(1, n_slots);
  h.n_up = n_up; h.n_down = n_down; h.elem_id = elem_id;
  out.n_slots = std::max(1, n_slots);
  h.off_parent = put(blob, parent_l, n);
  h.off_child = put(blob, child_l, 2 * n);
  h.off_flags = put(blob, flags, n);
  h.off_tiprow = put(blob, tiprow, n);
  h.off_orig = put(blob, orig, n);
  h.off_dist = put(blob, dist, n);
  h.off_int = put(blob, intlist, ni);
  h.off_intidx = put(blob, intidx, n);
  h.off_su = put(blob, su, ni);  // missing internal nodes are absent: entries beyond su.size() are -1
  {
    int16_t* p = (int16_t*)(blob.data() + h.off_su);
    for (int k = (int)su.size(); k < ni; k++) p[k] = -1;
  }
  h.off_slot = put(blob, slot, n);
  h.off_up_nodes = put(blob, up_nodes, n);
  h.off_up_off = put(blob, up_off, n_up + 1);
  h.off_down_nodes = put(blob, down_nodes, n);
  h.off_down_off = put(blob, down_off, n_down + 1);
  h.off_codes = put(blob, codes, codes.size());
  h.off_weights = put(blob, weights, Ppad);
  h.off_gap = put(blob, gap, gap.size());
  h.off_sitepat = put(blob, sitepat, GGpad);
  h.off_tipplanes = put(blob, tipplanes, tipplanes.size());
  h.off_tipamb = put(blob, tipamb, tipamb.size());
  blob.resize((blob.size() + 255) & ~size_t(255), 0);
  h.blob_bytes = (int32_t)blob.size();
  memcpy(blob.data(), &h, sizeof(h));
}

}  // namespace pa

extern "C" void pa_params_default(pa_params* p) {  // main.cpp:49-90
  memset(p, 0, sizeof(*p));
  p->seed = 5;
  p->init_conserve_rate = 0.5;
  p->init_acce_rate = 1;
  p->conserve_prior_a = 5; p->conserve_prior_b = 0.04;
  p->acce_prior_a = 10; p->acce_prior_b = 0.2;
  p->ropt = 1; p->cub = 1; p->nlb = 0.6;
  p->burnin = 200; p->mcmc = 800; p->adapt_freq = 500;
  p->init_lrate = 0.3; p->init_lrate2 = 0.1; p->init_grate = 0.5;
  p->hyper_lrate_a = 1; p->hyper_lrate_b = 9;
  p->hyper_grate_a = 3; p->hyper_grate_b = 1;
  p->hyper_lrate2_a = 1; p->hyper_lrate2_b = 9;
  p->conserve_prop = 0.8; p->gap_prop = 0.8; p->constomis = 0.01;
  p->gapchar = '-';
  p->prune_tree = 0;
  p->trim_gap_percent = 1;
  p->min_len = 50;
  p->num_thread = 1;
}
