// pa_fmt.h -- number formatting of the result files.  The reference writes every file through std::ofstream with the default (general)
// float format at precision 6, or 8 after `out.precision(8)` -- which libstdc++ implements as printf("%.*g").  Formatting 17 million
// probabilities of a genome-scale batch through operator<< took 5.3 s on one core; rows are formatted into buffers here (byte-identical
// output: tests/test_gpu_outputs.py compares the files with the reference's own writers), by all host threads.
#ifndef PA_FMT_H
#define PA_FMT_H
#include <charconv>
#include <cmath>
#include <cstdio>
#include <string>

namespace pa {

inline void fmt_g(std::string& out, double v, int prec = 6) {
  if (v == 0.0 && !std::signbit(v)) { out.push_back('0'); return; }
  if (v == 1.0) { out.push_back('1'); return; }
  char buf[40];
  const int n = snprintf(buf, sizeof(buf), "%.*g", prec, v);
  out.append(buf, n > 0 ? (size_t)n : 0);
}

inline void fmt_i(std::string& out, long long v) {
  char buf[24];
  const auto r = std::to_chars(buf, buf + sizeof(buf), v);
  out.append(buf, r.ptr - buf);
}

}  // namespace pa
#endif
