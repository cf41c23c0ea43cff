"""csrc/pa_fmt.h formats the numbers of every result file into buffers; the reference writes them through std::ofstream (default float
format, precision 6, or 8 after out.precision(8); ints through operator<<).  A small C++ program built here prints the same values both
ways -- the claim "byte-identical to iostream" is checked on edge values and on 200k random doubles without a GPU."""
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SRC = r"""
#include <cstdio>
#include <cstring>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include "pa_fmt.h"
int main() {
  std::vector<double> v;
  double x;
  while (fread(&x, sizeof(x), 1, stdin) == 1) v.push_back(x);
  size_t bad = 0;
  for (int prec : {6, 8}) {
    std::ostringstream os;
    os.precision(prec);
    std::string mine;
    for (double d : v) {
      os << d << "\t";
      pa::fmt_g(mine, d, prec);
      mine.push_back('\t');
    }
    if (os.str() != mine) {
      bad++;
      std::istringstream a(os.str()), b(mine);
      std::string ta, tb;
      int shown = 0;
      while (std::getline(a, ta, '\t') && std::getline(b, tb, '\t') && shown < 5)
        if (ta != tb) { fprintf(stderr, "prec %d: iostream '%s' fmt_g '%s'\n", prec, ta.c_str(), tb.c_str()); shown++; }
    }
  }
  std::ostringstream os;
  std::string mine;
  for (long long k : {0LL, -1LL, 2LL, 85LL, -2147483648LL, 9007199254740993LL}) {
    os << k << " ";
    pa::fmt_i(mine, k);
    mine.push_back(' ');
  }
  if (os.str() != mine) bad++;
  printf("%zu values, mismatching blocks: %zu\n", v.size(), bad);
  return bad ? 1 : 0;
}
"""


def test_fmt_g_equals_iostream(tmp_path):
    src = tmp_path / "fmt_check.cpp"
    src.write_text(SRC)
    exe = tmp_path / "fmt_check"
    subprocess.run(["g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "phyloacc_b200", "csrc"), str(src), "-o", str(exe)], check=True)
    rng = np.random.default_rng(5)
    edge = [0.0, -0.0, 1.0, -1.0, 0.5, 1e-5, 9.9999e-5, 1e-4, 123456.0, 1234567.0, 0.1, 1 / 3, 2 / 3, 0.0166667, 0.983333, 1e-300, 5e-324, 1e300,
            np.inf, -np.inf, np.nan, -np.nan, 999999.5, 99999.95, 0.99999949, 0.9999995, 1.0000005, -1231.0082, -2.0841616, 3.004331, 100.0, 1e15, 1e16]
    vals = np.concatenate([np.array(edge), rng.random(50000), rng.integers(0, 1001, 50000) / 1000.0, -rng.gamma(2.0, 700.0, 50000),
                           np.exp(rng.uniform(-40, 40, 50000)) * rng.choice([-1, 1], 50000)])
    r = subprocess.run([str(exe)], input=vals.astype(np.float64).tobytes(), capture_output=True)
    assert r.returncode == 0, (r.stdout.decode(), r.stderr.decode())
    assert f"{len(vals)} values".encode() in r.stdout
