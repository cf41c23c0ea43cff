// pa_main.cpp -- `PhyloAcc-ST-b200 <param_file>`: drop-in for the reference's PhyloAcc-ST binary
// (src/PhyloAcc-ST/main.cpp:324-528).  The driver probes `<st_path>` with NO arguments on every run and
// treats exit code <= 1 as "present" (phyloacc_lib/core.py:47-110), so argument checking happens before
// the CUDA library is even loaded.  Built a second time with -DPA_MAIN_GBGC as `PhyloAcc-GBGC-b200`, the drop-in for the PhyloAcc-ST-GBGC
// binary (src/PhyloAcc-ST-GBGC/SRC/main.cpp:240-445, entry pa_gbgc_run_from_config); without an argument that program falls back to a
// developer's hard-coded path and exits 1 with "Cannot open the parameters file" (main.cpp:93-106).
#include <dlfcn.h>
#include <libgen.h>
#include <limits.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <string>

int main(int argc, char* argv[]) {
#ifdef PA_MAIN_GBGC
  const char* entry = "pa_gbgc_run_from_config";
  if (argc <= 1) {
    std::cout << "Loading input data and running parameters......" << std::endl;
    std::cerr << "Cannot open the parameters file: (none given)" << std::endl;
    return 1;
  }
#else
  const char* entry = "pa_run_from_config";
  std::cout << "Loading input data and running parameters......" << std::endl;
  if (argc <= 1) {
    std::cout << "Please specify the path of the parameter file." << std::endl;  // main.cpp:108-113
    return 1;
  }
#endif
  char self[PATH_MAX];
  const ssize_t n = readlink("/proc/self/exe", self, sizeof(self) - 1);
  std::string dir = ".";
  if (n > 0) {
    self[n] = '\0';
    dir = dirname(self);
  }
  const char* env = getenv("PHYLOACC_B200_LIB");
  const std::string lib = env ? std::string(env) : dir + "/libphyloacc_b200.so";
  void* h = dlopen(lib.c_str(), RTLD_NOW | RTLD_LOCAL);
  if (!h) {
    std::cerr << "(Error. Cannot load " << lib << ": " << dlerror() << ")" << std::endl;
    return 1;
  }
  typedef int (*run_fn)(const char*);
  run_fn run = (run_fn)dlsym(h, entry);
  if (!run) {
    std::cerr << "(Error. " << lib << " does not export " << entry << ")" << std::endl;
    return 1;
  }
  return run(argv[1]);
}
