// pa_app.h -- pieces of the file-level front door shared by the PhyloAcc-ST (pa_app.cpp) and PhyloAcc-ST-GBGC (pa_gbgc_app.cpp) programs
#ifndef PA_APP_H
#define PA_APP_H
#include <string>
#include <vector>

#include "pa_engine.h"
#include "pa_internal.h"

namespace pa {

// what the parsers produce: LoadParams + LoadPhyloProfiles + LoadPhyloTree + id selection (main.cpp:331-358,406-439)
struct ConfigInputs {
  Params P;
  Tree tree;
  std::vector<std::string> element_names;
  std::vector<int64_t> elem_start;
  std::vector<int32_t> elem_len;
  std::vector<int32_t> og, cg;  // OUTGROUP / CONSERVE as leaf ids
  std::vector<int> ids;
  std::vector<uint8_t> X;       // [S][G]
  int64_t G = 0;
  int C = 0;
};

bool directory_exists(const std::string& p);
int load_inputs(ConfigInputs& in);  // in.P filled by the caller
int make_engine(const ConfigInputs& in, int device, int num_thread, pa_engine** out);
std::vector<int> pick_devices(int n_elems);

}  // namespace pa
#endif
