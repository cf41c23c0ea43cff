// pa_nvtx.h -- NVTX ranges around the host-visible stages of a run (load / sample / collect / write), for nsys / ncu timelines.
// NVTX v3 is header-only: without a profiler attached a range costs one branch.
#ifndef PA_NVTX_H
#define PA_NVTX_H
#if __has_include(<nvtx3/nvToolsExt.h>)
#include <nvtx3/nvToolsExt.h>
namespace pa {
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};
}  // namespace pa
#define PA_NVTX_CAT2(a, b) a##b
#define PA_NVTX_CAT(a, b) PA_NVTX_CAT2(a, b)
#define PA_NVTX_RANGE(name) ::pa::NvtxRange PA_NVTX_CAT(pa_nvtx_range_, __LINE__)(name)
#else
#define PA_NVTX_RANGE(name) do { } while (0)
#endif
#endif
