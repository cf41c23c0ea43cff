#include "pa_engine.h"
namespace pa {
int summarize_chain(Engine&, int, int, const ChainTraceHost&, ChainSummary&) { return PA_ERR_STATE; }
int Engine::sample(const int32_t*, int32_t) { set_error("not built"); return PA_ERR_STATE; }
int Engine::fetch_traces() { return PA_ERR_STATE; }
int Engine::collect(pa_results*) { return PA_ERR_STATE; }
int Engine::get_trace(int32_t, int32_t, ChainTraceHost&) { return PA_ERR_STATE; }
}
