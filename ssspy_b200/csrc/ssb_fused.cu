// Fused sm_100a fast path of the GaussILRMA iteration -- placeholder until the tensor-core sweep
// kernels land; every configuration currently takes the modular kernels.
#include "ssb_fused.h"

size_t ssb_fused_carve(ssb_fused_ws* ws, const ssb_config*, char* base) {
  ws->base = base;
  ws->bytes = 0;
  return 0;
}
int ssb_fused_supported(const ssb_config*) { return 0; }
int ssb_fused_prepare(ssb_fused_ws*, const ssb_config*, const cf*, cudaStream_t) { return 0; }
int ssb_fused_update_once(ssb_fused_ws*, const ssb_config*, const cf*, cf*, float*, float*, const cf*, cudaStream_t) {
  ssb_set_error("fused path not available");
  return 1;
}
