// Fused sm_100a fast path of the GaussILRMA iteration (see ssb_fused.cu).
#pragma once
#include "ssb_common.cuh"

struct ssb_fused_ws {
  char* base = nullptr;
  size_t bytes = 0;
};

// bytes of extra scratch (base may be NULL to only measure)
size_t ssb_fused_carve(ssb_fused_ws* ws, const ssb_config* cfg, char* base);
// 1 if the configuration is covered by the fused kernels
int ssb_fused_supported(const ssb_config* cfg);
int ssb_fused_prepare(ssb_fused_ws* ws, const ssb_config* cfg, const cf* X, cudaStream_t st);
int ssb_fused_update_once(ssb_fused_ws* ws, const ssb_config* cfg, const cf* X, cf* W, float* T, float* V,
                          const cf* C, cudaStream_t st);
