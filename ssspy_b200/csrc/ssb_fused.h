// Fused tensor-core fast path of the GaussILRMA iteration (see ssb_fused.cu).
#pragma once
#include "ssb_common.cuh"

struct ssb_fused_ws {
  char* base = nullptr;
  size_t bytes = 0;
  bool zeroed = false;    // the pre-split operand buffers rely on zero padding rows / columns
  bool vs_valid = false;  // Vs holds the split of the current V (set by the activation kernel, kept across the
                          // iterations of ssb_run only)
};

// bytes of extra scratch (base may be NULL to only measure)
size_t ssb_fused_carve(ssb_fused_ws* ws, const ssb_config* cfg, char* base);
// 1 if the configuration is covered by the fused kernels
int ssb_fused_supported(const ssb_config* cfg);
int ssb_fused_prepare(ssb_fused_ws* ws, const ssb_config* cfg, const cf* X, cudaStream_t st);
// MM source model (T then V, p = 2) followed by phi = 1/(T V) and the weighted covariances U
// P[B,N,I,J] f32 is scratch (power spectrogram handed from the basis to the activation update)
// phases: bit 0 = source model, bit 1 = covariances
int ssb_fused_source_and_cov(const ssb_config* cfg, ssb_fused_ws* ws, const cf* X, cf* W, float* T, float* V,
                             float* P, cf* U, cudaStream_t st, int phases = 3);
int ssb_fused_coop_enabled();
// N = 2, IP1, inside ssb_run only: [covariance + IP1 of iteration t, W unnormalised, q = mean_j |y|^2] fused with the
// basis update of iteration t + 1 (one pass over X from HBM), then the activation update; the caller normalises
// afterwards (ssb_fused_spatial_source in ssb_fused.cu explains why the order is legal); the kernel is the TMA tile
// kernel of ssb_tma.cu (SSB_TMA bit 2)
int ssb_fused_iter_fusable(const ssb_config* cfg, const ssb_fused_ws* ws);
int ssb_fused_spatial_source(const ssb_config* cfg, ssb_fused_ws* ws, const cf* X, cf* W, float* T, float* V, float* P,
                             double* q, cudaStream_t st);
// cooperative MM source model (ssb_coop.cu): basis kernel with one CTA = 16-bin tiles x all sources sharing the X
// slab in shared memory, then the activation kernel; ws is zero-initialised scratch of ssb_coop_ws_bytes() bytes
// (pre-split bf16 copies of V and T)
size_t ssb_coop_ws_bytes(const ssb_config* cfg);
// cooperative weighted covariance (N = 4, 8): reads the pre-split Vs left in ws by ssb_coop_source
int ssb_coop_cov_supported(const ssb_config* cfg);
int ssb_coop_cov(const ssb_config* cfg, const cf* X, const float* T, const void* ws, cf* U, cudaStream_t st);
// FastGaussMNMF: tensor-core multiplicative updates with elementwise factors given as arrays (which: 0 basis, 1 activation)
// the same updates with G / H formed in the kernel from Z2 = |Q x|^2 and D (four sources, K <= 16; ssb_coop.cu)
// (zscale[b * 4 + m], optional: Z2 is multiplied by it on load)
int ssb_coop_mnmf_update(const ssb_config* cfg, int which, const float* Z2, const float* zscale, const float* Dm, float* T,
                         float* V, void* ws, cudaStream_t st);
int ssb_coop_update_ab(const ssb_config* cfg, int which, const float* A, const float* Bm, float* T, float* V, void* ws,
                       cudaStream_t st);
int ssb_coop_source(const ssb_config* cfg, const cf* X, const cf* W, float* T, float* V, float* P, void* ws,
                    int vs_valid, cudaStream_t st);
// the pieces of ssb_coop_source on their own (the TMA basis kernel of ssb_tma.cu sits between them): pre-split of V
// into ws unless vs_valid, and the activation update from P and the pre-split basis in ws
int ssb_coop_vsplit(const ssb_config* cfg, const float* V, void* ws, int vs_valid, cudaStream_t st);
int ssb_coop_activation(const ssb_config* cfg, float* V, float* P, void* ws, cudaStream_t st);
// pre-split activation Vs / basis Ts inside the cooperative scratch ws
void* ssb_coop_vs(const ssb_config* cfg, void* ws);
void* ssb_coop_ts(const ssb_config* cfg, void* ws);
// TMA-fed tile kernels (ssb_tma.cu): cp.async.bulk.tensor + mbarrier rings; N = 2, 4, 8, n_frames % 16 == 0, K <= 32
int ssb_tma_mask(const ssb_config* cfg);
// tensor-core weighted covariance for N = 8 (ssb_covmma.cu): U = sum_j phi G as split-bf16 mma, X by TMA
int ssb_cov_mma_supported(const ssb_config* cfg, const cf* X);
int ssb_cov_mma(const ssb_config* cfg, const cf* X, const float* T, const void* Vs, cf* U, cudaStream_t st);
int ssb_tma_supported(const ssb_config* cfg);
int ssb_tma_basis(const ssb_config* cfg, const cf* X, const cf* W, float* T, const void* Vs, float* P, void* Ts,
                  cudaStream_t st);
int ssb_tma_cov_n2(const ssb_config* cfg, const cf* X, float* T, const void* Vs, cf* U, cudaStream_t st);
int ssb_tma_spatial_basis_n2(const ssb_config* cfg, const cf* X, cf* W, float* T, const void* Vs, float* P, void* Ts,
                             double* q, cudaStream_t st);
// closed-form IP1 for two sources; with C != NULL also q[mat,n] = Re(w_n C w_n^H) for the normalisation
int ssb_fused_ip1_n2(cf* W, const cf* U, const cf* C, double* q, int n_mat, int flooring, float eps,
                     cudaStream_t st);
// power normalisation from the per-bin terms q: psi_n = floor(sqrt(mean_i q)), T /= psi^p, W /= psi
int ssb_fused_normalize(const double* q, float* T, cf* W, int B, int N, int I, int K, float p, int flooring,
                        float eps, cudaStream_t st);
// weighted covariance with array weights phi[b*sb + s*sn + i*si + j] for s < n_src (AuxIVA: si = 0;
// FastGaussMNMF: per-bin weights); U[B,I,n_src,N,N];
// requires n_frames % 16 == 0
int ssb_fused_cov_w(const cf* X, const float* phi, long long sb, long long sn, long long si, int n_src, cf* U, int B,
                    int N, int I, int J, cudaStream_t st);
// FastGaussMNMF diagonaliser covariance with the weights 1 / (sum_n Lambda_n D[i, n, m]) formed in the kernel
int ssb_fused_cov_lambda(const cf* X, const float* Lam, const float* Dm, cf* U, int B, int N, int I, int J,
                         cudaStream_t st);
// ISS modes: MM source model (T then V, p = 2) with P = |Y|^2, and phi = 1/(T V) as an array
int ssb_fused_source_iss(const ssb_config* cfg, const cf* Y, float* T, float* V, float* P, cudaStream_t st);
// inverse = 1: phi = 1/(T V); inverse = 0: Lambda = T V
int ssb_fused_phi(const ssb_config* cfg, const float* T, const float* V, float* phi, int inverse, cudaStream_t st);
