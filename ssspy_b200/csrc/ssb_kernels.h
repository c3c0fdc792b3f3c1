// Internal launch-wrapper declarations shared by the translation units of libssb.so.
// Every ssbk_* function enqueues kernels on `st` and returns 0 / non-zero (ssb_last_error()).
#pragma once
#include "ssb_common.cuh"

#define SSB_DISPATCH_N(N_, ...)                                              \
  switch (N_) {                                                              \
    case 2: { constexpr int NN = 2; __VA_ARGS__; } break;                    \
    case 3: { constexpr int NN = 3; __VA_ARGS__; } break;                    \
    case 4: { constexpr int NN = 4; __VA_ARGS__; } break;                    \
    case 5: { constexpr int NN = 5; __VA_ARGS__; } break;                    \
    case 6: { constexpr int NN = 6; __VA_ARGS__; } break;                    \
    case 7: { constexpr int NN = 7; __VA_ARGS__; } break;                    \
    case 8: { constexpr int NN = 8; __VA_ARGS__; } break;                    \
    default:                                                                 \
      ssb_set_error("n_sources=%d unsupported (2..%d)", N_, SSB_MAX_SOURCES); \
      return 1;                                                              \
  }

// K (n_basis) is padded up to a compile-time KP for the register-resident NMF kernels
#define SSB_DISPATCH_K(K_, ...)                                            \
  if ((K_) <= 8) { constexpr int KP = 8; __VA_ARGS__; }                    \
  else if ((K_) <= 16) { constexpr int KP = 16; __VA_ARGS__; }             \
  else if ((K_) <= 32) { constexpr int KP = 32; __VA_ARGS__; }             \
  else if ((K_) <= 64) { constexpr int KP = 64; __VA_ARGS__; }             \
  else {                                                                   \
    ssb_set_error("n_basis=%d unsupported (1..%d)", K_, SSB_MAX_BASIS);    \
    return 1;                                                              \
  }

static inline int blocks_for(long long items, int per_block) { return (int)((items + per_block - 1) / per_block); }

// ---- ssb_spatial.cu ----------------------------------------------------------------------------
int ssbk_separate(const cf* X, const cf* W, cf* Y, float* P, int B, int N, int I, int J, cudaStream_t st);
int ssbk_abs2(const cf* Y, float* P, size_t n, cudaStream_t st);
int ssbk_wcov(const cf* X, const float* phi, long long sb, long long sn, long long si, const int* src, int n_src,
              cf* U, int B, int N, int I, int J, cudaStream_t st);
// C, q (optional): also emit q[mat, n] = Re(w_n C_mat w_n^H), the per-bin term of the power normalisation
int ssbk_ip1(cf* W, const cf* U, int n_mat, int N, int flooring, float eps, cudaStream_t st, const cf* C = nullptr,
             double* q = nullptr);
// uidx (host, 2*n_pairs, may be NULL => U slice index = source index): which of the n_u slices of U
// each pair member uses
int ssbk_ip2(cf* W, const cf* U, int n_mat, int N, const int* pairs, int n_pairs, int n_u, const int* uidx,
             int flooring, float eps, cudaStream_t st, const cf* C = nullptr, double* q = nullptr);
// r2part / r2 (optional, AuxIVA): the apply sweep of the covariance-domain kernel also emits r2[b,n,j] = sum_i |y_new|^2
// through the partials r2part[B, ssbk_iss1_r2_groups(I), N, J] (only when ssbk_iss1_emits_r2(N, J))
int ssbk_iss1(cf* Y, const float* phi, long long sb, long long sn, long long si, int B, int N, int I, int J,
              int flooring, float eps, cudaStream_t st, float* r2part = nullptr, float* r2 = nullptr);
int ssbk_iss1_r2_groups(int I);
int ssbk_iss1_emits_r2(int N, int J);
// pairs (host, 2*n_pairs, already wrapped into [0, N)); _update_spatial_model.py:197-314
int ssbk_iss2(cf* Y, const float* phi, long long sb, long long sn, long long si, int B, int N, int I, int J,
              const int* pairs, int n_pairs, int flooring, float eps, cudaStream_t st);
// AuxLaplaceFDICA (ssb_fdica.cu): per-(source, bin, frame) weights for the sources in `src` (NULL = all, in order)
// written as phi[B, n_src, I, J]; per-bin loss term; correlation-based permutation solver
int ssbk_fdica_phi(const cf* X, const cf* W, float* phi, const int* src, int n_src, int B, int N, int I, int J,
                   int flooring, float eps, cudaStream_t st);
int ssbk_fdica_rowloss(const cf* X, const cf* W, double* rowloss, int B, int N, int I, int J, cudaStream_t st);
int ssbk_perm_corr(const cf* Y, double* corr, int B, int N, int I, int J, int flooring, float eps, cudaStream_t st);
int ssbk_perm_align(cf* Y, cf* W, const int* order, int* perms, int B, int N, int I, int J, int flooring, float eps,
                    cudaStream_t st);
int ssbk_ipa(cf* Y, const float* phi, long long sb, long long sn, long long si, int B, int N, int I, int J,
             int normalization, int max_iter, int flooring, float eps, cudaStream_t st);
int ssbk_pb_w(const cf* W, cf* Wout, cf* scale_out, int n_mat, int N, int ref, cudaStream_t st);
int ssbk_cross_solve(const cf* A, const cf* Bm, cf* S, int B, int N, int I, int J, cudaStream_t st);
int ssbk_scale_rows(const cf* Y, const cf* S, cf* Yout, int B, int N, int I, int J, int ref, cudaStream_t st);
int ssbk_logdet(const cf* W, double* out, int n_mat, int N, cudaStream_t st);
int ssbk_mdp(const cf* Y, const cf* X, cf* Yout, int B, int N, int I, int J, int ref, cudaStream_t st);

// ---- ssb_nmf.cu --------------------------------------------------------------------------------
// MM/ME multiplicative updates from the power spectrogram P[B,N,I,J] (ssspy/bss/ilrma.py:1116-1126,
// :1192-1202, :1311-1323, :1387-1399)
// `model` = SSB_MODEL_ILRMA_{GAUSS,T,GGD}, `prm` = nu (TILRMA, ilrma.py:2620-2640, :2868-2886) or beta
// (GGDILRMA, ilrma.py:3700-3720)
// vdiv = 1: V[BN,K,J] per source; vdiv = N: V[B,K,J] shared by the sources (partitioning function).  With
// num_out / den_out the raw sums are written ([BN,I,K] resp. [BN,K,J]) instead of updating T / V.
int ssbk_nmf_basis(const float* P, float* T, const float* V, int BN, int I, int J, int K, float p, int source,
                   int model, float prm, int flooring, float eps, cudaStream_t st, int vdiv = 1,
                   float* num_out = nullptr, float* den_out = nullptr);
int ssbk_nmf_activation(const float* P, const float* T, float* V, int BN, int I, int J, int K, float p, int source,
                        int model, float prm, int flooring, float eps, cudaStream_t st, int vdiv = 1,
                        float* num_out = nullptr, float* den_out = nullptr);
// partitioning function (ilrma.py:201-245, :1007-1049, :1098-1126, :1174-1202, :424-430)
int ssbk_part_teff(const float* Z, const float* T, float* Teff, int B, int N, int I, int K, cudaStream_t st);
int ssbk_part_latent(const float* gnum, const float* gden, const float* T, float* Z, int B, int N, int I, int K,
                     float p, int source, int model, float prm, cudaStream_t st);
int ssbk_part_basis(const float* gnum, const float* gden, const float* Z, float* T, int B, int N, int I, int K, float p,
                    int source, int model, float prm, int flooring, float eps, cudaStream_t st);
int ssbk_part_activation(const float* hnum, const float* hden, float* V, int B, int N, int K, int J, float p,
                         int source, int model, float prm, int flooring, float eps, cudaStream_t st);
int ssbk_part_normalize(const double* psi2, float* Z, float* T, int B, int N, int I, int K, float p, int flooring,
                        float eps, cudaStream_t st);
// phi[B,N,I,J] = (T V)^(-2/p)   (ilrma.py:1494-1498); Student-t: 1/(nu/(nu+2) R^(2/p) + 2/(nu+2) P)
// (ilrma.py:2920-2934); GGD: 1/((2/beta) floor(P^((2-beta)/2)) R^(beta/p)) (ilrma.py:3992-4010).  P (only read for
// T / GGD) may alias phi.
int ssbk_nmf_phi(const float* T, const float* V, const float* P, float* phi, int BN, int I, int J, int K, float p,
                 int model, float prm, int flooring, float eps, cudaStream_t st, int vdiv = 1);
// rowloss[B,N,I] = mean_j( P / R^(2/p) + (2/p) log R ),  R = T V   (ilrma.py:1957-1964; t :3265-3280;
// GGD :4340-4355)
int ssbk_nmf_rowloss(const float* P, const float* T, const float* V, double* rowloss, int BN, int I, int J, int K,
                     float p, int model, float prm, cudaStream_t st, int vdiv = 1);
// loss[b] = sum_{n,i} rowloss - 2 sum_i logdet   (ilrma.py:1964-1965)
int ssbk_ilrma_loss_reduce(const double* rowloss, const double* logdet, double* loss, int B, int N, int I,
                           cudaStream_t st);
// power normalisation (ilrma.py:412-444)
int ssbk_psi_from_cov(const cf* W, const cf* C, double* psi2, int B, int N, int I, cudaStream_t st);
int ssbk_psi_from_y(const cf* Y, double* psi2, int B, int N, int I, int J, cudaStream_t st);
int ssbk_apply_psi(const double* psi2, float* T, cf* W, cf* Y, int B, int N, int I, int J, int K, float p,
                   int flooring, float eps, cudaStream_t st);
// projection-back normalisation: T[b,n,i,:] *= |s[b,i,n]|^p, s with element stride (ilrma.py:510-514)
int ssbk_scale_basis(float* T, const cf* s, long long s_mat_stride, long long s_src_stride, int B, int N, int I,
                     int K, float p, cudaStream_t st);

// ---- ssb_iva.cu --------------------------------------------------------------------------------
// r2[b,s,j] = sum_i |y_{src[s]}|^2 with y = W x (W != NULL) or y = Y (iva.py:1787, :1903, :1962)
int ssbk_iva_norm2(const cf* X, const cf* W, const cf* Y, const int* src, int n_src, float* r2, int B, int N, int I,
                   int J, cudaStream_t st);
// phi[b,s,j] = G'(r)/floor(2 r) (iva.py:1788-1789); Gauss uses variance[b,src[s],j] (iva.py:3273-3289);
// set_variance: variance = r2 / I first (update_source_model, iva.py:3465-3473; needs all sources)
int ssbk_iva_phi(const float* r2, float* variance, int set_variance, const int* src, int n_src, float* phi,
                 int model, int B, int N, int I, int J, int flooring, float eps, cudaStream_t st);
// loss[b] = sum_n mean_j G - 2 sum_i logdet (iva.py:215-220, :3093-3103, :3256-3271)
int ssbk_iva_loss(const float* r2, const float* variance, const double* logdet, double* loss, int model, int B,
                  int N, int I, int J, cudaStream_t st);

// ---- FastGaussMNMF (ssb_mnmf.cu, ssb_nmf.cu) --------------------------------------------------------
int ssbk_nmf_basis_ab(const float* A, const float* Bm, float* T, const float* V, int BN, int I, int J, int K,
                      int flooring, float eps, cudaStream_t st);
int ssbk_nmf_activation_ab(const float* A, const float* Bm, const float* T, float* V, int BN, int I, int J, int K,
                           int flooring, float eps, cudaStream_t st);
int ssbk_mnmf_gh(const cf* X, const float* T, const float* V, const float* Lam, const cf* Q, const float* D, float* G, float* H, int B,
                 int N, int I, int J, int K, cudaStream_t st);
int ssbk_mnmf_z2(const cf* X, const cf* Q, float* Z2, int B, int N, int I, int J, cudaStream_t st);
int ssbk_mnmf_phi(const cf* X, const float* T, const float* V, const float* Lam, const cf* Q, const float* D, float* phi, int B, int N,
                  int I, int J, int K, cudaStream_t st);
int ssbk_mnmf_spatial(const cf* X, const float* T, const float* V, const float* Lam, const cf* Q, float* D, double* zsum, int B, int N,
                      int I, int J, int K, int update_d, cudaStream_t st, float* z2out = nullptr);
int ssbk_mnmf_normalize(const double* zsum, cf* Q, float* D, int B, int N, int I, int J, int flooring, float eps,
                        cudaStream_t st, float* zscale = nullptr);
int ssbk_mnmf_rowloss(const cf* X, const float* T, const float* V, const float* Lam, const cf* Q, const float* D, double* rowloss, int B,
                      int N, int I, int J, int K, cudaStream_t st);
// Lleft (optional, [B,I,N,N] c128): Qinv <- Lleft Qinv before the filter (Q given in the whitened domain)
int ssbk_mnmf_separate(const cf* X, const float* T, const float* V, const cf* Q, const float* D, cd* Qinv, cf* Y, int B,
                       int N, int I, int J, int K, int ref, int flooring, float eps, cudaStream_t st,
                       const cd* Lleft = nullptr);

// ---- ssb_whiten.cu: whitened-domain iteration ---------------------------------------------------------------
// C64 = mean_j x x^H (fp64), C64 = L L^H, M = L^-1, Minv = L, ldM = log|det M|, Z = M X (complex64), wsync = 0
int ssbk_whiten_prepare(const cf* X, cd* C64, cd* M, cd* Minv, double* ldM, cf* Z, int* wsync, int B, int N, int I, int J,
                        cudaStream_t st);
// Ww = W Minv for the matrices whose W differs from Wexp (or that were never imported); Wexp <- W
int ssbk_w_import(const cf* W, cf* Wexp, cf* Ww, const cd* Minv, int* wsync, int n_mat, int N, cudaStream_t st);
// W = Ww M; Wexp <- W
int ssbk_w_export(const cf* Ww, const cd* M, cf* W, cf* Wexp, int* wsync, int n_mat, int N, cudaStream_t st);
// projection back on the whitened filter: s_n = (Minv Ww^-1)[ref, n], Ww[n,:] *= s_n, scale_out[mat,n] = s_n
int ssbk_pb_whitened(cf* Ww, const cd* Minv, cf* scale_out, int n_mat, int N, int ref, cudaStream_t st);
int ssbk_add_logdet(double* logdet, const double* ldM, int n, cudaStream_t st);
