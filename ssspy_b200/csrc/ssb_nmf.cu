// NMF source-model kernels of GaussILRMA (modular variants): MM / ME multiplicative updates of the
// basis T[B,N,I,K] and activation V[B,N,K,J] from the power spectrogram P = |Y|^2, the IP/ISS weight
// phi = (T V)^(-2/p), the per-row loss terms and the power / projection-back normalisation.
// K is padded to a compile-time KP so that T row / V column / numerator / denominator live in
// registers; contractions over K are plain FMA chains (the tensor-core variant is ssb_fused_*.cu).
#include "ssb_kernels.h"

namespace {

constexpr int WPB = 4;
constexpr int ACT_NW = 8;
// elementwise numerator factor A(P, R) of the multiplicative updates (the denominator factor is 1/R for all)
//   PM_MM2     Gauss MM, p = 2 : P / R^2                      exponent 1/2
//   PM_ME      Gauss ME (p = 2): P / R^2                      exponent 1
//   PM_GENERIC Gauss MM, any p : P / R^((p+2)/p)              exponent p/(p+2)
//   PM_T       Student-t       : P / (R~ R), R~ = c0 R^(2/p) + c1 P, c0 = nu/(nu+2), c1 = 2/(nu+2)
//   PM_GGD     GGD             : (beta/2) P^(beta/2) / R^((beta+p)/p)      exponent p/(beta+p)
enum { PM_MM2 = 0, PM_ME = 1, PM_GENERIC = 2, PM_T = 3, PM_GGD = 4 };

struct SrcParam {
  int mode;
  float aexp;  // PM_GENERIC: (p+2)/p;  PM_T: 2/p;  PM_GGD: (beta+p)/p
  float bexp;  // exponent of the ratio
  float c0, c1;  // PM_T: nu/(nu+2), 2/(nu+2);  PM_GGD: c0 = beta/2
};

__device__ __forceinline__ float upd_pow(float ratio, const SrcParam& sp) {
  if (sp.bexp == 0.5f) return sqrtf(ratio);
  if (sp.bexp == 1.0f) return ratio;
  return powf(ratio, sp.bexp);
}

__device__ __forceinline__ float src_factor(float P, float R, float inv, const SrcParam& sp) {
  switch (sp.mode) {
    case PM_MM2:
    case PM_ME: return P * inv * inv;
    case PM_GENERIC: return P / powf(R, sp.aexp);
    case PM_T: {
      const float r2p = sp.aexp == 1.0f ? R : powf(R, sp.aexp);
      return P * inv / fmaf(sp.c0, r2p, sp.c1 * P);
    }
    default: return sp.c0 * powf(P, sp.c0) / powf(R, sp.aexp);
  }
}


// T <- floor(T * (sum_j V P / R^a / sum_j V / R)^b); one warp per (b,n,i) row, lanes over frames.
template <int KP>
__global__ void __launch_bounds__(WPB * 32) k_nmf_basis(const float* __restrict__ P, float* __restrict__ T,
                                                        const float* __restrict__ V, int rows, int I, int J,
                                                        int K, SrcParam sp, int flooring, float eps, int vdiv,
                                                        float* __restrict__ num_out,
                                                        float* __restrict__ den_out) {
  const int row = blockIdx.x * WPB + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int bn = row / I;
  float t[KP], num[KP], den[KP];
#pragma unroll
  for (int k = 0; k < KP; ++k) {
    t[k] = k < K ? T[(size_t)row * K + k] : 0.f;
    num[k] = den[k] = 0.f;
  }
  const float* Vb = V + (size_t)(bn / vdiv) * K * J;  // vdiv = N: V shared by the sources (partitioning)
  const float* Pr = P + (size_t)row * J;
  for (int j = lane; j < J; j += 32) {
    float v[KP];
    float R = 0.f;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      v[k] = k < K ? Vb[(size_t)k * J + j] : 0.f;
      R = fmaf(t[k], v[k], R);
    }
    const float inv = 1.0f / R;
    const float p_ = Pr[j];
    const float A = src_factor(p_, R, inv, sp);
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      num[k] = fmaf(v[k], A, num[k]);
      den[k] = fmaf(v[k], inv, den[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < KP; ++k) {
    if (k < K) {
      float nu = warp_sum(num[k]), de = warp_sum(den[k]);
      if ((k & 31) == lane) {
        if (num_out) {  // raw sums for the partitioned updates (combined over sources / bins later)
          num_out[(size_t)row * K + k] = nu;
          den_out[(size_t)row * K + k] = de;
        } else {
          T[(size_t)row * K + k] = ssb_floor(upd_pow(nu / de, sp) * t[k], flooring, eps);
        }
      }
    }
  }
}

// V <- floor(V * (sum_i T P / R^a / sum_i T / R)^b); one block per (b,n, 32-frame tile), NW warps
// split the bins, lane = frame; cross-warp reduction through shared memory in fixed order.
template <int KP>
__global__ void __launch_bounds__(ACT_NW * 32) k_nmf_activation(const float* __restrict__ P,
                                                               const float* __restrict__ T, float* __restrict__ V,
                                                               int I, int J, int K, SrcParam sp, int flooring,
                                                               float eps, int vdiv, float* __restrict__ num_out,
                                                               float* __restrict__ den_out) {
  __shared__ float s_acc[2 * KP][32];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int jt = blockIdx.x, bn = blockIdx.y;
  const int j = jt * 32 + lane;
  const bool valid = j < J;
  float v[KP], num[KP], den[KP];
#pragma unroll
  for (int k = 0; k < KP; ++k) {
    v[k] = (valid && k < K) ? V[((size_t)(bn / vdiv) * K + k) * J + j] : 0.f;
    num[k] = den[k] = 0.f;
  }
  for (int i = w; i < I; i += ACT_NW) {
    const float* Tr = T + ((size_t)bn * I + i) * K;
    float t[KP];
    float R = 0.f;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      t[k] = k < K ? __ldg(Tr + k) : 0.f;
      R = fmaf(t[k], v[k], R);
    }
    if (!valid) R = 1.f;
    const float inv = 1.0f / R;
    const float p_ = valid ? P[((size_t)bn * I + i) * J + j] : 0.f;
    const float A = src_factor(p_, R, inv, sp);
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      num[k] = fmaf(t[k], A, num[k]);
      den[k] = fmaf(t[k], inv, den[k]);
    }
  }
  for (int ww = 0; ww < ACT_NW; ++ww) {
    if (w == ww) {
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        if (ww == 0) {
          s_acc[k][lane] = num[k];
          s_acc[KP + k][lane] = den[k];
        } else {
          s_acc[k][lane] += num[k];
          s_acc[KP + k][lane] += den[k];
        }
      }
    }
    __syncthreads();
  }
  if (w == 0 && valid) {
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      if (k < K) {
        if (num_out) {
          num_out[((size_t)bn * K + k) * J + j] = s_acc[k][lane];
          den_out[((size_t)bn * K + k) * J + j] = s_acc[KP + k][lane];
        } else {
          float ratio = s_acc[k][lane] / s_acc[KP + k][lane];
          V[((size_t)bn * K + k) * J + j] = ssb_floor(upd_pow(ratio, sp) * v[k], flooring, eps);
        }
      }
    }
  }
}

// Variants with the two elementwise factors given as arrays (FastGaussMNMF, ssspy/bss/mnmf.py:1351-1358,
// :1408-1415):  T <- floor(T sqrt(sum_j V A / sum_j V Bm)),  V <- floor(V sqrt(sum_i T A / sum_i T Bm)).
template <int KP>
__global__ void __launch_bounds__(WPB * 32) k_nmf_basis_ab(const float* __restrict__ A, const float* __restrict__ Bm,
                                                           float* __restrict__ T, const float* __restrict__ V,
                                                           int rows, int I, int J, int K, int flooring, float eps) {
  const int row = blockIdx.x * WPB + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int bn = row / I;
  float num[KP], den[KP];
#pragma unroll
  for (int k = 0; k < KP; ++k) num[k] = den[k] = 0.f;
  const float* Vb = V + (size_t)bn * K * J;
  for (int j = lane; j < J; j += 32) {
    const float a = A[(size_t)row * J + j], bb = Bm[(size_t)row * J + j];
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      const float v = k < K ? Vb[(size_t)k * J + j] : 0.f;
      num[k] = fmaf(v, a, num[k]);
      den[k] = fmaf(v, bb, den[k]);
    }
  }
#pragma unroll
  for (int k = 0; k < KP; ++k) {
    if (k < K) {
      float nu = warp_sum(num[k]), de = warp_sum(den[k]);
      if ((k & 31) == lane) {
        const float t = T[(size_t)row * K + k];
        T[(size_t)row * K + k] = ssb_floor(t * sqrtf(nu / de), flooring, eps);
      }
    }
  }
}

template <int KP>
__global__ void __launch_bounds__(ACT_NW * 32) k_nmf_activation_ab(const float* __restrict__ A,
                                                                  const float* __restrict__ Bm,
                                                                  const float* __restrict__ T, float* __restrict__ V,
                                                                  int I, int J, int K, int flooring, float eps) {
  __shared__ float s_acc[2 * KP][32];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int jt = blockIdx.x, bn = blockIdx.y;
  const int j = jt * 32 + lane;
  const bool valid = j < J;
  float num[KP], den[KP];
#pragma unroll
  for (int k = 0; k < KP; ++k) num[k] = den[k] = 0.f;
  for (int i = w; i < I; i += ACT_NW) {
    const float* Tr = T + ((size_t)bn * I + i) * K;
    const float a = valid ? A[((size_t)bn * I + i) * J + j] : 0.f;
    const float bb = valid ? Bm[((size_t)bn * I + i) * J + j] : 0.f;
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      const float t = k < K ? __ldg(Tr + k) : 0.f;
      num[k] = fmaf(t, a, num[k]);
      den[k] = fmaf(t, bb, den[k]);
    }
  }
  for (int ww = 0; ww < ACT_NW; ++ww) {
    if (w == ww) {
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        if (ww == 0) {
          s_acc[k][lane] = num[k];
          s_acc[KP + k][lane] = den[k];
        } else {
          s_acc[k][lane] += num[k];
          s_acc[KP + k][lane] += den[k];
        }
      }
    }
    __syncthreads();
  }
  if (w == 0 && valid) {
#pragma unroll
    for (int k = 0; k < KP; ++k) {
      if (k < K) {
        const float v = V[((size_t)bn * K + k) * J + j];
        V[((size_t)bn * K + k) * J + j] = ssb_floor(v * sqrtf(s_acc[k][lane] / s_acc[KP + k][lane]), flooring, eps);
      }
    }
  }
}

// phi = (T V)^(-2/p) (Gauss), 1/R~ (Student-t), 1/((2/beta) floor(P^((2-beta)/2)) R^(beta/p)) (GGD); one warp per
// row.  P and phi may be the same buffer (each element is read then written by the same thread).
template <int KP>
__global__ void __launch_bounds__(WPB * 32) k_nmf_phi(const float* __restrict__ T, const float* __restrict__ V,
                                                      const float* P, float* phi, int rows, int I, int J, int K,
                                                      float p, int model, float prm, int flooring, float eps,
                                                      int vdiv) {
  const int row = blockIdx.x * WPB + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int bn = row / I;
  float t[KP];
#pragma unroll
  for (int k = 0; k < KP; ++k) t[k] = k < K ? T[(size_t)row * K + k] : 0.f;
  const float* Vb = V + (size_t)(bn / vdiv) * K * J;
  const bool p2 = (p == 2.0f);
  const float e = -2.0f / p;
  for (int j = lane; j < J; j += 32) {
    float R = 0.f;
#pragma unroll
    for (int k = 0; k < KP; ++k)
      if (k < K) R = fmaf(t[k], Vb[(size_t)k * J + j], R);
    float out;
    if (model == SSB_MODEL_ILRMA_GAUSS) {
      out = p2 ? 1.0f / R : powf(R, e);
    } else {
      const float pw = P[(size_t)row * J + j];
      if (model == SSB_MODEL_ILRMA_T) {
        const float c0 = prm / (prm + 2.0f);
        out = 1.0f / fmaf(c0, p2 ? R : powf(R, -e), (1.0f - c0) * pw);
      } else {
        out = 1.0f / ((2.0f / prm) * ssb_floor(powf(pw, 0.5f * (2.0f - prm)), flooring, eps) * powf(R, prm / p));
      }
    }
    phi[(size_t)row * J + j] = out;
  }
}

// rowloss[row] = mean_j( P / R^(2/p) + (2/p) log R )
template <int KP>
__global__ void __launch_bounds__(WPB * 32) k_nmf_rowloss(const float* __restrict__ P, const float* __restrict__ T,
                                                          const float* __restrict__ V, double* __restrict__ rowloss,
                                                          int rows, int I, int J, int K, float p, int model,
                                                          float prm, int vdiv) {
  const int row = blockIdx.x * WPB + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int bn = row / I;
  float t[KP];
#pragma unroll
  for (int k = 0; k < KP; ++k) t[k] = k < K ? T[(size_t)row * K + k] : 0.f;
  const float* Vb = V + (size_t)(bn / vdiv) * K * J;
  const bool p2 = (p == 2.0f);
  const float e = 2.0f / p;
  double acc = 0.0;
  for (int j = lane; j < J; j += 32) {
    float R = 0.f;
#pragma unroll
    for (int k = 0; k < KP; ++k)
      if (k < K) R = fmaf(t[k], Vb[(size_t)k * J + j], R);
    const float pw = P[(size_t)row * J + j];
    float term;
    if (model == SSB_MODEL_ILRMA_GAUSS) {
      term = p2 ? pw / R + logf(R) : pw / powf(R, e) + e * logf(R);
    } else if (model == SSB_MODEL_ILRMA_T) {
      term = (1.0f + 0.5f * prm) * log1pf((2.0f / prm) * pw / (p2 ? R : powf(R, e))) + e * logf(R);
    } else {
      term = powf(pw, 0.5f * prm) / powf(R, prm / p) + e * logf(R);
    }
    acc += (double)term;
  }
  acc = warp_sum(acc);
  if (lane == 0) rowloss[row] = acc / (double)J;
}

__device__ __forceinline__ double block_sum(double v, double* sh) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = lane < (int)(blockDim.x >> 5) ? sh[lane] : 0.0;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;  // valid on warp 0
}

__global__ void k_ilrma_loss_reduce(const double* __restrict__ rowloss, const double* __restrict__ logdet,
                                    double* __restrict__ loss, int N, int I) {
  __shared__ double sh[32];
  const int b = blockIdx.x;
  double acc = 0.0;
  for (int e = threadIdx.x; e < N * I; e += blockDim.x) acc += rowloss[(size_t)b * N * I + e];
  for (int i = threadIdx.x; i < I; i += blockDim.x) acc -= 2.0 * logdet[(size_t)b * I + i];
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) loss[b] = acc;
}

// psi2[b,n] = mean_i Re( w_in^H-row C_i (w_in^H-row)^H ),  C_i = mean_j x x^H    
__global__ void k_psi_from_cov(const cf* __restrict__ W, const cf* __restrict__ C, double* __restrict__ psi2, int N,
                               int I) {
  __shared__ double sh[32];
  const int n = blockIdx.x, b = blockIdx.y;
  double acc = 0.0;
  for (int i = threadIdx.x; i < I; i += blockDim.x) {
    const cf* w = W + (((size_t)b * I + i) * N + n) * N;
    const cf* c = C + ((size_t)b * I + i) * N * N;
    double s = 0.0;
    for (int a = 0; a < N; ++a) {
      cd wa = cf2cd(w[a]);
      for (int cc = 0; cc < N; ++cc) {
        cd t = cd_mul(wa, cf2cd(c[a * N + cc]));
        s += cd_mulc(t, cf2cd(w[cc])).x;
      }
    }
    acc += s;
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) psi2[b * N + n] = acc / (double)I;
}

__global__ void k_psi_from_y(const cf* __restrict__ Y, double* __restrict__ psi2, long long per_src) {
  __shared__ double sh[32];
  const int bn = blockIdx.x;
  const cf* y = Y + (size_t)bn * per_src;
  double acc = 0.0;
  for (long long e = threadIdx.x; e < per_src; e += blockDim.x) {
    cf v = y[e];
    acc += (double)(v.x * v.x + v.y * v.y);
  }
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) psi2[bn] = acc / (double)per_src;
}

// T[b,n,i,k] /= psi^p ; W[b,i,n,:] /= psi ; Y[b,n,:,:] /= psi      (ilrma.py:434-444)
__global__ void k_apply_psi(const double* __restrict__ psi2, float* __restrict__ T, cf* __restrict__ W,
                            cf* __restrict__ Y, int B, int N, int I, int J, int K, float p, int flooring, double eps) {
  const size_t nT = T ? (size_t)B * N * I * K : 0;
  const size_t nW = W ? (size_t)B * I * N * N : 0;
  const size_t nY = Y ? (size_t)B * N * I * J : 0;
  const size_t total = nT + nW + nY;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    if (e < nT) {
      const int bn = (int)(e / ((size_t)I * K));
      const double psi = ssb_floor(sqrt(psi2[bn]), flooring, eps);
      const double sc = (p == 2.0f) ? psi * psi : pow(psi, (double)p);
      T[e] = (float)((double)T[e] / sc);
    } else if (e < nT + nW) {
      const size_t q = e - nT;
      const int n = (int)((q / N) % N);
      const int b = (int)(q / ((size_t)I * N * N));
      const double psi = ssb_floor(sqrt(psi2[b * N + n]), flooring, eps);
      cf w = W[q];
      W[q] = make_float2((float)(w.x / psi), (float)(w.y / psi));
    } else {
      const size_t q = e - nT - nW;
      const int bn = (int)(q / ((size_t)I * J));
      const double psi = ssb_floor(sqrt(psi2[bn]), flooring, eps);
      cf y = Y[q];
      Y[q] = make_float2((float)(y.x / psi), (float)(y.y / psi));
    }
  }
}

__global__ void k_scale_basis(float* __restrict__ T, const cf* __restrict__ s, long long s_mat, long long s_src,
                              int N, int I, int K, float p, size_t total) {
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t row = e / K;  // (b, n, i)
    const int i = (int)(row % I);
    const int n = (int)((row / I) % N);
    const int b = (int)(row / ((size_t)I * N));
    cf sv = s[((size_t)b * I + i) * s_mat + (size_t)n * s_src];
    double a = sqrt((double)sv.x * sv.x + (double)sv.y * sv.y);
    double sc = (p == 2.0f) ? a * a : pow(a, (double)p);
    T[e] = (float)((double)T[e] * sc);
  }
}

// ---- partitioning function (latent Z[B,N,K], shared T[B,I,K], V[B,K,J]; ilrma.py:201-245, :297-331) ----------
// The per-source model is R_n = Teff_n V with Teff[b,n,i,k] = z_nk t_ik, so the sweeps above are reused with
// Teff as the basis and V shared (vdiv = N); they emit the raw sums and the kernels below combine them.
__global__ void k_part_teff(const float* __restrict__ Z, const float* __restrict__ T, float* __restrict__ Teff, int N,
                            int I, int K, size_t total) {
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(e % K);
    const int i = (int)((e / K) % I);
    const size_t bn = e / ((size_t)K * I);
    const size_t b = bn / N;
    Teff[e] = Z[bn * K + k] * T[(b * I + i) * K + k];
  }
}

// Z <- Z (sum_i t_ik gnum[n,i,k] / sum_i t_ik gden[n,i,k])^b, then Z /= sum_n Z (ilrma.py:1007-1049): one block per
// mixture, one warp per (n, k)
__global__ void __launch_bounds__(256) k_part_latent(const float* __restrict__ gnum, const float* __restrict__ gden,
                                                     const float* __restrict__ T, float* __restrict__ Z, int N, int I,
                                                     int K, float bexp) {
  __shared__ float zs[SSB_MAX_SOURCES * SSB_MAX_BASIS];
  const int b = blockIdx.x, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int pair = w; pair < N * K; pair += 8) {
    const int n = pair / K, k = pair - n * K;
    float nu = 0.f, de = 0.f;
    for (int i = lane; i < I; i += 32) {
      const float t = T[((size_t)b * I + i) * K + k];
      const size_t g = (((size_t)b * N + n) * I + i) * K + k;
      nu = fmaf(t, gnum[g], nu);
      de = fmaf(t, gden[g], de);
    }
    nu = warp_sum(nu);
    de = warp_sum(de);
    if (lane == 0) {
      const float r = nu / de;
      zs[pair] = Z[((size_t)b * N + n) * K + k] * (bexp == 0.5f ? sqrtf(r) : bexp == 1.0f ? r : powf(r, bexp));
    }
  }
  __syncthreads();
  for (int pair = threadIdx.x; pair < N * K; pair += blockDim.x) {
    const int k = pair % K;
    float s = 0.f;
    for (int n = 0; n < N; ++n) s += zs[n * K + k];
    Z[(size_t)b * N * K + pair] = zs[pair] / s;
  }
}

// T[b,i,k] <- floor(T (sum_n z_nk gnum / sum_n z_nk gden)^b)   (ilrma.py:1098-1126)
__global__ void k_part_basis(const float* __restrict__ gnum, const float* __restrict__ gden,
                             const float* __restrict__ Z, float* __restrict__ T, int N, int I, int K, float bexp,
                             int flooring, float eps, size_t total) {
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(e % K);
    const int i = (int)((e / K) % I);
    const size_t b = e / ((size_t)K * I);
    float nu = 0.f, de = 0.f;
    for (int n = 0; n < N; ++n) {
      const float z = Z[(b * N + n) * K + k];
      const size_t g = ((b * N + n) * I + i) * K + k;
      nu = fmaf(z, gnum[g], nu);
      de = fmaf(z, gden[g], de);
    }
    const float r = nu / de;
    T[e] = ssb_floor(T[e] * (bexp == 0.5f ? sqrtf(r) : bexp == 1.0f ? r : powf(r, bexp)), flooring, eps);
  }
}

// V[b,k,j] <- floor(V (sum_n hnum[n,k,j] / sum_n hden[n,k,j])^b), hnum = sum_i z_nk t_ik A (ilrma.py:1174-1202)
__global__ void k_part_activation(const float* __restrict__ hnum, const float* __restrict__ hden,
                                  float* __restrict__ V, int N, int K, int J, float bexp, int flooring, float eps,
                                  size_t total) {
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t kj = e % ((size_t)K * J);
    const size_t b = e / ((size_t)K * J);
    float nu = 0.f, de = 0.f;
    for (int n = 0; n < N; ++n) {
      const size_t g = (b * N + n) * (size_t)K * J + kj;
      nu += hnum[g];
      de += hden[g];
    }
    const float r = nu / de;
    V[e] = ssb_floor(V[e] * (bexp == 0.5f ? sqrtf(r) : bexp == 1.0f ? r : powf(r, bexp)), flooring, eps);
  }
}

// power normalisation with the partitioning function (ilrma.py:424-430): Zp = Z / psi^p, scale_k = sum_n Zp,
// T[:,k] *= scale_k, Z = Zp / scale_k.  One block per mixture.
__global__ void __launch_bounds__(256) k_part_normalize(const double* __restrict__ psi2, float* __restrict__ Z,
                                                        float* __restrict__ T, int N, int I, int K, float p,
                                                        int flooring, double eps) {
  __shared__ double zp[SSB_MAX_SOURCES * SSB_MAX_BASIS];
  __shared__ double scale[SSB_MAX_BASIS];
  const int b = blockIdx.x;
  for (int pair = threadIdx.x; pair < N * K; pair += blockDim.x) {
    const int n = pair / K;
    const double psi = ssb_floor(sqrt(psi2[b * N + n]), flooring, eps);
    zp[pair] = (double)Z[(size_t)b * N * K + pair] / ((p == 2.0f) ? psi * psi : pow(psi, (double)p));
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    double s = 0.0;
    for (int n = 0; n < N; ++n) s += zp[n * K + k];
    scale[k] = s;
  }
  __syncthreads();
  for (int pair = threadIdx.x; pair < N * K; pair += blockDim.x)
    Z[(size_t)b * N * K + pair] = (float)(zp[pair] / scale[pair % K]);
  for (int e = threadIdx.x; e < I * K; e += blockDim.x) {
    float* t = T + (size_t)b * I * K + e;
    *t = (float)((double)*t * scale[e % K]);
  }
}

SrcParam src_param(float p, int source, int model, float prm) {
  SrcParam sp{};
  const bool me = source == SSB_SOURCE_ME;
  if (model == SSB_MODEL_ILRMA_T) {
    sp.mode = PM_T;
    sp.aexp = 2.0f / p;
    sp.bexp = me ? 1.0f : p / (p + 2.0f);
    sp.c0 = prm / (prm + 2.0f);
    sp.c1 = 2.0f / (prm + 2.0f);
  } else if (model == SSB_MODEL_ILRMA_GGD) {
    sp.mode = PM_GGD;
    sp.aexp = (prm + p) / p;
    sp.bexp = p / (prm + p);
    sp.c0 = 0.5f * prm;
  } else {
    sp.mode = me ? PM_ME : (p == 2.0f ? PM_MM2 : PM_GENERIC);
    sp.aexp = (p + 2.0f) / p;
    sp.bexp = me ? 1.0f : p / (p + 2.0f);
  }
  return sp;
}

}  // namespace

int ssbk_nmf_basis(const float* P, float* T, const float* V, int BN, int I, int J, int K, float p, int source,
                   int model, float prm, int flooring, float eps, cudaStream_t st, int vdiv, float* num_out,
                   float* den_out) {
  const int rows = BN * I;
  const SrcParam sp = src_param(p, source, model, prm);
  SSB_DISPATCH_K(K, k_nmf_basis<KP><<<blocks_for(rows, WPB), WPB * 32, 0, st>>>(P, T, V, rows, I, J, K, sp, flooring,
                                                                                eps, vdiv, num_out, den_out));
  return ssb_check_launch("nmf_basis", st);
}

int ssbk_nmf_activation(const float* P, const float* T, float* V, int BN, int I, int J, int K, float p, int source,
                        int model, float prm, int flooring, float eps, cudaStream_t st, int vdiv, float* num_out,
                        float* den_out) {
  const SrcParam sp = src_param(p, source, model, prm);
  dim3 grid((J + 31) / 32, BN);
  SSB_DISPATCH_K(K, k_nmf_activation<KP><<<grid, ACT_NW * 32, 0, st>>>(P, T, V, I, J, K, sp, flooring, eps, vdiv,
                                                                       num_out, den_out));
  return ssb_check_launch("nmf_activation", st);
}

int ssbk_nmf_basis_ab(const float* A, const float* Bm, float* T, const float* V, int BN, int I, int J, int K,
                      int flooring, float eps, cudaStream_t st) {
  const int rows = BN * I;
  SSB_DISPATCH_K(K, k_nmf_basis_ab<KP><<<blocks_for(rows, WPB), WPB * 32, 0, st>>>(A, Bm, T, V, rows, I, J, K, flooring,
                                                                                    eps));
  return ssb_check_launch("nmf_basis_ab", st);
}

int ssbk_nmf_activation_ab(const float* A, const float* Bm, const float* T, float* V, int BN, int I, int J, int K,
                           int flooring, float eps, cudaStream_t st) {
  dim3 grid((J + 31) / 32, BN);
  SSB_DISPATCH_K(K, k_nmf_activation_ab<KP><<<grid, ACT_NW * 32, 0, st>>>(A, Bm, T, V, I, J, K, flooring, eps));
  return ssb_check_launch("nmf_activation_ab", st);
}

int ssbk_nmf_phi(const float* T, const float* V, const float* P, float* phi, int BN, int I, int J, int K, float p,
                 int model, float prm, int flooring, float eps, cudaStream_t st, int vdiv) {
  const int rows = BN * I;
  SSB_DISPATCH_K(K, k_nmf_phi<KP><<<blocks_for(rows, WPB), WPB * 32, 0, st>>>(T, V, P, phi, rows, I, J, K, p, model,
                                                                              prm, flooring, eps, vdiv));
  return ssb_check_launch("nmf_phi", st);
}

int ssbk_nmf_rowloss(const float* P, const float* T, const float* V, double* rowloss, int BN, int I, int J, int K,
                     float p, int model, float prm, cudaStream_t st, int vdiv) {
  const int rows = BN * I;
  SSB_DISPATCH_K(K, k_nmf_rowloss<KP><<<blocks_for(rows, WPB), WPB * 32, 0, st>>>(P, T, V, rowloss, rows, I, J, K, p,
                                                                                  model, prm, vdiv));
  return ssb_check_launch("nmf_rowloss", st);
}

static int grid_for(size_t total) {
  size_t blocks = (total + 255) / 256;
  return (int)(blocks > 148 * 16 ? 148 * 16 : (blocks ? blocks : 1));
}

int ssbk_part_teff(const float* Z, const float* T, float* Teff, int B, int N, int I, int K, cudaStream_t st) {
  const size_t total = (size_t)B * N * I * K;
  k_part_teff<<<grid_for(total), 256, 0, st>>>(Z, T, Teff, N, I, K, total);
  return ssb_check_launch("part_teff", st);
}

int ssbk_part_latent(const float* gnum, const float* gden, const float* T, float* Z, int B, int N, int I, int K,
                     float p, int source, int model, float prm, cudaStream_t st) {
  k_part_latent<<<B, 256, 0, st>>>(gnum, gden, T, Z, N, I, K, src_param(p, source, model, prm).bexp);
  return ssb_check_launch("part_latent", st);
}

int ssbk_part_basis(const float* gnum, const float* gden, const float* Z, float* T, int B, int N, int I, int K, float p,
                    int source, int model, float prm, int flooring, float eps, cudaStream_t st) {
  const size_t total = (size_t)B * I * K;
  k_part_basis<<<grid_for(total), 256, 0, st>>>(gnum, gden, Z, T, N, I, K, src_param(p, source, model, prm).bexp,
                                                 flooring, eps, total);
  return ssb_check_launch("part_basis", st);
}

int ssbk_part_activation(const float* hnum, const float* hden, float* V, int B, int N, int K, int J, float p,
                         int source, int model, float prm, int flooring, float eps, cudaStream_t st) {
  const size_t total = (size_t)B * K * J;
  k_part_activation<<<grid_for(total), 256, 0, st>>>(hnum, hden, V, N, K, J, src_param(p, source, model, prm).bexp,
                                                      flooring, eps, total);
  return ssb_check_launch("part_activation", st);
}

int ssbk_part_normalize(const double* psi2, float* Z, float* T, int B, int N, int I, int K, float p, int flooring,
                        float eps, cudaStream_t st) {
  k_part_normalize<<<B, 256, 0, st>>>(psi2, Z, T, N, I, K, p, flooring, (double)eps);
  return ssb_check_launch("part_normalize", st);
}

int ssbk_ilrma_loss_reduce(const double* rowloss, const double* logdet, double* loss, int B, int N, int I,
                           cudaStream_t st) {
  k_ilrma_loss_reduce<<<B, 256, 0, st>>>(rowloss, logdet, loss, N, I);
  return ssb_check_launch("ilrma_loss_reduce", st);
}

int ssbk_psi_from_cov(const cf* W, const cf* C, double* psi2, int B, int N, int I, cudaStream_t st) {
  dim3 grid(N, B);
  k_psi_from_cov<<<grid, 256, 0, st>>>(W, C, psi2, N, I);
  return ssb_check_launch("psi_from_cov", st);
}

int ssbk_psi_from_y(const cf* Y, double* psi2, int B, int N, int I, int J, cudaStream_t st) {
  k_psi_from_y<<<B * N, 1024, 0, st>>>(Y, psi2, (long long)I * J);
  return ssb_check_launch("psi_from_y", st);
}

int ssbk_apply_psi(const double* psi2, float* T, cf* W, cf* Y, int B, int N, int I, int J, int K, float p,
                   int flooring, float eps, cudaStream_t st) {
  size_t total = (T ? (size_t)B * N * I * K : 0) + (W ? (size_t)B * I * N * N : 0) + (Y ? (size_t)B * N * I * J : 0);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_apply_psi<<<blocks, 256, 0, st>>>(psi2, T, W, Y, B, N, I, J, K, p, flooring, (double)eps);
  return ssb_check_launch("apply_psi", st);
}

int ssbk_scale_basis(float* T, const cf* s, long long s_mat_stride, long long s_src_stride, int B, int N, int I,
                     int K, float p, cudaStream_t st) {
  size_t total = (size_t)B * N * I * K;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_scale_basis<<<blocks, 256, 0, st>>>(T, s, s_mat_stride, s_src_stride, N, I, K, p, total);
  return ssb_check_launch("scale_basis", st);
}

// ------------------------------------------------------------------------------------------------
// ILRMABase.reconstruct_nmf (ssspy/bss/ilrma.py:297-328): one thread per (mixture, source, bin, frame)
__global__ void k_reconstruct_nmf(const float* __restrict__ T, const float* __restrict__ V, const float* __restrict__ Z,
                                  float* __restrict__ R, int N, int I, int J, int K, size_t total) {
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int j = (int)(idx % J);
  const int i = (int)((idx / J) % I);
  const size_t bn = idx / ((size_t)I * J);
  float acc = 0.f;
  if (Z == nullptr) {
    const float* t = T + (bn * I + i) * K;
    const float* v = V + bn * (size_t)K * J + j;
    for (int k = 0; k < K; ++k) acc = fmaf(t[k], v[(size_t)k * J], acc);
  } else {
    const size_t b = bn / N;
    const float* z = Z + bn * K;
    const float* t = T + (b * I + i) * K;
    const float* v = V + b * (size_t)K * J + j;
    for (int k = 0; k < K; ++k) acc = fmaf(z[k] * t[k], v[(size_t)k * J], acc);
  }
  R[idx] = acc;
}

extern "C" int ssb_reconstruct_nmf(const float* T, const float* V, const float* Z, float* R, int B, int N, int I, int J,
                                   int K, void* stream) {
  SSB_REQUIRE(T != nullptr && V != nullptr && R != nullptr, "reconstruct_nmf: NULL argument");
  SSB_REQUIRE(B >= 0 && N >= 1 && I >= 1 && J >= 1 && K >= 1, "reconstruct_nmf: invalid shape");
  const size_t total = (size_t)B * N * I * J;
  if (total == 0) return 0;
  k_reconstruct_nmf<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(T, V, Z, R, N, I, J, K, total);
  return ssb_check_launch("reconstruct_nmf", (cudaStream_t)stream);
}

