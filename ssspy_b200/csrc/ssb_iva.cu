// AuxIVA source-model kernels: cross-bin norms r[b,n,j] = ||y_n[:,j]||_2, auxiliary weights
// phi = G'(r)/floor(2r) for the Laplace / Gauss contrasts, and the IVA loss.
#include "ssb_kernels.h"

namespace {

constexpr int NW = 8;

struct SrcList {
  int n;
  int idx[SSB_MAX_SOURCES];
};

// r2[b,s,j] = sum_i |y_{src[s]}[i,j]|^2.  One block per (32-frame tile, b); warps split the bins.
template <int N>
__global__ void __launch_bounds__(NW * 32) k_iva_norm2(const cf* __restrict__ X, const cf* __restrict__ W,
                                                       const cf* __restrict__ Y, SrcList src,
                                                       float* __restrict__ r2, int I, int J) {
  __shared__ float s_acc[N][32];
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * 32 + lane, b = blockIdx.y;
  const bool valid = j < J;
  const size_t cs = (size_t)I * J;
  float acc[N];
#pragma unroll
  for (int s = 0; s < N; ++s) acc[s] = 0.f;
  if (valid) {
    for (int i = w; i < I; i += NW) {
      const size_t base = ((size_t)b * N * I + i) * J + j;
      if (W) {
        cf x[N];
#pragma unroll
        for (int m = 0; m < N; ++m) x[m] = X[base + m * cs];
#pragma unroll
        for (int s = 0; s < N; ++s) {
          if (s < src.n) {
            const cf* wr = W + (((size_t)b * I + i) * N + src.idx[s]) * N;
            float yr = 0.f, yi = 0.f;
#pragma unroll
            for (int m = 0; m < N; ++m) {
              cf ww = __ldg(wr + m);
              yr = fmaf(ww.x, x[m].x, fmaf(-ww.y, x[m].y, yr));
              yi = fmaf(ww.x, x[m].y, fmaf(ww.y, x[m].x, yi));
            }
            acc[s] = fmaf(yr, yr, fmaf(yi, yi, acc[s]));
          }
        }
      } else {
#pragma unroll
        for (int s = 0; s < N; ++s) {
          if (s < src.n) {
            cf y = Y[base + (size_t)src.idx[s] * cs];
            acc[s] = fmaf(y.x, y.x, fmaf(y.y, y.y, acc[s]));
          }
        }
      }
    }
  }
  for (int ww = 0; ww < NW; ++ww) {
    if (w == ww) {
#pragma unroll
      for (int s = 0; s < N; ++s) {
        if (ww == 0) s_acc[s][lane] = acc[s];
        else s_acc[s][lane] += acc[s];
      }
    }
    __syncthreads();
  }
  if (w == 0 && valid) {
#pragma unroll
    for (int s = 0; s < N; ++s)
      if (s < src.n) r2[((size_t)b * src.n + s) * J + j] = s_acc[s][lane];
  }
}

__global__ void k_iva_phi(const float* __restrict__ r2, float* __restrict__ variance, int set_variance, SrcList src,
                          float* __restrict__ phi, int model, int B, int N, int I, int J, int flooring, float eps) {
  const size_t total = (size_t)B * src.n * J;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(e % J);
    const int s = (int)((e / J) % src.n);
    const int b = (int)(e / ((size_t)J * src.n));
    const float q = r2[e];
    const float r = sqrtf(q);
    float dG = 2.0f;
    if (model == SSB_MODEL_IVA_GAUSS) {
      const size_t vi = ((size_t)b * N + src.idx[s]) * J + j;
      float alpha;
      if (set_variance) {
        alpha = q / (float)I;
        variance[vi] = alpha;
      } else {
        alpha = variance[vi];
      }
      dG = 2.0f * r / alpha;
    }
    if (phi) phi[e] = dG / ssb_floor(2.0f * r, flooring, eps);
  }
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
  return r;
}

__global__ void k_iva_loss(const float* __restrict__ r2, const float* __restrict__ variance,
                           const double* __restrict__ logdet, double* __restrict__ loss, int model, int N, int I,
                           int J) {
  __shared__ double sh[32];
  const int b = blockIdx.x;
  double acc = 0.0;
  for (int e = threadIdx.x; e < N * J; e += blockDim.x) {
    const double q = (double)r2[(size_t)b * N * J + e];
    double G;
    if (model == SSB_MODEL_IVA_GAUSS) {
      const double a = (double)variance[(size_t)b * N * J + e];
      G = (double)I * log(a) + q / a;
    } else {
      G = 2.0 * sqrt(q);
    }
    acc += G;
  }
  acc /= (double)J;
  for (int i = threadIdx.x; i < I; i += blockDim.x) acc -= 2.0 * logdet[(size_t)b * I + i];
  acc = block_sum(acc, sh);
  if (threadIdx.x == 0) loss[b] = acc;
}

}  // namespace

int ssbk_iva_norm2(const cf* X, const cf* W, const cf* Y, const int* src, int n_src, float* r2, int B, int N, int I,
                   int J, cudaStream_t st) {
  SSB_REQUIRE(n_src >= 1 && n_src <= N, "iva_norm2: n_src=%d out of range", n_src);
  SrcList sl;
  sl.n = n_src;
  for (int s = 0; s < n_src; ++s) sl.idx[s] = src ? src[s] : s;
  dim3 grid((J + 31) / 32, B);
  SSB_DISPATCH_N(N, k_iva_norm2<NN><<<grid, NW * 32, 0, st>>>(X, W, Y, sl, r2, I, J));
  return ssb_check_launch("iva_norm2", st);
}

int ssbk_iva_phi(const float* r2, float* variance, int set_variance, const int* src, int n_src, float* phi,
                 int model, int B, int N, int I, int J, int flooring, float eps, cudaStream_t st) {
  SrcList sl;
  sl.n = n_src;
  for (int s = 0; s < n_src; ++s) sl.idx[s] = src ? src[s] : s;
  size_t total = (size_t)B * n_src * J;
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_iva_phi<<<blocks, 256, 0, st>>>(r2, variance, set_variance, sl, phi, model, B, N, I, J, flooring, eps);
  return ssb_check_launch("iva_phi", st);
}

int ssbk_iva_loss(const float* r2, const float* variance, const double* logdet, double* loss, int model, int B,
                  int N, int I, int J, cudaStream_t st) {
  k_iva_loss<<<B, 256, 0, st>>>(r2, variance, logdet, loss, model, N, I, J);
  return ssb_check_launch("iva_loss", st);
}
