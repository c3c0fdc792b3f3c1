// FastGaussMNMF (ssspy/bss/mnmf.py:1076-1675), determined case n_sources == n_channels == N.
// State per mixture: T[N,I,K], V[N,K,J] (NMF), Q[I,N,N] complex64 (diagonaliser, rows q_m^H),
// D[I,N,N] f32 (spatial[i, source n, channel m]).  With Lambda = T V, L[i,j,m] = sum_n Lambda_n D[n,m],
// Z2[i,j,m] = |q_m^H x|^2:
//   basis / activation : multiplicative updates with G_n = sum_m D[n,m] Z2/L^2, H_n = sum_m D[n,m]/L
//   diagonaliser       : U[i,m] = mean_j x x^H / L[:,m], then the same IP1 / IP2 as ILRMA on Q
//   spatial            : D <- D sqrt(sum_j Lambda Z2/L^2 / sum_j Lambda/L)
//   normalisation, loss, and the multichannel Wiener filter with a per-(bin, frame) Hermitian
//   eigendecomposition (to_psd) in `separate`.
// One warp per (mixture, bin) with lanes over frames for the update kernels (correctness-first
// CUDA-core kernels; the K-contractions are plain FMA chains), one thread per (bin, frame) for the
// Wiener filter.
#include "ssb_group.cuh"
#include "ssb_kernels.h"

namespace {

constexpr int MW = 4;  // warps (bins) per block

// per-warp shared-memory context of one bin
template <int N>
struct BinCtx {
  float T[N][SSB_MAX_BASIS];
  cf Q[N][N];
  float D[N][N];
};

template <int N>
__device__ __forceinline__ void load_ctx(BinCtx<N>& c, const float* __restrict__ T, const cf* __restrict__ Q,
                                         const float* __restrict__ D, int b, int i, int I, int K, int lane) {
  for (int e = lane; e < N * K; e += 32) {
    const int n = e / K, k = e - n * K;
    c.T[n][k] = T[(((size_t)b * N + n) * I + i) * K + k];
  }
  for (int e = lane; e < N * N; e += 32) {
    c.Q[e / N][e % N] = Q[((size_t)b * I + i) * N * N + e];
    c.D[e / N][e % N] = D[((size_t)b * I + i) * N * N + e];
  }
  __syncwarp();
}

// Q and D of the bin in registers: they are loop invariants of the frame sweeps (48 registers at N = 4); reading them
// from shared memory in every frame cost one LDS per FMA operand (km_spatial ran at 20 % issue utilisation)
template <int N>
struct RegCtx {
  static constexpr bool REG = N <= 4;  // larger N: the copy would spill, keep reading shared memory
  static constexpr int NR = REG ? N : 1;
  cf Qr[NR][NR];
  float Dr[NR][NR];
  const BinCtx<N>& s;
  const float (*T)[SSB_MAX_BASIS];
  __device__ __forceinline__ explicit RegCtx(const BinCtx<N>& c) : s(c), T(c.T) {
    if (REG) {
#pragma unroll
      for (int a = 0; a < NR; ++a)
#pragma unroll
        for (int b = 0; b < NR; ++b) {
          Qr[a][b] = c.Q[a][b];
          Dr[a][b] = c.D[a][b];
        }
    }
  }
  __device__ __forceinline__ cf q(int a, int b) const { return REG ? Qr[REG ? a : 0][REG ? b : 0] : s.Q[a][b]; }
  __device__ __forceinline__ float d(int a, int b) const { return REG ? Dr[REG ? a : 0][REG ? b : 0] : s.D[a][b]; }
};

// Row pointers of one (mixture, bin): computed once per warp, so that the frame loops index with `plane * cs + j` only
// (the index arithmetic of the round-1 kernels, recomputed from (b, n, i, j) for every load, was most of their
// instructions: 767 IMAD + 737 IADD3 + 455 LEA against 677 FFMA in km_spatial<4>)
struct RowBase {
  const cf* x;        // X[b, 0, i, 0]
  const float* lam;   // Lam[b, 0, i, 0] or NULL
  const float* v;     // V[b, 0, 0, 0]
  size_t cs;          // I * J: stride between the channel / source planes of X and Lam
  size_t vs;          // K * J: stride between the sources of V
  int J, K;
};
__device__ __forceinline__ RowBase row_base(const cf* X, const float* V, const float* Lam, int b, int i, int N, int I,
                                            int J, int K) {
  RowBase r;
  const size_t o = ((size_t)b * N * I + i) * J;
  r.x = X + o;
  r.lam = Lam != nullptr ? Lam + o : nullptr;
  r.v = V + (size_t)b * N * K * J;
  r.cs = (size_t)I * J;
  r.vs = (size_t)K * J;
  r.J = J;
  r.K = K;
  return r;
}
__device__ __forceinline__ float mn_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// Lambda_n, L_m and Z2_m of one (bin, frame)
template <int N, bool LAM>
__device__ __forceinline__ void frame_stats(const RegCtx<N>& c, const RowBase& rb, int j, float (&lam)[N],
                                            float (&L)[N], float (&Z2)[N]) {
  cf x[N];
#pragma unroll
  for (int m = 0; m < N; ++m) x[m] = rb.x[m * rb.cs + j];
#pragma unroll
  for (int n = 0; n < N; ++n) {
    if (LAM) {  // Lambda = T V precomputed on the tensor pipe (compile-time: keeps the K loop out of the hot version)
      lam[n] = rb.lam[n * rb.cs + j];
    } else {
      const float* v = rb.v + n * rb.vs + j;
      float s = 0.f;
      for (int k = 0; k < rb.K; ++k) s = fmaf(c.T[n][k], v[(size_t)k * rb.J], s);
      lam[n] = s;
    }
  }
#pragma unroll
  for (int m = 0; m < N; ++m) {
    float l = 0.f, zr = 0.f, zi = 0.f;
#pragma unroll
    for (int n = 0; n < N; ++n) l = fmaf(lam[n], c.d(n, m), l);
#pragma unroll
    for (int cc = 0; cc < N; ++cc) {
      const cf q = c.q(m, cc);
      zr = fmaf(q.x, x[cc].x, fmaf(-q.y, x[cc].y, zr));
      zi = fmaf(q.x, x[cc].y, fmaf(q.y, x[cc].x, zi));
    }
    L[m] = l;
    Z2[m] = zr * zr + zi * zi;
  }
}

// G[b,n,i,j] = sum_m D[n,m] Z2_m / L_m^2,  H[b,n,i,j] = sum_m D[n,m] / L_m     (mnmf.py:1348-1350)
template <int N, bool LAM>
__global__ void __launch_bounds__(MW * 32) km_gh(const cf* __restrict__ X, const float* __restrict__ T,
                                                 const float* __restrict__ V, const float* __restrict__ Lam,
                                                 const cf* __restrict__ Q,
                                                 const float* __restrict__ D, float* __restrict__ G,
                                                 float* __restrict__ H, int B, int I, int J, int K) {
  __shared__ BinCtx<N> ctx[MW];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bin = blockIdx.x * MW + wib;
  if (bin >= B * I) return;
  const int b = bin / I, i = bin - b * I;
  BinCtx<N>& cs = ctx[wib];
  load_ctx<N>(cs, T, Q, D, b, i, I, K, lane);
  const RegCtx<N> c(cs);
  const RowBase rb = row_base(X, V, Lam, b, i, N, I, J, K);
  const size_t gh0 = ((size_t)b * N * I + i) * J;
#pragma unroll 2
  for (int j = lane; j < J; j += 32) {
    float lam[N], L[N], Z2[N];
    frame_stats<N, LAM>(c, rb, j, lam, L, Z2);
    float r[N], r2[N];
#pragma unroll
    for (int m = 0; m < N; ++m) {
      r[m] = mn_rcp(L[m]);
      r2[m] = Z2[m] * r[m] * r[m];
    }
#pragma unroll
    for (int n = 0; n < N; ++n) {
      float g = 0.f, h = 0.f;
#pragma unroll
      for (int m = 0; m < N; ++m) {
        g = fmaf(c.d(n, m), r2[m], g);
        h = fmaf(c.d(n, m), r[m], h);
      }
      const size_t o = gh0 + n * rb.cs + j;
      G[o] = g;
      H[o] = h;
    }
  }
}

// Z2[b,m,i,j] = |q_m^H x|^2: the only per-point input of the fused source-model kernels (kf_mnmf_update, ssb_coop.cu)
template <int N>
__global__ void __launch_bounds__(MW * 32) km_z2(const cf* __restrict__ X, const cf* __restrict__ Q,
                                                 float* __restrict__ Z2, int B, int I, int J) {
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bin = blockIdx.x * MW + wib;
  if (bin >= B * I) return;
  const int b = bin / I, i = bin - b * I;
  cf q[N][N];
#pragma unroll
  for (int m = 0; m < N; ++m)
#pragma unroll
    for (int c = 0; c < N; ++c) q[m][c] = Q[((size_t)b * I + i) * N * N + m * N + c];
  const size_t o = ((size_t)b * N * I + i) * J, cs = (size_t)I * J;
  for (int j = 2 * lane; j < J; j += 64) {  // two frames per lane: 16-byte loads, 8-byte stores (J is even here)
    float4 x[N];
#pragma unroll
    for (int c = 0; c < N; ++c) x[c] = *reinterpret_cast<const float4*>(X + o + c * cs + j);
#pragma unroll
    for (int m = 0; m < N; ++m) {
      float r0 = 0.f, i0 = 0.f, r1 = 0.f, i1 = 0.f;
#pragma unroll
      for (int c = 0; c < N; ++c) {
        r0 = fmaf(q[m][c].x, x[c].x, fmaf(-q[m][c].y, x[c].y, r0));
        i0 = fmaf(q[m][c].x, x[c].y, fmaf(q[m][c].y, x[c].x, i0));
        r1 = fmaf(q[m][c].x, x[c].z, fmaf(-q[m][c].y, x[c].w, r1));
        i1 = fmaf(q[m][c].x, x[c].w, fmaf(q[m][c].y, x[c].z, i1));
      }
      *reinterpret_cast<float2*>(Z2 + o + m * cs + j) = make_float2(fmaf(r0, r0, i0 * i0), fmaf(r1, r1, i1 * i1));
    }
  }
}

// phi[b,m,i,j] = 1 / L[i,j,m]      (mnmf.py:1504-1510)
template <int N, bool LAM>
__global__ void __launch_bounds__(MW * 32) km_phi(const cf* __restrict__ X, const float* __restrict__ T,
                                                  const float* __restrict__ V, const float* __restrict__ Lam,
                                                 const cf* __restrict__ Q,
                                                  const float* __restrict__ D, float* __restrict__ phi, int B, int I,
                                                  int J, int K) {
  __shared__ BinCtx<N> ctx[MW];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bin = blockIdx.x * MW + wib;
  if (bin >= B * I) return;
  const int b = bin / I, i = bin - b * I;
  BinCtx<N>& cs = ctx[wib];
  load_ctx<N>(cs, T, Q, D, b, i, I, K, lane);
  const RegCtx<N> c(cs);
  const RowBase rb = row_base(X, V, Lam, b, i, N, I, J, K);
#pragma unroll 2
  for (int j = lane; j < J; j += 32) {
    float lam[N], L[N], Z2[N];
    frame_stats<N, LAM>(c, rb, j, lam, L, Z2);
#pragma unroll
    for (int m = 0; m < N; ++m) phi[((size_t)b * N * I + i) * J + m * rb.cs + j] = mn_rcp(L[m]);
  }
}

// spatial update (mnmf.py:1660-1675) and the per-bin sums zsum[b,i,m] = sum_j Z2_m for the normalisation.
// Every source uses the D of the previous iteration (L is formed from the context loaded up front), so GS sources
// share one pass over the frames: GS = N for N <= 4 (one pass), 1 otherwise (registers).  update_d = 0: zsum only.
template <int N, bool LAM>
__global__ void __launch_bounds__(MW * 32, (N <= 4 ? 4 : 1)) km_spatial(const cf* __restrict__ X, const float* __restrict__ T,
                                                      const float* __restrict__ V, const float* __restrict__ Lam,
                                                 const cf* __restrict__ Q,
                                                      float* __restrict__ D, double* __restrict__ zsum, int B, int I,
                                                      int J, int K, int update_d, float* __restrict__ z2out) {
  constexpr int GS = N <= 4 ? N : 1;
  __shared__ BinCtx<N> ctx[MW];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bin = blockIdx.x * MW + wib;
  if (bin >= B * I) return;
  const int b = bin / I, i = bin - b * I;
  BinCtx<N>& cs = ctx[wib];
  load_ctx<N>(cs, T, Q, D, b, i, I, K, lane);
  const RegCtx<N> c(cs);
  const RowBase rb = row_base(X, V, Lam, b, i, N, I, J, K);
  for (int n0 = 0; n0 < (update_d ? N : 1); n0 += GS) {
    float num[GS][N], den[GS][N], zs[N];
#pragma unroll
    for (int m = 0; m < N; ++m) {
      zs[m] = 0.f;
#pragma unroll
      for (int gs = 0; gs < GS; ++gs) num[gs][m] = den[gs][m] = 0.f;
    }
    // accumulate one frame: lam[n] = Lambda_n, Z2[m], L[m]
    auto accumulate = [&](const float (&lam)[N], const float (&L)[N], const float (&Z2)[N]) {
      float ln[GS];
#pragma unroll
      for (int gs = 0; gs < GS; ++gs) {
        ln[gs] = lam[0];
#pragma unroll
        for (int q = 1; q < N; ++q) ln[gs] = (q == n0 + gs) ? lam[q] : ln[gs];
      }
#pragma unroll
      for (int m = 0; m < N; ++m) {
        const float r = mn_rcp(L[m]);
        const float a = r * r * Z2[m];
        zs[m] += Z2[m];
#pragma unroll
        for (int gs = 0; gs < GS; ++gs) {
          num[gs][m] = fmaf(ln[gs], a, num[gs][m]);
          den[gs][m] = fmaf(ln[gs], r, den[gs][m]);
        }
      }
    };
    if (LAM && N <= 4 && (J & 1) == 0) {
      // two frames per lane: 16-byte loads of X, 8-byte loads of Lambda, and (inside ssb_run) 8-byte stores of the Z2 of
      // the filters just updated, which the source model of the next iteration streams (kf_mnmf_update)
      for (int j = 2 * lane; j < J; j += 64) {
        float4 x[N];
        float2 l2[N];
#pragma unroll
        for (int m = 0; m < N; ++m) {
          x[m] = *reinterpret_cast<const float4*>(rb.x + m * rb.cs + j);
          l2[m] = *reinterpret_cast<const float2*>(rb.lam + m * rb.cs + j);
        }
        float za[N], zb[N];
#pragma unroll
        for (int f = 0; f < 2; ++f) {
          float lam[N], L[N], Z2[N];
#pragma unroll
          for (int n = 0; n < N; ++n) lam[n] = f ? l2[n].y : l2[n].x;
#pragma unroll
          for (int m = 0; m < N; ++m) {
            float l = 0.f, zr = 0.f, zi = 0.f;
#pragma unroll
            for (int n = 0; n < N; ++n) l = fmaf(lam[n], c.d(n, m), l);
#pragma unroll
            for (int cc = 0; cc < N; ++cc) {
              const cf q = c.q(m, cc);
              const float xr = f ? x[cc].z : x[cc].x, xi = f ? x[cc].w : x[cc].y;
              zr = fmaf(q.x, xr, fmaf(-q.y, xi, zr));
              zi = fmaf(q.x, xi, fmaf(q.y, xr, zi));
            }
            L[m] = l;
            Z2[m] = zr * zr + zi * zi;
            if (f) zb[m] = Z2[m];
            else za[m] = Z2[m];
          }
          accumulate(lam, L, Z2);
        }
        if (z2out != nullptr && n0 == 0) {
#pragma unroll
          for (int m = 0; m < N; ++m)
            *reinterpret_cast<float2*>(z2out + ((size_t)b * N * I + i) * J + m * rb.cs + j) = make_float2(za[m], zb[m]);
        }
      }
    } else {
#pragma unroll 2
      for (int j = lane; j < J; j += 32) {
        float lam[N], L[N], Z2[N];
        frame_stats<N, LAM>(c, rb, j, lam, L, Z2);
        accumulate(lam, L, Z2);
        if (z2out != nullptr && n0 == 0) {
#pragma unroll
          for (int m = 0; m < N; ++m) z2out[((size_t)b * N * I + i) * J + m * rb.cs + j] = Z2[m];
        }
      }
    }
#pragma unroll
    for (int m = 0; m < N; ++m) {
      const float z = warp_sum(zs[m]);
      if (n0 == 0 && lane == 0) zsum[((size_t)b * I + i) * N + m] = (double)z;
#pragma unroll
      for (int gs = 0; gs < GS; ++gs) {
        const float nu = warp_sum(num[gs][m]), de = warp_sum(den[gs][m]);
        if (update_d && lane == 0 && n0 + gs < N)
          D[(((size_t)b * I + i) * N + n0 + gs) * N + m] = sqrtf(nu / de) * c.d(n0 + gs, m);
      }
    }
  }
}

// psi_m = floor(sqrt(mean_ij Z2_m)); Q[:,m,:] /= psi_m; D[:,:,m] /= psi_m^2   (mnmf.py:666-678)
__global__ void km_normalize(const double* __restrict__ zsum, cf* __restrict__ Q, float* __restrict__ D, int N, int I,
                             int J, int flooring, double eps, float* __restrict__ zscale) {
  __shared__ double sh[8];
  __shared__ double s_psi[SSB_MAX_SOURCES];
  const int b = blockIdx.x;
  for (int m = 0; m < N; ++m) {
    double acc = 0.0;
    for (int i = threadIdx.x; i < I; i += blockDim.x) acc += zsum[((size_t)b * I + i) * N + m];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
      s_psi[m] = ssb_floor(sqrt(t / ((double)I * (double)J)), flooring, eps);
      if (zscale != nullptr) zscale[b * N + m] = (float)(1.0 / (s_psi[m] * s_psi[m]));  // Z2_m scales with |q_m|^2
    }
    __syncthreads();
  }
  for (int e = threadIdx.x; e < I * N * N; e += blockDim.x) {
    const int r = (e / N) % N, cc = e % N;  // Q[i, m=r, c], D[i, n=r, m=cc]
    const size_t o = (size_t)b * I * N * N + e;
    const double pq = s_psi[r], pd = s_psi[cc];
    cf q = Q[o];
    Q[o] = make_float2((float)(q.x / pq), (float)(q.y / pq));
    D[o] = (float)((double)D[o] / (pd * pd));
  }
}

// rowloss[b,i] = mean_j sum_m (Z2/L + log L)     (mnmf.py:1255-1258)
template <int N, bool LAM>
__global__ void __launch_bounds__(MW * 32) km_rowloss(const cf* __restrict__ X, const float* __restrict__ T,
                                                      const float* __restrict__ V, const float* __restrict__ Lam,
                                                 const cf* __restrict__ Q,
                                                      const float* __restrict__ D, double* __restrict__ rowloss, int B,
                                                      int I, int J, int K) {
  __shared__ BinCtx<N> ctx[MW];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bin = blockIdx.x * MW + wib;
  if (bin >= B * I) return;
  const int b = bin / I, i = bin - b * I;
  BinCtx<N>& cs = ctx[wib];
  load_ctx<N>(cs, T, Q, D, b, i, I, K, lane);
  const RegCtx<N> c(cs);
  const RowBase rb = row_base(X, V, Lam, b, i, N, I, J, K);
  double acc = 0.0;
#pragma unroll 2
  for (int j = lane; j < J; j += 32) {
    float lam[N], L[N], Z2[N];
    frame_stats<N, LAM>(c, rb, j, lam, L, Z2);
    float s = 0.f;
#pragma unroll
    for (int m = 0; m < N; ++m) s += Z2[m] / L[m] + logf(L[m]);
    acc += (double)s;
  }
  acc = warp_sum(acc);
  if (lane == 0) rowloss[bin] = acc / (double)J;
}

// Qinv[b,i] = Q[b,i]^-1 (complex128), one lane group per bin
template <int N>
__global__ void __launch_bounds__(MW * 32) km_qinv(const cf* __restrict__ Q, cd* __restrict__ Qinv, int n_mat,
                                                   int* __restrict__ status) {
  constexpr int GS = GroupShape<N>::GS, GW = GroupShape<N>::GW;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane / GS, r = lane - grp * GS, gbase = grp * GS;
  const int mat_raw = (blockIdx.x * MW + warp) * GW + grp;
  const bool valid = mat_raw < n_mat;
  const int mat = valid ? mat_raw : n_mat - 1;
  cd a[N], rhs[N];
#pragma unroll
  for (int c = 0; c < N; ++c) {
    a[c] = (r < N) ? cf2cd(Q[((size_t)mat * N + r) * N + c]) : cd_make(0, 0);
    rhs[c] = cd_make(r == c ? 1.0 : 0.0, 0);
  }
  bool sing;
  group_solve<N, N, GS>(a, rhs, r, gbase, nullptr, &sing);
  if (sing && valid && r == 0) atomicOr(status, SSB_STATUS_SINGULAR);  // np.linalg.inv(Q) raises (mnmf.py:1198)
  if (valid && r < N) {
#pragma unroll
    for (int c = 0; c < N; ++c) Qinv[((size_t)mat * N + r) * N + c] = rhs[c];
  }
}

// Qinv[b,i] <- Lm[b,i] Qinv[b,i]: the inverse of a diagonaliser kept in the whitened domain (Q = Q~ M, Q^-1 = M^-1 Q~^-1)
template <int N>
__global__ void km_leftmul(const cd* __restrict__ Lm, cd* __restrict__ Qinv, int n_mat) {
  const int mat = blockIdx.x * blockDim.x + threadIdx.x;
  if (mat >= n_mat) return;
  const cd* L = Lm + (size_t)mat * N * N;
  cd* Q = Qinv + (size_t)mat * N * N;
  cd q[N * N], o[N * N];
#pragma unroll
  for (int e = 0; e < N * N; ++e) q[e] = Q[e];
#pragma unroll
  for (int r = 0; r < N; ++r)
#pragma unroll
    for (int c = 0; c < N; ++c) {
      cd s = cd_make(0, 0);
#pragma unroll
      for (int k = 0; k < N; ++k) s = cd_fma(L[r * N + k], q[k * N + c], s);
      o[r * N + c] = s;
    }
#pragma unroll
  for (int e = 0; e < N * N; ++e) Q[e] = o[e];
}

// Multichannel Wiener filter (mnmf.py:1186-1217), one thread per (bin, frame), one block per bin:
//   R = Qinv diag(L) Qinv^H -> to_psd (Hermitian eigendecomposition, eigenvalues floored, rebuilt)
//   u_n = R^-1 R_n[:, ref],  R_n[:, ref] = Qinv diag(Lambda_n D[n,:]) conj(Qinv[ref, :])
//   y_n = u_n^H x
template <int N>
__global__ void __launch_bounds__(128) km_separate(const cf* __restrict__ X, const float* __restrict__ T,
                                                   const float* __restrict__ V, const float* __restrict__ D,
                                                   const cd* __restrict__ Qinv, cf* __restrict__ Y, int I, int J,
                                                   int K, int ref, int flooring, double eps) {
  __shared__ float sT[N][SSB_MAX_BASIS];
  __shared__ float sD[N][N];
  __shared__ cd sQi[N][N];
  const int bin = blockIdx.x;
  const int b = bin / I, i = bin - b * I;
  for (int e = threadIdx.x; e < N * K; e += blockDim.x) {
    const int n = e / K, k = e - n * K;
    sT[n][k] = T[(((size_t)b * N + n) * I + i) * K + k];
  }
  for (int e = threadIdx.x; e < N * N; e += blockDim.x) {
    sD[e / N][e % N] = D[(size_t)bin * N * N + e];
    sQi[e / N][e % N] = Qinv[(size_t)bin * N * N + e];
  }
  __syncthreads();
  for (int j = threadIdx.x; j < J; j += blockDim.x) {
    double lam[N], L[N];
#pragma unroll
    for (int n = 0; n < N; ++n) {
      const float* v = V + ((size_t)b * N + n) * K * J + j;
      float s = 0.f;
      for (int k = 0; k < K; ++k) s = fmaf(sT[n][k], v[(size_t)k * J], s);
      lam[n] = (double)s;
    }
#pragma unroll
    for (int m = 0; m < N; ++m) {
      double l = 0.0;
#pragma unroll
      for (int n = 0; n < N; ++n) l += lam[n] * (double)sD[n][m];
      L[m] = l;
    }
    cd R[N * N], P[N * N];
    for (int a = 0; a < N; ++a)
      for (int c = a; c < N; ++c) {
        cd s = cd_make(0, 0);
        for (int m = 0; m < N; ++m) s = cd_add(s, cd_scale(cd_mulc(sQi[a][m], sQi[c][m]), L[m]));
        if (a == c) s.y = 0.0;
        R[a * N + c] = s;
        R[c * N + a] = cd_conj(s);
      }
    jacobi_herm(R, P, N);  // eigenvalues on the diagonal of R, eigenvectors in the columns of P
    double il[N];
    for (int m = 0; m < N; ++m) il[m] = 1.0 / ssb_floor(R[m * N + m].x, flooring, eps);
    cd x[N];
#pragma unroll
    for (int m = 0; m < N; ++m) x[m] = cf2cd(X[(((size_t)b * N + m) * I + i) * J + j]);
    // px = P^H x ; for each source: y_n = r_n^H R^-1 x = sum_e conj(P^H r_n)_e (1/lambda_e) (P^H x)_e
    cd px[N];
    for (int e = 0; e < N; ++e) {
      cd s = cd_make(0, 0);
      for (int a = 0; a < N; ++a) s = cd_add(s, cd_mul(cd_conj(P[a * N + e]), x[a]));
      px[e] = s;
    }
    for (int n = 0; n < N; ++n) {
      cd rn[N];  // R_n[:, ref]
      for (int a = 0; a < N; ++a) {
        cd s = cd_make(0, 0);
        for (int m = 0; m < N; ++m) s = cd_add(s, cd_scale(cd_mulc(sQi[a][m], sQi[ref][m]), lam[n] * (double)sD[n][m]));
        rn[a] = s;
      }
      cd y = cd_make(0, 0);
      for (int e = 0; e < N; ++e) {
        cd pr = cd_make(0, 0);
        for (int a = 0; a < N; ++a) pr = cd_add(pr, cd_mul(cd_conj(P[a * N + e]), rn[a]));
        y = cd_add(y, cd_scale(cd_mul(cd_conj(pr), px[e]), il[e]));
      }
      Y[(((size_t)b * N + n) * I + i) * J + j] = cd2cf(y);
    }
  }
}

}  // namespace

int ssbk_mnmf_gh(const cf* X, const float* T, const float* V, const float* Lam, const cf* Q, const float* D, float* G, float* H, int B,
                 int N, int I, int J, int K, cudaStream_t st) {
  if (Lam != nullptr) { SSB_DISPATCH_N(N, (km_gh<NN, true><<<blocks_for((long long)B * I, MW), MW * 32, 0, st>>>(X, T, V, Lam, Q, D, G, H, B, I, J, K))); }
  else { SSB_DISPATCH_N(N, (km_gh<NN, false><<<blocks_for((long long)B * I, MW), MW * 32, 0, st>>>(X, T, V, Lam, Q, D, G, H, B, I, J, K))); }
  return ssb_check_launch("mnmf_gh", st);
}
int ssbk_mnmf_z2(const cf* X, const cf* Q, float* Z2, int B, int N, int I, int J, cudaStream_t st) {
  SSB_REQUIRE((J % 2) == 0, "mnmf_z2 needs an even number of frames");
  SSB_DISPATCH_N(N, km_z2<NN><<<blocks_for((long long)B * I, MW), MW * 32, 0, st>>>(X, Q, Z2, B, I, J));
  return ssb_check_launch("mnmf_z2", st);
}
int ssbk_mnmf_phi(const cf* X, const float* T, const float* V, const float* Lam, const cf* Q, const float* D, float* phi, int B, int N,
                  int I, int J, int K, cudaStream_t st) {
  if (Lam != nullptr) { SSB_DISPATCH_N(N, (km_phi<NN, true><<<blocks_for((long long)B * I, MW), MW * 32, 0, st>>>(X, T, V, Lam, Q, D, phi, B, I, J, K))); }
  else { SSB_DISPATCH_N(N, (km_phi<NN, false><<<blocks_for((long long)B * I, MW), MW * 32, 0, st>>>(X, T, V, Lam, Q, D, phi, B, I, J, K))); }
  return ssb_check_launch("mnmf_phi", st);
}
int ssbk_mnmf_spatial(const cf* X, const float* T, const float* V, const float* Lam, const cf* Q, float* D, double* zsum, int B, int N,
                      int I, int J, int K, int update_d, cudaStream_t st, float* z2out) {
  if (Lam != nullptr) {
    SSB_DISPATCH_N(N, (km_spatial<NN, true><<<blocks_for((long long)B * I, MW), MW * 32, 0, st>>>(X, T, V, Lam, Q, D, zsum,
                                                                                                   B, I, J, K, update_d, z2out)));
  } else {
    SSB_DISPATCH_N(N, (km_spatial<NN, false><<<blocks_for((long long)B * I, MW), MW * 32, 0, st>>>(X, T, V, Lam, Q, D, zsum,
                                                                                                    B, I, J, K, update_d, z2out)));
  }
  return ssb_check_launch("mnmf_spatial", st);
}
int ssbk_mnmf_normalize(const double* zsum, cf* Q, float* D, int B, int N, int I, int J, int flooring, float eps,
                        cudaStream_t st, float* zscale) {
  km_normalize<<<B, 256, 0, st>>>(zsum, Q, D, N, I, J, flooring, (double)eps, zscale);
  return ssb_check_launch("mnmf_normalize", st);
}
int ssbk_mnmf_rowloss(const cf* X, const float* T, const float* V, const float* Lam, const cf* Q, const float* D, double* rowloss, int B,
                      int N, int I, int J, int K, cudaStream_t st) {
  if (Lam != nullptr) {
    SSB_DISPATCH_N(N, (km_rowloss<NN, true><<<blocks_for((long long)B * I, MW), MW * 32, 0, st>>>(X, T, V, Lam, Q, D, rowloss,
                                                                                                   B, I, J, K)));
  } else {
    SSB_DISPATCH_N(N, (km_rowloss<NN, false><<<blocks_for((long long)B * I, MW), MW * 32, 0, st>>>(X, T, V, Lam, Q, D, rowloss,
                                                                                                    B, I, J, K)));
  }
  return ssb_check_launch("mnmf_rowloss", st);
}
int ssbk_mnmf_separate(const cf* X, const float* T, const float* V, const cf* Q, const float* D, cd* Qinv, cf* Y, int B,
                       int N, int I, int J, int K, int ref, int flooring, float eps, cudaStream_t st, const cd* Lleft) {
  SSB_REQUIRE(ref >= 0 && ref < N, "reference_id=%d out of range for N=%d", ref, N);
  SSB_DISPATCH_N(N, km_qinv<NN><<<blocks_for((long long)B * I, MW * GroupShape<NN>::GW), MW * 32, 0, st>>>(Q, Qinv, B * I, ssb_status_word()));
  if (ssb_check_launch("mnmf_qinv", st)) return 1;
  if (Lleft != nullptr) {
    SSB_DISPATCH_N(N, km_leftmul<NN><<<blocks_for((long long)B * I, 64), 64, 0, st>>>(Lleft, Qinv, B * I));
    if (ssb_check_launch("mnmf_qinv_unwhiten", st)) return 1;
  }
  SSB_DISPATCH_N(N, km_separate<NN><<<B * I, 128, 0, st>>>(X, T, V, D, Qinv, Y, I, J, K, ref, flooring, (double)eps));
  return ssb_check_launch("mnmf_separate", st);
}
