// Per-bin spatial kernels of the ILRMA / AuxIVA path (modular variants): separate, weighted
// covariance, IP1, IP2, ISS1, projection back, cross-solve (W recovery), log|det W|.
// One warp owns one (mixture, bin); the N x N complex linear algebra is done in fp64 in shared
// memory by the warp (ssb_common.cuh), the frame reductions in fp32 with a warp tree reduce.
#include "ssb_group.cuh"
#include <stdlib.h>

#include "ssb_kernels.h"

namespace {

constexpr int WPB = 4;  // warps (= bins) per block for the warp-per-bin kernels

// ------------------------------------------------------------------------------------------------
// separate: Y[b,n,i,j] = sum_m W[b,i,n,m] X[b,m,i,j]        (ssspy/bss/ilrma.py:292-295)
// optionally P[b,n,i,j] = |Y|^2.  One block per (b,i); threads stride over frames.
template <int N>
__global__ void __launch_bounds__(128) k_separate(const cf* __restrict__ X, const cf* __restrict__ W,
                                                  cf* __restrict__ Y, float* __restrict__ P, int I, int J) {
  __shared__ cf w[N * N];
  const int bi = blockIdx.x;
  const int b = bi / I, i = bi - b * I;
  if (threadIdx.x < N * N) w[threadIdx.x] = W[(size_t)bi * N * N + threadIdx.x];
  __syncthreads();
  const size_t base = ((size_t)b * N * I + i) * J;  // + m*I*J
  const size_t cs = (size_t)I * J;
  for (int j = threadIdx.x; j < J; j += blockDim.x) {
    cf x[N];
#pragma unroll
    for (int m = 0; m < N; ++m) x[m] = X[base + m * cs + j];
#pragma unroll
    for (int n = 0; n < N; ++n) {
      float yr = 0.f, yi = 0.f;
#pragma unroll
      for (int m = 0; m < N; ++m) {
        cf ww = w[n * N + m];
        yr = fmaf(ww.x, x[m].x, fmaf(-ww.y, x[m].y, yr));
        yi = fmaf(ww.x, x[m].y, fmaf(ww.y, x[m].x, yi));
      }
      if (Y) Y[base + n * cs + j] = make_float2(yr, yi);
      if (P) P[base + n * cs + j] = yr * yr + yi * yi;
    }
  }
}

// P = |Y|^2 elementwise (ISS modes, where the state is Y itself)
__global__ void k_abs2(const cf* __restrict__ Y, float* __restrict__ P, size_t n) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; t < n; t += stride) {
    cf y = Y[t];
    P[t] = y.x * y.x + y.y * y.y;
  }
}

// ------------------------------------------------------------------------------------------------
// weighted covariance  U[b,i,s,a,c] = (1/J) sum_j phi[b,src[s],i,j] X[b,a,i,j] conj(X[b,c,i,j])
// (ssspy/bss/ilrma.py:1500-1505, ssspy/bss/iva.py:1785-1791).  One warp per (b,i); NSB sources are
// accumulated per pass over the bin's (channel x frame) slab; Hermitian => N^2 real accumulators
// per source (upper triangle: re at [a][c], im at [c][a]).
struct SrcList {
  int n;
  int idx[SSB_MAX_SOURCES];
};

template <int N>
struct CovCfg {
  static constexpr int raw = 64 / (N * N);
  static constexpr int NSB = raw < 1 ? 1 : (raw > N ? N : raw);
};

template <int N>
__global__ void __launch_bounds__(WPB * 32) k_wcov(const cf* __restrict__ X, const float* __restrict__ phi,
                                                   long long sb, long long sn, long long si, SrcList src,
                                                   cf* __restrict__ U, int B, int I, int J) {
  constexpr int NSB = CovCfg<N>::NSB;
  const int warp = blockIdx.x * WPB + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (warp >= B * I) return;
  const int b = warp / I, i = warp - b * I;
  const size_t base = ((size_t)b * N * I + i) * J;
  const size_t cs = (size_t)I * J;
  const float invJ = 1.0f / (float)J;
  for (int s0 = 0; s0 < src.n; s0 += NSB) {
    float acc[NSB][N * N];
#pragma unroll
    for (int s = 0; s < NSB; ++s)
#pragma unroll
      for (int e = 0; e < N * N; ++e) acc[s][e] = 0.f;
    for (int j = lane; j < J; j += 32) {
      cf x[N];
#pragma unroll
      for (int m = 0; m < N; ++m) x[m] = X[base + m * cs + j];
#pragma unroll
      for (int s = 0; s < NSB; ++s) {
        if (s0 + s < src.n) {
          float ph = phi ? phi[(size_t)b * sb + (size_t)src.idx[s0 + s] * sn + (size_t)i * si + j] : 1.0f;
#pragma unroll
          for (int a = 0; a < N; ++a) {
            float xr = ph * x[a].x, xi = ph * x[a].y;
            acc[s][a * N + a] = fmaf(xr, x[a].x, fmaf(xi, x[a].y, acc[s][a * N + a]));
#pragma unroll
            for (int c = a + 1; c < N; ++c) {
              // x_a conj(x_c) = (ar cr + ai ci) + i (ai cr - ar ci)
              acc[s][a * N + c] = fmaf(xr, x[c].x, fmaf(xi, x[c].y, acc[s][a * N + c]));
              acc[s][c * N + a] = fmaf(xi, x[c].x, fmaf(-xr, x[c].y, acc[s][c * N + a]));
            }
          }
        }
      }
    }
#pragma unroll
    for (int s = 0; s < NSB; ++s) {
      if (s0 + s < src.n) {
#pragma unroll
        for (int e = 0; e < N * N; ++e) acc[s][e] = warp_sum(acc[s][e]) * invJ;
        cf* u = U + ((size_t)warp * src.n + (s0 + s)) * N * N;
        if (lane == 0) {
#pragma unroll
          for (int a = 0; a < N; ++a) {
            u[a * N + a] = make_float2(acc[s][a * N + a], 0.f);
#pragma unroll
            for (int c = a + 1; c < N; ++c) {
              u[a * N + c] = make_float2(acc[s][a * N + c], acc[s][c * N + a]);
              u[c * N + a] = make_float2(acc[s][a * N + c], -acc[s][c * N + a]);
            }
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// IP1 (ssspy/bss/_update_spatial_model.py:63-76).  One warp per matrix, Gauss-Seidel over sources.
template <int N>
__global__ void __launch_bounds__(WPB * 32) k_ip1(cf* __restrict__ W, const cf* __restrict__ U, int n_mat,
                                                  int flooring, double eps) {
  __shared__ cd sW[WPB][N * N];
  __shared__ cd sU[WPB][N * N];
  __shared__ cd sA[WPB][N * (N + 1)];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mat = blockIdx.x * WPB + wib;
  if (mat >= n_mat) return;
  cd* w = sW[wib];
  cd* u = sU[wib];
  cd* A = sA[wib];
  for (int e = lane; e < N * N; e += 32) w[e] = cf2cd(W[(size_t)mat * N * N + e]);
  for (int n = 0; n < N; ++n) {
    __syncwarp();
    for (int e = lane; e < N * N; e += 32) u[e] = cf2cd(U[((size_t)mat * N + n) * N * N + e]);
    __syncwarp();
    for (int e = lane; e < N * (N + 1); e += 32) {
      int r = e / (N + 1), c = e - r * (N + 1);
      cd s = cd_make(0, 0);
      if (c < N) {
#pragma unroll
        for (int k = 0; k < N; ++k) s = cd_fma(w[r * N + k], u[k * N + c], s);
      } else {
        s = cd_make(r == n ? 1.0 : 0.0, 0);
      }
      A[e] = s;
    }
    __syncwarp();
    warp_gauss_jordan(A, N, 1, N + 1, lane);
    // wUw = Re( w^H U_n w ), w[a] = A[a][N]
    double part = 0.0;
    for (int e = lane; e < N * N; e += 32) {
      int a = e / N, c = e - a * N;
      cd t = cd_mul(u[e], A[c * (N + 1) + N]);
      part += cd_mulc(t, A[a * (N + 1) + N]).x;  // Re( conj(w_a) * U[a][c] w_c )
    }
    double wUw = warp_sum(part);
    double d = ssb_floor(sqrt(fmax(wUw, 0.0)), flooring, eps);
    __syncwarp();
    if (lane < N) w[n * N + lane] = cd_scale(cd_conj(A[lane * (N + 1) + N]), 1.0 / d);
  }
  __syncwarp();
  for (int e = lane; e < N * N; e += 32) W[(size_t)mat * N * N + e] = cd2cf(w[e]);
}

// ------------------------------------------------------------------------------------------------
// IP2 (ssspy/bss/_update_spatial_model.py:137-141, :353-395).  One warp per matrix; pairs in order.
struct PairList {
  int n;
  int n_u;  // sources per matrix stored in U
  short m[SSB_MAX_PAIRS], nn[SSB_MAX_PAIRS];    // rows of W
  short um[SSB_MAX_PAIRS], un[SSB_MAX_PAIRS];   // which U slices
};

template <int N>
__device__ __forceinline__ void quad2(const cd* __restrict__ P, const cd* __restrict__ u, int lane, cd G[4]) {
  // G[r][c] = sum_{a,b} conj(P[a][r]) U[a][b] P[b][c],  P is N x 2 (row major, ld 2)
#pragma unroll
  for (int rc = 0; rc < 4; ++rc) {
    const int r = rc >> 1, c = rc & 1;
    double pr = 0.0, pi = 0.0;
    for (int e = lane; e < N * N; e += 32) {
      int a = e / N, bb = e - a * N;
      cd t = cd_mul(u[e], P[bb * 2 + c]);
      cd v = cd_mul(cd_conj(P[a * 2 + r]), t);
      pr += v.x;
      pi += v.y;
    }
    G[rc] = cd_make(warp_sum(pr), warp_sum(pi));
  }
}

template <int N>
__global__ void __launch_bounds__(WPB * 32) k_ip2(cf* __restrict__ W, const cf* __restrict__ U, int n_mat,
                                                  PairList pl, int flooring, double eps) {
  __shared__ cd sW[WPB][N * N];
  __shared__ cd sUm[WPB][N * N];
  __shared__ cd sUn[WPB][N * N];
  __shared__ cd sA[WPB][N * (N + 2)];
  __shared__ cd sPm[WPB][N * 2];
  __shared__ cd sPn[WPB][N * 2];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mat = blockIdx.x * WPB + wib;
  if (mat >= n_mat) return;
  cd *w = sW[wib], *um = sUm[wib], *un = sUn[wib], *A = sA[wib], *Pm = sPm[wib], *Pn = sPn[wib];
  for (int e = lane; e < N * N; e += 32) w[e] = cf2cd(W[(size_t)mat * N * N + e]);
  for (int q = 0; q < pl.n; ++q) {
    const int m = pl.m[q], n = pl.nn[q];
    __syncwarp();
    for (int e = lane; e < N * N; e += 32) {
      um[e] = cf2cd(U[((size_t)mat * pl.n_u + pl.um[q]) * N * N + e]);
      un[e] = cf2cd(U[((size_t)mat * pl.n_u + pl.un[q]) * N * N + e]);
    }
    __syncwarp();
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      const cd* u = which ? un : um;
      cd* P = which ? Pn : Pm;
      for (int e = lane; e < N * (N + 2); e += 32) {
        int r = e / (N + 2), c = e - r * (N + 2);
        cd s = cd_make(0, 0);
        if (c < N) {
#pragma unroll
          for (int k = 0; k < N; ++k) s = cd_fma(w[r * N + k], u[k * N + c], s);
        } else {
          s = cd_make((c == N ? r == m : r == n) ? 1.0 : 0.0, 0);
        }
        A[e] = s;
      }
      __syncwarp();
      warp_gauss_jordan(A, N, 2, N + 2, lane);
      for (int e = lane; e < N * 2; e += 32) P[e] = A[(e >> 1) * (N + 2) + N + (e & 1)];
      __syncwarp();
    }
    cd Gm[4], Gn[4];
    quad2<N>(Pm, um, lane, Gm);
    quad2<N>(Pn, un, lane, Gn);
    // generalised eigenproblem Gm h = l Gn h  (ssspy/linalg/eigh.py:173-201), all lanes redundantly
    const double b00 = Gn[0].x, b11 = Gn[3].x;
    const cd b10 = Gn[2];
    const double l00 = sqrt(b00);
    const cd l10 = cd_scale(b10, 1.0 / l00);
    const double l11 = sqrt(b11 - cd_abs2(l10));
    // Linv = [[i00, 0], [i10, i11]]
    const double i00 = 1.0 / l00, i11 = 1.0 / l11;
    const cd i10 = cd_scale(l10, -i00 * i11);
    const double a00 = Gm[0].x, a11 = Gm[3].x;
    const cd a01 = Gm[1];
    // C = Linv A Linv^H
    const double c00 = i00 * a00 * i00;
    // c01 = (Linv A)[0,:] . conj(Linv[1,:]) = i00*a00*conj(i10) + i00*a01*i11
    const cd c01 = cd_add(cd_scale(cd_conj(i10), i00 * a00), cd_scale(a01, i00 * i11));
    // c11 = i10 a00 conj(i10) + i10 a01 i11 + i11 conj(a01) conj(i10) + i11 a11 i11
    const double c11 = cd_abs2(i10) * a00 + 2.0 * i11 * cd_mul(i10, a01).x + i11 * i11 * a11;
    double lam[2];
    cd y0[2], y1[2];
    herm_eig2(c00, c11, c01, lam, y0, y1);
    // z = Linv^H y ;  Linv^H = [[i00, conj(i10)], [0, i11]]
    cd hm[2], hn[2];  // h_m <- larger eigenvalue, h_n <- smaller (_update_spatial_model.py:370-373)
    hm[0] = cd_add(cd_scale(y1[0], i00), cd_mul(cd_conj(i10), y1[1]));
    hm[1] = cd_scale(y1[1], i11);
    hn[0] = cd_add(cd_scale(y0[0], i00), cd_mul(cd_conj(i10), y0[1]));
    hn[1] = cd_scale(y0[1], i11);
    // normalise: h / floor(sqrt(max(Re h^H G h, 0)))
    auto quad = [](const cd* G, const cd* h) {
      cd t0 = cd_add(cd_mul(G[0], h[0]), cd_mul(G[1], h[1]));
      cd t1 = cd_add(cd_mul(G[2], h[0]), cd_mul(G[3], h[1]));
      return cd_mulc(t0, h[0]).x + cd_mulc(t1, h[1]).x;
    };
    const double dm = ssb_floor(sqrt(fmax(quad(Gm, hm), 0.0)), flooring, eps);
    const double dn = ssb_floor(sqrt(fmax(quad(Gn, hn), 0.0)), flooring, eps);
    hm[0] = cd_scale(hm[0], 1.0 / dm);
    hm[1] = cd_scale(hm[1], 1.0 / dm);
    hn[0] = cd_scale(hn[0], 1.0 / dn);
    hn[1] = cd_scale(hn[1], 1.0 / dn);
    __syncwarp();
    if (lane < N) {
      cd wm = cd_add(cd_mul(Pm[lane * 2], hm[0]), cd_mul(Pm[lane * 2 + 1], hm[1]));
      cd wn = cd_add(cd_mul(Pn[lane * 2], hn[0]), cd_mul(Pn[lane * 2 + 1], hn[1]));
      w[m * N + lane] = cd_conj(wm);
      w[n * N + lane] = cd_conj(wn);
    }
  }
  __syncwarp();
  for (int e = lane; e < N * N; e += 32) W[(size_t)mat * N * N + e] = cd2cf(w[e]);
}

// ------------------------------------------------------------------------------------------------
// Lane-group IP1 / IP2: GS lanes own one bin, lane r holds row r of W (and of every intermediate)
// in fp64 registers (ssb_group.cuh); U_n is staged per group in shared memory for broadcast reads.
constexpr int QW = 4;  // warps per block

template <int N>
__device__ __forceinline__ void group_load_u(cf* su, const cf* __restrict__ Ug, int r) {
  constexpr int GS = GroupShape<N>::GS;
#pragma unroll
  for (int e = r; e < N * N; e += GS) su[e] = Ug[e];
}

// a[c] = sum_k wrow[k] * U[k][c]  (row r of W U)
template <int N>
__device__ __forceinline__ void group_row_times(const cd (&wrow)[N], const cf* su, cd (&a)[N]) {
#pragma unroll
  for (int c = 0; c < N; ++c) {
    cd s = cd_make(0, 0);
#pragma unroll
    for (int k = 0; k < N; ++k) s = cd_fma(wrow[k], cf2cd(su[k * N + c]), s);
    a[c] = s;
  }
}


// q[mat, r] = Re(w_r C w_r^H) for the row held by lane r (the per-bin term of the power normalisation
// psi_n^2 = mean_i q); C_i staged in `su`.  The stored complex64 filter is what counts.
template <int N>
__device__ __forceinline__ void group_emit_q(const cd (&wrow)[N], cf* su, const cf* __restrict__ Cg, int r, bool valid,
                                             double* __restrict__ q) {
  __syncwarp();
  group_load_u<N>(su, Cg, r);
  __syncwarp();
  if (!valid || r >= N) return;
  cd wf[N];
#pragma unroll
  for (int c = 0; c < N; ++c) wf[c] = cf2cd(cd2cf(wrow[c]));
  double acc = 0.0;
#pragma unroll
  for (int c = 0; c < N; ++c) {
    cd s = cd_make(0, 0);
#pragma unroll
    for (int a = 0; a < N; ++a) s = cd_fma(wf[a], cf2cd(su[a * N + c]), s);
    acc += cd_mulc(s, wf[c]).x;
  }
  *q = acc;
}

template <int N>
__global__ void __launch_bounds__(QW * 32) kq_ip1(cf* __restrict__ W, const cf* __restrict__ U, int n_mat,
                                                  int flooring, double eps, const cf* __restrict__ Cn,
                                                  double* __restrict__ qn, int* __restrict__ status) {
  constexpr int GS = GroupShape<N>::GS, GW = GroupShape<N>::GW;
  __shared__ cf s_u[QW][GW][N * N];
  bool any_sing = false;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane / GS, r = lane - grp * GS, gbase = grp * GS;
  const int mat_raw = (blockIdx.x * QW + warp) * GW + grp;
  const bool valid = mat_raw < n_mat;
  const int mat = valid ? mat_raw : n_mat - 1;
  cf* su = s_u[warp][grp];
  cd wrow[N];
#pragma unroll
  for (int c = 0; c < N; ++c) wrow[c] = (r < N) ? cf2cd(W[((size_t)mat * N + r) * N + c]) : cd_make(0, 0);
#pragma unroll 1
  for (int n = 0; n < N; ++n) {
    __syncwarp();
    group_load_u<N>(su, U + ((size_t)mat * N + n) * N * N, r);
    __syncwarp();
    cd a[N], rhs[1];
    group_row_times<N>(wrow, su, a);
    rhs[0] = cd_make(r == n ? 1.0 : 0.0, 0);
    bool sing;
    group_solve<N, 1, GS>(a, rhs, r, gbase, nullptr, &sing);
    any_sing = any_sing || sing;
    // w = rhs (component r on lane r);  wUw = Re(w^H U_n w)
    cd t = cd_make(0, 0);
#pragma unroll
    for (int c = 0; c < N; ++c) {
      const cd wc = shfl_cd(rhs[0], gbase + c);
      if (r < N) t = cd_fma(cf2cd(su[r * N + c]), wc, t);
    }
    const double part = (r < N) ? cd_mulc(t, rhs[0]).x : 0.0;
    const double wUw = group_sum<GS>(part);
    const double d = ssb_floor(sqrt(fmax(wUw, 0.0)), flooring, eps);
    const cd mine = cd_scale(cd_conj(rhs[0]), 1.0 / d);
#pragma unroll
    for (int c = 0; c < N; ++c) {
      const cd v = shfl_cd(mine, gbase + c);
      if (r == n) wrow[c] = v;
    }
  }
  if (valid && r < N) {
#pragma unroll
    for (int c = 0; c < N; ++c) W[((size_t)mat * N + r) * N + c] = cd2cf(wrow[c]);
  }
  if (any_sing && valid && r == 0) atomicOr(status, SSB_STATUS_SINGULAR);
  if (Cn != nullptr) group_emit_q<N>(wrow, su, Cn + (size_t)mat * N * N, r, valid, qn + (size_t)mat * N + r);
}

// G[x][y] = sum_{a,c} conj(P[a][x]) U[a][c] P[c][y] with P row-distributed (lane a holds P[a][0..1])
template <int N, int GS>
__device__ __forceinline__ void group_quad2(const cd (&P)[2], const cf* su, int r, int gbase, cd (&G)[4]) {
  cd t[2] = {cd_make(0, 0), cd_make(0, 0)};
#pragma unroll
  for (int c = 0; c < N; ++c) {
    const cd p0 = shfl_cd(P[0], gbase + c), p1 = shfl_cd(P[1], gbase + c);
    if (r < N) {
      const cd u = cf2cd(su[r * N + c]);
      t[0] = cd_fma(u, p0, t[0]);
      t[1] = cd_fma(u, p1, t[1]);
    }
  }
#pragma unroll
  for (int x = 0; x < 2; ++x)
#pragma unroll
    for (int y = 0; y < 2; ++y) {
      const cd v = (r < N) ? cd_mul(cd_conj(P[x]), t[y]) : cd_make(0, 0);
      G[x * 2 + y] = cd_make(group_sum<GS>(v.x), group_sum<GS>(v.y));
    }
}

template <int N, int MINB = (N <= 4 ? 4 : 3)>
__global__ void __launch_bounds__(QW * 32, MINB) kq_ip2(cf* __restrict__ W, const cf* __restrict__ U, int n_mat,
                                                  PairList pl, int flooring, double eps, const cf* __restrict__ Cn,
                                                  double* __restrict__ qn, int* __restrict__ status) {
  constexpr int GS = GroupShape<N>::GS, GW = GroupShape<N>::GW;
  bool any_sing = false;
  __shared__ cf s_um[QW][GW][N * N];
  __shared__ cf s_un[QW][GW][N * N];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = lane / GS, r = lane - grp * GS, gbase = grp * GS;
  const int mat_raw = (blockIdx.x * QW + warp) * GW + grp;
  const bool valid = mat_raw < n_mat;
  const int mat = valid ? mat_raw : n_mat - 1;
  cf* sum_ = s_um[warp][grp];
  cf* sun_ = s_un[warp][grp];
  cd wrow[N];
#pragma unroll
  for (int c = 0; c < N; ++c) wrow[c] = (r < N) ? cf2cd(W[((size_t)mat * N + r) * N + c]) : cd_make(0, 0);
#pragma unroll 1
  for (int q = 0; q < pl.n; ++q) {
    const int m = pl.m[q], n = pl.nn[q];
    __syncwarp();
    group_load_u<N>(sum_, U + ((size_t)mat * pl.n_u + pl.um[q]) * N * N, r);
    group_load_u<N>(sun_, U + ((size_t)mat * pl.n_u + pl.un[q]) * N * N, r);
    __syncwarp();
    cd a[N], Pm[2], Pn[2];
    group_row_times<N>(wrow, sum_, a);
    Pm[0] = cd_make(r == m ? 1.0 : 0.0, 0);
    Pm[1] = cd_make(r == n ? 1.0 : 0.0, 0);
    bool sing_m, sing_n;
    group_solve<N, 2, GS>(a, Pm, r, gbase, nullptr, &sing_m);
    group_row_times<N>(wrow, sun_, a);
    Pn[0] = cd_make(r == m ? 1.0 : 0.0, 0);
    Pn[1] = cd_make(r == n ? 1.0 : 0.0, 0);
    group_solve<N, 2, GS>(a, Pn, r, gbase, nullptr, &sing_n);
    any_sing = any_sing || sing_m || sing_n;
    cd Gm[4], Gn[4];
    group_quad2<N, GS>(Pm, sum_, r, gbase, Gm);
    group_quad2<N, GS>(Pn, sun_, r, gbase, Gn);
    // generalised eigenproblem Gm h = l Gn h (ssspy/linalg/eigh.py:173-201), redundantly on every lane
    const double b00 = Gn[0].x, b11 = Gn[3].x;
    const cd b10 = Gn[2];
    const double l00 = sqrt(b00);
    const cd l10 = cd_scale(b10, 1.0 / l00);
    const double l11 = sqrt(b11 - cd_abs2(l10));
    const double i00 = 1.0 / l00, i11 = 1.0 / l11;
    const cd i10 = cd_scale(l10, -i00 * i11);
    const double a00 = Gm[0].x, a11 = Gm[3].x;
    const cd a01 = Gm[1];
    const double c00 = i00 * a00 * i00;
    const cd c01 = cd_add(cd_scale(cd_conj(i10), i00 * a00), cd_scale(a01, i00 * i11));
    const double c11 = cd_abs2(i10) * a00 + 2.0 * i11 * cd_mul(i10, a01).x + i11 * i11 * a11;
    double lam[2];
    cd y0[2], y1[2];
    herm_eig2(c00, c11, c01, lam, y0, y1);
    cd hm[2], hn[2];  // h_m <- larger eigenvalue, h_n <- smaller (_update_spatial_model.py:370-373)
    hm[0] = cd_add(cd_scale(y1[0], i00), cd_mul(cd_conj(i10), y1[1]));
    hm[1] = cd_scale(y1[1], i11);
    hn[0] = cd_add(cd_scale(y0[0], i00), cd_mul(cd_conj(i10), y0[1]));
    hn[1] = cd_scale(y0[1], i11);
    auto quad = [](const cd* G, const cd* h) {
      cd t0 = cd_add(cd_mul(G[0], h[0]), cd_mul(G[1], h[1]));
      cd t1 = cd_add(cd_mul(G[2], h[0]), cd_mul(G[3], h[1]));
      return cd_mulc(t0, h[0]).x + cd_mulc(t1, h[1]).x;
    };
    const double dm = 1.0 / ssb_floor(sqrt(fmax(quad(Gm, hm), 0.0)), flooring, eps);
    const double dn = 1.0 / ssb_floor(sqrt(fmax(quad(Gn, hn), 0.0)), flooring, eps);
    // w_m = P_m h_m / d_m (component r on lane r); rows m, n of W are their conjugates
    const cd wm = cd_conj(cd_scale(cd_add(cd_mul(Pm[0], hm[0]), cd_mul(Pm[1], hm[1])), dm));
    const cd wn = cd_conj(cd_scale(cd_add(cd_mul(Pn[0], hn[0]), cd_mul(Pn[1], hn[1])), dn));
#pragma unroll
    for (int c = 0; c < N; ++c) {
      const cd vm = shfl_cd(wm, gbase + c), vn = shfl_cd(wn, gbase + c);
      if (r == m) wrow[c] = vm;
      if (r == n) wrow[c] = vn;
    }
  }
  if (valid && r < N) {
#pragma unroll
    for (int c = 0; c < N; ++c) W[((size_t)mat * N + r) * N + c] = cd2cf(wrow[c]);
  }
  if (any_sing && valid && r == 0) atomicOr(status, SSB_STATUS_SINGULAR);
  if (Cn != nullptr) group_emit_q<N>(wrow, sum_, Cn + (size_t)mat * N * N, r, valid, qn + (size_t)mat * N + r);
}

// ------------------------------------------------------------------------------------------------
// ISS1 (ssspy/bss/_update_spatial_model.py:181-192).  One warp per (b,i), in place on Y; each of the
// N sequential steps reads the bin's slab twice (statistics, then rank-1 update).
template <int N>
__global__ void __launch_bounds__(WPB * 32) k_iss1(cf* __restrict__ Y, const float* __restrict__ phi,
                                                   long long sb, long long sn, long long si, int B, int I,
                                                   int J, int flooring, float eps) {
  const int warp = blockIdx.x * WPB + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (warp >= B * I) return;
  const int b = warp / I, i = warp - b * I;
  const size_t base = ((size_t)b * N * I + i) * J;
  const size_t cs = (size_t)I * J;
  const float* ph0 = phi + (size_t)b * sb + (size_t)i * si;
  const float invJ = 1.0f / (float)J;
  for (int n = 0; n < N; ++n) {
    float nr[N], ni[N], dn[N];
#pragma unroll
    for (int m = 0; m < N; ++m) nr[m] = ni[m] = dn[m] = 0.f;
    for (int j = lane; j < J; j += 32) {
      cf yn = Y[base + n * cs + j];
      float a2 = yn.x * yn.x + yn.y * yn.y;
#pragma unroll
      for (int m = 0; m < N; ++m) {
        float ph = ph0[(size_t)m * sn + j];
        cf ym = Y[base + m * cs + j];
        float pr = ph * ym.x, pi = ph * ym.y;
        nr[m] = fmaf(pr, yn.x, fmaf(pi, yn.y, nr[m]));   // Re(ph ym conj(yn))
        ni[m] = fmaf(pi, yn.x, fmaf(-pr, yn.y, ni[m]));  // Im
        dn[m] = fmaf(ph, a2, dn[m]);
      }
    }
    cf v[N];
#pragma unroll
    for (int m = 0; m < N; ++m) {
      float r_ = warp_sum(nr[m]) * invJ, i_ = warp_sum(ni[m]) * invJ;
      float d_ = ssb_floor(warp_sum(dn[m]) * invJ, flooring, eps);
      if (m == n)
        v[m] = make_float2(1.0f - 1.0f / sqrtf(d_), 0.f);
      else
        v[m] = make_float2(r_ / d_, i_ / d_);
    }
    for (int j = lane; j < J; j += 32) {
      cf yn = Y[base + n * cs + j];
#pragma unroll
      for (int m = 0; m < N; ++m) {
        cf ym = Y[base + m * cs + j];
        ym.x -= v[m].x * yn.x - v[m].y * yn.y;
        ym.y -= v[m].x * yn.y + v[m].y * yn.x;
        Y[base + m * cs + j] = ym;
      }
    }
    __syncwarp();
  }
}

// ISS1 with the bin's (source x frame) slab of Y and of the weights resident in shared memory: Y is read
// and written once per iteration instead of 2N times.  One warp per (b,i); dynamic shared memory
// holds WPBS slabs of N*J complex64 + N*J float.
template <int N>
__global__ void k_iss1_smem(cf* __restrict__ Y, const float* __restrict__ phi, long long sb, long long sn,
                            long long si, int B, int I, int J, int flooring, float eps, int wpb) {
  extern __shared__ __align__(16) unsigned char iss_smem[];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp = blockIdx.x * wpb + wib;
  if (warp >= B * I) return;
  cf* ys = reinterpret_cast<cf*>(iss_smem) + (size_t)wib * N * J;
  float* ps = reinterpret_cast<float*>(reinterpret_cast<cf*>(iss_smem) + (size_t)wpb * N * J) + (size_t)wib * N * J;
  const int b = warp / I, i = warp - b * I;
  const size_t base = ((size_t)b * N * I + i) * J;
  const size_t cs = (size_t)I * J;
  const float* ph0 = phi + (size_t)b * sb + (size_t)i * si;
  const float invJ = 1.0f / (float)J;
#pragma unroll
  for (int m = 0; m < N; ++m)
    for (int j = lane; j < J; j += 32) {
      ys[m * J + j] = Y[base + m * cs + j];
      ps[m * J + j] = ph0[(size_t)m * sn + j];
    }
  __syncwarp();
  for (int n = 0; n < N; ++n) {
    float nr[N], ni[N], dn[N];
#pragma unroll
    for (int m = 0; m < N; ++m) nr[m] = ni[m] = dn[m] = 0.f;
    for (int j = lane; j < J; j += 32) {
      const cf yn = ys[n * J + j];
      const float a2 = yn.x * yn.x + yn.y * yn.y;
#pragma unroll
      for (int m = 0; m < N; ++m) {
        const float ph = ps[m * J + j];
        const cf ym = ys[m * J + j];
        const float pr = ph * ym.x, pi = ph * ym.y;
        nr[m] = fmaf(pr, yn.x, fmaf(pi, yn.y, nr[m]));
        ni[m] = fmaf(pi, yn.x, fmaf(-pr, yn.y, ni[m]));
        dn[m] = fmaf(ph, a2, dn[m]);
      }
    }
    cf v[N];
#pragma unroll
    for (int m = 0; m < N; ++m) {
      const float r_ = warp_sum(nr[m]) * invJ, i_ = warp_sum(ni[m]) * invJ;
      const float d_ = ssb_floor(warp_sum(dn[m]) * invJ, flooring, eps);
      if (m == n) v[m] = make_float2(1.0f - 1.0f / sqrtf(d_), 0.f);
      else v[m] = make_float2(r_ / d_, i_ / d_);
    }
    for (int j = lane; j < J; j += 32) {
      const cf yn = ys[n * J + j];
#pragma unroll
      for (int m = 0; m < N; ++m) {
        cf ym = ys[m * J + j];
        ym.x -= v[m].x * yn.x - v[m].y * yn.y;
        ym.y -= v[m].x * yn.y + v[m].y * yn.x;
        ys[m * J + j] = ym;
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int m = 0; m < N; ++m)
    for (int j = lane; j < J; j += 32) Y[base + m * cs + j] = ys[m * J + j];
}

// ISS1, one CTA (ISS_NW warps) per (b,i): the bin's Y and weight slabs live in shared memory, the
// frames are split over all threads and the 3N statistics of each of the N sequential steps are
// combined through shared memory (one __syncthreads per step).  Small per-CTA footprint (N*J*12 bytes)
// => many resident warps, unlike the warp-per-bin variant above.
constexpr int ISS_NW = 4;
template <int N>
__global__ void __launch_bounds__(ISS_NW * 32) k_iss1_cta(cf* __restrict__ Y, const float* __restrict__ phi,
                                                          long long sb, long long sn, long long si, int I, int J,
                                                          int flooring, float eps) {
  extern __shared__ __align__(16) unsigned char iss_smem[];
  __shared__ float red[2][ISS_NW][3 * N];
  cf* ys = reinterpret_cast<cf*>(iss_smem);
  float* ps = reinterpret_cast<float*>(ys + (size_t)N * J);
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int bi = blockIdx.x;
  const int b = bi / I, i = bi - b * I;
  const size_t base = ((size_t)b * N * I + i) * J;
  const size_t cs = (size_t)I * J;
  const float* ph0 = phi + (size_t)b * sb + (size_t)i * si;
  const float invJ = 1.0f / (float)J;
#pragma unroll
  for (int m = 0; m < N; ++m)
    for (int j = tid; j < J; j += ISS_NW * 32) {
      ys[m * J + j] = Y[base + m * cs + j];
      ps[m * J + j] = ph0[(size_t)m * sn + j];
    }
  // every thread only ever touches its own frames of the slab, so no barrier is needed for ys/ps
  for (int n = 0; n < N; ++n) {
    float nr[N], ni[N], dn[N];
#pragma unroll
    for (int m = 0; m < N; ++m) nr[m] = ni[m] = dn[m] = 0.f;
    for (int j = tid; j < J; j += ISS_NW * 32) {
      const cf yn = ys[n * J + j];
      const float a2 = yn.x * yn.x + yn.y * yn.y;
#pragma unroll
      for (int m = 0; m < N; ++m) {
        const float ph = ps[m * J + j];
        const cf ym = ys[m * J + j];
        const float pr = ph * ym.x, pi = ph * ym.y;
        nr[m] = fmaf(pr, yn.x, fmaf(pi, yn.y, nr[m]));
        ni[m] = fmaf(pi, yn.x, fmaf(-pr, yn.y, ni[m]));
        dn[m] = fmaf(ph, a2, dn[m]);
      }
    }
#pragma unroll
    for (int m = 0; m < N; ++m) {
      const float a = warp_sum(nr[m]), c = warp_sum(ni[m]), d = warp_sum(dn[m]);
      if (lane == 0) {
        red[n & 1][w][3 * m] = a;
        red[n & 1][w][3 * m + 1] = c;
        red[n & 1][w][3 * m + 2] = d;
      }
    }
    __syncthreads();
    cf v[N];
#pragma unroll
    for (int m = 0; m < N; ++m) {
      float a = 0.f, c = 0.f, d = 0.f;
#pragma unroll
      for (int ww = 0; ww < ISS_NW; ++ww) {  // fixed order: deterministic
        a += red[n & 1][ww][3 * m];
        c += red[n & 1][ww][3 * m + 1];
        d += red[n & 1][ww][3 * m + 2];
      }
      const float d_ = ssb_floor(d * invJ, flooring, eps);
      if (m == n) v[m] = make_float2(1.0f - 1.0f / sqrtf(d_), 0.f);
      else v[m] = make_float2(a * invJ / d_, c * invJ / d_);
    }
    for (int j = tid; j < J; j += ISS_NW * 32) {
      const cf yn = ys[n * J + j];
#pragma unroll
      for (int m = 0; m < N; ++m) {
        cf ym = ys[m * J + j];
        ym.x -= v[m].x * yn.x - v[m].y * yn.y;
        ym.y -= v[m].x * yn.y + v[m].y * yn.x;
        ys[m * J + j] = ym;
      }
    }
  }
#pragma unroll
  for (int m = 0; m < N; ++m)
    for (int j = tid; j < J; j += ISS_NW * 32) Y[base + m * cs + j] = ys[m * J + j];
}

// ------------------------------------------------------------------------------------------------
// ISS1 in the covariance domain (N <= 4), one WARP per (b,i).  The N sequential steps of
// ssspy/bss/_update_spatial_model.py:181-192 only need second-order statistics of the slab: with y = A y_old,
//   num_m = mean_j phi_m y_m conj(y_n) = a_m U_m a_n^H,   den_m = mean_j phi_m |y_n|^2 = a_n U_m a_n^H,
//   U_m = mean_j phi_m y_old y_old^H,   a_m = row m of A,
// and the update y_m -= v_m y_n is A[m,:] -= v_m A[n,:].  So: one sweep over the slab for the N weighted
// covariances (Hermitian products of a frame shared by the N sources, one reduce-scatter over the lanes instead of
// N x 3N warp reductions), the N rank-one updates of A in fp64 with A distributed over the lanes (lane (m, r) owns
// A[m][r] and row r of U_m), and one sweep applying A (the slab is re-read from L2: the footprint of all resident
// warps is a few tens of MB).  No shared-memory slab, no CTA barrier, any n_frames.
constexpr int ISSC_W = 4;  // warps (= bins) per block
//
// r2part != NULL (AuxIVA inside ssb_run): the apply sweep also accumulates |y_new|^2 per (source, frame) over the bins a
// warp walks (grid = (groups, mixtures); warp w of group gx takes bins gx * W + w, + groups * W, ...), in shared memory
// owned by the warp (lane j-strided: no atomics), and the block writes one partial r2part[b, gx, n, j]: the next
// iteration's r[n, j] = ||y_n[:, j]|| (ssspy/bss/iva.py:1961-1964) then needs no pass over Y at all.
template <int N>
__global__ void __launch_bounds__(ISSC_W * 32) k_iss1_cov(cf* __restrict__ Y, const float* __restrict__ phi,
                                                          long long sb, long long sn, long long si, int I, int J,
                                                          int flooring, float eps, float* __restrict__ r2part) {
  constexpr int NO = N * (N - 1) / 2;      // off-diagonal pairs (a < c)
  constexpr int NV = N * N;                // reals per source: NO (re, im) pairs + N diagonals
  constexpr int NVAL = N * NV;             // statistics per bin
  constexpr int NVP = (NVAL + 31) / 32 * 32;
  constexpr int PER = NVP / 32;
  __shared__ float s_red[ISSC_W][NVP];
  __shared__ cf s_A[ISSC_W][N * N];
  extern __shared__ float s_r2[];  // [ISSC_W][N][J] when r2part != NULL
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const size_t cs = (size_t)I * J;
  float* my_r2 = s_r2 + (size_t)w * N * J;
  if (r2part != nullptr) {
    for (int e = lane; e < N * J; e += 32) my_r2[e] = 0.f;
    __syncwarp();
  }
  for (int i = blockIdx.x * ISSC_W + w; i < I; i += gridDim.x * ISSC_W) {
  const size_t base = ((size_t)b * N * I + i) * J;
  const float* ph0 = phi + (size_t)b * sb + (size_t)i * si;
  // (prefetch.global.L2 of the warp's next bin at this point was measured: 2.83 vs 2.78 ms at config 3, no gain)
  // ---- sweep 1: weighted covariances ---------------------------------------------------------------------------
  // accumulators as (re, im) pairs [and pairs of diagonal entries], one packed FFMA2 per pair and source
  constexpr int NDP = (N + 1) / 2;
  float2 ao[N][NO], ad[N][NDP];
#pragma unroll
  for (int m = 0; m < N; ++m) {
#pragma unroll
    for (int e = 0; e < NO; ++e) ao[m][e] = make_float2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < NDP; ++k) ad[m][k] = make_float2(0.f, 0.f);
  }
#pragma unroll 2
  for (int j = lane; j < J; j += 32) {
    cf y[N];
    float2 pp[N];
#pragma unroll
    for (int m = 0; m < N; ++m) {
      y[m] = Y[base + m * cs + j];
      const float ph = ph0[(size_t)m * sn + j];
      pp[m] = make_float2(ph, ph);
    }
    int e = 0;
#pragma unroll
    for (int a = 0; a < N; ++a)
#pragma unroll
      for (int c = a + 1; c < N; ++c, ++e) {
        const float2 pr = make_float2(fmaf(y[a].x, y[c].x, y[a].y * y[c].y), fmaf(y[a].y, y[c].x, -(y[a].x * y[c].y)));
#pragma unroll
        for (int m = 0; m < N; ++m) ao[m][e] = __ffma2_rn(pp[m], pr, ao[m][e]);
      }
#pragma unroll
    for (int k = 0; k < NDP; ++k) {
      const int a = 2 * k, c = (2 * k + 1 < N) ? 2 * k + 1 : 2 * k;
      const float2 pd = make_float2(fmaf(y[a].x, y[a].x, y[a].y * y[a].y),
                                    (2 * k + 1 < N) ? fmaf(y[c].x, y[c].x, y[c].y * y[c].y) : 0.f);
#pragma unroll
      for (int m = 0; m < N; ++m) ad[m][k] = __ffma2_rn(pp[m], pd, ad[m][k]);
    }
  }
  float acc[NVP];  // per source m: [m*NV + 2e], [m*NV + 2e + 1] = Re, Im of pair e; [m*NV + 2*NO + a] = diagonal a
#pragma unroll
  for (int e = 0; e < NVP; ++e) acc[e] = 0.f;
#pragma unroll
  for (int m = 0; m < N; ++m) {
#pragma unroll
    for (int e = 0; e < NO; ++e) {
      acc[m * NV + 2 * e] = ao[m][e].x;
      acc[m * NV + 2 * e + 1] = ao[m][e].y;
    }
#pragma unroll
    for (int a = 0; a < N; ++a) acc[m * NV + 2 * NO + a] = (a & 1) ? ad[m][a >> 1].y : ad[m][a >> 1].x;
  }
  // reduce-scatter over the lanes: after the step with offset o a lane keeps the half of its values selected by
  // bit o of its id; lane l ends with PER consecutive values of the warp sum
  {
    int cnt = NVP;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      cnt >>= 1;
      const bool up = (lane & o) != 0;
#pragma unroll
      for (int e = 0; e < NVP / 2; ++e) {
        if (e < cnt) {
          const float send = up ? acc[e] : acc[e + cnt];
          const float keep = up ? acc[e + cnt] : acc[e];
          acc[e] = keep + __shfl_xor_sync(SSB_FULL, send, o);
        }
      }
    }
    int off = 0, span = NVP;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      span >>= 1;
      if (lane & o) off += span;
    }
#pragma unroll
    for (int q = 0; q < PER; ++q) s_red[w][off + q] = acc[q];
  }
  __syncwarp();
  // ---- N rank-one updates of A: lane (m, r) owns A[m][r] and row r of U_m (fp64) ----------------------------------
  {
    const int m = lane / N, r = lane - m * N;
    const bool act = lane < N * N;
    const double invJ = 1.0 / (double)J;
    cd Um[N];  // U_m[r][s], s = 0..N-1
#pragma unroll
    for (int s_ = 0; s_ < N; ++s_) {
      double re = 0.0, im = 0.0;
      if (act) {
        if (s_ == r) {
          re = (double)s_red[w][m * NV + 2 * NO + r];
        } else {
          const int a = r < s_ ? r : s_, c = r < s_ ? s_ : r;
          const int e = a * N - a * (a + 1) / 2 + (c - a - 1);
          re = (double)s_red[w][m * NV + 2 * e];
          im = (double)s_red[w][m * NV + 2 * e + 1];
          if (r > s_) im = -im;  // U[c][a] = conj(U[a][c])
        }
      }
      Um[s_] = cd_make(re * invJ, im * invJ);
    }
    cd amr = cd_make((act && m == r) ? 1.0 : 0.0, 0.0);  // A[m][r]
#pragma unroll
    for (int n = 0; n < N; ++n) {
      // row n of A, gathered from lanes (n, s)
      cd an[N];
#pragma unroll
      for (int s_ = 0; s_ < N; ++s_) an[s_] = shfl_cd(amr, n * N + s_);
      // t_r = sum_s U_m[r][s] conj(A[n][s]);  num_m = sum_r A[m][r] t_r;  den_m = sum_r A[n][r] t_r
      cd tr = cd_make(0, 0);
#pragma unroll
      for (int s_ = 0; s_ < N; ++s_) tr = cd_fma(Um[s_], cd_conj(an[s_]), tr);
      cd anr = an[0];
#pragma unroll
      for (int s_ = 1; s_ < N; ++s_)
        if (r == s_) anr = an[s_];
      cd pn = act ? cd_mul(amr, tr) : cd_make(0, 0);
      double pden = act ? cd_mul(anr, tr).x : 0.0;
      if (N == 3) {  // the three lanes of a source are not an xor group
        cd sacc = cd_make(0, 0);
        double dacc = 0.0;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
          const int src = (m * 3 + q) & 31;
          sacc.x += __shfl_sync(SSB_FULL, pn.x, src);
          sacc.y += __shfl_sync(SSB_FULL, pn.y, src);
          dacc += __shfl_sync(SSB_FULL, pden, src);
        }
        pn = sacc;
        pden = dacc;
      } else {
#pragma unroll
        for (int o = 1; o < N; o <<= 1) {
          pn.x += __shfl_xor_sync(SSB_FULL, pn.x, o);
          pn.y += __shfl_xor_sync(SSB_FULL, pn.y, o);
          pden += __shfl_xor_sync(SSB_FULL, pden, o);
        }
      }
      // same rounding points as the step-by-step kernel: fp32 statistics, floor, fp32 quotient
      const float d_ = ssb_floor((float)pden, flooring, eps);
      cd v;
      if (m == n) v = cd_make((double)(1.0f - 1.0f / sqrtf(d_)), 0.0);
      else v = cd_make((double)((float)pn.x / d_), (double)((float)pn.y / d_));
      amr = cd_sub(amr, cd_mul(v, anr));  // A[m][r] -= v_m A[n][r]
    }
    if (act) s_A[w][lane] = cd2cf(amr);
  }
  __syncwarp();
  // ---- sweep 2: y <- A y (slab re-read through L2) -----------------------------------------------------------------
  // out_p = sum_q a_pq y_q as (re, im) pairs: (y.x, y.x) * (a.x, a.y) + (y.y, y.y) * (-a.y, a.x), two FFMA2 per term
  float2 a_[N * N], as_[N * N];
#pragma unroll
  for (int e = 0; e < N * N; ++e) {
    a_[e] = s_A[w][e];
    as_[e] = make_float2(-a_[e].y, a_[e].x);
  }
#pragma unroll 2
  for (int j = lane; j < J; j += 32) {
    float2 yx[N], yy[N];
#pragma unroll
    for (int m = 0; m < N; ++m) {
      const cf y = Y[base + m * cs + j];
      yx[m] = make_float2(y.x, y.x);
      yy[m] = make_float2(y.y, y.y);
    }
#pragma unroll
    for (int p = 0; p < N; ++p) {
      float2 o = make_float2(0.f, 0.f);
#pragma unroll
      for (int q = 0; q < N; ++q) {
        o = __ffma2_rn(yx[q], a_[p * N + q], o);
        o = __ffma2_rn(yy[q], as_[p * N + q], o);
      }
      Y[base + p * cs + j] = o;
      if (r2part != nullptr) my_r2[p * J + j] += fmaf(o.x, o.x, o.y * o.y);
    }
  }
  __syncwarp();  // s_red / s_A are reused by this warp's next bin
  }
  if (r2part != nullptr) {
    __syncthreads();
    float* out = r2part + ((size_t)b * gridDim.x + blockIdx.x) * N * J;
    for (int e = threadIdx.x; e < N * J; e += ISSC_W * 32) {
      float v = 0.f;
#pragma unroll
      for (int ww = 0; ww < ISSC_W; ++ww) v += s_r2[(size_t)ww * N * J + e];
      out[e] = v;
    }
  }
}

// r2[b, n, j] = sum_g r2part[b, g, n, j] (fixed order)
__global__ void k_r2_sum(const float* __restrict__ part, float* __restrict__ r2, int G, int NJ, int B) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (e >= NJ) return;
  float v = 0.f;
  for (int g = 0; g < G; ++g) v += part[((size_t)b * G + g) * NJ + e];
  r2[(size_t)b * NJ + e] = v;
}

// ------------------------------------------------------------------------------------------------
// ISS2 (ssspy/bss/_update_spatial_model.py:197-314): pairwise iterative source steering, one CTA per (b,i).
// For every pair (m, n), u = (y_m, y_n):
//   all sources s:   G_s = mean_j phi_s u u^H                                   (2x2 Hermitian)
//   s not in pair:   f_s = mean_j phi_s u conj(y_s),  q_s = -G_s^-1 f_s,  y_s += q_s^H u       (:263-283)
//   pair:            G_m h = l G_n h; y_m <- p_0^H u with the smaller-eigenvalue vector normalised by G_m,
//                    y_n <- p_1^H u with the larger one normalised by G_n      (:288-303)
// Like k_iss1_cta every thread owns its frames of the (source x frame) slab, which lives in shared memory when it
// fits (ys / ps point into it) and is updated in place in global memory otherwise; the 8N statistics of a pair are
// combined through shared memory in a fixed order.
struct Iss2Pairs {
  int n_pairs;
  int m[SSB_MAX_PAIRS], n[SSB_MAX_PAIRS];
};

template <int N>
__global__ void __launch_bounds__(ISS_NW * 32) k_iss2_cta(cf* __restrict__ Y, const float* __restrict__ phi,
                                                          long long sb, long long sn, long long si, int I, int J,
                                                          int flooring, float eps, Iss2Pairs pl, int use_smem) {
  extern __shared__ __align__(16) unsigned char iss_smem[];
  __shared__ float red[2][ISS_NW][8 * N];
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int bi = blockIdx.x;
  const int b = bi / I, i = bi - b * I;
  const size_t base = ((size_t)b * N * I + i) * J;
  const size_t cs = (size_t)I * J;
  const float* ph0 = phi + (size_t)b * sb + (size_t)i * si;
  const double invJ = 1.0 / (double)J;
  cf* ys = Y + base;           // row stride ystr
  const float* ps = ph0;       // row stride pstr
  size_t ystr = cs, pstr = (size_t)sn;
  if (use_smem) {
    cf* ys_s = reinterpret_cast<cf*>(iss_smem);
    float* ps_s = reinterpret_cast<float*>(ys_s + (size_t)N * J);
#pragma unroll
    for (int m = 0; m < N; ++m)
      for (int j = tid; j < J; j += ISS_NW * 32) {
        ys_s[m * J + j] = Y[base + m * cs + j];
        ps_s[m * J + j] = ph0[(size_t)m * sn + j];
      }
    ys = ys_s;
    ps = ps_s;
    ystr = pstr = (size_t)J;
  }
  for (int q = 0; q < pl.n_pairs; ++q) {
    const int pm = pl.m[q], pn = pl.n[q];
    float st[N][8];
#pragma unroll
    for (int s_ = 0; s_ < N; ++s_)
#pragma unroll
      for (int e = 0; e < 8; ++e) st[s_][e] = 0.f;
    for (int j = tid; j < J; j += ISS_NW * 32) {
      const cf u0 = ys[pm * ystr + j], u1 = ys[pn * ystr + j];
      const float a00 = u0.x * u0.x + u0.y * u0.y, a11 = u1.x * u1.x + u1.y * u1.y;
      const float a01r = u0.x * u1.x + u0.y * u1.y, a01i = u0.y * u1.x - u0.x * u1.y;  // u0 conj(u1)
#pragma unroll
      for (int s_ = 0; s_ < N; ++s_) {
        const float ph = ps[s_ * pstr + j];
        const cf y = ys[s_ * ystr + j];
        st[s_][0] = fmaf(ph, a00, st[s_][0]);
        st[s_][1] = fmaf(ph, a11, st[s_][1]);
        st[s_][2] = fmaf(ph, a01r, st[s_][2]);
        st[s_][3] = fmaf(ph, a01i, st[s_][3]);
        const float pr = ph * y.x, pi = ph * y.y;  // phi conj(y_s) = (pr, -pi)
        st[s_][4] = fmaf(u0.x, pr, fmaf(u0.y, pi, st[s_][4]));
        st[s_][5] = fmaf(u0.y, pr, fmaf(-u0.x, pi, st[s_][5]));
        st[s_][6] = fmaf(u1.x, pr, fmaf(u1.y, pi, st[s_][6]));
        st[s_][7] = fmaf(u1.y, pr, fmaf(-u1.x, pi, st[s_][7]));
      }
    }
#pragma unroll
    for (int s_ = 0; s_ < N; ++s_)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float v = warp_sum(st[s_][e]);
        if (lane == 0) red[q & 1][w][s_ * 8 + e] = v;
      }
    __syncthreads();
    // coefficients: y_s' = keep_s y_s + conj(c0_s) u0 + conj(c1_s) u1
    cf c0[N], c1[N];
    cd Gm[4], Gn[4];
#pragma unroll
    for (int s_ = 0; s_ < N; ++s_) {
      double t[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float a = 0.f;
#pragma unroll
        for (int ww = 0; ww < ISS_NW; ++ww) a += red[q & 1][ww][s_ * 8 + e];  // fixed order: deterministic
        t[e] = (double)a * invJ;
      }
      if (s_ == pm || s_ == pn) {
        cd* G = s_ == pm ? Gm : Gn;
        G[0] = cd_make(t[0], 0);
        G[1] = cd_make(t[2], t[3]);
        G[2] = cd_make(t[2], -t[3]);
        G[3] = cd_make(t[1], 0);
        c0[s_] = c1[s_] = make_float2(0.f, 0.f);
      } else {
        // q = -G^-1 f with G = [[g00, g01], [conj(g01), g11]]
        const cd g01 = cd_make(t[2], t[3]), f0 = cd_make(t[4], t[5]), f1 = cd_make(t[6], t[7]);
        const double det = t[0] * t[1] - cd_abs2(g01);
        const double idet = -1.0 / det;
        const cd q0 = cd_scale(cd_sub(cd_scale(f0, t[1]), cd_mul(g01, f1)), idet);
        const cd q1 = cd_scale(cd_sub(cd_scale(f1, t[0]), cd_mul(cd_conj(g01), f0)), idet);
        c0[s_] = cd2cf(q0);
        c1[s_] = cd2cf(q1);
      }
    }
    {
      cd hs[2], hl[2];
      gen_eig2(Gm, Gn, hs, hl);
      const double dm = 1.0 / ssb_floor(sqrt(fmax(quad2(Gm, hs), 0.0)), flooring, (double)eps);
      const double dn = 1.0 / ssb_floor(sqrt(fmax(quad2(Gn, hl), 0.0)), flooring, (double)eps);
#pragma unroll
      for (int s_ = 0; s_ < N; ++s_) {
        if (s_ == pm) {
          c0[s_] = cd2cf(cd_scale(hs[0], dm));
          c1[s_] = cd2cf(cd_scale(hs[1], dm));
        } else if (s_ == pn) {
          c0[s_] = cd2cf(cd_scale(hl[0], dn));
          c1[s_] = cd2cf(cd_scale(hl[1], dn));
        }
      }
    }
    for (int j = tid; j < J; j += ISS_NW * 32) {
      const cf u0 = ys[pm * ystr + j], u1 = ys[pn * ystr + j];
#pragma unroll
      for (int s_ = 0; s_ < N; ++s_) {
        cf y = ys[s_ * ystr + j];
        if (s_ == pm || s_ == pn) y = make_float2(0.f, 0.f);
        // conj(c) u = (c.x u.x + c.y u.y, c.x u.y - c.y u.x)
        y.x += c0[s_].x * u0.x + c0[s_].y * u0.y + c1[s_].x * u1.x + c1[s_].y * u1.y;
        y.y += c0[s_].x * u0.y - c0[s_].y * u0.x + c1[s_].x * u1.y - c1[s_].y * u1.x;
        ys[s_ * ystr + j] = y;
      }
    }
  }
  if (use_smem) {
#pragma unroll
    for (int m = 0; m < N; ++m)
      for (int j = tid; j < J; j += ISS_NW * 32) Y[base + m * cs + j] = ys[m * J + j];
  }
}

// ------------------------------------------------------------------------------------------------
// projection back, filter form (ssspy/algorithm/projection_back.py:87-99): one warp per matrix.
// scale_out[mat*N + n] (optional) receives (W^-1)[ref, n] for the projection-back normalisation.
template <int N>
__global__ void __launch_bounds__(WPB * 32) k_pb_w(const cf* __restrict__ W, cf* __restrict__ Wout,
                                                   cf* __restrict__ scale_out, int n_mat, int ref,
                                                   int* __restrict__ status) {
  __shared__ cd sA[WPB][N * 2 * N];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mat = blockIdx.x * WPB + wib;
  if (mat >= n_mat) return;
  cd* A = sA[wib];
  for (int e = lane; e < N * 2 * N; e += 32) {
    int r = e / (2 * N), c = e - r * 2 * N;
    A[e] = c < N ? cf2cd(W[(size_t)mat * N * N + r * N + c]) : cd_make(c - N == r ? 1.0 : 0.0, 0);
  }
  __syncwarp();
  bool sing;
  warp_gauss_jordan(A, N, N, 2 * N, lane, nullptr, &sing);
  if (sing && lane == 0) atomicOr(status, SSB_STATUS_SINGULAR);  // np.linalg.inv raises here (projection_back.py:89)
  // scale[n] = Winv[ref][n] = A[ref][N + n]
  for (int e = lane; e < N * N; e += 32) {
    int n = e / N;
    cd s = A[ref * 2 * N + N + n];
    cd wv = cf2cd(W[(size_t)mat * N * N + e]);
    Wout[(size_t)mat * N * N + e] = cd2cf(cd_mul(wv, s));
  }
  if (scale_out && lane < N) scale_out[(size_t)mat * N + lane] = cd2cf(A[ref * 2 * N + N + lane]);
}

// ------------------------------------------------------------------------------------------------
// cross-solve  S[b,i] = (A_i Bm_i^H) (Bm_i Bm_i^H)^-1  with A_i, Bm_i the (N x J) slabs of bin i.
//   projection back, spectrogram form: A = X, Bm = Y (projection_back.py:104-110)
//   W recovery for the ISS-mode loss:  A = Y, Bm = X (ilrma.py:1939-1944, iva.py:2180-2185)
// One warp per (b,i); N+1 passes over the slab (Gram matrix, then one row of A Bm^H per pass).
template <int N>
__global__ void __launch_bounds__(WPB * 32) k_cross_solve(const cf* __restrict__ Am, const cf* __restrict__ Bm,
                                                          cf* __restrict__ S, int B, int I, int J,
                                                          int* __restrict__ status) {
  __shared__ cd sG[WPB][N * 2 * N];  // [Gram | I] -> Gram^-1
  __shared__ cd sC[WPB][N * N];      // A Bm^H
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp = blockIdx.x * WPB + wib;
  if (warp >= B * I) return;
  const int b = warp / I, i = warp - b * I;
  const size_t base = ((size_t)b * N * I + i) * J;
  const size_t cs = (size_t)I * J;
  cd* G = sG[wib];
  cd* C = sC[wib];
  {
    float acc[N * N];
#pragma unroll
    for (int e = 0; e < N * N; ++e) acc[e] = 0.f;
    for (int j = lane; j < J; j += 32) {
      cf x[N];
#pragma unroll
      for (int m = 0; m < N; ++m) x[m] = Bm[base + m * cs + j];
#pragma unroll
      for (int a = 0; a < N; ++a) {
        acc[a * N + a] = fmaf(x[a].x, x[a].x, fmaf(x[a].y, x[a].y, acc[a * N + a]));
#pragma unroll
        for (int c = a + 1; c < N; ++c) {
          acc[a * N + c] = fmaf(x[a].x, x[c].x, fmaf(x[a].y, x[c].y, acc[a * N + c]));
          acc[c * N + a] = fmaf(x[a].y, x[c].x, fmaf(-x[a].x, x[c].y, acc[c * N + a]));
        }
      }
    }
#pragma unroll
    for (int e = 0; e < N * N; ++e) acc[e] = warp_sum(acc[e]);
    if (lane == 0) {
#pragma unroll
      for (int a = 0; a < N; ++a) {
#pragma unroll
        for (int c = 0; c < N; ++c) {
          cd v;
          if (a == c) v = cd_make(acc[a * N + a], 0);
          else if (a < c) v = cd_make(acc[a * N + c], acc[c * N + a]);
          else v = cd_make(acc[c * N + a], -acc[a * N + c]);
          G[a * 2 * N + c] = v;
          G[a * 2 * N + N + c] = cd_make(a == c ? 1.0 : 0.0, 0);
        }
      }
    }
  }
  for (int r = 0; r < N; ++r) {
    float cr[N], ci[N];
#pragma unroll
    for (int m = 0; m < N; ++m) cr[m] = ci[m] = 0.f;
    for (int j = lane; j < J; j += 32) {
      cf a = Am[base + r * cs + j];
#pragma unroll
      for (int m = 0; m < N; ++m) {
        cf x = Bm[base + m * cs + j];
        cr[m] = fmaf(a.x, x.x, fmaf(a.y, x.y, cr[m]));
        ci[m] = fmaf(a.y, x.x, fmaf(-a.x, x.y, ci[m]));
      }
    }
#pragma unroll
    for (int m = 0; m < N; ++m) {
      float vr = warp_sum(cr[m]), vi = warp_sum(ci[m]);
      if (lane == 0) C[r * N + m] = cd_make(vr, vi);
    }
  }
  __syncwarp();
  bool sing;
  warp_gauss_jordan(G, N, N, 2 * N, lane, nullptr, &sing);
  if (sing && lane == 0) atomicOr(status, SSB_STATUS_SINGULAR);  // np.linalg.inv of the Gram matrix (projection_back.py:110)
  for (int e = lane; e < N * N; e += 32) {
    int r = e / N, c = e - r * N;
    cd s = cd_make(0, 0);
#pragma unroll
    for (int k = 0; k < N; ++k) s = cd_fma(C[r * N + k], G[k * 2 * N + N + c], s);
    S[(size_t)warp * N * N + e] = cd2cf(s);
  }
}

// Yout[b,n,i,:] = Y[b,n,i,:] * S[b,i,ref,n]      (projection_back.py:117-119)
__global__ void k_scale_rows(const cf* __restrict__ Y, const cf* __restrict__ S, cf* __restrict__ Yout, int N,
                             int I, int J, int ref) {
  const int row = blockIdx.x;  // (b, n, i)
  const int i = row % I;
  const int n = (row / I) % N;
  const int b = row / (I * N);
  const cf s = S[(((size_t)b * I + i) * N + ref) * N + n];
  const size_t base = (size_t)row * J;
  for (int j = threadIdx.x; j < J; j += blockDim.x) {
    cf y = Y[base + j];
    Yout[base + j] = make_float2(y.x * s.x - y.y * s.y, y.x * s.y + y.y * s.x);
  }
}

// minimal distortion principle (ssspy/algorithm/minimal_distortion_principle.py:31-43): per (b,n,i) row
//   z = sum_j y conj(x_ref) / sum_j |y|^2,  y <- conj(z) y.     One warp per row; ref < 0 is invalid.
__global__ void __launch_bounds__(WPB * 32) k_mdp(const cf* __restrict__ Y, const cf* __restrict__ X,
                                                  cf* __restrict__ Yout, int rows, int N, int I, int J, int ref) {
  const int row = blockIdx.x * WPB + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int i = row % I;
  const int b = row / (I * N);
  const cf* y = Y + (size_t)row * J;
  const cf* x = X + (((size_t)b * N + ref) * I + i) * J;
  float nr = 0.f, ni = 0.f, de = 0.f;
  for (int j = lane; j < J; j += 32) {
    const cf a = y[j], c = x[j];
    nr = fmaf(a.x, c.x, fmaf(a.y, c.y, nr));   // Re(y conj(x))
    ni = fmaf(a.y, c.x, fmaf(-a.x, c.y, ni));  // Im
    de = fmaf(a.x, a.x, fmaf(a.y, a.y, de));
  }
  nr = warp_sum(nr);
  ni = warp_sum(ni);
  de = warp_sum(de);
  const float zr = nr / de, zi = ni / de;  // z; the scale applied is conj(z)
  cf* o = Yout + (size_t)row * J;
  for (int j = lane; j < J; j += 32) {
    const cf a = y[j];
    o[j] = make_float2(zr * a.x + zi * a.y, zr * a.y - zi * a.x);
  }
}

// log|det W| per matrix (np.linalg.slogdet call sites ilrma.py:534, iva.py:234); one warp per matrix.
template <int N>
__global__ void __launch_bounds__(WPB * 32) k_logdet(const cf* __restrict__ W, double* __restrict__ out,
                                                     int n_mat) {
  __shared__ cd sA[WPB][N * N];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mat = blockIdx.x * WPB + wib;
  if (mat >= n_mat) return;
  cd* A = sA[wib];
  for (int e = lane; e < N * N; e += 32) A[e] = cf2cd(W[(size_t)mat * N * N + e]);
  __syncwarp();
  double lad;
  warp_gauss_jordan(A, N, 0, N, lane, &lad);
  if (lane == 0) out[mat] = lad;
}

}  // namespace


int ssbk_separate(const cf* X, const cf* W, cf* Y, float* P, int B, int N, int I, int J, cudaStream_t st) {
  SSB_DISPATCH_N(N, k_separate<NN><<<B * I, 128, 0, st>>>(X, W, Y, P, I, J));
  return ssb_check_launch("separate", st);
}

int ssbk_abs2(const cf* Y, float* P, size_t n, cudaStream_t st) {
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 32) blocks = 148 * 32;
  k_abs2<<<blocks, 256, 0, st>>>(Y, P, n);
  return ssb_check_launch("abs2", st);
}

int ssbk_wcov(const cf* X, const float* phi, long long sb, long long sn, long long si, const int* src, int n_src,
              cf* U, int B, int N, int I, int J, cudaStream_t st) {
  SSB_REQUIRE(n_src >= 1 && n_src <= SSB_MAX_SOURCES, "weighted_covariance: n_src=%d out of range", n_src);
  SrcList sl;
  sl.n = n_src;
  for (int s = 0; s < n_src; ++s) sl.idx[s] = src ? src[s] : s;
  SSB_DISPATCH_N(N, k_wcov<NN><<<blocks_for((long long)B * I, WPB), WPB * 32, 0, st>>>(X, phi, sb, sn, si, sl, U,
                                                                                           B, I, J));
  return ssb_check_launch("weighted_covariance", st);
}

int ssbk_ip1(cf* W, const cf* U, int n_mat, int N, int flooring, float eps, cudaStream_t st, const cf* C, double* q) {
  SSB_DISPATCH_N(N, kq_ip1<NN><<<blocks_for(n_mat, QW * GroupShape<NN>::GW), QW * 32, 0, st>>>(W, U, n_mat, flooring,
                                                                                                 (double)eps, C, q,
                                                                                                 ssb_status_word()));
  return ssb_check_launch("update_by_ip1", st);
}

int ssbk_ip2(cf* W, const cf* U, int n_mat, int N, const int* pairs, int n_pairs, int n_u, const int* uidx,
             int flooring, float eps, cudaStream_t st, const cf* C, double* q) {
  SSB_REQUIRE(n_pairs >= 0 && n_pairs <= SSB_MAX_PAIRS, "update_by_ip2: n_pairs=%d exceeds %d", n_pairs,
              SSB_MAX_PAIRS);
  PairList pl;
  pl.n = n_pairs;
  pl.n_u = n_u;
  for (int q = 0; q < n_pairs; ++q) {
    int m = pairs[2 * q], n = pairs[2 * q + 1];
    SSB_REQUIRE(m >= 0 && m < N && n >= 0 && n < N && m != n, "update_by_ip2: invalid pair (%d, %d) for N=%d", m, n, N);
    pl.m[q] = (short)m;
    pl.nn[q] = (short)n;
    pl.um[q] = (short)(uidx ? uidx[2 * q] : m);
    pl.un[q] = (short)(uidx ? uidx[2 * q + 1] : n);
  }
  if (n_pairs == 0) return 0;
  // N = 8: the kernel capped at 128 registers (4 blocks per SM instead of 3; 28 bytes of spills): 2.62 -> 2.45 ms at
  // BASELINE config 4 (gpurun_out/r2p_c4_occ.json); SSB_IP2_OCC=0 restores the 168-register build
  static const int occ = getenv("SSB_IP2_OCC") != nullptr ? atoi(getenv("SSB_IP2_OCC")) : 1;
  if (N == 8 && occ == 1) {
    kq_ip2<8, 4><<<blocks_for(n_mat, QW * GroupShape<8>::GW), QW * 32, 0, st>>>(W, U, n_mat, pl, flooring, (double)eps, C,
                                                                              q, ssb_status_word());
    return ssb_check_launch("update_by_ip2", st);
  }
  SSB_DISPATCH_N(N, kq_ip2<NN><<<blocks_for(n_mat, QW * GroupShape<NN>::GW), QW * 32, 0, st>>>(W, U, n_mat, pl, flooring,
                                                                                                 (double)eps, C, q,
                                                                                                 ssb_status_word()));
  return ssb_check_launch("update_by_ip2", st);
}

int ssbk_iss1_r2_groups(int I) { return (I + ISSC_W * 16 - 1) / (ISSC_W * 16) < 1 ? 1 : (I + ISSC_W * 16 - 1) / (ISSC_W * 16); }

int ssbk_iss1_emits_r2(int N, int J) {
  const char* e = getenv("SSB_ISS_COV");
  return N <= 4 && (e == nullptr || atoi(e) != 0) && (size_t)ISSC_W * N * J * sizeof(float) <= 160 * 1024;
}

int ssbk_iss1(cf* Y, const float* phi, long long sb, long long sn, long long si, int B, int N, int I, int J,
              int flooring, float eps, cudaStream_t st, float* r2part, float* r2) {
  // CTA-per-bin shared-memory variants when the (source x frame) slab of Y + weights fits one CTA
  const size_t slab = (size_t)N * J * (sizeof(cf) + sizeof(float));
  static int cov_mode = -1;  // SSB_ISS_COV: 1 (default) covariance-domain kernel for N <= 4, 0 step-by-step kernels
  if (cov_mode < 0) {
    const char* e = getenv("SSB_ISS_COV");
    cov_mode = e ? atoi(e) : 1;
  }
  if (N <= 4 && cov_mode) {
    const bool emit = r2part != nullptr && r2 != nullptr && ssbk_iss1_emits_r2(N, J);
    // emitting: ~16 bins per warp (one partial per block and mixture); else one bin per warp as before
    const int groups = emit ? ssbk_iss1_r2_groups(I) : (I + ISSC_W - 1) / ISSC_W;
    const size_t sm = emit ? (size_t)ISSC_W * N * J * sizeof(float) : 0;
    float* part = emit ? r2part : nullptr;
    dim3 grid(groups, B);
    static bool attr_dev[SSB_MAX_DEVICES] = {};
    bool& attr_set = attr_dev[ssb_current_device()];
    if (!attr_set) {
      SSB_CUDA(cudaFuncSetAttribute(k_iss1_cov<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      SSB_CUDA(cudaFuncSetAttribute(k_iss1_cov<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      SSB_CUDA(cudaFuncSetAttribute(k_iss1_cov<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      attr_set = true;
    }
    if (N == 2) k_iss1_cov<2><<<grid, ISSC_W * 32, sm, st>>>(Y, phi, sb, sn, si, I, J, flooring, eps, part);
    else if (N == 3) k_iss1_cov<3><<<grid, ISSC_W * 32, sm, st>>>(Y, phi, sb, sn, si, I, J, flooring, eps, part);
    else k_iss1_cov<4><<<grid, ISSC_W * 32, sm, st>>>(Y, phi, sb, sn, si, I, J, flooring, eps, part);  // (capping it at 128
    // registers for four blocks per SM changes nothing: 2.83 vs 2.72 ms at config 3, gpurun_out/r2q_c3_*.json)
    if (ssb_check_launch("update_by_iss1", st)) return 1;
    if (emit) {
      dim3 g2(blocks_for((long long)N * J, 256), B);
      k_r2_sum<<<g2, 256, 0, st>>>(r2part, r2, groups, N * J, B);
      return ssb_check_launch("iva_r2_from_iss1", st);
    }
    return 0;
  }
  if (slab <= 200 * 1024) {
    SSB_DISPATCH_N(N, {
      static bool attr_dev[SSB_MAX_DEVICES] = {};  // function attributes are per device
  bool& attr_set = attr_dev[ssb_current_device()];
      if (!attr_set) {
        SSB_CUDA(cudaFuncSetAttribute(k_iss1_cta<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
      }
      k_iss1_cta<NN><<<B * I, ISS_NW * 32, slab, st>>>(Y, phi, sb, sn, si, I, J, flooring, eps);
    });
    return ssb_check_launch("update_by_iss1", st);
  }
  SSB_DISPATCH_N(N, k_iss1<NN><<<blocks_for((long long)B * I, WPB), WPB * 32, 0, st>>>(Y, phi, sb, sn, si, B, I, J,
                                                                                           flooring, eps));
  return ssb_check_launch("update_by_iss1", st);
}

int ssbk_iss2(cf* Y, const float* phi, long long sb, long long sn, long long si, int B, int N, int I, int J,
              const int* pairs, int n_pairs, int flooring, float eps, cudaStream_t st) {
  SSB_REQUIRE(n_pairs >= 0 && n_pairs <= SSB_MAX_PAIRS, "n_pairs=%d exceeds %d", n_pairs, SSB_MAX_PAIRS);
  Iss2Pairs pl{};
  pl.n_pairs = n_pairs;
  for (int q = 0; q < n_pairs; ++q) {
    pl.m[q] = pairs[2 * q];
    pl.n[q] = pairs[2 * q + 1];
    SSB_REQUIRE(pl.m[q] >= 0 && pl.m[q] < N && pl.n[q] >= 0 && pl.n[q] < N && pl.m[q] != pl.n[q],
                "invalid pair (%d, %d) for n_sources=%d", pl.m[q], pl.n[q], N);
  }
  const size_t slab = (size_t)N * J * (sizeof(cf) + sizeof(float));
  const int use_smem = slab <= 200 * 1024;
  SSB_DISPATCH_N(N, {
    static bool attr_dev[SSB_MAX_DEVICES] = {};  // function attributes are per device
  bool& attr_set = attr_dev[ssb_current_device()];
    if (!attr_set) {
      SSB_CUDA(cudaFuncSetAttribute(k_iss2_cta<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_set = true;
    }
    k_iss2_cta<NN><<<B * I, ISS_NW * 32, use_smem ? slab : 0, st>>>(Y, phi, sb, sn, si, I, J, flooring, eps, pl,
                                                                      use_smem);
  });
  return ssb_check_launch("update_by_iss2", st);
}

int ssbk_pb_w(const cf* W, cf* Wout, cf* scale_out, int n_mat, int N, int ref, cudaStream_t st) {
  SSB_REQUIRE(ref >= 0 && ref < N, "projection_back: reference_id=%d out of range for N=%d", ref, N);
  SSB_DISPATCH_N(N, k_pb_w<NN><<<blocks_for(n_mat, WPB), WPB * 32, 0, st>>>(W, Wout, scale_out, n_mat, ref,
                                                                            ssb_status_word()));
  return ssb_check_launch("projection_back_w", st);
}

int ssbk_cross_solve(const cf* A, const cf* Bm, cf* S, int B, int N, int I, int J, cudaStream_t st) {
  SSB_DISPATCH_N(N, k_cross_solve<NN><<<blocks_for((long long)B * I, WPB), WPB * 32, 0, st>>>(A, Bm, S, B, I, J,
                                                                                              ssb_status_word()));
  return ssb_check_launch("cross_solve", st);
}

int ssbk_scale_rows(const cf* Y, const cf* S, cf* Yout, int B, int N, int I, int J, int ref, cudaStream_t st) {
  SSB_REQUIRE(ref >= 0 && ref < N, "projection_back: reference_id=%d out of range for N=%d", ref, N);
  k_scale_rows<<<B * N * I, 128, 0, st>>>(Y, S, Yout, N, I, J, ref);
  return ssb_check_launch("scale_rows", st);
}

int ssbk_mdp(const cf* Y, const cf* X, cf* Yout, int B, int N, int I, int J, int ref, cudaStream_t st) {
  SSB_REQUIRE(ref >= 0 && ref < N, "minimal_distortion_principle: reference_id=%d out of range for N=%d", ref, N);
  const int rows = B * N * I;
  k_mdp<<<blocks_for(rows, WPB), WPB * 32, 0, st>>>(Y, X, Yout, rows, N, I, J, ref);
  return ssb_check_launch("minimal_distortion_principle", st);
}

int ssbk_logdet(const cf* W, double* out, int n_mat, int N, cudaStream_t st) {
  SSB_DISPATCH_N(N, k_logdet<NN><<<blocks_for(n_mat, WPB), WPB * 32, 0, st>>>(W, out, n_mat));
  return ssb_check_launch("logdet", st);
}
