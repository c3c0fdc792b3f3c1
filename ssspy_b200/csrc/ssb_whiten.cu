// Whitened-domain iteration: the numerical conditioning layer of the demixing-filter (IP1 / IP2) modes.
//
// The reference keeps every quantity in fp64, so the additive rounding of the weighted covariances U_n (1e-16) times
// the condition number of the mixture covariance C_i = mean_j x x^H (1e2 .. 1e6 per bin for N = 2 .. 8) is invisible.
// With complex64 device state the same product is 6e-8 x cond(C_i): measured on a B200 against the fp64 oracle,
// GaussILRMA-IP2 at N = 8 lost two digits (3.8e-2 on Y) and even IP1 at N = 8 sat at 1e-4.  The updates
//   w_n = (W U_n)^-1 e_n                       (ssspy/bss/_update_spatial_model.py:63-76)
//   P_q = (W U_q)^-1 [e_m e_n], A = P_m^H U_m P_m, ...    (:317-395)
// are equivariant under a per-bin change of basis of the observations: with z = M x, W~ = W M^-1 and
// U~_n = mean_j phi z z^H = M U_n M^H they read w~_n = (W~ U~_n)^-1 e_n with the same normalisation
// (w^H U w = w~^H U~ w~), and y = W x = W~ z is unchanged.  Choosing M = L^-1, C = L L^H (fp64 Cholesky of the fp64
// covariance) makes U~_n a weighted average of a unit-covariance signal: cond(U~) ~ max(phi) / min(phi) instead of
// cond(C) x that, so fp32 accumulation / complex64 storage errors are no longer amplified.  The oracle confirms it
// (same 1e-7 noise on U: 1.2e-1 -> 1.2e-5 on Y for IP2 at N = 8).
//
//   k_cov64          C[b,i] = mean_j x x^H, fp64 accumulation, complex128 out
//   k_whiten_factor  M = L^-1, M^-1 = L, log|det M| per bin (identity for a bin whose C is not positive definite)
//   k_whiten_apply   Z = M X in fp64, stored complex64: the slab every iteration kernel reads instead of X
//   k_w_import       W~ = W M^-1 for the bins whose user-visible W changed since the last export
//   k_w_export       W = W~ M (the reference's demix_filter, ssspy/bss/ilrma.py:186-188)
//   k_pb_whitened    projection back from the whitened filter: s_n = (M^-1 W~^-1)[ref, n]  (projection_back.py:87-99)
//   k_add_logdet     log|det W| = log|det W~| + log|det M|   (ilrma.py:524-536)
#include "ssb_kernels.h"

namespace {

constexpr int WB = 4;  // warps per block of the per-bin reductions

template <int N>
__global__ void __launch_bounds__(WB * 32) k_cov64(const cf* __restrict__ X, cd* __restrict__ C, int B, int I, int J) {
  const int warp = blockIdx.x * WB + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (warp >= B * I) return;
  const int b = warp / I, i = warp - b * I;
  const size_t base = ((size_t)b * N * I + i) * J, cs = (size_t)I * J;
  double re[N * (N + 1) / 2], im[N * (N + 1) / 2];
#pragma unroll
  for (int e = 0; e < N * (N + 1) / 2; ++e) re[e] = im[e] = 0.0;
  for (int j = lane; j < J; j += 32) {
    double xr[N], xi[N];
#pragma unroll
    for (int m = 0; m < N; ++m) {
      const cf v = X[base + m * cs + j];
      xr[m] = (double)v.x;
      xi[m] = (double)v.y;
    }
    int e = 0;
#pragma unroll
    for (int a = 0; a < N; ++a)
#pragma unroll
      for (int c = a; c < N; ++c, ++e) {  // x_a conj(x_c)
        re[e] = fma(xr[a], xr[c], fma(xi[a], xi[c], re[e]));
        im[e] = fma(xi[a], xr[c], fma(-xr[a], xi[c], im[e]));
      }
  }
  const double invJ = 1.0 / (double)J;
  cd* out = C + (size_t)warp * N * N;
  int e = 0;
#pragma unroll
  for (int a = 0; a < N; ++a)
#pragma unroll
    for (int c = a; c < N; ++c, ++e) {
      double r = re[e], q = im[e];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        r += __shfl_xor_sync(SSB_FULL, r, o);
        q += __shfl_xor_sync(SSB_FULL, q, o);
      }
      if (lane == 0) {
        out[a * N + c] = cd_make(r * invJ, a == c ? 0.0 : q * invJ);
        if (a != c) out[c * N + a] = cd_make(r * invJ, -q * invJ);
      }
    }
}

// One thread per bin: C = L L^H (lower), M = L^-1 by forward substitution; bins whose covariance is not numerically
// positive definite (a silent or duplicated channel) keep M = I, i.e. run unwhitened exactly as before.
template <int N>
__global__ void k_whiten_factor(const cd* __restrict__ C, cd* __restrict__ M, cd* __restrict__ Minv,
                                double* __restrict__ ldM, int n_mat) {
  const int mat = blockIdx.x * blockDim.x + threadIdx.x;
  if (mat >= n_mat) return;
  cd L[N][N], R[N][N];
#pragma unroll
  for (int a = 0; a < N; ++a)
#pragma unroll
    for (int c = 0; c < N; ++c) {
      L[a][c] = cd_make(0, 0);
      R[a][c] = cd_make(0, 0);
    }
  const cd* Cm = C + (size_t)mat * N * N;
  bool ok = true;
  double tr = 0.0;
  for (int a = 0; a < N; ++a) tr += Cm[a * N + a].x;
  ok = isfinite(tr) && tr > 0.0;
  double ld = 0.0;
  for (int c = 0; c < N && ok; ++c) {
    double d = Cm[c * N + c].x;
    for (int k = 0; k < c; ++k) d -= cd_abs2(L[c][k]);
    // relative floor: a pivot below 1e-12 of the mean channel power means a (numerically) rank-deficient bin
    if (!(d > 1e-12 * tr / N)) {
      ok = false;
      break;
    }
    const double l = sqrt(d);
    L[c][c] = cd_make(l, 0);
    ld += log(l);
    for (int a = c + 1; a < N; ++a) {
      cd s = Cm[a * N + c];
      for (int k = 0; k < c; ++k) s = cd_sub(s, cd_mulc(L[a][k], L[c][k]));
      L[a][c] = cd_scale(s, 1.0 / l);
    }
  }
  if (ok) {  // R = L^-1 (lower triangular), column by column
    for (int c = 0; c < N; ++c) {
      R[c][c] = cd_make(1.0 / L[c][c].x, 0);
      for (int a = c + 1; a < N; ++a) {
        cd s = cd_make(0, 0);
        for (int k = c; k < a; ++k) s = cd_fma(L[a][k], R[k][c], s);
        R[a][c] = cd_scale(s, -1.0 / L[a][a].x);
      }
    }
  } else {
    ld = 0.0;
#pragma unroll
    for (int a = 0; a < N; ++a)
#pragma unroll
      for (int c = 0; c < N; ++c) L[a][c] = R[a][c] = cd_make(a == c ? 1.0 : 0.0, 0);
  }
  cd* Mo = M + (size_t)mat * N * N;
  cd* Mi = Minv + (size_t)mat * N * N;
#pragma unroll
  for (int a = 0; a < N; ++a)
#pragma unroll
    for (int c = 0; c < N; ++c) {
      Mo[a * N + c] = R[a][c];
      Mi[a * N + c] = L[a][c];
    }
  ldM[mat] = -ld;  // log|det M| = -log det L
}

// Z[b,:,i,j] = M[b,i] X[b,:,i,j]: fp64 multiply-adds on the complex64 inputs, one rounding to complex64 at the end
template <int N>
__global__ void __launch_bounds__(256) k_whiten_apply(const cf* __restrict__ X, const cd* __restrict__ M,
                                                      cf* __restrict__ Z, int I, int J) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y, b = blockIdx.z;
  if (j >= J) return;
  const size_t base = ((size_t)b * N * I + i) * J + j, cs = (size_t)I * J;
  const cd* Mm = M + ((size_t)b * I + i) * N * N;
  cd x[N];
#pragma unroll
  for (int m = 0; m < N; ++m) x[m] = cf2cd(X[base + m * cs]);
#pragma unroll
  for (int a = 0; a < N; ++a) {
    cd s = cd_make(0, 0);
#pragma unroll
    for (int m = 0; m <= a; ++m) s = cd_fma(Mm[a * N + m], x[m], s);  // M is lower triangular
    Z[base + a * cs] = cd2cf(s);
  }
}

// W~ = W M^-1 where the user-visible W differs (bitwise) from what the library exported last, or was never imported
template <int N>
__global__ void k_w_import(const cf* __restrict__ W, cf* __restrict__ Wexp, cf* __restrict__ Ww,
                           const cd* __restrict__ Minv, int* __restrict__ wsync, int n_mat) {
  const int mat = blockIdx.x * blockDim.x + threadIdx.x;
  if (mat >= n_mat) return;
  const cf* w = W + (size_t)mat * N * N;
  cf* we = Wexp + (size_t)mat * N * N;
  bool same = wsync[mat] != 0;
  cf wv[N * N];
#pragma unroll
  for (int e = 0; e < N * N; ++e) {
    wv[e] = w[e];
    const cf o = we[e];
    same = same && (__float_as_uint(wv[e].x) == __float_as_uint(o.x)) && (__float_as_uint(wv[e].y) == __float_as_uint(o.y));
  }
  if (same) return;
  const cd* L = Minv + (size_t)mat * N * N;
  cf* ww = Ww + (size_t)mat * N * N;
#pragma unroll
  for (int r = 0; r < N; ++r)
#pragma unroll
    for (int c = 0; c < N; ++c) {
      cd s = cd_make(0, 0);
#pragma unroll
      for (int k = c; k < N; ++k) s = cd_fma(cf2cd(wv[r * N + k]), L[k * N + c], s);  // L lower triangular
      ww[r * N + c] = cd2cf(s);
    }
#pragma unroll
  for (int e = 0; e < N * N; ++e) we[e] = wv[e];
  wsync[mat] = 1;
}

// W = W~ M, remembered bit for bit in Wexp so that an untouched W is recognised by the next import
template <int N>
__global__ void k_w_export(const cf* __restrict__ Ww, const cd* __restrict__ M, cf* __restrict__ W,
                           cf* __restrict__ Wexp, int* __restrict__ wsync, int n_mat) {
  const int mat = blockIdx.x * blockDim.x + threadIdx.x;
  if (mat >= n_mat) return;
  const cf* ww = Ww + (size_t)mat * N * N;
  const cd* Mm = M + (size_t)mat * N * N;
  cf* w = W + (size_t)mat * N * N;
  cf* we = Wexp + (size_t)mat * N * N;
#pragma unroll
  for (int r = 0; r < N; ++r)
#pragma unroll
    for (int c = 0; c < N; ++c) {
      cd s = cd_make(0, 0);
#pragma unroll
      for (int k = c; k < N; ++k) s = cd_fma(cf2cd(ww[r * N + k]), Mm[k * N + c], s);
      const cf v = cd2cf(s);
      w[r * N + c] = v;
      we[r * N + c] = v;
    }
  wsync[mat] = 1;
}

// Projection back from the whitened filter (projection_back.py:87-99 with W = W~ M):
//   s_n = (W^-1)[ref, n] = sum_k Minv[ref, k] (W~^-1)[k, n];   W~[n, :] *= s_n
// One thread per bin, fp64 Gauss-Jordan with partial pivoting on [W~ | I].  scale_out[mat, n] (optional) = s_n.
template <int N>
__global__ void k_pb_whitened(cf* __restrict__ Ww, const cd* __restrict__ Minv, cf* __restrict__ scale_out, int n_mat,
                              int ref, int* __restrict__ status) {
  const int mat = blockIdx.x * blockDim.x + threadIdx.x;
  if (mat >= n_mat) return;
  cd A[N][2 * N];
  cf* ww = Ww + (size_t)mat * N * N;
#pragma unroll
  for (int r = 0; r < N; ++r)
#pragma unroll
    for (int c = 0; c < N; ++c) {
      A[r][c] = cf2cd(ww[r * N + c]);
      A[r][N + c] = cd_make(r == c ? 1.0 : 0.0, 0);
    }
  for (int p = 0; p < N; ++p) {
    int piv = p;
    double best = cd_abs2(A[p][p]);
    for (int r = p + 1; r < N; ++r) {
      const double v = cd_abs2(A[r][p]);
      if (v > best) {
        best = v;
        piv = r;
      }
    }
    if (piv != p)
      for (int c = 0; c < 2 * N; ++c) {
        const cd t = A[p][c];
        A[p][c] = A[piv][c];
        A[piv][c] = t;
      }
    if (cd_abs2(A[p][p]) == 0.0) atomicOr(status, SSB_STATUS_SINGULAR);  // np.linalg.inv raises (projection_back.py:89)
    const cd ipv = cd_inv(A[p][p]);
    for (int c = 0; c < 2 * N; ++c) A[p][c] = cd_mul(A[p][c], ipv);
    for (int r = 0; r < N; ++r) {
      if (r == p) continue;
      const cd f = A[r][p];
      for (int c = 0; c < 2 * N; ++c) A[r][c] = cd_sub(A[r][c], cd_mul(f, A[p][c]));
    }
  }
  const cd* L = Minv + (size_t)mat * N * N;
#pragma unroll
  for (int n = 0; n < N; ++n) {
    cd s = cd_make(0, 0);
#pragma unroll
    for (int k = 0; k < N; ++k) s = cd_fma(L[ref * N + k], A[k][N + n], s);
#pragma unroll
    for (int c = 0; c < N; ++c) ww[n * N + c] = cd2cf(cd_mul(cf2cd(ww[n * N + c]), s));
    if (scale_out) scale_out[(size_t)mat * N + n] = cd2cf(s);
  }
}

__global__ void k_add_logdet(double* __restrict__ logdet, const double* __restrict__ ldM, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) logdet[i] += ldM[i];
}

__global__ void k_zero_int(int* p, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0;
}

}  // namespace

int ssbk_whiten_prepare(const cf* X, cd* C64, cd* M, cd* Minv, double* ldM, cf* Z, int* wsync, int B, int N, int I, int J,
                        cudaStream_t st) {
  const int n_mat = B * I;
  SSB_DISPATCH_N(N, k_cov64<NN><<<blocks_for(n_mat, WB), WB * 32, 0, st>>>(X, C64, B, I, J));
  if (ssb_check_launch("whiten_cov64", st)) return 1;
  SSB_DISPATCH_N(N, k_whiten_factor<NN><<<blocks_for(n_mat, 64), 64, 0, st>>>(C64, M, Minv, ldM, n_mat));
  if (ssb_check_launch("whiten_factor", st)) return 1;
  dim3 grid(blocks_for(J, 256), I, B);
  SSB_DISPATCH_N(N, k_whiten_apply<NN><<<grid, 256, 0, st>>>(X, M, Z, I, J));
  if (ssb_check_launch("whiten_apply", st)) return 1;
  k_zero_int<<<blocks_for(n_mat, 256), 256, 0, st>>>(wsync, n_mat);
  return ssb_check_launch("whiten_reset", st);
}

int ssbk_w_import(const cf* W, cf* Wexp, cf* Ww, const cd* Minv, int* wsync, int n_mat, int N, cudaStream_t st) {
  SSB_DISPATCH_N(N, k_w_import<NN><<<blocks_for(n_mat, 64), 64, 0, st>>>(W, Wexp, Ww, Minv, wsync, n_mat));
  return ssb_check_launch("whiten_w_import", st);
}

int ssbk_w_export(const cf* Ww, const cd* M, cf* W, cf* Wexp, int* wsync, int n_mat, int N, cudaStream_t st) {
  SSB_DISPATCH_N(N, k_w_export<NN><<<blocks_for(n_mat, 64), 64, 0, st>>>(Ww, M, W, Wexp, wsync, n_mat));
  return ssb_check_launch("whiten_w_export", st);
}

int ssbk_pb_whitened(cf* Ww, const cd* Minv, cf* scale_out, int n_mat, int N, int ref, cudaStream_t st) {
  SSB_DISPATCH_N(N, k_pb_whitened<NN><<<blocks_for(n_mat, 64), 64, 0, st>>>(Ww, Minv, scale_out, n_mat, ref,
                                                                            ssb_status_word()));
  return ssb_check_launch("whiten_projection_back", st);
}

int ssbk_add_logdet(double* logdet, const double* ldM, int n, cudaStream_t st) {
  k_add_logdet<<<blocks_for(n, 256), 256, 0, st>>>(logdet, ldM, n);
  return ssb_check_launch("whiten_logdet", st);
}
