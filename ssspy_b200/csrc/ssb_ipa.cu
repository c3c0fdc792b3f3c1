// Iterative projection with adjustment (IPA), ssspy/bss/_update_spatial_model.py:398-513, with the
// log-quadratically penalised quadratic minimisation of ssspy/linalg/lqpqm.py:13-292 (lqpqm2,
// solve_equation, _find_largest_root) and to_psd / _psd_inv (ssspy/special/psd.py:11-71,
// _update_spatial_model.py:611-645).
//
// One CTA per (mixture, bin), like k_iss2_cta: the (source x frame) slab and its weights live in shared
// memory, every thread owns its frames.  For n = 0..N-1 (Gauss-Seidel on the current Y):
//   all threads : U_s = mean_j phi_s y y^H for every s (fp32 partials, fixed-order combine in fp64)
//   lanes s < N : to_psd(U_s) by cyclic Jacobi (fp64); lane n also forms psd_inv(U_n)
//   lane 0      : a, b, C, d, z, H, v  ->  lqpqm2 (Jacobi eigh of H, Cardano start + Newton)  ->  q, p
//   all threads : y_n <- p^H y,  y_s <- y_s + conj(q_s) y_n(old)
// All N x N algebra is fp64; the reference's batch-global early exit of the Newton loop
// (lqpqm.py:190-193, `np.all`) is not reproduced: every bin always runs max_iter updates, which moves
// a converged root by <= floor(0).
#include "ssb_kernels.h"

namespace {

constexpr int IPA_NW = 4;

__device__ __forceinline__ double floor_d(double x, int mode, double eps) { return ssb_floor(x, mode, eps); }

// Hermitian A (N x N, row-major) -> P diag(g(lambda)) P^H, symmetrised; lam_out/P_out keep floor(lambda), P
template <int N>
__device__ void psd_rebuild(cd* X, double* lam_out, cd* P_out, int flooring, double eps) {
  cd A[N * N], Vv[N * N];
  for (int r = 0; r < N; ++r)
    for (int c = 0; c < N; ++c) {
      // (X + X^H) / 2  (psd.py:49-52)
      const cd a = X[r * N + c], bt = X[c * N + r];
      A[r * N + c] = cd_make(0.5 * (a.x + bt.x), 0.5 * (a.y - bt.y));
    }
  jacobi_herm(A, Vv, N);
  double lam[N];
  for (int k = 0; k < N; ++k) lam[k] = floor_d(A[k * N + k].x, flooring, eps);
  for (int r = 0; r < N; ++r)
    for (int c = r; c < N; ++c) {
      cd acc = cd_make(0, 0);
      for (int k = 0; k < N; ++k) acc = cd_add(acc, cd_scale(cd_mulc(Vv[r * N + k], Vv[c * N + k]), lam[k]));
      if (r == c) acc.y = 0.0;
      X[r * N + c] = acc;
      X[c * N + r] = cd_conj(acc);
    }
  if (lam_out) {
    for (int k = 0; k < N; ++k) lam_out[k] = lam[k];
    for (int e = 0; e < N * N; ++e) P_out[e] = Vv[e];
  }
}

// largest real root of x^3 + A x^2 + B x + C (lqpqm.py:222-292)
__device__ double largest_cubic_root(double A, double B, double C) {
  const double P = -(A * A) / 3.0 + B;
  const double Q = (2.0 * A * A * A) / 27.0 - (A * B) / 3.0 + C;
  const double disc = (Q * 0.5) * (Q * 0.5) + (P / 3.0) * (P / 3.0) * (P / 3.0);
  // w = -Q/2 + sqrt(disc) in C
  cd w = disc >= 0.0 ? cd_make(-0.5 * Q + sqrt(disc), 0.0) : cd_make(-0.5 * Q, sqrt(-disc));
  const double amp = sqrt(cd_abs2(w));
  double x1, x2 = -INFINITY, x3 = -INFINITY;
  if (amp == 0.0) {
    x1 = cbrt(-Q);  // U == 0 => P == 0
    // V = -P / 3 with U := 1 (lqpqm.py:262-263): X2, X3 = Re(omega + V omega^*), both real parts -(1 + V)/2
    const double Vr = -P / 3.0;
    x2 = x3 = -0.5 * (1.0 + Vr);
  } else {
    const double ph = atan2(w.y, w.x) / 3.0;
    const double ca = cbrt(amp);
    const cd U = cd_make(ca * cos(ph), ca * sin(ph));
    const cd V = cd_scale(cd_inv(U), -P / 3.0);
    x1 = U.x + V.x;
    const cd om = cd_make(-0.5, 0.8660254037844386);
    x2 = cd_add(cd_mul(U, om), cd_mul(V, cd_conj(om))).x;
    x3 = cd_add(cd_mul(U, cd_conj(om)), cd_mul(V, om)).x;
  }
  const bool mono = P >= 0.0;
  const bool drop = mono || disc > 0.0;
  double root = x1;
  if (!drop) root = fmax(root, fmax(x2, x3));
  return root - A / 3.0;
}

// lqpqm2 (lqpqm.py:13-119) with singular_fn = (x < floor(0)) as update_by_ipa calls it (:484-490)
template <int M>
__device__ void lqpqm2(const cd* H, const cd* v, double z, int flooring, double eps, int max_iter, cd* y,
                       int sing_mode = 0) {
  cd A[M * M], S[M * M];
  for (int e = 0; e < M * M; ++e) A[e] = H[e];
  jacobi_herm(A, S, M);
  double phi[M];
  for (int k = 0; k < M; ++k) phi[k] = A[k * M + k].x;
  const double f0 = floor_d(0.0, flooring, eps);
  double nv = 0.0;
  for (int r = 0; r < M; ++r) nv += cd_abs2(v[r]);
  nv = sqrt(nv);
  if (sing_mode == 0 ? nv < f0 : nv == 0.0) {  // v = 0 (lqpqm.py:78-89)
    int kmax = 0;
    for (int k = 1; k < M; ++k)
      if (phi[k] > phi[kmax]) kmax = k;
    const double lam = fmax(z, phi[kmax]);
    const double scale = sqrt(fmax((lam - z) / phi[kmax], 0.0));
    for (int r = 0; r < M; ++r) y[r] = cd_scale(S[r * M + kmax], scale);
    return;
  }
  cd vt[M];
  for (int k = 0; k < M; ++k) {
    cd acc = cd_make(0, 0);
    for (int r = 0; r < M; ++r) acc = cd_add(acc, cd_mul(cd_conj(S[r * M + k]), v[r]));
    vt[k] = acc;
  }
  // solve_equation (lqpqm.py:122-219), normalization=True
  double ph[M], av2[M];
  int idx = 0;
  for (int k = 0; k < M; ++k) {
    const bool keep = phi[k] * cd_abs2(vt[k]) >= f0;
    ph[k] = keep ? phi[k] : 0.0;
    av2[k] = keep ? cd_abs2(vt[k]) : 0.0;
    if (ph[k] > ph[idx]) idx = k;
  }
  const double phi_max = floor_d(ph[idx], flooring, eps);
  const double vmax2 = av2[idx] / (phi_max * phi_max);
  const double ip = 1.0 / phi_max;
  for (int k = 0; k < M; ++k) {
    ph[k] *= ip;
    av2[k] *= ip * ip;
  }
  const double zn = z * ip;
  double lamb = largest_cubic_root(-(vmax2 + 2.0 + zn), 1.0 + 2.0 * zn, -zn);
  if (!(lamb > 1.0)) lamb = 1.0 + f0;
  lamb = fmax(lamb, zn);
  for (int it = 0; it < max_iter; ++it) {
    double s2 = 0.0, s3 = 0.0;
    for (int k = 0; k < M; ++k) {
      const double d = lamb - ph[k];
      s2 += ph[k] * av2[k] / (d * d);
      s3 += ph[k] * ph[k] * av2[k] / (d * d * d);
    }
    const double f = lamb * lamb * s2 - lamb + zn;
    const double df = -2.0 * lamb * s3 - 1.0;
    const double mu = lamb - f / df;
    lamb = mu > 1.0 ? mu : 0.5 * (1.0 + lamb);
  }
  const double lam = lamb * phi_max;
  for (int r = 0; r < M; ++r) {
    cd acc = cd_make(0, 0);
    for (int k = 0; k < M; ++k) acc = cd_add(acc, cd_mul(S[r * M + k], cd_scale(vt[k], phi[k] / (lam - phi[k]))));
    y[r] = acc;
  }
}

template <int N>
__global__ void __launch_bounds__(IPA_NW * 32) k_ipa_cta(cf* __restrict__ Y, const float* __restrict__ phi,
                                                         long long sb, long long sn, long long si, int I, int J,
                                                         int flooring, float epsf, int normalization, int max_iter,
                                                         int use_smem) {
  constexpr int M = N - 1;
  extern __shared__ __align__(16) unsigned char ipa_smem[];
  __shared__ float red[2][IPA_NW][N * N];
  __shared__ cd Us[N][N * N];   // U_s, then to_psd(U_s)
  __shared__ cd Ui[N * N];      // psd_inv(U_n)
  __shared__ cf pcoef[N];       // conj(p_s)
  __shared__ cf qcoef[N];       // Eq_s = conj(q_s) (0 for s = n)
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int bi = blockIdx.x;
  const int b = bi / I, i = bi - b * I;
  const size_t base = ((size_t)b * N * I + i) * J;
  const size_t cs = (size_t)I * J;
  const float* ph0 = phi + (size_t)b * sb + (size_t)i * si;
  const double invJ = 1.0 / (double)J;
  const double eps = (double)epsf;
  cf* ys = Y + base;
  const float* ps = ph0;
  size_t ystr = cs, pstr = (size_t)sn;
  if (use_smem) {
    cf* ys_s = reinterpret_cast<cf*>(ipa_smem);
    float* ps_s = reinterpret_cast<float*>(ys_s + (size_t)N * J);
#pragma unroll
    for (int m = 0; m < N; ++m)
      for (int j = tid; j < J; j += IPA_NW * 32) {
        ys_s[m * J + j] = Y[base + m * cs + j];
        ps_s[m * J + j] = ph0[(size_t)m * sn + j];
      }
    ys = ys_s;
    ps = ps_s;
    ystr = pstr = (size_t)J;
  }
  int rb = 0;
  for (int n = 0; n < N; ++n) {
    // ---- U_s = mean_j phi_s y y^H  (:441-445) ---------------------------------------------------------
    for (int s = 0; s < N; ++s, rb ^= 1) {
      float acc[N * N];
#pragma unroll
      for (int e = 0; e < N * N; ++e) acc[e] = 0.f;
      for (int j = tid; j < J; j += IPA_NW * 32) {
        const float ph = ps[s * pstr + j];
        cf y[N];
#pragma unroll
        for (int a = 0; a < N; ++a) y[a] = ys[a * ystr + j];
#pragma unroll
        for (int a = 0; a < N; ++a) {
          const float pr = ph * y[a].x, pi = ph * y[a].y;
          acc[a * N + a] = fmaf(pr, y[a].x, fmaf(pi, y[a].y, acc[a * N + a]));
#pragma unroll
          for (int c = a + 1; c < N; ++c) {  // y_a conj(y_c)
            acc[a * N + c] = fmaf(pr, y[c].x, fmaf(pi, y[c].y, acc[a * N + c]));
            acc[c * N + a] = fmaf(pi, y[c].x, fmaf(-pr, y[c].y, acc[c * N + a]));
          }
        }
      }
#pragma unroll
      for (int e = 0; e < N * N; ++e) {
        const float v = warp_sum(acc[e]);
        if (lane == 0) red[rb][w][e] = v;
      }
      __syncthreads();
      if (tid < N * N) {
        const int a = tid / N, c = tid - a * N;
        if (a <= c) {
          double re = 0.0, im = 0.0;
#pragma unroll
          for (int ww = 0; ww < IPA_NW; ++ww) {  // fixed order: deterministic
            re += (double)red[rb][ww][a * N + c];
            if (a != c) im += (double)red[rb][ww][c * N + a];
          }
          Us[s][a * N + c] = cd_make(re * invJ, im * invJ);
          if (a != c) Us[s][c * N + a] = cd_make(re * invJ, -im * invJ);
        }
      }
    }
    __syncthreads();
    // ---- small algebra (fp64) --------------------------------------------------------------------------
    if (w == 0) {
      if (lane < N) {
        if (lane == n) {
          double lam[N];
          cd P[N * N];
          psd_rebuild<N>(Us[lane], lam, P, flooring, eps);
          // psd_inv(to_psd(U_n)): same eigenvectors, eigenvalues floored once more (:635-645)
          for (int r = 0; r < N; ++r)
            for (int c = r; c < N; ++c) {
              cd acc = cd_make(0, 0);
              for (int k = 0; k < N; ++k)
                acc = cd_add(acc, cd_scale(cd_mulc(P[r * N + k], P[c * N + k]), 1.0 / floor_d(lam[k], flooring, eps)));
              Ui[r * N + c] = acc;
              Ui[c * N + r] = cd_conj(acc);
            }
        } else {
          psd_rebuild<N>(Us[lane], nullptr, nullptr, flooring, eps);
        }
      }
      __syncwarp();
      if (lane == 0) {
        int oth[M];
        for (int r = 0, s = 0; s < N; ++s)
          if (s != n) oth[r++] = s;
        double a[M], sa[M];
        cd bvec[M], G[M * (M + 1)], Cm[M * M], d[M];
        for (int r = 0; r < M; ++r) {
          a[r] = Us[oth[r]][n * N + n].x;       // (:453-455)
          bvec[r] = Us[oth[r]][n * N + oth[r]];  // (:456-457)
          sa[r] = sqrt(a[r]);
          for (int c = 0; c < M; ++c) {
            Cm[r * M + c] = cd_conj(Ui[oth[r] * N + oth[c]]);  // (:458-460)
            G[r * (M + 1) + c] = Cm[r * M + c];
          }
          d[r] = cd_conj(Ui[oth[r] * N + n]);
          G[r * (M + 1) + M] = d[r];
        }
        thread_gauss_jordan(G, M, 1, M + 1);  // Cd = C^-1 d (:462)
        double dCd = 0.0;
        cd Cd[M], Hn[M * M], v[M];
        for (int r = 0; r < M; ++r) {
          Cd[r] = G[r * (M + 1) + M];
          dCd += cd_mulc(Cd[r], d[r]).x;  // Re conj(d) Cd
        }
        double z = Ui[n * N + n].x - dCd;
        for (int r = 0; r < M; ++r) {
          for (int c = 0; c < M; ++c) Hn[r * M + c] = cd_scale(Cm[r * M + c], 1.0 / (sa[r] * sa[c]));
          v[r] = cd_sub(cd_scale(bvec[r], -1.0 / sa[r]), cd_scale(Cd[r], sa[r]));
        }
        if (normalization) {  // (:477-481)
          double tr = 0.0;
          for (int r = 0; r < M; ++r) tr += Hn[r * M + r].x;
          for (int e = 0; e < M * M; ++e) Hn[e] = cd_scale(Hn[e], 1.0 / tr);
          z /= tr;
        }
        cd yq[M];
        lqpqm2<M>(Hn, v, z, flooring, eps, max_iter, yq);
        cd qt[N], G2[N * (N + 1)];
        for (int s = 0; s < N; ++s) qt[s] = cd_make(s == n ? 1.0 : 0.0, 0.0);
        for (int r = 0; r < M; ++r) {
          const cd q = cd_sub(cd_scale(yq[r], 1.0 / sa[r]), cd_scale(bvec[r], 1.0 / a[r]));  // (:492)
          const cd eq = cd_conj(q);
          qcoef[oth[r]] = cd2cf(eq);
          qt[oth[r]] = cd_make(-eq.x, -eq.y);
        }
        qcoef[n] = make_float2(0.f, 0.f);
        for (int r = 0; r < N; ++r) {
          for (int c = 0; c < N; ++c) G2[r * (N + 1) + c] = Us[n][r * N + c];
          G2[r * (N + 1) + N] = qt[r];
        }
        thread_gauss_jordan(G2, N, 1, N + 1);  // U_n^-1 qt (:497)
        double qUq = 0.0;
        for (int r = 0; r < N; ++r) qUq += cd_mulc(G2[r * (N + 1) + N], qt[r]).x;
        const double den = floor_d(sqrt(fmax(qUq, 0.0)), flooring, eps);
        for (int r = 0; r < N; ++r) pcoef[r] = cd2cf(cd_conj(cd_scale(G2[r * (N + 1) + N], 1.0 / den)));
      }
    }
    __syncthreads();
    // ---- y_n <- p^H y,  y_s <- y_s + conj(q_s) y_n(old)   (:505-511) ---------------------------------
    cf pc[N], qc[N];
#pragma unroll
    for (int s = 0; s < N; ++s) {
      pc[s] = pcoef[s];
      qc[s] = qcoef[s];
    }
    for (int j = tid; j < J; j += IPA_NW * 32) {
      cf y[N];
#pragma unroll
      for (int s = 0; s < N; ++s) y[s] = ys[s * ystr + j];
      const cf yn = y[n];
      float nr = 0.f, ni = 0.f;
#pragma unroll
      for (int s = 0; s < N; ++s) {
        nr += pc[s].x * y[s].x - pc[s].y * y[s].y;
        ni += pc[s].x * y[s].y + pc[s].y * y[s].x;
      }
#pragma unroll
      for (int s = 0; s < N; ++s) {
        cf o;
        if (s == n) o = make_float2(nr, ni);
        else o = make_float2(y[s].x + qc[s].x * yn.x - qc[s].y * yn.y, y[s].y + qc[s].x * yn.y + qc[s].y * yn.x);
        ys[s * ystr + j] = o;
      }
    }
    __syncthreads();
  }
  if (use_smem) {
#pragma unroll
    for (int m = 0; m < N; ++m)
      for (int j = tid; j < J; j += IPA_NW * 32) Y[base + m * cs + j] = ys[m * J + j];
  }
}


// ---- standalone operators of ssspy.linalg used by the IPA path (complex128 / float64 on the wire) -------------
// cbrt of a complex number as the reference defines it: cbrt(|x|) exp(i arg(x) / 3)   (ssspy/linalg/cubic.py:4-22)
__device__ __forceinline__ cd cbrt_c(cd x) {
  const double a = cbrt(sqrt(cd_abs2(x))), ph = atan2(x.y, x.x) / 3.0;
  return cd_make(a * cos(ph), a * sin(ph));
}
// principal square root (numpy.sqrt on complex128)
__device__ __forceinline__ cd sqrt_c(cd x) {
  const double r = sqrt(cd_abs2(x));
  if (r == 0.0) return cd_make(0.0, 0.0);
  const double re = sqrt(0.5 * (r + fabs(x.x)));
  const double im = 0.5 * fabs(x.y) / re;
  if (x.x >= 0.0) return cd_make(re, x.y < 0.0 ? -im : im);
  return cd_make(im, x.y < 0.0 ? -re : re);
}

__global__ void k_cbrt(const cd* __restrict__ x, cd* __restrict__ y, long long n) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e < n) y[e] = cbrt_c(x[e]);
}

// roots of x^3 + A x^2 + B x + C (ssspy/linalg/polynomial.py:43-104), out[3][n]
__global__ void k_solve_cubic(const cd* __restrict__ A, const cd* __restrict__ B, const cd* __restrict__ C,
                              cd* __restrict__ out, long long n) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= n) return;
  const cd a = A[e], b = B[e], c = C[e];
  // divisions, not multiplications by reciprocals: P == 0 decides the branch below, and coefficients constructed as
  // B = A^2 / 3 must give P = 0 exactly as they do in the reference's -(A**2) / 3 + B
  const cd a2 = cd_mul(a, a), a3 = cd_mul(a2, a), ab = cd_mul(a, b);
  const cd P = cd_make(-a2.x / 3.0 + b.x, -a2.y / 3.0 + b.y);
  const cd Q = cd_make((2.0 * a3.x) / 27.0 - ab.x / 3.0 + c.x, (2.0 * a3.y) / 27.0 - ab.y / 3.0 + c.y);
  const cd hq = cd_scale(Q, 0.5), tp = cd_make(P.x / 3.0, P.y / 3.0);
  const cd disc = cd_add(cd_mul(hq, hq), cd_mul(cd_mul(tp, tp), tp));
  const bool sing = P.x == 0.0 && P.y == 0.0;
  const cd U = sing ? cd_make(1.0, 0.0) : cbrt_c(cd_add(cd_make(-hq.x, -hq.y), sqrt_c(disc)));
  const cd V = cd_mul(cd_scale(P, -1.0 / 3.0), cd_inv(U));
  const cd om = cd_make(-0.5, 0.8660254037844386), omc = cd_make(-0.5, -0.8660254037844386);
  const cd X1 = sing ? cbrt_c(cd_make(-Q.x, -Q.y)) : cd_add(U, V);
  const cd X2 = sing ? cd_mul(X1, om) : cd_add(cd_mul(U, om), cd_mul(V, omc));
  const cd X3 = sing ? cd_mul(X1, omc) : cd_add(cd_mul(U, omc), cd_mul(V, om));
  const cd sh = cd_scale(a, 1.0 / 3.0);
  out[e] = cd_sub(X1, sh);
  out[n + e] = cd_sub(X2, sh);
  out[2 * n + e] = cd_sub(X3, sh);
}

// lqpqm2 per bin (ssspy/linalg/lqpqm.py:13-119); sing_mode 0: ||v|| < floor(0) ("flooring"), 1: ||v|| == 0 (None)
template <int M>
__global__ void k_lqpqm2(const cd* __restrict__ H, const cd* __restrict__ v, const double* __restrict__ z,
                         cd* __restrict__ y, int n_bins, int flooring, double eps, int sing_mode, int max_iter) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_bins) return;
  cd h[M * M], vv[M], yy[M];
  for (int e = 0; e < M * M; ++e) h[e] = H[(size_t)i * M * M + e];
  for (int r = 0; r < M; ++r) vv[r] = v[(size_t)i * M + r];
  lqpqm2<M>(h, vv, z[i], flooring, eps, max_iter, yy, sing_mode);
  for (int r = 0; r < M; ++r) y[(size_t)i * M + r] = yy[r];
}

}  // namespace

int ssbk_ipa(cf* Y, const float* phi, long long sb, long long sn, long long si, int B, int N, int I, int J,
             int normalization, int max_iter, int flooring, float eps, cudaStream_t st) {
  SSB_REQUIRE(N >= 2 && N <= SSB_MAX_SOURCES, "n_sources=%d out of range", N);
  SSB_REQUIRE(max_iter >= 0, "newton_iter=%d must be non-negative", max_iter);
  const size_t slab = (size_t)N * J * (sizeof(cf) + sizeof(float));
  const int use_smem = slab <= 180 * 1024;
  SSB_DISPATCH_N(N, {
    static bool attr_dev[SSB_MAX_DEVICES] = {};  // function attributes are per device
  bool& attr_set = attr_dev[ssb_current_device()];
    if (!attr_set) {
      SSB_CUDA(cudaFuncSetAttribute(k_ipa_cta<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 180 * 1024));
      attr_set = true;
    }
    k_ipa_cta<NN><<<B * I, IPA_NW * 32, use_smem ? slab : 0, st>>>(Y, phi, sb, sn, si, I, J, flooring, eps,
                                                                    normalization, max_iter, use_smem);
  });
  return ssb_check_launch("update_by_ipa", st);
}

// ---- C-ABI: standalone cbrt / solve_cubic / lqpqm2 (include/ssb.h) -------------------------------------------
extern "C" int ssb_cbrt(const void* x, void* y, long long n, void* stream) {
  if (n <= 0) return 0;
  SSB_REQUIRE(x != nullptr && y != nullptr, "cbrt: NULL buffer");
  k_cbrt<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const cd*)x, (cd*)y, n);
  return ssb_check_launch("cbrt", (cudaStream_t)stream);
}

extern "C" int ssb_solve_cubic(const void* A, const void* B, const void* C, void* roots, long long n, void* stream) {
  if (n <= 0) return 0;
  SSB_REQUIRE(A != nullptr && B != nullptr && C != nullptr && roots != nullptr, "solve_cubic: NULL buffer");
  k_solve_cubic<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>((const cd*)A, (const cd*)B, (const cd*)C,
                                                                               (cd*)roots, n);
  return ssb_check_launch("solve_cubic", (cudaStream_t)stream);
}

extern "C" int ssb_lqpqm2(const void* H, const void* v, const double* z, void* y, int n_bins, int M, int flooring,
                          double eps, int singular_mode, int max_iter, void* stream) {
  if (n_bins <= 0) return 0;
  SSB_REQUIRE(M >= 1 && M <= SSB_MAX_SOURCES - 1, "lqpqm2: matrices of size %d (1..%d supported)", M, SSB_MAX_SOURCES - 1);
  SSB_REQUIRE(max_iter >= 0, "lqpqm2: max_iter=%d must be non-negative", max_iter);
  SSB_REQUIRE(H != nullptr && v != nullptr && z != nullptr && y != nullptr, "lqpqm2: NULL buffer");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((n_bins + 63) / 64);
#define SSB_LQ(MM) \
  case MM: k_lqpqm2<MM><<<grid, 64, 0, st>>>((const cd*)H, (const cd*)v, z, (cd*)y, n_bins, flooring, eps, singular_mode, max_iter); break;
  switch (M) {
    SSB_LQ(1) SSB_LQ(2) SSB_LQ(3) SSB_LQ(4) SSB_LQ(5) SSB_LQ(6) SSB_LQ(7)
    default: break;
  }
#undef SSB_LQ
  return ssb_check_launch("lqpqm2", st);
}
