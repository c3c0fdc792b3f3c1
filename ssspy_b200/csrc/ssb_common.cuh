// Shared device helpers: complex arithmetic, flooring, warp reductions and the warp-cooperative
// small dense solve used by IP1 / IP2 / projection back / loss.  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ssb.h"

#define SSB_WARP 32
#define SSB_FULL 0xffffffffu

typedef double2 cd;  // complex128 (x = re, y = im)
typedef float2 cf;   // complex64

// ---- host side error plumbing ---------------------------------------------------------------
void ssb_set_error(const char* fmt, ...);
int ssb_check_launch(const char* what, cudaStream_t st);
#define SSB_REQUIRE(cond, ...)        \
  do {                                \
    if (!(cond)) {                    \
      ssb_set_error(__VA_ARGS__);     \
      return 1;                       \
    }                                 \
  } while (0)
#define SSB_CUDA(call)                                                                  \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) {                                                            \
      ssb_set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      return 1;                                                                         \
    }                                                                                   \
  } while (0)

// Function attributes (dynamic shared memory limits) are per device: launch sites keep one "configured" flag per
// device, indexed by the calling thread's current device.
#define SSB_MAX_DEVICES 64
static inline int ssb_current_device() {
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= SSB_MAX_DEVICES) d = 0;
  return d;
}

// Device-side status word of the library (one per device, lazily allocated): bit 0 = a pivoting solver met an exactly
// zero pivot, i.e. the matrix numpy.linalg.solve / inv would reject with LinAlgError("Singular matrix")
// (ssspy/linalg/_solve.py:15, ssspy/algorithm/projection_back.py:89).  ssb_status_fetch reads and clears it.
int* ssb_status_word();

// ---- complex double ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ cd cd_make(double r, double i) { return make_double2(r, i); }
__host__ __device__ __forceinline__ cd cd_add(cd a, cd b) { return make_double2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ cd cd_sub(cd a, cd b) { return make_double2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ cd cd_mul(cd a, cd b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// a * conj(b)
__host__ __device__ __forceinline__ cd cd_mulc(cd a, cd b) {
  return make_double2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
__host__ __device__ __forceinline__ cd cd_conj(cd a) { return make_double2(a.x, -a.y); }
__host__ __device__ __forceinline__ cd cd_scale(cd a, double s) { return make_double2(a.x * s, a.y * s); }
__host__ __device__ __forceinline__ double cd_abs2(cd a) { return a.x * a.x + a.y * a.y; }
__host__ __device__ __forceinline__ cd cd_inv(cd a) {
  double d = 1.0 / (a.x * a.x + a.y * a.y);
  return make_double2(a.x * d, -a.y * d);
}
__host__ __device__ __forceinline__ cd cd_div(cd a, cd b) { return cd_mul(a, cd_inv(b)); }
__host__ __device__ __forceinline__ cd cd_fma(cd a, cd b, cd c) {  // a*b + c
  return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}
__device__ __forceinline__ cd cf2cd(cf a) { return make_double2((double)a.x, (double)a.y); }
__device__ __forceinline__ cf cd2cf(cd a) { return make_float2((float)a.x, (float)a.y); }

// ---- flooring (ssspy/special/flooring.py:6-18) ------------------------------------------------
template <typename T>
__device__ __forceinline__ T ssb_floor(T x, int mode, T eps) {
  if (mode == SSB_FLOOR_MAX) return x > eps ? x : eps;  // np.maximum(x, eps)
  if (mode == SSB_FLOOR_ADD) return x + eps;
  return x;
}

// ---- warp reductions ---------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SSB_FULL, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(SSB_FULL, v, o);
  return v;
}

// ---- warp-cooperative Gauss-Jordan with partial pivoting --------------------------------------
// A is an n x ld complex128 matrix in shared memory holding [M | RHS] with nrhs right-hand sides
// in columns n..n+nrhs-1.  On return columns n.. hold M^-1 RHS.  All 32 lanes of the warp must
// call it.  Pivoting matches LAPACK gesv's choice (largest |a| in the column, first on ties by
// cabs1 is not replicated -- results agree to rounding).  Returns log|det M| through *logabsdet
// when non-null (used by the loss).
__device__ __forceinline__ void warp_gauss_jordan(cd* A, int n, int nrhs, int ld, int lane,
                                                  double* logabsdet = nullptr, bool* singular = nullptr) {
  const int ncol = n + nrhs;
  double lad = 0.0;
  bool sing = false;
  for (int p = 0; p < n; ++p) {
    // pivot search (redundant on all lanes)
    int piv = p;
    double best = cd_abs2(A[p * ld + p]);
    for (int r = p + 1; r < n; ++r) {
      double v = cd_abs2(A[r * ld + p]);
      if (v > best) {
        best = v;
        piv = r;
      }
    }
    __syncwarp();
    if (piv != p) {
      for (int c = lane; c < ncol; c += SSB_WARP) {
        cd t = A[p * ld + c];
        A[p * ld + c] = A[piv * ld + c];
        A[piv * ld + c] = t;
      }
    }
    __syncwarp();
    cd pv = A[p * ld + p];
    sing = sing || (cd_abs2(pv) == 0.0);
    lad += 0.5 * log(cd_abs2(pv));
    cd ipv = cd_inv(pv);
    __syncwarp();
    for (int c = p + 1 + lane; c < ncol; c += SSB_WARP) A[p * ld + c] = cd_mul(A[p * ld + c], ipv);
    __syncwarp();
    // eliminate column p from every other row (entries with c > p only; column p itself is dead)
    const int w = ncol - (p + 1);
    const int total = (n - 1) * w;
    for (int e = lane; e < total; e += SSB_WARP) {
      int rr = e / w;
      int c = p + 1 + (e - rr * w);
      int r = rr < p ? rr : rr + 1;
      cd f = A[r * ld + p];
      cd a = A[r * ld + c];
      cd b = A[p * ld + c];
      A[r * ld + c] = make_double2(a.x - (f.x * b.x - f.y * b.y), a.y - (f.x * b.y + f.y * b.x));
    }
    __syncwarp();
  }
  if (logabsdet) *logabsdet = lad;
  if (singular) *singular = sing;
}

// Single-thread variant on a thread-private (local memory) matrix; used by the standalone linalg
// helpers where one thread owns one matrix.
__device__ __forceinline__ void thread_gauss_jordan(cd* A, int n, int nrhs, int ld, bool* singular = nullptr) {
  const int ncol = n + nrhs;
  bool sing = false;
  for (int p = 0; p < n; ++p) {
    int piv = p;
    double best = cd_abs2(A[p * ld + p]);
    for (int r = p + 1; r < n; ++r) {
      double v = cd_abs2(A[r * ld + p]);
      if (v > best) {
        best = v;
        piv = r;
      }
    }
    if (piv != p)
      for (int c = 0; c < ncol; ++c) {
        cd t = A[p * ld + c];
        A[p * ld + c] = A[piv * ld + c];
        A[piv * ld + c] = t;
      }
    sing = sing || (cd_abs2(A[p * ld + p]) == 0.0);
    cd ipv = cd_inv(A[p * ld + p]);
    for (int c = p + 1; c < ncol; ++c) A[p * ld + c] = cd_mul(A[p * ld + c], ipv);
    for (int r = 0; r < n; ++r) {
      if (r == p) continue;
      cd f = A[r * ld + p];
      for (int c = p + 1; c < ncol; ++c) A[r * ld + c] = cd_sub(A[r * ld + c], cd_mul(f, A[p * ld + c]));
    }
  }
  if (singular) *singular = sing;
}

// Closed-form generalised 2x2 Hermitian eigenproblem A h = l B h (type 1), following
// ssspy/linalg/eigh.py:173-201: B = L L^H, C = L^-1 A L^-H, eig(C) ascending, z = L^-H y with
// unit-norm y.  Eigenvector phase is free; we fix y[0] real >= 0 when possible.
// Outputs: lam[0] <= lam[1]; z0, z1 (each 2 entries) the corresponding columns.
__device__ __forceinline__ void herm_eig2(double c00, double c11, cd c01, double* lam, cd* y0, cd* y1) {
  double tr = 0.5 * (c00 + c11);
  double df = 0.5 * (c00 - c11);
  double off2 = cd_abs2(c01);
  double rad = sqrt(df * df + off2);
  lam[0] = tr - rad;
  lam[1] = tr + rad;
  // eigenvector for lam[1] (largest): (c01, lam1 - c00) or (lam1 - c11, conj(c01)); pick the better
  // conditioned form.  lam1 - c00 = rad - df, lam1 - c11 = rad + df.
  cd a0, a1;
  if (off2 == 0.0) {
    // diagonal C: eigenvectors are the unit vectors
    if (c00 >= c11) {
      a0 = cd_make(1, 0);
      a1 = cd_make(0, 0);
    } else {
      a0 = cd_make(0, 0);
      a1 = cd_make(1, 0);
    }
  } else if (df >= 0) {
    a0 = cd_make(rad + df, 0);
    a1 = cd_conj(c01);
  } else {
    a0 = c01;
    a1 = cd_make(rad - df, 0);
  }
  double nrm = 1.0 / sqrt(cd_abs2(a0) + cd_abs2(a1));
  a0 = cd_scale(a0, nrm);
  a1 = cd_scale(a1, nrm);
  y1[0] = a0;
  y1[1] = a1;
  // orthogonal complement for lam[0]: (-conj(a1), conj(a0))
  y0[0] = cd_make(-a1.x, a1.y);
  y0[1] = cd_conj(a0);
}

// Generalised 2x2 Hermitian eigenproblem Gm h = l Gn h (row-major 2x2 inputs), ssspy/linalg/eigh.py:173-201:
// h_small / h_large are the eigenvectors of the smaller / larger eigenvalue (z = L^-H y, Gn = L L^H).
__device__ __forceinline__ void gen_eig2(const cd* Gm, const cd* Gn, cd* h_small, cd* h_large) {
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
  h_large[0] = cd_add(cd_scale(y1[0], i00), cd_mul(cd_conj(i10), y1[1]));
  h_large[1] = cd_scale(y1[1], i11);
  h_small[0] = cd_add(cd_scale(y0[0], i00), cd_mul(cd_conj(i10), y0[1]));
  h_small[1] = cd_scale(y0[1], i11);
}

// Re(h^H G h) for a row-major 2x2 G
__device__ __forceinline__ double quad2(const cd* G, const cd* h) {
  const cd t0 = cd_add(cd_mul(G[0], h[0]), cd_mul(G[1], h[1]));
  const cd t1 = cd_add(cd_mul(G[2], h[0]), cd_mul(G[3], h[1]));
  return cd_mulc(t0, h[0]).x + cd_mulc(t1, h[1]).x;
}

// Hermitian eigendecomposition by cyclic Jacobi: A (N x N, overwritten) -> eigenvalues on the
// diagonal, Vv accumulates the eigenvectors (columns).
__device__ __forceinline__ void jacobi_herm(cd* A, cd* Vv, int N) {
  for (int r = 0; r < N; ++r)
    for (int c = 0; c < N; ++c) Vv[r * N + c] = cd_make(r == c ? 1.0 : 0.0, 0.0);
  for (int sweep = 0; sweep < 40; ++sweep) {
    double off = 0.0, dia = 0.0;
    for (int r = 0; r < N; ++r)
      for (int c = 0; c < N; ++c) {
        if (r == c) dia += A[r * N + c].x * A[r * N + c].x;
        else off += cd_abs2(A[r * N + c]);
      }
    if (off <= 1e-32 * (dia + off) || off == 0.0) break;
    for (int p = 0; p < N - 1; ++p)
      for (int q = p + 1; q < N; ++q) {
        const cd apq = A[p * N + q];
        const double mag = sqrt(cd_abs2(apq));
        if (mag == 0.0) continue;
        const cd ph = cd_scale(apq, 1.0 / mag);  // e^{i phi}
        const double theta = (A[q * N + q].x - A[p * N + p].x) / (2.0 * mag);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        // J = D R, D = diag(1, e^{-i phi}) on (p,q), R = [[c, s], [-s, c]]
        const cd Jpp = cd_make(c, 0), Jpq = cd_make(s, 0);
        const cd Jqp = cd_scale(cd_conj(ph), -s), Jqq = cd_scale(cd_conj(ph), c);
        for (int k = 0; k < N; ++k) {  // A <- A J (columns p, q)
          cd akp = A[k * N + p], akq = A[k * N + q];
          A[k * N + p] = cd_add(cd_mul(akp, Jpp), cd_mul(akq, Jqp));
          A[k * N + q] = cd_add(cd_mul(akp, Jpq), cd_mul(akq, Jqq));
          cd vkp = Vv[k * N + p], vkq = Vv[k * N + q];
          Vv[k * N + p] = cd_add(cd_mul(vkp, Jpp), cd_mul(vkq, Jqp));
          Vv[k * N + q] = cd_add(cd_mul(vkp, Jpq), cd_mul(vkq, Jqq));
        }
        for (int k = 0; k < N; ++k) {  // A <- J^H A (rows p, q)
          cd apk = A[p * N + k], aqk = A[q * N + k];
          A[p * N + k] = cd_add(cd_mul(cd_conj(Jpp), apk), cd_mul(cd_conj(Jqp), aqk));
          A[q * N + k] = cd_add(cd_mul(cd_conj(Jpq), apk), cd_mul(cd_conj(Jqq), aqk));
        }
        A[p * N + q] = cd_make(0, 0);
        A[q * N + p] = cd_make(0, 0);
        A[p * N + p].y = 0.0;
        A[q * N + q].y = 0.0;
      }
  }
}

