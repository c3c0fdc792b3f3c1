// ssspy.linalg-equivalent batched helpers, exported on their own: inverse, solve, (generalised)
// Hermitian eigendecomposition.  complex128 in / out, one thread per small matrix (N <= 8); these are
// the standalone forms of the device routines the hot kernels use inline.
//   inv / solve : Gauss-Jordan with partial pivoting (np.linalg.inv / np.linalg.solve call sites,
//                 ssspy/linalg/_solve.py:15, ssspy/linalg/inv.py:39-54)
//   eigh        : cyclic complex Jacobi; generalised problem by Cholesky reduction exactly as
//                 ssspy/linalg/eigh.py:164-207 (types 1, 2, 3), ascending eigenvalues.
#include "ssb_kernels.h"

namespace {

constexpr int MAXN = SSB_MAX_SOURCES;

__global__ void k_inv(const cd* __restrict__ A, cd* __restrict__ Ainv, int n_mat, int N, int* __restrict__ status) {
  const int mat = blockIdx.x * blockDim.x + threadIdx.x;
  if (mat >= n_mat) return;
  cd M[MAXN * 2 * MAXN];
  const int ld = 2 * N;
  for (int r = 0; r < N; ++r)
    for (int c = 0; c < N; ++c) {
      M[r * ld + c] = A[(size_t)mat * N * N + r * N + c];
      M[r * ld + N + c] = cd_make(r == c ? 1.0 : 0.0, 0.0);
    }
  bool sing;
  thread_gauss_jordan(M, N, N, ld, &sing);
  if (sing) atomicOr(status, SSB_STATUS_SINGULAR);
  for (int r = 0; r < N; ++r)
    for (int c = 0; c < N; ++c) Ainv[(size_t)mat * N * N + r * N + c] = M[r * ld + N + c];
}

__global__ void k_solve(const cd* __restrict__ A, const cd* __restrict__ Bm, cd* __restrict__ X, int n_mat, int N,
                        int R, int* __restrict__ status) {
  const int mat = blockIdx.x * blockDim.x + threadIdx.x;
  if (mat >= n_mat) return;
  cd M[MAXN * 2 * MAXN];
  const int ld = N + R;
  for (int r = 0; r < N; ++r) {
    for (int c = 0; c < N; ++c) M[r * ld + c] = A[(size_t)mat * N * N + r * N + c];
    for (int c = 0; c < R; ++c) M[r * ld + N + c] = Bm[(size_t)mat * N * R + r * R + c];
  }
  bool sing;
  thread_gauss_jordan(M, N, R, ld, &sing);
  if (sing) atomicOr(status, SSB_STATUS_SINGULAR);
  for (int r = 0; r < N; ++r)
    for (int c = 0; c < R; ++c) X[(size_t)mat * N * R + r * R + c] = M[r * ld + N + c];
}

__global__ void k_eigh(const cd* __restrict__ Ain, const cd* __restrict__ Bin, int type, double* __restrict__ lamb,
                       cd* __restrict__ Z, int n_mat, int N) {
  const int mat = blockIdx.x * blockDim.x + threadIdx.x;
  if (mat >= n_mat) return;
  cd A[MAXN * MAXN], Vv[MAXN * MAXN], L[MAXN * MAXN], Tm[MAXN * MAXN], Li[MAXN * 2 * MAXN];
  for (int e = 0; e < N * N; ++e) A[e] = Ain[(size_t)mat * N * N + e];
  // symmetrise from the lower triangle as LAPACK's 'L' does
  for (int r = 0; r < N; ++r) {
    A[r * N + r].y = 0.0;
    for (int c = r + 1; c < N; ++c) A[r * N + c] = cd_conj(A[c * N + r]);
  }
  const bool gen = Bin != nullptr;
  if (gen) {
    // Cholesky B = L L^H (lower)
    for (int e = 0; e < N * N; ++e) L[e] = cd_make(0, 0);
    const cd* Bm = Bin + (size_t)mat * N * N;
    for (int c = 0; c < N; ++c) {
      double d = Bm[c * N + c].x;
      for (int k = 0; k < c; ++k) d -= cd_abs2(L[c * N + k]);
      const double lcc = sqrt(d);
      L[c * N + c] = cd_make(lcc, 0);
      for (int r = c + 1; r < N; ++r) {
        cd s = Bm[r * N + c];
        for (int k = 0; k < c; ++k) s = cd_sub(s, cd_mulc(L[r * N + k], L[c * N + k]));
        L[r * N + c] = cd_scale(s, 1.0 / lcc);
      }
    }
    // Li = L^-1 (needed for types 1 and 2)
    const int ld = 2 * N;
    for (int r = 0; r < N; ++r)
      for (int c = 0; c < N; ++c) {
        Li[r * ld + c] = L[r * N + c];
        Li[r * ld + N + c] = cd_make(r == c ? 1.0 : 0.0, 0);
      }
    thread_gauss_jordan(Li, N, N, ld);
    if (type == 1) {
      // C = L^-1 A L^-H
      for (int r = 0; r < N; ++r)
        for (int c = 0; c < N; ++c) {
          cd s = cd_make(0, 0);
          for (int k = 0; k < N; ++k) s = cd_fma(Li[r * ld + N + k], A[k * N + c], s);
          Tm[r * N + c] = s;
        }
      for (int r = 0; r < N; ++r)
        for (int c = 0; c < N; ++c) {
          cd s = cd_make(0, 0);
          for (int k = 0; k < N; ++k) s = cd_add(s, cd_mulc(Tm[r * N + k], Li[c * ld + N + k]));
          A[r * N + c] = s;
        }
    } else {
      // C = L^H A L
      for (int r = 0; r < N; ++r)
        for (int c = 0; c < N; ++c) {
          cd s = cd_make(0, 0);
          for (int k = 0; k < N; ++k) s = cd_add(s, cd_mul(cd_conj(L[k * N + r]), A[k * N + c]));
          Tm[r * N + c] = s;
        }
      for (int r = 0; r < N; ++r)
        for (int c = 0; c < N; ++c) {
          cd s = cd_make(0, 0);
          for (int k = 0; k < N; ++k) s = cd_fma(Tm[r * N + k], L[k * N + c], s);
          A[r * N + c] = s;
        }
    }
    for (int r = 0; r < N; ++r) A[r * N + r].y = 0.0;
  }
  jacobi_herm(A, Vv, N);
  // ascending order (selection sort on the diagonal)
  int order[MAXN];
  for (int k = 0; k < N; ++k) order[k] = k;
  for (int a = 0; a < N - 1; ++a) {
    int best = a;
    for (int c = a + 1; c < N; ++c)
      if (A[order[c] * N + order[c]].x < A[order[best] * N + order[best]].x) best = c;
    int t = order[a];
    order[a] = order[best];
    order[best] = t;
  }
  for (int k = 0; k < N; ++k) lamb[(size_t)mat * N + k] = A[order[k] * N + order[k]].x;
  const int ld = 2 * N;
  for (int r = 0; r < N; ++r)
    for (int k = 0; k < N; ++k) {
      cd z;
      if (!gen) {
        z = Vv[r * N + order[k]];
      } else if (type == 3) {  // z = L y
        z = cd_make(0, 0);
        for (int q = 0; q < N; ++q) z = cd_fma(L[r * N + q], Vv[q * N + order[k]], z);
      } else {  // z = L^-H y
        z = cd_make(0, 0);
        for (int q = 0; q < N; ++q) z = cd_add(z, cd_mul(cd_conj(Li[q * ld + N + r]), Vv[q * N + order[k]]));
      }
      Z[(size_t)mat * N * N + r * N + k] = z;
    }
}

}  // namespace

extern "C" int ssb_inv(const void* A, void* Ainv, int n_mat, int N, void* stream) {
  SSB_REQUIRE(N >= 1 && N <= MAXN, "inv: N=%d unsupported (1..%d)", N, MAXN);
  if (n_mat <= 0) return 0;
  k_inv<<<blocks_for(n_mat, 64), 64, 0, (cudaStream_t)stream>>>((const cd*)A, (cd*)Ainv, n_mat, N, ssb_status_word());
  return ssb_check_launch("inv", (cudaStream_t)stream);
}

extern "C" int ssb_solve(const void* A, const void* B, void* X, int n_mat, int N, int R, void* stream) {
  SSB_REQUIRE(N >= 1 && N <= MAXN && R >= 1 && R <= MAXN, "solve: N=%d R=%d unsupported (1..%d)", N, R, MAXN);
  if (n_mat <= 0) return 0;
  k_solve<<<blocks_for(n_mat, 64), 64, 0, (cudaStream_t)stream>>>((const cd*)A, (const cd*)B, (cd*)X, n_mat, N, R,
                                                               ssb_status_word());
  return ssb_check_launch("solve", (cudaStream_t)stream);
}

extern "C" int ssb_eigh(const void* A, const void* B, int type, double* lamb, void* Z, int n_mat, int N,
                        void* stream) {
  SSB_REQUIRE(N >= 1 && N <= MAXN, "eigh: N=%d unsupported (1..%d)", N, MAXN);
  SSB_REQUIRE(type >= 1 && type <= 3, "Invalid type=%d is given.", type);
  if (n_mat <= 0) return 0;
  k_eigh<<<blocks_for(n_mat, 32), 32, 0, (cudaStream_t)stream>>>((const cd*)A, (const cd*)B, type, lamb, (cd*)Z, n_mat,
                                                               N);
  return ssb_check_launch("eigh", (cudaStream_t)stream);
}

// ILRMABase.compute_logdet (ssspy/bss/ilrma.py:524-536)
extern "C" int ssb_logdet(const void* W, double* out, int n_mat, int N, void* stream) {
  SSB_REQUIRE(N >= 1 && N <= SSB_MAX_SOURCES, "logdet: N=%d unsupported (1..%d)", N, SSB_MAX_SOURCES);
  SSB_REQUIRE(W != nullptr && out != nullptr, "logdet: NULL argument");
  if (n_mat <= 0) return 0;
  return ssbk_logdet((const cf*)W, out, n_mat, N, (cudaStream_t)stream);
}

