// AuxLaplaceFDICA (ssspy/bss/fdica.py:846-1245, :1527-1667) and the correlation-based permutation solver
// (ssspy/algorithm/permutation_alignment.py:12-121).
//
// FDICA is the per-bin member of the family: the weights phi[n,i,j] = G'(|y|) / floor(2 |y|) depend on one
// (source, bin, frame) only, so the iteration is  phi -> weighted covariance (array weights, the kernels of
// ssb_fused.cu / ssb_spatial.cu) -> IP1 / IP2 (lane-group solvers).  What is new here: the weight kernel, the
// loss, and the permutation alignment that follows the iterations:
//   P = |y| / floor(sqrt(sum_n |y_n|^2))           per (bin, frame)
//   bins are visited in ascending order of  corr_i = sum_j (sum_n P_n)^2  (argsort on the host, as numpy does)
//   for every bin the permutation maximising  sum_n <crit_n, P_perm(n)>  is applied and crit += P_perm
// The sequence over bins is inherently serial (greedy accumulation), so one CTA owns one mixture; per bin the N x N
// matrix C[n][n'] = <crit_n, P_n'> is reduced once and the N! permutations are scored on it (first maximum in
// itertools.permutations order = lowest lexicographic rank).
#include "ssb_kernels.h"

namespace {

constexpr int FW_ = 4;  // warps per block for the per-bin kernels

// phi[b, s, i, j] = 2 / floor(2 |w_{src[s]}^H x|)       (fdica.py:1100-1106, Laplace contrast :1630-1651)
struct FdicaSrc {
  int n_src;
  int src[SSB_MAX_SOURCES];
};

template <int N>
__global__ void __launch_bounds__(FW_ * 32) k_fdica_phi(const cf* __restrict__ X, const cf* __restrict__ W,
                                                        float* __restrict__ phi, FdicaSrc sl, int n_bins_total, int I,
                                                        int J, int flooring, float eps) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bi = blockIdx.x * FW_ + w;
  if (bi >= n_bins_total) return;
  const int b = bi / I, i = bi - b * I;
  const size_t cs = (size_t)I * J;
  const cf* xb = X + ((size_t)b * N * I + i) * J;
  for (int s = 0; s < sl.n_src; ++s) {
    cf wr[N];
#pragma unroll
    for (int m = 0; m < N; ++m) wr[m] = W[(((size_t)b * I + i) * N + sl.src[s]) * N + m];
    float* out = phi + (((size_t)b * sl.n_src + s) * I + i) * J;
    for (int j = lane; j < J; j += 32) {
      float yr = 0.f, yi = 0.f;
#pragma unroll
      for (int m = 0; m < N; ++m) {
        const cf x = xb[m * cs + j];
        yr = fmaf(wr[m].x, x.x, fmaf(-wr[m].y, x.y, yr));
        yi = fmaf(wr[m].x, x.y, fmaf(wr[m].y, x.x, yi));
      }
      out[j] = 2.0f / ssb_floor(2.0f * sqrtf(yr * yr + yi * yi), flooring, eps);
    }
  }
}

// rowloss[b, i] = mean_j sum_n 2 |y_nij|   (fdica.py:216-220), fp64 accumulation
template <int N>
__global__ void __launch_bounds__(FW_ * 32) k_fdica_rowloss(const cf* __restrict__ X, const cf* __restrict__ W,
                                                            double* __restrict__ rowloss, int n_bins_total, int I,
                                                            int J) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bi = blockIdx.x * FW_ + w;
  if (bi >= n_bins_total) return;
  const int b = bi / I, i = bi - b * I;
  const size_t cs = (size_t)I * J;
  const cf* xb = X + ((size_t)b * N * I + i) * J;
  cf wm[N][N];
#pragma unroll
  for (int n = 0; n < N; ++n)
#pragma unroll
    for (int m = 0; m < N; ++m) wm[n][m] = W[(((size_t)b * I + i) * N + n) * N + m];
  double acc = 0.0;
  for (int j = lane; j < J; j += 32) {
    cf x[N];
#pragma unroll
    for (int m = 0; m < N; ++m) x[m] = xb[m * cs + j];
    float s = 0.f;
#pragma unroll
    for (int n = 0; n < N; ++n) {
      float yr = 0.f, yi = 0.f;
#pragma unroll
      for (int m = 0; m < N; ++m) {
        yr = fmaf(wm[n][m].x, x[m].x, fmaf(-wm[n][m].y, x[m].y, yr));
        yi = fmaf(wm[n][m].x, x[m].y, fmaf(wm[n][m].y, x[m].x, yi));
      }
      s += 2.0f * sqrtf(yr * yr + yi * yi);
    }
    acc += (double)s;
  }
  acc = warp_sum(acc);
  if (lane == 0) rowloss[bi] = acc / (double)J;
}

// corr[b, i] = sum_j (sum_n P_n)^2,  P_n = |y_n| / floor(sqrt(sum_n |y_n|^2))   (permutation_alignment.py:89-93)
template <int N>
__global__ void __launch_bounds__(FW_ * 32) k_perm_corr(const cf* __restrict__ Y, double* __restrict__ corr,
                                                        int n_bins_total, int I, int J, int flooring, float eps) {
  const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int bi = blockIdx.x * FW_ + w;
  if (bi >= n_bins_total) return;
  const int b = bi / I, i = bi - b * I;
  const size_t cs = (size_t)I * J;
  const cf* yb = Y + ((size_t)b * N * I + i) * J;
  double acc = 0.0;
  for (int j = lane; j < J; j += 32) {
    double a[N], s2 = 0.0, s1 = 0.0;
#pragma unroll
    for (int n = 0; n < N; ++n) {
      const cf y = yb[n * cs + j];
      a[n] = sqrt((double)y.x * y.x + (double)y.y * y.y);
      s2 += a[n] * a[n];
      s1 += a[n];
    }
    const double nrm = ssb_floor(sqrt(s2), flooring, (double)eps);
    const double t = s1 / nrm;
    acc += t * t;
  }
  acc = warp_sum(acc);
  if (lane == 0) corr[bi] = acc;
}

// lexicographic unranking of a permutation of 0..N-1
template <int N>
__device__ __forceinline__ void perm_unrank(int rank, int (&perm)[N]) {
  int fact = 1;
#pragma unroll
  for (int q = 2; q < N; ++q) fact *= q;  // (N-1)!
  unsigned used = 0;
#pragma unroll
  for (int pos = 0; pos < N; ++pos) {
    const int idx = rank / fact;
    rank -= idx * fact;
    if (pos < N - 1) fact /= (N - 1 - pos);
    int cnt = 0, pick = 0;
#pragma unroll
    for (int v = 0; v < N; ++v) {
      if (!(used & (1u << v))) {
        if (cnt == idx) pick = v;
        ++cnt;
      }
    }
    used |= 1u << pick;
    perm[pos] = pick;
  }
}

constexpr int PT = 256;  // threads of the alignment CTA

// One CTA per mixture; dynamic shared memory: crit[N][J] f32, P[N][J] f32, ybuf[N][J] c64.
template <int N>
__global__ void __launch_bounds__(PT) k_perm_align(cf* __restrict__ Y, cf* __restrict__ W, const int* __restrict__ order,
                                                   int* __restrict__ perms, int I, int J, int flooring, float eps) {
  extern __shared__ __align__(16) unsigned char psm[];
  float* crit = reinterpret_cast<float*>(psm);
  float* P = crit + (size_t)N * J;
  cf* ybuf = reinterpret_cast<cf*>(P + (size_t)N * J);
  __shared__ double s_part[PT / 32][N * N];
  __shared__ double s_C[N * N];
  __shared__ double s_best[PT / 32];
  __shared__ int s_rank[PT / 32];
  __shared__ int s_perm[N];
  const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x;
  const size_t cs = (size_t)I * J;
  int nperm = 1;
#pragma unroll
  for (int q = 2; q <= N; ++q) nperm *= q;
  for (int tstep = 0; tstep < I; ++tstep) {
    const int i = order[(size_t)b * I + tstep];
    cf* yb = Y + ((size_t)b * N * I + i) * J;
    // P of this bin (and the raw rows, for the permuted write-back)
    for (int j = tid; j < J; j += PT) {
      float a[N], s2 = 0.f;
#pragma unroll
      for (int n = 0; n < N; ++n) {
        const cf y = yb[n * cs + j];
        ybuf[n * J + j] = y;
        a[n] = sqrtf(y.x * y.x + y.y * y.y);
        s2 += a[n] * a[n];
      }
      const float inv = 1.0f / ssb_floor(sqrtf(s2), flooring, eps);
#pragma unroll
      for (int n = 0; n < N; ++n) {
        P[n * J + j] = a[n] * inv;
        if (tstep == 0) crit[n * J + j] = a[n] * inv;  // P_criteria = P[indices[0]]
      }
    }
    if (tstep == 0) {
      if (tid < N) perms[((size_t)b * I + i) * N + tid] = tid;
      __syncthreads();
      continue;
    }
    // C[n][n'] = <crit_n, P_n'>
    double c[N * N];
#pragma unroll
    for (int e = 0; e < N * N; ++e) c[e] = 0.0;
    for (int j = tid; j < J; j += PT) {
      float cr[N], pv[N];
#pragma unroll
      for (int n = 0; n < N; ++n) {
        cr[n] = crit[n * J + j];
        pv[n] = P[n * J + j];
      }
#pragma unroll
      for (int n = 0; n < N; ++n)
#pragma unroll
        for (int m = 0; m < N; ++m) c[n * N + m] += (double)cr[n] * (double)pv[m];
    }
#pragma unroll
    for (int e = 0; e < N * N; ++e) {
      const double v = warp_sum(c[e]);
      if (lane == 0) s_part[w][e] = v;
    }
    __syncthreads();
    if (tid < N * N) {
      double v = 0.0;
#pragma unroll
      for (int ww = 0; ww < PT / 32; ++ww) v += s_part[ww][tid];  // fixed order: deterministic
      s_C[tid] = v;
    }
    __syncthreads();
    // score every permutation; first maximum in lexicographic order wins (permutation_alignment.py:101-106)
    double best = -1.0;
    int best_rank = 0x7fffffff;
    for (int r = tid; r < nperm; r += PT) {
      int pm[N];
      perm_unrank<N>(r, pm);
      double sc = 0.0;
#pragma unroll
      for (int n = 0; n < N; ++n) {
        double v = 0.0;
#pragma unroll
        for (int m = 0; m < N; ++m) v = (pm[n] == m) ? s_C[n * N + m] : v;
        sc += v;
      }
      if (sc > best) {  // ranks ascend within a thread
        best = sc;
        best_rank = r;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(SSB_FULL, best, o);
      const int orank = __shfl_xor_sync(SSB_FULL, best_rank, o);
      if (ob > best || (ob == best && orank < best_rank)) {
        best = ob;
        best_rank = orank;
      }
    }
    if (lane == 0) {
      s_best[w] = best;
      s_rank[w] = best_rank;
    }
    __syncthreads();
    if (tid == 0) {
      double bb = s_best[0];
      int br = s_rank[0];
      for (int ww = 1; ww < PT / 32; ++ww)
        if (s_best[ww] > bb || (s_best[ww] == bb && s_rank[ww] < br)) {
          bb = s_best[ww];
          br = s_rank[ww];
        }
      int pm[N];
      perm_unrank<N>(br, pm);
#pragma unroll
      for (int n = 0; n < N; ++n) {
        s_perm[n] = pm[n];
        perms[((size_t)b * I + i) * N + n] = pm[n];
      }
    }
    __syncthreads();
    int pm[N];
    bool ident = true;
#pragma unroll
    for (int n = 0; n < N; ++n) {
      pm[n] = s_perm[n];
      ident = ident && (pm[n] == n);
    }
    // crit += P[perm]; Y[i] <- Y[i, perm]; W[i] <- W[i, perm]
    for (int j = tid; j < J; j += PT) {
#pragma unroll
      for (int n = 0; n < N; ++n) {
        float pv = 0.f;
        cf yv = make_float2(0.f, 0.f);
#pragma unroll
        for (int m = 0; m < N; ++m) {
          if (pm[n] == m) {
            pv = P[m * J + j];
            yv = ybuf[m * J + j];
          }
        }
        crit[n * J + j] += pv;
        if (!ident) yb[n * cs + j] = yv;
      }
    }
    if (W != nullptr && !ident) {
      cf* wb = W + ((size_t)b * I + i) * N * N;
      cf wv = make_float2(0.f, 0.f);
      if (tid < N * N) {
        const int n = tid / N, m = tid - n * N;
        wv = wb[pm[0] * N + m];
#pragma unroll
        for (int q = 1; q < N; ++q)
          if (n == q) wv = wb[pm[q] * N + m];
      }
      __syncthreads();
      if (tid < N * N) wb[tid] = wv;
    }
    __syncthreads();
  }
}

}  // namespace

int ssbk_fdica_phi(const cf* X, const cf* W, float* phi, const int* src, int n_src, int B, int N, int I, int J,
                   int flooring, float eps, cudaStream_t st) {
  FdicaSrc sl{};
  sl.n_src = n_src;
  for (int s = 0; s < n_src; ++s) sl.src[s] = src ? src[s] : s;
  const int nb = B * I;
  SSB_DISPATCH_N(N, k_fdica_phi<NN><<<(nb + FW_ - 1) / FW_, FW_ * 32, 0, st>>>(X, W, phi, sl, nb, I, J, flooring, eps));
  return ssb_check_launch("fdica_phi", st);
}

int ssbk_fdica_rowloss(const cf* X, const cf* W, double* rowloss, int B, int N, int I, int J, cudaStream_t st) {
  const int nb = B * I;
  SSB_DISPATCH_N(N, k_fdica_rowloss<NN><<<(nb + FW_ - 1) / FW_, FW_ * 32, 0, st>>>(X, W, rowloss, nb, I, J));
  return ssb_check_launch("fdica_rowloss", st);
}

int ssbk_perm_corr(const cf* Y, double* corr, int B, int N, int I, int J, int flooring, float eps, cudaStream_t st) {
  const int nb = B * I;
  SSB_DISPATCH_N(N, k_perm_corr<NN><<<(nb + FW_ - 1) / FW_, FW_ * 32, 0, st>>>(Y, corr, nb, I, J, flooring, eps));
  return ssb_check_launch("permutation_correlation", st);
}

int ssbk_perm_align(cf* Y, cf* W, const int* order, int* perms, int B, int N, int I, int J, int flooring, float eps,
                    cudaStream_t st) {
  const size_t sm = (size_t)N * J * (2 * sizeof(float) + sizeof(cf));
  SSB_REQUIRE(sm <= 200 * 1024, "permutation solver: n_sources * n_frames = %d exceeds the shared-memory slab", N * J);
  SSB_DISPATCH_N(N, {
    static bool attr_dev[SSB_MAX_DEVICES] = {};  // function attributes are per device
  bool& attr_set = attr_dev[ssb_current_device()];
    if (!attr_set) {
      SSB_CUDA(cudaFuncSetAttribute(k_perm_align<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr_set = true;
    }
    k_perm_align<NN><<<B, PT, sm, st>>>(Y, W, order, perms, I, J, flooring, eps);
  });
  return ssb_check_launch("permutation_align", st);
}
