// STFT / inverse STFT on the device: the step either side of the demixing path.  The reference
// has no transform of its own; its notebooks call scipy.signal.stft / istft (notebooks/BSS/ILRMA/GaussILRMA-IP1-MM.ipynb:
// window="hann", nperseg=n_fft, noverlap=n_fft-hop), so these kernels follow scipy's conventions: zero boundary
// extension by nperseg / 2 on both sides, zero padding to a whole number of hops, one-sided spectrum scaled by
// 1 / sum(window); the inverse is the weighted overlap-add  x[t] = sum(window) * sum_f w[t - f h] y_f[t - f h] /
// sum_f w^2[t - f h]  with the boundary samples removed.
//
//   k_stft   CTA = one frame of one signal: windowed samples -> shared memory (complex fp64), in-place radix-2 FFT
//            (bit-reversed load, log2 n butterfly passes, twiddles by sincospi), bins 0 .. n/2 written to
//            Z[row][bin][frame]
//   k_istft  CTA = one frame: Hermitian extension of the n/2 + 1 bins, inverse FFT, window -> seg[row][frame][n]
//   k_ola    thread = one output sample: fixed-order sum over the (at most n / hop) frames that cover it: deterministic
// fp64 throughout (the transform is a few percent of one separator iteration; accuracy against scipy is 1e-13).
#include "ssb_kernels.h"

namespace {

__device__ __forceinline__ unsigned bitrev(unsigned v, int bits) { return __brev(v) >> (32 - bits); }

// in-place radix-2 decimation-in-time FFT of s[0 .. n) (already in bit-reversed order); sign = -1 forward, +1 inverse
__device__ void fft_inplace(cd* s, int n, int logn, double sign) {
  for (int st = 1; st <= logn; ++st) {
    const int half = 1 << (st - 1);
    for (int e = threadIdx.x; e < n / 2; e += blockDim.x) {
      const int grp = e >> (st - 1), k = e & (half - 1);
      const int i0 = (grp << st) + k, i1 = i0 + half;
      double sn, cs;
      sincospi(sign * (double)k / (double)half, &sn, &cs);
      const cd a = s[i0], b = s[i1];
      const cd tb = cd_make(b.x * cs - b.y * sn, b.x * sn + b.y * cs);
      s[i0] = cd_add(a, tb);
      s[i1] = cd_sub(a, tb);
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) k_stft(const double* __restrict__ x, const double* __restrict__ win,
                                              cd* __restrict__ Z, long long n_samples, int n, int logn, int hop,
                                              int n_frames, double scale) {
  extern __shared__ __align__(16) unsigned char stft_smem[];
  cd* s = reinterpret_cast<cd*>(stft_smem);
  const int f = blockIdx.x;
  const size_t row = blockIdx.y;
  const long long t0 = (long long)f * hop - n / 2;  // position of the frame in the un-extended signal
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const long long t = t0 + e;
    const double v = (t >= 0 && t < n_samples) ? x[row * n_samples + t] * win[e] : 0.0;
    s[bitrev((unsigned)e, logn)] = cd_make(v, 0.0);
  }
  __syncthreads();
  fft_inplace(s, n, logn, -1.0);
  const int n_bins = n / 2 + 1;
  for (int k = threadIdx.x; k < n_bins; k += blockDim.x)
    Z[(row * n_bins + k) * (size_t)n_frames + f] = cd_scale(s[k], scale);
}

__global__ void __launch_bounds__(256) k_istft(const cd* __restrict__ Z, const double* __restrict__ win,
                                               double* __restrict__ seg, int n, int logn, int n_frames) {
  extern __shared__ __align__(16) unsigned char stft_smem[];
  cd* s = reinterpret_cast<cd*>(stft_smem);
  const int f = blockIdx.x;
  const size_t row = blockIdx.y;
  const int n_bins = n / 2 + 1;
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    // numpy.fft.irfft: imaginary parts of the DC and Nyquist bins are ignored, the rest is the Hermitian extension
    const int k = e < n_bins ? e : n - e;
    cd v = Z[(row * n_bins + k) * (size_t)n_frames + f];
    if (e >= n_bins) v.y = -v.y;
    if (e == 0 || e == n / 2) v.y = 0.0;
    s[bitrev((unsigned)e, logn)] = v;
  }
  __syncthreads();
  fft_inplace(s, n, logn, 1.0);
  const double inv_n = 1.0 / (double)n;
  for (int e = threadIdx.x; e < n; e += blockDim.x)
    seg[(row * n_frames + f) * (size_t)n + e] = s[e].x * inv_n * win[e];
}

__global__ void k_ola(const double* __restrict__ seg, const double* __restrict__ win, double* __restrict__ y,
                      long long n_out, int n, int hop, int n_frames, double scale) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const size_t row = blockIdx.y;
  if (t >= n_out) return;
  const long long p = t + n / 2;  // position in the extended signal
  long long f_hi = p / hop;
  if (f_hi > n_frames - 1) f_hi = n_frames - 1;
  long long f_lo = (p - n + hop) / hop;  // smallest f with p - f hop < n
  if (p - n + 1 <= 0) f_lo = 0;
  if (f_lo < 0) f_lo = 0;
  double acc = 0.0, norm = 0.0;
  for (long long f = f_lo; f <= f_hi; ++f) {
    const long long e = p - f * hop;
    if (e < 0 || e >= n) continue;
    acc += seg[(row * n_frames + f) * (size_t)n + e];
    norm += win[e] * win[e];
  }
  acc *= scale;
  y[row * n_out + t] = norm > 1e-10 ? acc / norm : acc;  // scipy.signal.istft: x / where(norm > 1e-10, norm, 1)
}

int check_fft_size(int n, int* logn) {
  int l = 0;
  while ((1 << l) < n) ++l;
  SSB_REQUIRE(n >= 16 && n <= 8192 && (1 << l) == n, "stft: nperseg=%d must be a power of two in [16, 8192]", n);
  *logn = l;
  return 0;
}

template <typename KernelT>
int set_smem(KernelT kernel, size_t bytes) {
  if (bytes > 48 * 1024) SSB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  return 0;
}

}  // namespace

extern "C" int ssb_stft_frames(long long n_samples, int nperseg, int hop, int* n_frames) {
  SSB_REQUIRE(n_frames != nullptr && n_samples >= 0 && nperseg > 0 && hop > 0 && hop <= nperseg,
              "stft: invalid sizes (n_samples=%lld, nperseg=%d, hop=%d)", n_samples, nperseg, hop);
  const long long ext = n_samples + 2 * (long long)(nperseg / 2);          // boundary="zeros"
  const long long nadd = ((-(ext - nperseg)) % hop + hop) % hop % nperseg;  // padded=True
  *n_frames = (int)((ext + nadd - (nperseg - hop)) / hop);
  return 0;
}

extern "C" int ssb_stft(const double* x, const double* window, double window_sum, void* Z, int n_rows,
                        long long n_samples, int nperseg, int hop, void* stream) {
  int logn, n_frames;
  if (check_fft_size(nperseg, &logn)) return 1;
  if (ssb_stft_frames(n_samples, nperseg, hop, &n_frames)) return 1;
  if (n_rows <= 0 || n_frames <= 0) return 0;
  SSB_REQUIRE(x != nullptr && window != nullptr && Z != nullptr, "stft: NULL buffer");
  SSB_REQUIRE(window_sum != 0.0, "stft: the window sums to zero");
  const size_t sm = (size_t)nperseg * sizeof(cd);
  if (set_smem(k_stft, sm)) return 1;
  const int n_bins = nperseg / 2 + 1;
  for (int r0 = 0; r0 < n_rows; r0 += 32768) {  // grid.y is limited to 65535
    dim3 grid(n_frames, n_rows - r0 < 32768 ? n_rows - r0 : 32768);
    k_stft<<<grid, 256, sm, (cudaStream_t)stream>>>(x + (size_t)r0 * n_samples, window,
                                                    (cd*)Z + (size_t)r0 * n_bins * n_frames, n_samples, nperseg, logn,
                                                    hop, n_frames, 1.0 / fabs(window_sum));  // scaling="spectrum"
    if (ssb_check_launch("stft", (cudaStream_t)stream)) return 1;
  }
  return 0;
}

extern "C" int ssb_istft(const void* Z, const double* window, double window_sum, double* y, double* seg, int n_rows,
                         int n_frames, int nperseg, int hop, void* stream) {
  int logn;
  if (check_fft_size(nperseg, &logn)) return 1;
  SSB_REQUIRE(hop > 0 && hop <= nperseg, "istft: invalid hop=%d for nperseg=%d", hop, nperseg);
  if (n_rows <= 0 || n_frames <= 0) return 0;
  SSB_REQUIRE(Z != nullptr && window != nullptr && y != nullptr && seg != nullptr, "istft: NULL buffer");
  const size_t sm = (size_t)nperseg * sizeof(cd);
  if (set_smem(k_istft, sm)) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  const int n_bins = nperseg / 2 + 1;
  const long long n_out = (long long)nperseg + (long long)(n_frames - 1) * hop - 2 * (long long)(nperseg / 2);
  for (int r0 = 0; r0 < n_rows; r0 += 32768) {  // grid.y is limited to 65535
    const int nr = n_rows - r0 < 32768 ? n_rows - r0 : 32768;
    dim3 grid(n_frames, nr);
    double* segr = seg + (size_t)r0 * n_frames * nperseg;
    k_istft<<<grid, 256, sm, st>>>((const cd*)Z + (size_t)r0 * n_bins * n_frames, window, segr, nperseg, logn, n_frames);
    if (ssb_check_launch("istft", st)) return 1;
    if (n_out <= 0) continue;
    dim3 g2((unsigned)((n_out + 255) / 256), nr);
    k_ola<<<g2, 256, 0, st>>>(segr, window, y + (size_t)r0 * n_out, n_out, nperseg, hop, n_frames, window_sum);
    if (ssb_check_launch("istft_overlap_add", st)) return 1;
  }
  return 0;
}
