// Fused tensor-core path of the GaussILRMA iteration (source_algorithm="MM", domain=2, IP1/IP2).
//
// The K-contractions of the MM update (R = T V, num = A V^T, den = B V^T and their transposes,
// ssspy/bss/ilrma.py:1116-1123, :1192-1199, :1494-1498) are GEMMs with a tiny inner/outer dimension
// (K = n_basis).  On CUDA cores the iteration is FMA-bound at ~40x the HBM floor (profiles/
// r1_bench_v1_modular.json); here they run on the tensor pipe as mma.sync.m16n8k16 bf16 with an
// fp32 -> (hi, lo) bf16 split and three products (hi*hi + hi*lo + lo*hi, fp32 accumulate), which keeps
// ~16 mantissa bits (measured final-Y error 2e-6 vs 7e-4 for plain bf16).  The
// elementwise stage between the GEMMs (|y|^2, 1/R, P/R^2) stays in registers: the accumulator
// fragment of the first GEMM is re-packed as the A operand of the second (no shared-memory round
// trip), and X is read straight from global memory as 16-byte vectors.
//
//   kf_basis      : one warp = 16 bins x all frames of one (mixture, source):  T <- T sqrt(num/den)
//   kf_activation : one warp = 16 frames x all bins of one (mixture, source):  V <- V sqrt(num/den)
//                   (owns its V entries completely: no partial sums, no atomics, deterministic)
//   kf_phi_cov    : one warp = 16 bins x all frames, all sources: phi = 1/(T V) on the tensor pipe,
//                   U[b,i,n] = mean_j phi x x^H accumulated in registers
//   kf_ip1_n2     : closed-form 2x2 IP1 (fp64), one thread per bin
// followed by the modular normalisation kernels.  Operation order is exactly the reference's.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "ssb_fused.h"
#include "ssb_kernels.h"

namespace {

constexpr int FW = 8;      // warps per CTA
constexpr int JC = 256;    // frames of V staged in shared memory at a time
constexpr int PADH = 8;    // bf16 padding of shared-memory rows (bank-conflict-free fragment loads)

struct Split {
  uint32_t hi, lo;
};

// (a -> low half, b -> high half): hi = bf16(rn) pair, lo = bf16(rn) of the exact residuals: |x - hi - lo| <= 2^-18 |x| with
// a random sign.  (A truncated hi leaves every lo >= 0, so the dropped lo * lo product biases each contraction by up to
// 2^-16: measured as a 1e-5 relative error of T against the fp64 oracle, 6e-7 with this split; same instruction count.)
__device__ __forceinline__ Split split2(float a, float b) {
  Split s;
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  s.hi = *reinterpret_cast<uint32_t*>(&h);
  const float ra = a - __uint_as_float(s.hi << 16);
  const float rb = b - __uint_as_float(s.hi & 0xffff0000u);
  __nv_bfloat162 l = __floats2bfloat162_rn(ra, rb);
  s.lo = *reinterpret_cast<uint32_t*>(&l);
  return s;
}

__device__ __forceinline__ void split1(float a, __nv_bfloat16* hi, __nv_bfloat16* lo) {
  const __nv_bfloat16 h = __float2bfloat16_rn(a);
  *hi = h;
  *lo = __float2bfloat16_rn(a - __bfloat162float(h));
}

// Three-way splits (hi + mid + lo, residual 2^-27): kf_phi writes the weights the ISS kernels consume as an array, and
// the ISS updates of an ill-conditioned bin amplify their relative error by 1e3 (measured: one bin of 1025 at 2.9e-3
// with two-way operands, which alone put the whole Y at 1.2e-4 from the oracle; tools/diag_iss.py).
struct Split3 {
  uint32_t hi, mid, lo;
};
__device__ __forceinline__ Split3 split3(float a, float b) {
  Split3 s;
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  s.hi = *reinterpret_cast<uint32_t*>(&h);
  const float ra = a - __uint_as_float(s.hi << 16);
  const float rb = b - __uint_as_float(s.hi & 0xffff0000u);
  __nv_bfloat162 m = __floats2bfloat162_rn(ra, rb);
  s.mid = *reinterpret_cast<uint32_t*>(&m);
  const float qa = ra - __uint_as_float(s.mid << 16);
  const float qb = rb - __uint_as_float(s.mid & 0xffff0000u);
  __nv_bfloat162 l = __floats2bfloat162_rn(qa, qb);
  s.lo = *reinterpret_cast<uint32_t*>(&l);
  return s;
}
__device__ __forceinline__ void split1_3(float a, __nv_bfloat16* hi, __nv_bfloat16* mid, __nv_bfloat16* lo) {
  const __nv_bfloat16 h = __float2bfloat16_rn(a);
  const float r = a - __bfloat162float(h);
  const __nv_bfloat16 m = __float2bfloat16_rn(r);
  *hi = h;
  *mid = m;
  *lo = __float2bfloat16_rn(r - __bfloat162float(m));
}

// D += A(16x16, row) * B(16x8, col), bf16 inputs, fp32 accumulate
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// 3-term split product: c += (ah + al) * (bh + bl) without the al*bl term
__device__ __forceinline__ void mma_split(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                          uint32_t bh0, uint32_t bh1, uint32_t bl0, uint32_t bl1) {
  mma16816(c, ah, bh0, bh1);
  mma16816(c, ah, bl0, bl1);
  mma16816(c, al, bh0, bh1);
}

__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

__device__ __forceinline__ uint32_t lds32(const __nv_bfloat16* p) { return *reinterpret_cast<const uint32_t*>(p); }

// ---- cp.async staging of the X tile a warp consumes one step later ---------------------------------
// Each lane copies exactly the 16-byte (8-byte) vectors it will itself consume into a lane-private
// slot of a warp-private ring in shared memory and reads them back after cp.async.wait_group, so no
// warp-level barrier is needed; the global-load latency is overlapped with the MMAs of the current
// step without holding the data in registers.
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* dst, const void* src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NPEND>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(NPEND) : "memory");
}
constexpr int XSTAGES = 2;

// |sum_m w[m] x[m]|^2 for two consecutive frames held in a float4 per channel
template <int N>
__device__ __forceinline__ void power2(const float4 (&x)[N], const cf (&w)[N], float& p0, float& p1) {
  float r0 = 0.f, i0 = 0.f, r1 = 0.f, i1 = 0.f;
#pragma unroll
  for (int m = 0; m < N; ++m) {
    r0 = fmaf(w[m].x, x[m].x, fmaf(-w[m].y, x[m].y, r0));
    i0 = fmaf(w[m].x, x[m].y, fmaf(w[m].y, x[m].x, i0));
    r1 = fmaf(w[m].x, x[m].z, fmaf(-w[m].y, x[m].w, r1));
    i1 = fmaf(w[m].x, x[m].w, fmaf(w[m].y, x[m].z, i1));
  }
  p0 = fmaf(r0, r0, i0 * i0);
  p1 = fmaf(r1, r1, i1 * i1);
}


// ---- Hermitian rank-one accumulation shared by the covariance kernels -------------------------------
// U_s += phi_s x x^H for G sources of one (bin, frame): the N(N-1)/2 complex products x_a conj(x_c) and the N
// powers |x_a|^2 are formed once and shared by the G sources; each accumulator is a (re, im) [or a pair of
// diagonal entries] float2 updated with one packed FFMA2 (fma.rn.f32x2, sm_100+), i.e. N^2/2 issue slots per
// source instead of 2N^2 + 2N scalar FMA/MUL.
template <int N, int G>
struct HermAcc {
  static constexpr int NO = N * (N - 1) / 2, ND = (N + 1) / 2;
  float2 o[G][NO], d[G][ND];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int gs = 0; gs < G; ++gs) {
#pragma unroll
      for (int e = 0; e < NO; ++e) o[gs][e] = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k < ND; ++k) d[gs][k] = make_float2(0.f, 0.f);
    }
  }
  __device__ __forceinline__ void add(const float (&xr)[N], const float (&xi)[N], const float (&ph)[G]) {
    float2 pp[G];
#pragma unroll
    for (int gs = 0; gs < G; ++gs) pp[gs] = make_float2(ph[gs], ph[gs]);
    int e = 0;
#pragma unroll
    for (int a = 0; a < N; ++a)
#pragma unroll
      for (int c = a + 1; c < N; ++c, ++e) {
        // x_a conj(x_c)
        const float2 pr = make_float2(fmaf(xr[a], xr[c], xi[a] * xi[c]), fmaf(xi[a], xr[c], -(xr[a] * xi[c])));
#pragma unroll
        for (int gs = 0; gs < G; ++gs) o[gs][e] = __ffma2_rn(pp[gs], pr, o[gs][e]);
      }
#pragma unroll
    for (int k = 0; k < ND; ++k) {
      const int a = 2 * k, c = (2 * k + 1 < N) ? 2 * k + 1 : 2 * k;
      const float2 pr = make_float2(fmaf(xr[a], xr[a], xi[a] * xi[a]),
                                    (2 * k + 1 < N) ? fmaf(xr[c], xr[c], xi[c] * xi[c]) : 0.f);
#pragma unroll
      for (int gs = 0; gs < G; ++gs) d[gs][k] = __ffma2_rn(pp[gs], pr, d[gs][k]);
    }
  }
  // sum over the four lanes of a row group (xor 1, 2), scale, and write U[N x N] of source slot gs
  __device__ __forceinline__ void reduce4(float scale) {
#pragma unroll
    for (int gs = 0; gs < G; ++gs) {
#pragma unroll
      for (int e = 0; e < NO; ++e) {
        float2 v = o[gs][e];
        v.x += __shfl_xor_sync(SSB_FULL, v.x, 1);
        v.y += __shfl_xor_sync(SSB_FULL, v.y, 1);
        v.x += __shfl_xor_sync(SSB_FULL, v.x, 2);
        v.y += __shfl_xor_sync(SSB_FULL, v.y, 2);
        o[gs][e] = make_float2(v.x * scale, v.y * scale);
      }
#pragma unroll
      for (int k = 0; k < ND; ++k) {
        float2 v = d[gs][k];
        v.x += __shfl_xor_sync(SSB_FULL, v.x, 1);
        v.y += __shfl_xor_sync(SSB_FULL, v.y, 1);
        v.x += __shfl_xor_sync(SSB_FULL, v.x, 2);
        v.y += __shfl_xor_sync(SSB_FULL, v.y, 2);
        d[gs][k] = make_float2(v.x * scale, v.y * scale);
      }
    }
  }
  __device__ __forceinline__ void store(int gs, cf* u) const {
    int e = 0;
#pragma unroll
    for (int a = 0; a < N; ++a) {
      u[a * N + a] = make_float2((a & 1) ? d[gs][a >> 1].y : d[gs][a >> 1].x, 0.f);
#pragma unroll
      for (int c = a + 1; c < N; ++c, ++e) {
        u[a * N + c] = o[gs][e];
        u[c * N + a] = make_float2(o[gs][e].x, -o[gs][e].y);
      }
    }
  }
};

// ------------------------------------------------------------------------------------------------
// kf_basis.  CTA = (bin group of FW*16 bins, source n, mixture b); warp = 16 bins.
// Fragment conventions (PTX m16n8k16): g = lane/4, t = lane%4;
//   C: c0,c1 = (row g, cols 2t,2t+1), c2,c3 = (row g+8, same cols)
//   A: a0 = (row g, k 2t..), a1 = (row g+8, k 2t..), a2 = (row g, k 2t+8..), a3 = (row g+8, k 2t+8..)
//   B: b0 = (k 2t.., n g), b1 = (k 2t+8.., n g)
// Two C tiles over frames [j0, j0+8) and [j0+8, j0+16) are exactly the A operand of the next MMA.
template <int N, int KS, bool STG, int PFD>
__global__ void __launch_bounds__(FW * 32) kf_basis(const cf* __restrict__ X, const cf* __restrict__ W,
                                                    float* __restrict__ T, const float* __restrict__ V,
                                                    float* __restrict__ Pout, int I, int J, int K, int flooring,
                                                    float eps) {
  constexpr int KP = 16 * KS;
  constexpr int JKS = KP + PADH;   // row stride (halfs) of the [frame][basis] layout
  constexpr int KJS = JC + PADH;   // row stride (halfs) of the [basis][frame] layout
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __nv_bfloat16* vjk_hi = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* vjk_lo = vjk_hi + JC * JKS;
  __nv_bfloat16* vkj_hi = vjk_lo + JC * JKS;
  __nv_bfloat16* vkj_lo = vkj_hi + KP * KJS;
  constexpr int NLD = 4 * N;  // 16-byte vectors per lane per 16-frame step
  float4* xring = reinterpret_cast<float4*>(vkj_lo + KP * KJS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int n = blockIdx.y, b = blockIdx.z;
  float4* xw = xring + (size_t)warp * XSTAGES * NLD * 32 + lane;  // lane-private slots, stride 32
  const int i0 = (blockIdx.x * FW + warp) * 16;
  const bool warp_active = i0 < I;
  const int row[2] = {i0 + g, i0 + g + 8};
  const bool rvalid[2] = {row[0] < I, row[1] < I};
  const int rowc[2] = {min(row[0], I - 1), min(row[1], I - 1)};
  const size_t bn = (size_t)b * N + n;

  // T tile -> A fragments (hi, lo), and the fp32 values needed for the final update
  uint32_t Thi[KS][4], Tlo[KS][4];
  float Told[KS][2][2][2];  // [ks][nb][rr][e]: basis ks*16 + nb*8 + 2t + e, row rr
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const float* tr = T + (bn * I + rowc[rr]) * K;
#pragma unroll
      for (int nb = 0; nb < 2; ++nb) {
        const int k0 = ks * 16 + nb * 8 + 2 * t;
        const float v0 = (k0 < K && rvalid[rr]) ? tr[k0] : 0.f;
        const float v1 = (k0 + 1 < K && rvalid[rr]) ? tr[k0 + 1] : 0.f;
        Told[ks][nb][rr][0] = v0;
        Told[ks][nb][rr][1] = v1;
        const Split s = split2(v0, v1);
        Thi[ks][nb * 2 + rr] = s.hi;  // a0: (g, k lo) a1: (g+8, k lo) a2: (g, k hi) a3: (g+8, k hi)
        Tlo[ks][nb * 2 + rr] = s.lo;
      }
    }
  }
  cf w[2][N];
#pragma unroll
  for (int rr = 0; rr < 2; ++rr)
#pragma unroll
    for (int m = 0; m < N; ++m)
      w[rr][m] = W ? W[(((size_t)b * I + rowc[rr]) * N + n) * N + m] : make_float2(m == n ? 1.f : 0.f, 0.f);

  float num[2 * KS][4], den[2 * KS][4];
#pragma unroll
  for (int q = 0; q < 2 * KS; ++q)
#pragma unroll
    for (int c = 0; c < 4; ++c) num[q][c] = den[q][c] = 0.f;

  const float* Vb = V + bn * K * J;
  const size_t xrow[2] = {((size_t)b * N * I + rowc[0]) * J, ((size_t)b * N * I + rowc[1]) * J};
  const size_t cs = (size_t)I * J;

  // prefetch of the X tile of frames [f0, f0+16) into ring slot `stage`
  auto issue = [&](int f0, int stage) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int rr = 0; rr < 2; ++rr)
#pragma unroll
        for (int m = 0; m < N; ++m)
          cp_async16(xw + (stage * NLD + (h * 2 + rr) * N + m) * 32, X + xrow[rr] + f0 + 8 * h + 2 * t + m * cs);
  };
  // software prefetch (discarded 4-byte L2-level loads, one per 128-byte line) of the X tile of frames
  // [f0, f0+16): 16 rows x N channels; the cp.async of that step then hits L2
  auto prefetch = [&](int f0) {
#pragma unroll
    for (int q = lane; q < 16 * N; q += 32) {
      const int i = min(i0 + (q & 15), I - 1);
      const cf* p = X + (((size_t)b * N + (q >> 4)) * I + i) * J + f0;
      unsigned dummy;
      asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(dummy) : "l"(p));
    }
  };
  int step = 0;
  if (STG && warp_active) {
    issue(0, 0);
    cp_async_commit();
  }
  if (PFD > 0 && warp_active) {
#pragma unroll
    for (int d = 0; d < PFD; ++d)
      if (16 * d < J) prefetch(16 * d);
  }

  for (int jc0 = 0; jc0 < J; jc0 += JC) {
    __syncthreads();
    // stage V[:, jc0 : jc0+JC] as bf16 hi/lo in both layouts
    {
      // all loads first (independent, one L2 round trip), then split + scatter into both layouts
      constexpr int NIT = KP * JC / (FW * 32);
      float vals[NIT];
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const int e = threadIdx.x + it * FW * 32;
        const int k = e / JC, jj = e - k * JC;
        vals[it] = (k < K && jc0 + jj < J) ? __ldg(Vb + (size_t)k * J + jc0 + jj) : 0.f;
      }
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const int e = threadIdx.x + it * FW * 32;
        const int k = e / JC, jj = e - k * JC;
        __nv_bfloat16 h, l;
        split1(vals[it], &h, &l);
        vkj_hi[k * KJS + jj] = h;
        vkj_lo[k * KJS + jj] = l;
        vjk_hi[jj * JKS + k] = h;
        vjk_lo[jj * JKS + k] = l;
      }
    }
    __syncthreads();
    if (!warp_active) continue;
    const int jend = min(JC, J - jc0);
    for (int jj = 0; jj < jend; jj += 16, ++step) {
      if (STG) {
        if (jc0 + jj + 16 < J) issue(jc0 + jj + 16, (step + 1) & 1);
        cp_async_commit();
      }
      if (PFD > 0 && jc0 + jj + 16 * PFD < J) prefetch(jc0 + jj + 16 * PFD);
      // ---- GEMM1: R[16 bins x 16 frames] = T V ------------------------------------------------
      float R[2][4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int c = 0; c < 4; ++c) R[h][c] = 0.f;
        const int fr = jj + 8 * h + g;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const uint32_t bh0 = lds32(vjk_hi + fr * JKS + ks * 16 + 2 * t);
          const uint32_t bh1 = lds32(vjk_hi + fr * JKS + ks * 16 + 2 * t + 8);
          const uint32_t bl0 = lds32(vjk_lo + fr * JKS + ks * 16 + 2 * t);
          const uint32_t bl1 = lds32(vjk_lo + fr * JKS + ks * 16 + 2 * t + 8);
          mma_split(R[h], Thi[ks], Tlo[ks], bh0, bh1, bl0, bl1);
        }
      }
      // ---- elementwise: P = |w^H x|^2, A = P / R^2, B = 1 / R ----------------------------------
      if (STG) cp_async_wait<1>();  // this step's tile has landed (only the next one may be in flight)
      uint32_t Ahi[4], Alo[4], Bhi[4], Blo[4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          float4 x[N];
          const size_t off = xrow[rr] + jc0 + jj + 8 * h + 2 * t;
#pragma unroll
          for (int m = 0; m < N; ++m)
            x[m] = STG ? xw[((step & 1) * NLD + (h * 2 + rr) * N + m) * 32]
                       : *reinterpret_cast<const float4*>(X + off + m * cs);
          float p0, p1;
          power2<N>(x, w[rr], p0, p1);
          // the power spectrogram is kept for the activation update (same W => same P, ilrma.py:1169-1172)
          if (rvalid[rr])
            *reinterpret_cast<float2*>(Pout + (bn * I + row[rr]) * J + jc0 + jj + 8 * h + 2 * t) = make_float2(p0, p1);
          const float i0v = rvalid[rr] ? fast_rcp(R[h][rr * 2 + 0]) : 0.f;
          const float i1v = rvalid[rr] ? fast_rcp(R[h][rr * 2 + 1]) : 0.f;
          const Split sa = split2(p0 * i0v * i0v, p1 * i1v * i1v);
          const Split sb = split2(i0v, i1v);
          Ahi[h * 2 + rr] = sa.hi;
          Alo[h * 2 + rr] = sa.lo;
          Bhi[h * 2 + rr] = sb.hi;
          Blo[h * 2 + rr] = sb.lo;
        }
      }
      // ---- GEMM2: num += A V^T, den += B V^T  (contraction over the 16 frames) -------------------
#pragma unroll
      for (int q = 0; q < 2 * KS; ++q) {
        const __nv_bfloat16* ph = vkj_hi + (q * 8 + g) * KJS + jj + 2 * t;
        const __nv_bfloat16* pl = vkj_lo + (q * 8 + g) * KJS + jj + 2 * t;
        const uint32_t vh0 = lds32(ph), vh1 = lds32(ph + 8), vl0 = lds32(pl), vl1 = lds32(pl + 8);
        mma_split(num[q], Ahi, Alo, vh0, vh1, vl0, vl1);
        mma_split(den[q], Bhi, Blo, vh0, vh1, vl0, vl1);
      }
    }
  }
  if (!warp_active) return;
  // ---- T <- floor(T * sqrt(num / den))      (ilrma.py:1125-1126, p = 2) ----------------------------
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
#pragma unroll
    for (int nb = 0; nb < 2; ++nb)
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        if (!rvalid[rr]) continue;
        const int q = ks * 2 + nb;
        const int k0 = ks * 16 + nb * 8 + 2 * t;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          if (k0 + e < K) {
            const float ratio = num[q][rr * 2 + e] / den[q][rr * 2 + e];
            T[(bn * I + row[rr]) * K + k0 + e] = ssb_floor(sqrtf(ratio) * Told[ks][nb][rr][e], flooring, eps);
          }
        }
      }
}

// ------------------------------------------------------------------------------------------------
// kf_phi: phi[b,n,i,j] = 1 / (T V) written out (ISS modes consume the weights as an array,
// ssspy/bss/ilrma.py:1690-1696).  Same tiling as kf_basis, R on the tensor pipe.
template <int KS, bool INV>
__global__ void __launch_bounds__(FW * 32) kf_phi(const float* __restrict__ T, const float* __restrict__ V,
                                                  float* __restrict__ phi, int I, int J, int K) {
  constexpr int KP = 16 * KS;
  constexpr int JKS = KP + PADH;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __nv_bfloat16* vjk_hi = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* vjk_mid = vjk_hi + JC * JKS;
  __nv_bfloat16* vjk_lo = vjk_mid + JC * JKS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const size_t bn = blockIdx.y;
  const int i0 = (blockIdx.x * FW + warp) * 16;
  const bool warp_active = i0 < I;
  const int row[2] = {i0 + g, i0 + g + 8};
  const bool rvalid[2] = {row[0] < I, row[1] < I};
  const int rowc[2] = {min(row[0], I - 1), min(row[1], I - 1)};
  uint32_t Thi[KS][4], Tmid[KS][4], Tlo[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const float* tr = T + (bn * I + rowc[rr]) * K;
#pragma unroll
      for (int nb = 0; nb < 2; ++nb) {
        const int k0 = ks * 16 + nb * 8 + 2 * t;
        const Split3 s = split3((k0 < K) ? tr[k0] : 0.f, (k0 + 1 < K) ? tr[k0 + 1] : 0.f);
        Thi[ks][nb * 2 + rr] = s.hi;
        Tmid[ks][nb * 2 + rr] = s.mid;
        Tlo[ks][nb * 2 + rr] = s.lo;
      }
    }
  const float* Vb = V + bn * K * J;
  for (int jc0 = 0; jc0 < J; jc0 += JC) {
    __syncthreads();
    {
      constexpr int NIT = KP * JC / (FW * 32);
      float vals[NIT];
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const int e = threadIdx.x + it * FW * 32;
        const int k = e / JC, jj = e - k * JC;
        vals[it] = (k < K && jc0 + jj < J) ? __ldg(Vb + (size_t)k * J + jc0 + jj) : 0.f;
      }
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const int e = threadIdx.x + it * FW * 32;
        const int k = e / JC, jj = e - k * JC;
        __nv_bfloat16 h, m, l;
        split1_3(vals[it], &h, &m, &l);
        vjk_hi[jj * JKS + k] = h;
        vjk_mid[jj * JKS + k] = m;
        vjk_lo[jj * JKS + k] = l;
      }
    }
    __syncthreads();
    if (!warp_active) continue;
    const int jend = min(JC, J - jc0);
    for (int jj = 0; jj < jend; jj += 8) {
      // small terms first; every product down to 2^-18 of the result is kept (hi lo, lo hi, mid mid)
      float R[4] = {0.f, 0.f, 0.f, 0.f};
      const int fr = jj + g;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const __nv_bfloat16* ph = vjk_hi + fr * JKS + ks * 16 + 2 * t;
        const __nv_bfloat16* pm = vjk_mid + fr * JKS + ks * 16 + 2 * t;
        const __nv_bfloat16* pl = vjk_lo + fr * JKS + ks * 16 + 2 * t;
        const uint32_t h0 = lds32(ph), h1 = lds32(ph + 8), m0 = lds32(pm), m1 = lds32(pm + 8);
        mma16816(R, Tlo[ks], h0, h1);
        mma16816(R, Thi[ks], lds32(pl), lds32(pl + 8));
        mma16816(R, Tmid[ks], m0, m1);
      }
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const __nv_bfloat16* ph = vjk_hi + fr * JKS + ks * 16 + 2 * t;
        const __nv_bfloat16* pm = vjk_mid + fr * JKS + ks * 16 + 2 * t;
        const uint32_t h0 = lds32(ph), h1 = lds32(ph + 8);
        mma16816(R, Tmid[ks], h0, h1);
        mma16816(R, Thi[ks], lds32(pm), lds32(pm + 8));
        mma16816(R, Thi[ks], h0, h1);
      }
#pragma unroll
      for (int rr = 0; rr < 2; ++rr)
        if (rvalid[rr])
          *reinterpret_cast<float2*>(phi + (bn * I + row[rr]) * J + jc0 + jj + 2 * t) =
              INV ? make_float2(1.0f / R[rr * 2], 1.0f / R[rr * 2 + 1]) : make_float2(R[rr * 2], R[rr * 2 + 1]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// kf_activation.  CTA = (frame group of FW*16 frames, source n, mixture b); warp = 16 frames; all
// warps walk over the bins together, FW*16 bins of T staged in shared memory per round.  P = |y|^2 was
// written by kf_basis (the demixing filters do not change between the two updates).
// Orientation is transposed w.r.t. kf_basis: C rows = frames, C cols = bins, so that the
// accumulator fragment of R^T = V^T T^T is the A operand of num^T += (P/R^2)^T T.
constexpr int BCH = FW * 16;  // bins staged per round

__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(src) : "memory");
}

template <int KS>
__global__ void __launch_bounds__(FW * 32) kf_activation(const float* __restrict__ P, const float* __restrict__ T,
                                                         float* __restrict__ V, int NS, int I, int J, int K,
                                                         int flooring, float eps) {
  constexpr int KP = 16 * KS;
  constexpr int BKS = KP + PADH;    // [bin][basis] row stride (halfs)
  constexpr int KBS = BCH + PADH;   // [basis][bin] row stride (halfs)
  constexpr int NLD = 8;            // 4-byte P values per lane per 16-bin step
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __nv_bfloat16* tbk_hi = reinterpret_cast<__nv_bfloat16*>(smem_raw);
  __nv_bfloat16* tbk_lo = tbk_hi + BCH * BKS;
  __nv_bfloat16* tkb_hi = tbk_lo + BCH * BKS;
  __nv_bfloat16* tkb_lo = tkb_hi + KP * KBS;
  float* pring = reinterpret_cast<float*>(tkb_lo + KP * KBS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int n = blockIdx.y, b = blockIdx.z;
  float* pw = pring + (size_t)warp * XSTAGES * NLD * 32 + lane;
  const int j0 = (blockIdx.x * FW + warp) * 16;
  const bool warp_active = j0 < J;
  const size_t bn = (size_t)b * NS + n;
  float* Vb = V + bn * K * J;
  const float* Pb = P + bn * I * J;
  const int fr[2] = {min(j0 + g, J - 1), min(j0 + g + 8, J - 1)};
  const bool fvalid[2] = {j0 + g < J, j0 + g + 8 < J};

  // V^T tile (16 frames x KP) -> A fragments; fp32 copies for the final update
  uint32_t Vhi[KS][4], Vlo[KS][4];
  float Vold[KS][2][2][2];  // [ks][nb][rr][e]: basis ks*16+nb*8+2t+e, frame rr
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
#pragma unroll
    for (int nb = 0; nb < 2; ++nb)
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int k0 = ks * 16 + nb * 8 + 2 * t;
        const float v0 = (k0 < K && fvalid[rr]) ? Vb[(size_t)k0 * J + fr[rr]] : 0.f;
        const float v1 = (k0 + 1 < K && fvalid[rr]) ? Vb[(size_t)(k0 + 1) * J + fr[rr]] : 0.f;
        Vold[ks][nb][rr][0] = v0;
        Vold[ks][nb][rr][1] = v1;
        const Split s = split2(v0, v1);
        Vhi[ks][nb * 2 + rr] = s.hi;
        Vlo[ks][nb * 2 + rr] = s.lo;
      }

  float num[2 * KS][4], den[2 * KS][4];
#pragma unroll
  for (int q = 0; q < 2 * KS; ++q)
#pragma unroll
    for (int c = 0; c < 4; ++c) num[q][c] = den[q][c] = 0.f;

  // prefetch of P for bins [ibase, ibase+16) x this warp's 16 frames into ring slot `stage`.  The P
  // scratch is padded by 16 rows, so the (masked) rows past the last bin need no clamping; all
  // per-lane offsets are loop invariant.
  int poff[8];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int rr = 0; rr < 2; ++rr)
#pragma unroll
      for (int e = 0; e < 2; ++e) poff[(h * 2 + rr) * 2 + e] = (8 * h + 2 * t + e) * J + fr[rr];
  const uint32_t pw_s = (uint32_t)__cvta_generic_to_shared(pw);
  auto issue = [&](int ibase, int stage) {
    const float* p = Pb + (size_t)ibase * J;
    const uint32_t d = pw_s + stage * (NLD * 32 * 4);
#pragma unroll
    for (int k = 0; k < 8; ++k)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + k * 128), "l"(p + poff[k]) : "memory");
  };
  int step = 0;
  if (warp_active) {
    issue(0, 0);
    cp_async_commit();
  }
  const bool frames_full = fvalid[0] && fvalid[1];

  for (int ib0 = 0; ib0 < I; ib0 += BCH) {
    __syncthreads();
    // stage T[ib0 : ib0+BCH, :] (bf16 hi/lo, both layouts)
    {
      constexpr int NIT = BCH * KP / (FW * 32);
      float vals[NIT];
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const int e = threadIdx.x + it * FW * 32;
        const int bi = e / KP, k = e - bi * KP;
        const int i = ib0 + bi;
        vals[it] = (i < I && k < K) ? __ldg(T + (bn * I + i) * K + k) : 0.f;
      }
#pragma unroll
      for (int it = 0; it < NIT; ++it) {
        const int e = threadIdx.x + it * FW * 32;
        const int bi = e / KP, k = e - bi * KP;
        __nv_bfloat16 h, l;
        split1(vals[it], &h, &l);
        tbk_hi[bi * BKS + k] = h;
        tbk_lo[bi * BKS + k] = l;
        tkb_hi[k * KBS + bi] = h;
        tkb_lo[k * KBS + bi] = l;
      }
    }
    __syncthreads();
    if (!warp_active) continue;
    const int nbt = min(BCH, I - ib0);
    for (int bb = 0; bb < nbt; bb += 16, ++step) {
      if (ib0 + bb + 16 < I) issue(ib0 + bb + 16, (step + 1) & 1);
      cp_async_commit();
      // ---- GEMM1: R^T[16 frames x 16 bins] = V^T T^T ---------------------------------------------
      float R[2][4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {  // h: bins bb+8h .. bb+8h+7 (n-tile)
#pragma unroll
        for (int c = 0; c < 4; ++c) R[h][c] = 0.f;
        const int bi = bb + 8 * h + g;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          const uint32_t bh0 = lds32(tbk_hi + bi * BKS + ks * 16 + 2 * t);
          const uint32_t bh1 = lds32(tbk_hi + bi * BKS + ks * 16 + 2 * t + 8);
          const uint32_t bl0 = lds32(tbk_lo + bi * BKS + ks * 16 + 2 * t);
          const uint32_t bl1 = lds32(tbk_lo + bi * BKS + ks * 16 + 2 * t + 8);
          mma_split(R[h], Vhi[ks], Vlo[ks], bh0, bh1, bl0, bl1);
        }
      }
      // ---- elementwise at (frame rr, bin bb+8h+2t+e) -----------------------------------------------
      cp_async_wait<1>();
      uint32_t Ahi[4], Alo[4], Bhi[4], Blo[4];
      const float* pst = pw + (step & 1) * (NLD * 32);
      const int lim = I - (ib0 + bb);  // valid bins in this 16-bin step
      if (lim >= 16 && frames_full) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            const float iv0 = fast_rcp(R[h][rr * 2]), iv1 = fast_rcp(R[h][rr * 2 + 1]);
            const float p0 = pst[((h * 2 + rr) * 2) * 32], p1 = pst[((h * 2 + rr) * 2 + 1) * 32];
            const Split sa = split2(p0 * iv0 * iv0, p1 * iv1 * iv1);
            const Split sb = split2(iv0, iv1);
            Ahi[h * 2 + rr] = sa.hi;
            Alo[h * 2 + rr] = sa.lo;
            Bhi[h * 2 + rr] = sb.hi;
            Blo[h * 2 + rr] = sb.lo;
          }
      } else {
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            float a_[2], i_[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const bool ok = (8 * h + 2 * t + e < lim) && fvalid[rr];
              const float p = pst[((h * 2 + rr) * 2 + e) * 32];
              const float iv = ok ? fast_rcp(R[h][rr * 2 + e]) : 0.f;
              i_[e] = iv;
              a_[e] = ok ? p * iv * iv : 0.f;
            }
            const Split sa = split2(a_[0], a_[1]);
            const Split sb = split2(i_[0], i_[1]);
            Ahi[h * 2 + rr] = sa.hi;
            Alo[h * 2 + rr] = sa.lo;
            Bhi[h * 2 + rr] = sb.hi;
            Blo[h * 2 + rr] = sb.lo;
          }
      }
      // ---- GEMM2: num^T += A^T T, den^T += B^T T  (contraction over the 16 bins) -----------------
#pragma unroll
      for (int q = 0; q < 2 * KS; ++q) {
        const __nv_bfloat16* ph = tkb_hi + (q * 8 + g) * KBS + bb + 2 * t;
        const __nv_bfloat16* pl = tkb_lo + (q * 8 + g) * KBS + bb + 2 * t;
        const uint32_t th0 = lds32(ph), th1 = lds32(ph + 8), tl0 = lds32(pl), tl1 = lds32(pl + 8);
        mma_split(num[q], Ahi, Alo, th0, th1, tl0, tl1);
        mma_split(den[q], Bhi, Blo, th0, th1, tl0, tl1);
      }
    }
  }
  if (!warp_active) return;
  // ---- V <- floor(V * sqrt(num / den))      (ilrma.py:1201-1202, p = 2) ----------------------------
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
#pragma unroll
    for (int nb = 0; nb < 2; ++nb)
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        if (!fvalid[rr]) continue;
        const int q = ks * 2 + nb;
        const int k0 = ks * 16 + nb * 8 + 2 * t;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          if (k0 + e < K) {
            const float ratio = num[q][rr * 2 + e] / den[q][rr * 2 + e];
            Vb[(size_t)(k0 + e) * J + fr[rr]] = ssb_floor(sqrtf(ratio) * Vold[ks][nb][rr][e], flooring, eps);
          }
        }
      }
}

// ------------------------------------------------------------------------------------------------
// kf_phi_cov.  CTA = (bin group of FW*16 bins, mixture b); warp = 16 bins x all frames.  G sources
// are handled per pass over the bins' X slab (their V staged side by side in shared memory), so for
// N <= 4 the slab is read at most twice; RS splits the two row groups of the tile into separate
// passes when 2 N^2 accumulators per source would not fit the register file (N >= 6).
//   phi = 1 / (T V)                                   (ilrma.py:1494-1498, p = 2)
//   U[b,i,n,a,c] = (1/J) sum_j phi[n,i,j] x_a conj(x_c) (ilrma.py:1500-1505)
// Each thread accumulates its rows' Hermitian N x N (N^2 reals per row) over its frames; the four
// lanes of a row group are reduced with shuffles at the end.
template <int N>
struct CovShape {
  static constexpr int G = N <= 4 ? N : 2;   // sources per pass (they share the Hermitian products of a frame)
  static constexpr bool RS = N >= 4;   // one row group (8 bins) of the warp tile per pass over the frames
  static constexpr bool STG = true;    // cp.async ring for X in kf_phi_cov (direct loads leave the HBM latency exposed)
  static constexpr bool STG_W = true;  // ... in kf_cov_w
  // frames of V staged at a time in kf_phi_cov: G sources side by side must fit next to the X ring
  static constexpr int jcc(int KP) { return G * KP >= 128 ? 64 : (G * KP >= 64 ? 128 : 256); }
};

template <int N, int KS>
__global__ void __launch_bounds__(FW * 32) kf_phi_cov(const cf* __restrict__ X, const float* __restrict__ T,
                                                      const float* __restrict__ V, cf* __restrict__ U, int I, int J,
                                                      int K) {
  constexpr int KP = 16 * KS;
  constexpr int JKS = KP + PADH;
  constexpr int G = CovShape<N>::G;
  constexpr int JCC = CovShape<N>::jcc(KP);
  constexpr bool RS = CovShape<N>::RS;
  constexpr bool STG = CovShape<N>::STG;
  constexpr int NR = RS ? 1 : 2;   // row groups accumulated per pass
  constexpr int NLD = 2 * NR * N;  // 16-byte vectors per lane per step (rows of this pass only)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __nv_bfloat16* vjk_hi = reinterpret_cast<__nv_bfloat16*>(smem_raw);  // [G][JCC][JKS]
  __nv_bfloat16* vjk_lo = vjk_hi + G * JCC * JKS;
  float4* xring = reinterpret_cast<float4*>(vjk_lo + G * JCC * JKS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z;
  float4* xw = xring + (size_t)warp * XSTAGES * NLD * 32 + lane;
  const int i0 = (blockIdx.y * FW + warp) * 16;
  const bool warp_active = i0 < I;
  const int row[2] = {i0 + g, i0 + g + 8};
  const bool rvalid[2] = {row[0] < I, row[1] < I};
  const int rowc[2] = {min(row[0], I - 1), min(row[1], I - 1)};
  const size_t xrow[2] = {((size_t)b * N * I + rowc[0]) * J, ((size_t)b * N * I + rowc[1]) * J};
  const size_t cs = (size_t)I * J;
  const float invJ = 1.0f / (float)J;

  auto issue = [&](int f0, int stage, int rs) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const int rr = RS ? rs : r;
#pragma unroll
        for (int m = 0; m < N; ++m)
          cp_async16(xw + (stage * NLD + (h * NR + r) * N + m) * 32, X + xrow[rr] + f0 + 8 * h + 2 * t + m * cs);
      }
  };

  // source groups are spread over blockIdx.x (adjacent CTAs share the X tile through L2)
  {
    const int n0 = blockIdx.x * G;
    for (int rs = 0; rs < (RS ? 2 : 1); ++rs) {
      uint32_t Thi[G][KS][4], Tlo[G][KS][4];
#pragma unroll
      for (int gs = 0; gs < G; ++gs) {
        const int n = min(n0 + gs, N - 1);
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            const float* tr = T + (((size_t)b * N + n) * I + rowc[rr]) * K;
#pragma unroll
            for (int nb = 0; nb < 2; ++nb) {
              const int k0 = ks * 16 + nb * 8 + 2 * t;
              const float v0 = (k0 < K) ? tr[k0] : 0.f;
              const float v1 = (k0 + 1 < K) ? tr[k0 + 1] : 0.f;
              const Split s = split2(v0, v1);
              Thi[gs][ks][nb * 2 + rr] = s.hi;
              Tlo[gs][ks][nb * 2 + rr] = s.lo;
            }
          }
      }
      HermAcc<N, G> acc[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r) acc[r].zero();

      int step = 0;
      if (STG && warp_active) {
        issue(0, 0, rs);
        cp_async_commit();
      }
      for (int jc0 = 0; jc0 < J; jc0 += JCC) {
        __syncthreads();
#pragma unroll
        for (int gs = 0; gs < G; ++gs) {
          constexpr int NIT = KP * JCC / (FW * 32);
          const int n = n0 + gs;
          float vals[NIT];
#pragma unroll
          for (int it = 0; it < NIT; ++it) {
            const int r = threadIdx.x + it * FW * 32;
            const int k = r / JCC, jj = r - k * JCC;
            vals[it] = (n < N && k < K && jc0 + jj < J) ? __ldg(V + (((size_t)b * N + n) * K + k) * J + jc0 + jj) : 0.f;
          }
#pragma unroll
          for (int it = 0; it < NIT; ++it) {
            const int r = threadIdx.x + it * FW * 32;
            const int k = r / JCC, jj = r - k * JCC;
            __nv_bfloat16 h, l;
            split1(vals[it], &h, &l);
            vjk_hi[(gs * JCC + jj) * JKS + k] = h;
            vjk_lo[(gs * JCC + jj) * JKS + k] = l;
          }
        }
        __syncthreads();
        if (!warp_active) continue;
        const int jend = min(JCC, J - jc0);
        for (int jj = 0; jj < jend; jj += 16, ++step) {
          if (STG) {
            if (jc0 + jj + 16 < J) issue(jc0 + jj + 16, (step + 1) & 1, rs);
            cp_async_commit();
            cp_async_wait<1>();
          }
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            // X for this half step: rows (g, g+8) x frames (2t, 2t+1) of frames [jj+8h, jj+8h+8)
            float4 x[NR][N];
#pragma unroll
            for (int r = 0; r < NR; ++r) {
              const int rr = RS ? rs : r;
#pragma unroll
              for (int m = 0; m < N; ++m)
                x[r][m] = STG ? xw[((step & 1) * NLD + (h * NR + r) * N + m) * 32]
                              : *reinterpret_cast<const float4*>(X + xrow[rr] + jc0 + jj + 8 * h + 2 * t + m * cs);
            }
            const int fr = jj + 8 * h + g;
            // phi for every source of the pass: R = T V on the tensor pipe, then 1 / R
            float phv[NR][2][G];
#pragma unroll
            for (int gs = 0; gs < G; ++gs) {
              float R[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
              for (int ks = 0; ks < KS; ++ks) {
                const __nv_bfloat16* ph = vjk_hi + (gs * JCC + fr) * JKS + ks * 16 + 2 * t;
                const __nv_bfloat16* pl = vjk_lo + (gs * JCC + fr) * JKS + ks * 16 + 2 * t;
                mma_split(R, Thi[gs][ks], Tlo[gs][ks], lds32(ph), lds32(ph + 8), lds32(pl), lds32(pl + 8));
              }
#pragma unroll
              for (int r = 0; r < NR; ++r) {
                const int rr = RS ? rs : r;
                phv[r][0][gs] = fast_rcp(R[rr * 2 + 0]);
                phv[r][1][gs] = fast_rcp(R[rr * 2 + 1]);
              }
            }
#pragma unroll
            for (int r = 0; r < NR; ++r) {
              float xr[N], xi[N];
#pragma unroll
              for (int m = 0; m < N; ++m) {
                xr[m] = x[r][m].x;
                xi[m] = x[r][m].y;
              }
              acc[r].add(xr, xi, phv[r][0]);
#pragma unroll
              for (int m = 0; m < N; ++m) {
                xr[m] = x[r][m].z;
                xi[m] = x[r][m].w;
              }
              acc[r].add(xr, xi, phv[r][1]);
            }
          }
        }
      }
      if (STG) cp_async_wait<0>();
      if (warp_active) {
#pragma unroll
        for (int r = 0; r < NR; ++r) {
          const int rr = RS ? rs : r;
          acc[r].reduce4(invJ);
#pragma unroll
          for (int gs = 0; gs < G; ++gs) {
            const int n = n0 + gs;
            if (t == 0 && rvalid[rr] && n < N) acc[r].store(gs, U + (((size_t)b * I + row[rr]) * N + n) * N * N);
          }
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// kf_cov_w: weighted covariance with the weights given as an array phi[b, s, j] (AuxIVA: phi depends on
// (source, frame) only, ssspy/bss/iva.py:1785-1791).  Same tiling, staging and Hermitian register
// accumulation as kf_phi_cov, without the tensor-core stage.  `src` lists the source slots of phi / U.
struct CovSrc {
  int n;
};

template <int N, bool MN>
__global__ void __launch_bounds__(FW * 32) kf_cov_w(const cf* __restrict__ X, const float* __restrict__ phi,
                                                    long long sb, long long sn, long long si, int n_src,
                                                    cf* __restrict__ U, int I, int J, const float* __restrict__ Dm) {
  constexpr int G = CovShape<N>::G;
  constexpr bool RS = CovShape<N>::RS;
  constexpr bool STG = CovShape<N>::STG_W;
  constexpr int NR = RS ? 1 : 2;
  constexpr int NLD = 2 * NR * N;  // 16-byte vectors per lane per step (rows of this pass only)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* xring = reinterpret_cast<float4*>(smem_raw);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int b = blockIdx.z;
  float4* xw = xring + (size_t)warp * XSTAGES * NLD * 32 + lane;
  const int i0 = (blockIdx.y * FW + warp) * 16;
  if (i0 >= I) return;
  const int row[2] = {i0 + g, i0 + g + 8};
  const bool rvalid[2] = {row[0] < I, row[1] < I};
  const int rowc[2] = {min(row[0], I - 1), min(row[1], I - 1)};
  const size_t xrow[2] = {((size_t)b * N * I + rowc[0]) * J, ((size_t)b * N * I + rowc[1]) * J};
  const size_t cs = (size_t)I * J;
  const float invJ = 1.0f / (float)J;

  auto issue = [&](int f0, int stage, int rs) {
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const int rr = RS ? rs : r;
#pragma unroll
        for (int m = 0; m < N; ++m)
          cp_async16(xw + (stage * NLD + (h * NR + r) * N + m) * 32, X + xrow[rr] + f0 + 8 * h + 2 * t + m * cs);
      }
  };

  {
    const int s0 = blockIdx.x * G;
    for (int rs = 0; rs < (RS ? 2 : 1); ++rs) {
      HermAcc<N, G> acc[NR];
#pragma unroll
      for (int r = 0; r < NR; ++r) acc[r].zero();
      // MN (FastGaussMNMF, ssspy/bss/mnmf.py:1504-1514): `phi` holds Lambda[b, n, i, j] = (T V) and the weight of
      // "source" m is 1 / sum_n Lambda_n D[i, n, m], formed here instead of by a separate pass that wrote it out
      float dreg[MN ? NR : 1][MN ? N : 1][MN ? G : 1];
      if (MN) {
#pragma unroll
        for (int r = 0; r < NR; ++r)
#pragma unroll
          for (int n = 0; n < N; ++n)
#pragma unroll
            for (int gs = 0; gs < G; ++gs)
              dreg[MN ? r : 0][MN ? n : 0][MN ? gs : 0] =
                  Dm[(((size_t)b * I + rowc[RS ? rs : r]) * N + n) * N + min(s0 + gs, n_src - 1)];
      }
      int step = 0;
      if (STG) {
        issue(0, 0, rs);
        cp_async_commit();
      }
      for (int jj = 0; jj < J; jj += 16, ++step) {
        if (STG) {
          if (jj + 16 < J) issue(jj + 16, (step + 1) & 1, rs);
          cp_async_commit();
          cp_async_wait<1>();
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          float4 x[NR][N];
#pragma unroll
          for (int r = 0; r < NR; ++r) {
            const int rr = RS ? rs : r;
#pragma unroll
            for (int m = 0; m < N; ++m)
              x[r][m] = STG ? xw[((step & 1) * NLD + (h * NR + r) * N + m) * 32]
                            : *reinterpret_cast<const float4*>(X + xrow[rr] + jj + 8 * h + 2 * t + m * cs);
          }
#pragma unroll
          for (int r = 0; r < NR; ++r) {
            float ph0[G], ph1[G];
            if (MN) {
              float2 lam[N];
#pragma unroll
              for (int n = 0; n < N; ++n)
                lam[n] = *reinterpret_cast<const float2*>(phi + (size_t)b * sb + (size_t)n * sn +
                                                          (size_t)rowc[RS ? rs : r] * si + jj + 8 * h + 2 * t);
#pragma unroll
              for (int gs = 0; gs < G; ++gs) {
                float l0 = 0.f, l1 = 0.f;
#pragma unroll
                for (int n = 0; n < N; ++n) {
                  l0 = fmaf(lam[n].x, dreg[MN ? r : 0][MN ? n : 0][MN ? gs : 0], l0);
                  l1 = fmaf(lam[n].y, dreg[MN ? r : 0][MN ? n : 0][MN ? gs : 0], l1);
                }
                ph0[gs] = fast_rcp(l0);
                ph1[gs] = fast_rcp(l1);
              }
            } else {
#pragma unroll
              for (int gs = 0; gs < G; ++gs) {
                const int s = min(s0 + gs, n_src - 1);
                const float2 ph = *reinterpret_cast<const float2*>(phi + (size_t)b * sb + (size_t)s * sn +
                                                                   (size_t)rowc[RS ? rs : r] * si + jj + 8 * h + 2 * t);
                ph0[gs] = ph.x;
                ph1[gs] = ph.y;
              }
            }
            float xr[N], xi[N];
#pragma unroll
            for (int m = 0; m < N; ++m) {
              xr[m] = x[r][m].x;
              xi[m] = x[r][m].y;
            }
            acc[r].add(xr, xi, ph0);
#pragma unroll
            for (int m = 0; m < N; ++m) {
              xr[m] = x[r][m].z;
              xi[m] = x[r][m].w;
            }
            acc[r].add(xr, xi, ph1);
          }
        }
      }
      if (STG) cp_async_wait<0>();
#pragma unroll
      for (int r = 0; r < NR; ++r) {
        const int rr = RS ? rs : r;
        acc[r].reduce4(invJ);
#pragma unroll
        for (int gs = 0; gs < G; ++gs) {
          const int s = s0 + gs;
          if (t == 0 && rvalid[rr] && s < n_src) acc[r].store(gs, U + (((size_t)b * I + row[rr]) * n_src + s) * N * N);
        }
      }
    }
  }
}

template <int N>
int launch_cov_w(const cf* X, const float* phi, long long sb, long long sn, long long si, int n_src, cf* U, int B,
                 int I, int J, cudaStream_t st, const float* Dm = nullptr) {
  constexpr int NRC = CovShape<N>::RS ? 1 : 2;
  constexpr int G = CovShape<N>::G;
  const size_t sm = CovShape<N>::STG_W ? (size_t)FW * XSTAGES * 2 * NRC * N * 32 * sizeof(float4) : 0;
  static bool attr_dev[SSB_MAX_DEVICES] = {};  // function attributes are per device
  bool& attr_set = attr_dev[ssb_current_device()];
  if (!attr_set) {
    SSB_CUDA(cudaFuncSetAttribute(kf_cov_w<N, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    SSB_CUDA(cudaFuncSetAttribute(kf_cov_w<N, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    attr_set = true;
  }
  dim3 grid((n_src + G - 1) / G, (I + FW * 16 - 1) / (FW * 16), B);
  if (Dm != nullptr) {
    kf_cov_w<N, true><<<grid, FW * 32, sm, st>>>(X, phi, sb, sn, si, n_src, U, I, J, Dm);
    return ssb_check_launch("fused_cov_lambda", st);
  }
  kf_cov_w<N, false><<<grid, FW * 32, sm, st>>>(X, phi, sb, sn, si, n_src, U, I, J, nullptr);
  return ssb_check_launch("fused_cov_w", st);
}

// ------------------------------------------------------------------------------------------------
// kf_ip1_n2: IP1 for two sources in closed form, fp64, one thread per (mixture, bin)
// (ssspy/bss/_update_spatial_model.py:63-76).  When C (unweighted per-bin covariance) is given it
// also emits q[mat, n] = Re(w_n C_i w_n^H), the per-bin term of the power normalisation
// psi_n^2 = mean_{i,j} |y|^2 = mean_i q, so normalisation needs no pass over X.
__global__ void __launch_bounds__(128) kf_ip1_n2(cf* __restrict__ W, const cf* __restrict__ U,
                                                 const cf* __restrict__ C, double* __restrict__ q, int n_mat,
                                                 int flooring, double eps, int* __restrict__ status) {
  const int mat = blockIdx.x * blockDim.x + threadIdx.x;
  if (mat >= n_mat) return;
  cd w[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) w[e] = cf2cd(W[(size_t)mat * 4 + e]);
#pragma unroll
  for (int n = 0; n < 2; ++n) {
    cd u[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) u[e] = cf2cd(U[((size_t)mat * 2 + n) * 4 + e]);
    // A = W U_n
    cd a00 = cd_add(cd_mul(w[0], u[0]), cd_mul(w[1], u[2]));
    cd a01 = cd_add(cd_mul(w[0], u[1]), cd_mul(w[1], u[3]));
    cd a10 = cd_add(cd_mul(w[2], u[0]), cd_mul(w[3], u[2]));
    cd a11 = cd_add(cd_mul(w[2], u[1]), cd_mul(w[3], u[3]));
    // x = A^-1 e_n  (adjugate / det)
    cd det = cd_sub(cd_mul(a00, a11), cd_mul(a01, a10));
    if (det.x == 0.0 && det.y == 0.0) atomicOr(status, SSB_STATUS_SINGULAR);  // np.linalg.solve raises (_solve.py:15)
    cd idet = cd_inv(det);
    cd x0, x1;
    if (n == 0) {
      x0 = cd_mul(a11, idet);
      x1 = cd_mul(cd_make(-a10.x, -a10.y), idet);
    } else {
      x0 = cd_mul(cd_make(-a01.x, -a01.y), idet);
      x1 = cd_mul(a00, idet);
    }
    cd t0 = cd_add(cd_mul(u[0], x0), cd_mul(u[1], x1));
    cd t1 = cd_add(cd_mul(u[2], x0), cd_mul(u[3], x1));
    double qq = cd_mulc(t0, x0).x + cd_mulc(t1, x1).x;
    double d = ssb_floor(sqrt(fmax(qq, 0.0)), flooring, eps);
    w[n * 2 + 0] = cd_scale(cd_conj(x0), 1.0 / d);
    w[n * 2 + 1] = cd_scale(cd_conj(x1), 1.0 / d);
  }
  cf wf[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    wf[e] = cd2cf(w[e]);
    W[(size_t)mat * 4 + e] = wf[e];
  }
  if (C != nullptr) {
    cd c[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) c[e] = cf2cd(C[(size_t)mat * 4 + e]);
#pragma unroll
    for (int n = 0; n < 2; ++n) {
      // the stored (complex64) filter is what the normalisation sees
      const cd r0 = cf2cd(wf[n * 2]), r1 = cf2cd(wf[n * 2 + 1]);
      const cd s0 = cd_add(cd_mul(r0, c[0]), cd_mul(r1, c[2]));
      const cd s1 = cd_add(cd_mul(r0, c[1]), cd_mul(r1, c[3]));
      q[(size_t)mat * 2 + n] = cd_mulc(s0, r0).x + cd_mulc(s1, r1).x;
    }
  }
}

// kf_normalize: psi_n = floor(sqrt(mean_i q[b,i,n])), T[b,n] /= psi^p, W[b,:,n,:] /= psi
// (ssspy/bss/ilrma.py:412-444); one block per (mixture, source).
__global__ void __launch_bounds__(256) kf_normalize(const double* __restrict__ q, float* __restrict__ T,
                                                    cf* __restrict__ W, int N, int I, int K, float p, int flooring,
                                                    double eps) {
  __shared__ double sh[8];
  __shared__ double s_psi;
  const int n = blockIdx.x, b = blockIdx.y;
  double acc = 0.0;
  for (int i = threadIdx.x; i < I; i += blockDim.x) acc += q[((size_t)b * I + i) * N + n];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
    s_psi = ssb_floor(sqrt(t / (double)I), flooring, eps);
  }
  __syncthreads();
  const double psi = s_psi;
  const float it = (float)(1.0 / ((p == 2.0f) ? psi * psi : pow(psi, (double)p)));
  const float iw = (float)(1.0 / psi);
  float* Tb = T + ((size_t)b * N + n) * I * K;
  // every z-slice of the grid recomputes psi (1 K doubles) and scales its share of T and W
  const int tid = blockIdx.z * blockDim.x + threadIdx.x, nth = gridDim.z * blockDim.x;
  for (int e = tid; e < I * K; e += nth) Tb[e] *= it;
  for (int e = tid; e < I * N; e += nth) {
    const int i = e / N, m = e - i * N;
    cf* w = W + (((size_t)b * I + i) * N + n) * N + m;
    *w = make_float2(w->x * iw, w->y * iw);
  }
}

template <int N, int KS>
int launch_all(const ssb_config* c, const cf* X, cf* W, float* T, float* V, float* P, cf* U, ssb_fused_ws* vs,
               int phases, cudaStream_t st) {
  const int B = c->n_batch, I = c->n_bins, J = c->n_frames, K = c->n_basis;
  constexpr int KP = 16 * KS;
  constexpr bool STG = N <= 4;  // cp.async staging of X in the source-model kernels
  constexpr int G = CovShape<N>::G;
  const size_t ring16 = (size_t)FW * XSTAGES * 4 * N * 32 * sizeof(float4);
  const size_t sm_basis = (size_t)(2 * JC * (KP + PADH) + 2 * KP * (JC + PADH)) * sizeof(__nv_bfloat16) +
                          (STG ? ring16 : 0);
  const size_t sm_act = (size_t)(2 * BCH * (KP + PADH) + 2 * KP * (BCH + PADH)) * sizeof(__nv_bfloat16) +
                        (size_t)FW * XSTAGES * 8 * 32 * sizeof(float);
  constexpr int NRC = CovShape<N>::RS ? 1 : 2;
  const size_t ring_cov = (size_t)FW * XSTAGES * 2 * NRC * N * 32 * sizeof(float4);
  const size_t sm_cov = (size_t)(2 * G * CovShape<N>::jcc(KP) * (KP + PADH)) * sizeof(__nv_bfloat16) + (CovShape<N>::STG ? ring_cov : 0);
  // SSB_COOP: 1 (default) cooperative basis kernel (ssb_coop.cu), 0 one CTA per (mixture, source)
  const int coop = ssb_fused_coop_enabled();
  static bool attr_dev[SSB_MAX_DEVICES] = {};  // function attributes are per device
  bool& attr_set = attr_dev[ssb_current_device()];
  if (!attr_set) {
    SSB_CUDA(cudaFuncSetAttribute(kf_basis<N, KS, STG, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_basis));
    SSB_CUDA(cudaFuncSetAttribute(kf_activation<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_act));
    SSB_CUDA(cudaFuncSetAttribute(kf_phi_cov<N, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_cov));
    attr_set = true;
  }
  if (!(phases & 1)) {
    // covariance only (the source model of this iteration has already run)
  } else if (coop && vs != nullptr && W != nullptr) {
    if (ssb_tma_supported(c) && (ssb_tma_mask(c) & 1)) {
      // TMA-fed basis kernel (ssb_tma.cu) between the pre-split of V and the activation kernel of ssb_coop.cu
      if (ssb_coop_vsplit(c, V, vs->base, vs->vs_valid ? 1 : 0, st)) return 1;
      if (ssb_tma_basis(c, X, W, T, ssb_coop_vs(c, vs->base), P, ssb_coop_ts(c, vs->base), st)) return 1;
      if (ssb_coop_activation(c, V, P, vs->base, st)) return 1;
    } else if (ssb_coop_source(c, X, W, T, V, P, vs->base, vs->vs_valid ? 1 : 0, st)) {
      return 1;
    }
    vs->vs_valid = true;  // ssb_update_once clears it again unless it runs inside ssb_run
  } else {
    dim3 gb((I + FW * 16 - 1) / (FW * 16), N, B);
    kf_basis<N, KS, STG, 0><<<gb, FW * 32, sm_basis, st>>>(X, W, T, V, P, I, J, K, c->flooring, c->eps);
    if (ssb_check_launch("fused_basis", st)) return 1;
    dim3 ga((J + FW * 16 - 1) / (FW * 16), N, B);
    kf_activation<KS><<<ga, FW * 32, sm_act, st>>>(P, T, V, N, I, J, K, c->flooring, c->eps);
    if (ssb_check_launch("fused_activation", st)) return 1;
  }
  if (!(phases & 2)) return 0;
  // SSB_COOP_COV: 1 (default) cooperative covariance kernel for N = 4, 8 (needs the Vs left by ssb_coop_source)
  static int coop_cov = -1;
  if (coop_cov < 0) {
    const char* e = getenv("SSB_COOP_COV");
    coop_cov = e ? atoi(e) : 1;
  }
  // N = 2: TMA-fed covariance from the pre-split V left by the activation kernel (ssb_tma.cu)
  if (N == 2 && coop && vs != nullptr && vs->vs_valid && ssb_tma_supported(c) && (ssb_tma_mask(c) & 2))
    return ssb_tma_cov_n2(c, X, T, ssb_coop_vs(c, vs->base), U, st);
  // N = 4, 8: the frame reduction as a GEMM on the tensor pipe (ssb_covmma.cu), from the same pre-split V
  if ((N == 8 || N == 4) && coop && vs != nullptr && vs->vs_valid && ssb_cov_mma_supported(c, X))
    return ssb_cov_mma(c, X, T, ssb_coop_vs(c, vs->base), U, st);
  if (coop && coop_cov && vs != nullptr && W != nullptr && ssb_coop_cov_supported(c))
    return ssb_coop_cov(c, X, T, vs->base, U, st);
  dim3 gc((N + G - 1) / G, (I + FW * 16 - 1) / (FW * 16), B);
  kf_phi_cov<N, KS><<<gc, FW * 32, sm_cov, st>>>(X, T, V, U, I, J, K);
  return ssb_check_launch("fused_phi_cov", st);
}

}  // namespace

// ---- host side ------------------------------------------------------------------------------------
// SSB_COOP: 1 (default) cooperative source-model kernels (ssb_coop.cu), 0 one CTA per (mixture, source)
int ssb_fused_coop_enabled() {
  static int coop = -1;
  if (coop < 0) {
    const char* e = getenv("SSB_COOP");
    coop = e ? atoi(e) : 1;
  }
  return coop;
}

size_t ssb_fused_carve(ssb_fused_ws* ws, const ssb_config* c, char* base) {
  ws->base = base;
  const bool mnmf_ab = c->model == SSB_MODEL_FASTMNMF_GAUSS && c->n_basis <= 32 && (c->n_frames % 16) == 0;
  ws->bytes = (ssb_fused_supported(c) || mnmf_ab) ? ((ssb_coop_ws_bytes(c) + 255) & ~(size_t)255) : 0;
  ws->zeroed = false;
  ws->vs_valid = false;
  return ws->bytes;
}

int ssb_fused_supported(const ssb_config* c) {
  return c->model == SSB_MODEL_ILRMA_GAUSS && !c->partitioning && c->source == SSB_SOURCE_MM && c->domain == 2.0f &&
         c->n_basis <= 32 &&
         (c->n_frames % 16) == 0 && c->n_sources >= 2 && c->n_sources <= SSB_MAX_SOURCES;
}

int ssb_fused_prepare(ssb_fused_ws*, const ssb_config*, const cf*, cudaStream_t) { return 0; }

// source model (T then V) + weighted covariance U with the tensor-core kernels
int ssb_fused_source_and_cov(const ssb_config* c, ssb_fused_ws* ws, const cf* X, cf* W, float* T, float* V,
                             float* P, cf* U, cudaStream_t st, int phases) {
  const int KS = c->n_basis <= 16 ? 1 : 2;
  ssb_fused_ws* vs = (ws && ws->bytes) ? ws : nullptr;
  if (vs && !vs->zeroed) {
    SSB_CUDA(cudaMemsetAsync(vs->base, 0, vs->bytes, st));
    vs->zeroed = true;
    vs->vs_valid = false;
  }
  if (KS == 1) {
    SSB_DISPATCH_N(c->n_sources, return (launch_all<NN, 1>(c, X, W, T, V, P, U, vs, phases, st)));
  } else {
    SSB_DISPATCH_N(c->n_sources, return (launch_all<NN, 2>(c, X, W, T, V, P, U, vs, phases, st)));
  }
  return 0;
}

// iterations fused across the update_once boundary inside ssb_run (N = 2, IP1): see ssb_fused_spatial_source
int ssb_fused_iter_fusable(const ssb_config* c, const ssb_fused_ws* ws) {
  return ssb_fused_supported(c) && c->n_sources == 2 && c->spatial == SSB_SPATIAL_IP1 && ws != nullptr &&
         ws->bytes > 0 && ssb_fused_coop_enabled() && ssb_tma_supported(c) && (ssb_tma_mask(c) & 4) != 0 &&
         (c->normalization == SSB_NORM_POWER || c->normalization == SSB_NORM_NONE);
}

int ssb_fused_spatial_source(const ssb_config* c, ssb_fused_ws* ws, const cf* X, cf* W, float* T, float* V, float* P,
                             double* q, cudaStream_t st) {
  SSB_REQUIRE(ws != nullptr && ws->bytes > 0 && ws->zeroed && ws->vs_valid,
              "fused_spatial_source: the source model of the first iteration must have run in this ssb_run");
  SSB_REQUIRE(ssb_tma_supported(c) && (ssb_tma_mask(c) & 4),
              "fused_spatial_source: the fused covariance + IP1 + basis kernel is the TMA tile kernel (SSB_TMA bit 2)");
  // covariance + IP1 + next basis update on TMA-fed tiles whose frames are split over warps (the second pass over a tile
  // is served by L2), then the activation update.  The power normalisation of iteration t (W /= psi, T /= psi^2,
  // ilrma.py:412-444) commutes with the two source-model updates that follow it: with P -> P / psi^2, T -> T / psi^2 the
  // ratio of the basis update is unchanged and the activation update is invariant, so kf_normalize runs AFTER the
  // activation kernel and rescales the new T (and W); W is written back unnormalised and q[b,i,n] = mean_j |w_n^H x|^2 is
  // emitted for it (tests/test_oracle_golden.py::test_deferred_power_normalisation_commutes_with_the_source_model).
  if (ssb_tma_spatial_basis_n2(c, X, W, T, ssb_coop_vs(c, ws->base), P, ssb_coop_ts(c, ws->base), q, st)) return 1;
  return ssb_coop_activation(c, V, P, ws->base, st);
}

template <int KS>
int launch_source_iss(const ssb_config* c, const cf* Y, float* T, float* V, float* P, cudaStream_t st) {
  const int BN = c->n_batch * c->n_sources, I = c->n_bins, J = c->n_frames, K = c->n_basis;
  constexpr int KP = 16 * KS;
  const size_t ring16 = (size_t)FW * XSTAGES * 4 * 32 * sizeof(float4);
  const size_t sm_basis = (size_t)(2 * JC * (KP + PADH) + 2 * KP * (JC + PADH)) * sizeof(__nv_bfloat16) + ring16;
  const size_t sm_act = (size_t)(2 * BCH * (KP + PADH) + 2 * KP * (BCH + PADH)) * sizeof(__nv_bfloat16) +
                        (size_t)FW * XSTAGES * 8 * 32 * sizeof(float);
  static bool attr_dev[SSB_MAX_DEVICES] = {};  // function attributes are per device
  bool& attr_set = attr_dev[ssb_current_device()];
  if (!attr_set) {
    SSB_CUDA(cudaFuncSetAttribute(kf_basis<1, KS, true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_basis));
    SSB_CUDA(cudaFuncSetAttribute(kf_activation<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_act));
    attr_set = true;
  }
  // every (mixture, source) spectrogram is a one-channel "mixture" with the identity filter
  dim3 gb((I + FW * 16 - 1) / (FW * 16), 1, BN);
  kf_basis<1, KS, true, 0><<<gb, FW * 32, sm_basis, st>>>(Y, nullptr, T, V, P, I, J, K, c->flooring, c->eps);
  if (ssb_check_launch("fused_basis", st)) return 1;
  dim3 ga((J + FW * 16 - 1) / (FW * 16), 1, BN);
  kf_activation<KS><<<ga, FW * 32, sm_act, st>>>(P, T, V, 1, I, J, K, c->flooring, c->eps);
  return ssb_check_launch("fused_activation", st);
}

// MM source model (p = 2) for the ISS modes: P = |Y|^2 from the stored spectrograms
int ssb_fused_source_iss(const ssb_config* c, const cf* Y, float* T, float* V, float* P, cudaStream_t st) {
  if (c->n_basis <= 16) return launch_source_iss<1>(c, Y, T, V, P, st);
  return launch_source_iss<2>(c, Y, T, V, P, st);
}

// phi[B*N, I, J] = 1 / (T V)  (inverse = 1)  or  Lambda = T V  (inverse = 0)
int ssb_fused_phi(const ssb_config* c, const float* T, const float* V, float* phi, int inverse, cudaStream_t st) {
  const int BN = c->n_batch * c->n_sources, I = c->n_bins, J = c->n_frames, K = c->n_basis;
  dim3 grid((I + FW * 16 - 1) / (FW * 16), BN);
  if (K <= 16) {
    const size_t sm = (size_t)(3 * JC * (16 + PADH)) * sizeof(__nv_bfloat16);
    if (inverse) kf_phi<1, true><<<grid, FW * 32, sm, st>>>(T, V, phi, I, J, K);
    else kf_phi<1, false><<<grid, FW * 32, sm, st>>>(T, V, phi, I, J, K);
  } else {
    const size_t sm = (size_t)(3 * JC * (32 + PADH)) * sizeof(__nv_bfloat16);  // 60 KB: above the 48 KB default
    static bool attr_dev[SSB_MAX_DEVICES] = {};  // function attributes are per device
    bool& attr_set = attr_dev[ssb_current_device()];
    if (!attr_set) {
      SSB_CUDA(cudaFuncSetAttribute(kf_phi<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      SSB_CUDA(cudaFuncSetAttribute(kf_phi<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
      attr_set = true;
    }
    if (inverse) kf_phi<2, true><<<grid, FW * 32, sm, st>>>(T, V, phi, I, J, K);
    else kf_phi<2, false><<<grid, FW * 32, sm, st>>>(T, V, phi, I, J, K);
  }
  return ssb_check_launch(inverse ? "fused_phi" : "fused_lambda", st);
}

// weighted covariance with array weights phi[b*sb + s*sn + j] (n_frames % 16 == 0 required)
int ssb_fused_cov_w(const cf* X, const float* phi, long long sb, long long sn, long long si, int n_src, cf* U, int B,
                    int N, int I, int J, cudaStream_t st) {
  SSB_REQUIRE((J % 16) == 0, "fused_cov_w needs n_frames %% 16 == 0");
  SSB_DISPATCH_N(N, return (launch_cov_w<NN>(X, phi, sb, sn, si, n_src, U, B, I, J, st)));
  return 0;
}

// FastGaussMNMF: U[b, i, m] = mean_j x x^H / (sum_n Lambda[b, n, i, j] D[b, i, n, m])   (mnmf.py:1504-1514)
int ssb_fused_cov_lambda(const cf* X, const float* Lam, const float* Dm, cf* U, int B, int N, int I, int J,
                         cudaStream_t st) {
  SSB_REQUIRE((J % 16) == 0, "fused_cov_lambda needs n_frames %% 16 == 0");
  SSB_DISPATCH_N(N, return (launch_cov_w<NN>(X, Lam, (long long)N * I * J, (long long)I * J, J, N, U, B, I, J, st, Dm)));
  return 0;
}

int ssb_fused_ip1_n2(cf* W, const cf* U, const cf* C, double* q, int n_mat, int flooring, float eps,
                     cudaStream_t st) {
  kf_ip1_n2<<<blocks_for(n_mat, 128), 128, 0, st>>>(W, U, C, q, n_mat, flooring, (double)eps, ssb_status_word());
  return ssb_check_launch("fused_ip1_n2", st);
}

int ssb_fused_normalize(const double* q, float* T, cf* W, int B, int N, int I, int K, float p, int flooring,
                        float eps, cudaStream_t st) {
  dim3 grid(N, B, 8);
  kf_normalize<<<grid, 256, 0, st>>>(q, T, W, N, I, K, p, flooring, (double)eps);
  return ssb_check_launch("fused_normalize", st);
}
