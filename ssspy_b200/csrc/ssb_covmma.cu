// Weighted covariance for eight sources on the tensor pipe (GaussILRMA, p = 2):
//   phi[n,i,j] = 1 / (T V)[n,i,j]                                  (ssspy/bss/ilrma.py:1494-1498)
//   U[b,i,n,a,c] = (1/J) sum_j phi[n,i,j] x_a[i,j] conj(x_c[i,j])    (ilrma.py:1500-1505)
//
// kf_cov_coop (ssb_coop.cu) does the N^3 multiply-adds per (bin, frame) on the FP32 pipe: 2.06 ms at N = 8, I = 1025,
// J = 512, B = 64, bound by issued instructions (1.0 G warp instructions, profiles/r1_ncu_coop_summary.md).  The frame
// reduction is a GEMM once the Hermitian products are formed first (they do not depend on the source):
//   G[j, col]   = the 64 real numbers Re / Im x_a conj(x_c) (a < c) and |x_a|^2 of frame j      (CUDA cores, once)
//   U[n, col]   = sum_j phi[n, j] G[j, col]                                                    (tensor pipe)
// per bin an (8 sources) x (64 columns) x (J frames) product.  G is split into bf16 (hi, lo), phi into (hi, mid, lo); the
// sixteen rows of the m16n8k16 A operand hold phi_hi of the eight sources on rows 0-7 and phi_mid on rows 8-15, so ONE
// mma per (column tile, G part) yields phi_hi G and phi_mid G together; a second A operand carries phi_lo on rows 0-7
// and multiplies G_hi only: 24 mma per (bin, 16 frames) instead of 1152 FFMA per lane-frame group.
//
// CTA = one tile of 16 bins, 16 warps, one barrier per 16-frame step:
//   phase A  warp (source n, frame half h): R[16 bins x 8 frames] = T_n V_n on the tensor pipe (3 mma per 16 basis
//            vectors), phi = 1 / R, split, stored to shared memory as the A operand of every bin
//            ([bin][16 rows][16 frames] bf16, 48-byte rows and 784-byte bins: conflict-free stores and ldmatrix)
//   barrier  (the previous step's X stage and V chunk are free behind it: lane 0 of warp 0 / of warps 0-7 re-arm them)
//   phase B  warp = bin: A from ldmatrix, G from the bin's X slab (5 channels per lane: lane group g pairs channel g
//            with g+1, g+2, g+3, g+4; groups 4-7 use their last slot for two diagonal entries), 24 mma, fp32 accumulators
// X arrives by TMA (cp.async.bulk.tensor, tensor map with the plane axis INSIDE the bin axis so that a box lands as
// [bin][channel][64 bytes]: lanes of one quarter warp that read different channels hit different banks), V chunks by
// bulk copies, all on mbarriers; there is no empty-barrier: the per-step CTA barrier is the release.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include <mutex>

#include "ssb_fused.h"
#include "ssb_kernels.h"

namespace {

constexpr int PADH = 8, JCV = 32, XS = 3, N = 8;
constexpr int PHI_ROW = 48, PHI_BIN = 24 * PHI_ROW + 16;  // bytes: 1168 per bin (24 rows: hi, mid, lo of 8 sources)

struct Split {
  uint32_t hi, lo;
};
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
// Three-way split for the weights: hi + mid + lo with |x - hi - mid - lo| <= 2^-27 |x|.  The rounding of phi[n, j] is shared
// by every entry of U_n, so a two-way split (2^-18) perturbs the covariance coherently: in the fp64 oracle, rounding phi
// to hi + lo alone moves the final Y of an IP2 run by 2e-5 (N = 8, J = 1024, whitened, 5 iterations), the same split of
// the products G by 3e-6, and hi + mid + lo is indistinguishable from fp32 (2.4e-7).
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
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void ldsm_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr)
               : "memory");
}
__device__ __forceinline__ void ldsm_x2(uint32_t& r0, uint32_t& r1, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(1000000)  // suspend-time hint (ns)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

template <int KS>
struct CovShape8 {
  static constexpr int KP = 16 * KS, JKS = KP + PADH;
  static constexpr int CHB = 2 * JCV * JKS * 2;   // bytes of one source's V chunk (hi + lo)
  static constexpr int XHB = 16 * N * 64;         // one frame group (8 frames) of a tile: [bin][channel][64 bytes]
  static constexpr int XSB = 2 * XHB;             // one stage: 16 frames
  static constexpr int X_BYTES = XS * XSB;
  static constexpr int V_BYTES = N * 2 * CHB;
  static constexpr int PHI_BYTES = 2 * 16 * PHI_BIN;
  static constexpr int NBAR = XS + N * 2;
  static constexpr int SMEM = X_BYTES + V_BYTES + PHI_BYTES + NBAR * 8 + 128;
};

// PHI3: three-way weights (IP2, whose pairwise eigenproblems amplify the coherent rounding of phi); IP1 keeps the
// two-way split (its result is 1.4e-6 from the oracle at N = 8) and 16 mma per step: rows 8-15 then carry phi_lo.
template <int KS, bool PHI3>
__global__ void __launch_bounds__(512, 1)
    kc_cov_mma8(const __grid_constant__ CUtensorMap tmX, const float* __restrict__ T, const __nv_bfloat16* __restrict__ Vs,
                cf* __restrict__ U, int I, int J, int K, int nchunk) {
  using S = CovShape8<KS>;
  constexpr int JKS = S::JKS, CHB = S::CHB, XSB = S::XSB, XHB = S::XHB;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t smem_s = (uint32_t)__cvta_generic_to_shared(smem_raw);
  const uint32_t xring_s = smem_s;
  const uint32_t vring_s = smem_s + S::X_BYTES;                 // [source][2][CHB]
  const uint32_t phi_s = smem_s + S::X_BYTES + S::V_BYTES;      // [2][16 bins][PHI_BIN]
  const uint32_t bars_s = phi_s + S::PHI_BYTES;
  auto xfull = [&](int st) { return bars_s + (uint32_t)(st * 8); };
  auto vfull = [&](int n, int st) { return bars_s + (uint32_t)((XS + n * 2 + st) * 8); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int b = blockIdx.y, i0 = blockIdx.x * 16;
  const int nsteps = J >> 4;
  if (threadIdx.x == 0) {
    for (int e = 0; e < S::NBAR; ++e) mbar_init(bars_s + (uint32_t)(e * 8), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
  }
  __syncthreads();

  // ---- requests (issued by lane 0 of warp 0 for X, lane 0 of warp n < 8 for the V chunks of source n) ----
  auto x_request = [&](int s) {  // 16-frame stage s -> slot s % XS
    const uint32_t bar = xfull(s % XS), dst = xring_s + (uint32_t)((s % XS) * XSB);
    mbar_expect_tx(bar, XSB);
    tma_load_3d(dst, &tmX, 32 * s, b * N, i0, bar);
    tma_load_3d(dst + XHB, &tmX, 32 * s + 16, b * N, i0, bar);
  };
  const unsigned char* vsrc = reinterpret_cast<const unsigned char*>(Vs);
  auto v_request = [&](int n, int c) {  // chunk c (32 frames) of source n -> slot c & 1
    const uint32_t bar = vfull(n, c & 1);
    mbar_expect_tx(bar, CHB);
    bulk_load(vring_s + (uint32_t)((n * 2 + (c & 1)) * CHB), vsrc + (((size_t)b * N + n) * nchunk + c) * (size_t)CHB, CHB, bar);
  };
  if (warp == 0 && lane == 0)
    for (int s = 0; s < XS && s < nsteps; ++s) x_request(s);
  if (warp < N && lane == 0) {
    v_request(warp, 0);
    if (nchunk > 1) v_request(warp, 1);
  }

  // ---- phase A role: source nA, frame half hA of every step; T fragments of the tile's 16 bins ----
  const int nA = warp & 7, hA = warp >> 3;
  uint32_t Thi[KS][4], Tlo[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const float* tr = T + (((size_t)b * N + nA) * I + min(i0 + g + 8 * rr, I - 1)) * K;
#pragma unroll
      for (int nb = 0; nb < 2; ++nb) {
        const int k0 = ks * 16 + nb * 8 + 2 * t;
        const Split sp = split2((k0 < K) ? tr[k0] : 0.f, (k0 + 1 < K) ? tr[k0 + 1] : 0.f);
        Thi[ks][nb * 2 + rr] = sp.hi;
        Tlo[ks][nb * 2 + rr] = sp.lo;
      }
    }
  const int mid = lane >> 3, mrow = lane & 7;
  // V fragments of frames [8 hA, 8 hA + 8) of a step: matrices (hi k0-7, hi k8-15, lo k0-7, lo k8-15)
  const uint32_t vlane = vring_s + (uint32_t)(nA * 2 * CHB) + (mid >> 1) * (JCV * JKS * 2) +
                         ((8 * hA + mrow) * JKS + (mid & 1) * 8) * 2;
  // phi store: bins g, g + 8; rows nA (hi), 8 + nA (mid), 16 + nA (lo); frames 8 hA + 2 t, + 1
  const uint32_t plane_st = phi_s + (uint32_t)(g * PHI_BIN + nA * PHI_ROW + (8 * hA + 2 * t) * 2);

  // ---- phase B role: bin bb = warp ----
  const int bb = warp;
  const bool bin_valid = i0 + bb < I;
  // A operand (16 x 16 bf16): matrices (rows 0-7, k 0-7), (rows 8-15, k 0-7), (rows 0-7, k 8-15), (rows 8-15, k 8-15)
  const uint32_t plane_ld = phi_s + (uint32_t)(bb * PHI_BIN + (((mid & 1) * 8 + mrow) * PHI_ROW) + (mid >> 1) * 16);
  // second A operand: rows 0-7 = phi_lo (rows 16-23 of the bin), rows 8-15 = 0; matrices (k 0-7), (k 8-15)
  const uint32_t plane_ld2 = phi_s + (uint32_t)(bb * PHI_BIN + ((16 + mrow) * PHI_ROW) + (mid & 1) * 16);
  // X of this bin: channel c_k = (g + k) & 7, frames 8 h + 2 t, + 1 at  stage + h * XHB + bb * 512 + c_k * 64 + t * 16
  uint32_t xoff[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) xoff[k] = xring_s + (uint32_t)(bb * 512 + ((g + k) & 7) * 64 + t * 16);
  const bool diag_slot = g >= 4;  // the (g, g + 4) slot of lane groups 4-7 carries |x_g|^2 and |x_{g+4}|^2 instead
  // D: the accumulator fragments of the running mma chain; every FLUSH steps they are added into Ssum with ordinary
  // round-to-nearest FADDs and cleared.  The tensor pipe adds into its fp32 accumulator with truncation, so a chain over
  // all J / 16 steps would bias the (positive) diagonal sums by about (J / 16) * 2^-24; chains of FLUSH = 2 keep the
  // accumulation error at the level of the lane-partial sums of the FP32-pipe kernel (kf_cov_coop).
  constexpr int FLUSH = 2;
  float D[8][4], Ssum[8][2];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
#pragma unroll
    for (int c = 0; c < 4; ++c) D[q][c] = 0.f;
    Ssum[q][0] = Ssum[q][1] = 0.f;
  }

#pragma unroll 1
  for (int s = 0; s < nsteps; ++s) {
    // ======== phase A: phi of (source nA, frames 8 hA ..) for the 16 bins ========
    {
      const int c = s >> 1;
      mbar_wait(vfull(nA, c & 1), (c >> 1) & 1);  // (already complete on the second step of a chunk)
      const uint32_t vb = vlane + (c & 1) * CHB + (s & 1) * (16 * JKS * 2);
      float R[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t bh0, bh1, bl0, bl1;
        ldsm_x4(bh0, bh1, bl0, bl1, vb + (ks * 16) * 2);
        mma16816(R, Thi[ks], bh0, bh1);
        mma16816(R, Thi[ks], bl0, bl1);
        mma16816(R, Tlo[ks], bh0, bh1);
        mma16816(R, Tlo[ks], bl0, bl1);  // all four partial products: phase A is a small part of the step
      }
      // no floor on R (ilrma.py:1494-1498)
      const Split3 p0 = split3(fast_rcp(R[0]), fast_rcp(R[1]));  // bin g
      const Split3 p1 = split3(fast_rcp(R[2]), fast_rcp(R[3]));  // bin g + 8
      const uint32_t pd = plane_st + (s & 1) * (16 * PHI_BIN);
      sts32(pd, p0.hi);
      sts32(pd + 8 * PHI_BIN, p1.hi);
      if (PHI3) {
        sts32(pd + 8 * PHI_ROW, p0.mid);
        sts32(pd + 16 * PHI_ROW, p0.lo);
        sts32(pd + 8 * PHI_BIN + 8 * PHI_ROW, p1.mid);
        sts32(pd + 8 * PHI_BIN + 16 * PHI_ROW, p1.lo);
      } else {  // hi + mid is the two-way split
        sts32(pd + 8 * PHI_ROW, p0.mid);
        sts32(pd + 8 * PHI_BIN + 8 * PHI_ROW, p1.mid);
      }
    }
    __syncthreads();
    // behind the barrier every warp has left step s - 1: its X stage and (after an odd step) its V chunk are free
    if (lane == 0 && s >= 1) {
      if (warp == 0 && s - 1 + XS < nsteps) x_request(s - 1 + XS);
      if (warp < N && (s & 1) == 0) {
        const int cn = (s >> 1) + 1;
        if (cn < nchunk) v_request(warp, cn);
      }
    }
    // ======== phase B: U[:, cols] += phi[:, frames] G[frames, cols] for bin bb ========
    mbar_wait(xfull(s % XS), (s / XS) & 1);
    uint32_t A[4];
    ldsm_x4(A[0], A[1], A[2], A[3], plane_ld + (s & 1) * (16 * PHI_BIN));
    uint32_t A2[4] = {0u, 0u, 0u, 0u};
    if (PHI3) ldsm_x2(A2[0], A2[2], plane_ld2 + (s & 1) * (16 * PHI_BIN));
    const uint32_t xst = (uint32_t)((s % XS) * XSB);
    float4 x0[2];  // channel c_0 = g: (re, im) of frames 8 h + 2 t, + 1
#pragma unroll
    for (int h = 0; h < 2; ++h) x0[h] = lds128(xoff[0] + xst + h * XHB);
#pragma unroll
    for (int qp = 0; qp < 4; ++qp) {
      float re[2][2], im[2][2];  // [h][frame e]
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float4 xb = lds128(xoff[qp + 1] + xst + h * XHB);
        const float4 xa = x0[h];
        if (qp < 3) {  // x_a conj(x_c) = (ar cr + ai ci) + i (ai cr - ar ci)
          re[h][0] = fmaf(xa.x, xb.x, xa.y * xb.y);
          im[h][0] = fmaf(xa.y, xb.x, -(xa.x * xb.y));
          re[h][1] = fmaf(xa.z, xb.z, xa.w * xb.w);
          im[h][1] = fmaf(xa.w, xb.z, -(xa.z * xb.w));
        } else {
          // lane groups 0-3: the pair (g, g + 4); groups 4-7: |x_g|^2 in the Re column, |x_{g+4}|^2 in the Im column
          const float ux0 = diag_slot ? xa.x : xb.x, uy0 = diag_slot ? xa.y : xb.y;
          const float ux1 = diag_slot ? xa.z : xb.z, uy1 = diag_slot ? xa.w : xb.w;
          re[h][0] = fmaf(xa.x, ux0, xa.y * uy0);
          re[h][1] = fmaf(xa.z, ux1, xa.w * uy1);
          const float v10 = diag_slot ? xb.x : xa.y, v20 = diag_slot ? xb.y : -xa.x;
          const float v11 = diag_slot ? xb.z : xa.w, v21 = diag_slot ? xb.w : -xa.z;
          im[h][0] = fmaf(v10, xb.x, v20 * xb.y);
          im[h][1] = fmaf(v11, xb.z, v21 * xb.w);
        }
      }
      const Split r0 = split2(re[0][0], re[0][1]), r1 = split2(re[1][0], re[1][1]);
      mma16816(D[2 * qp], A, r0.hi, r1.hi);
      mma16816(D[2 * qp], A, r0.lo, r1.lo);
      if (PHI3) mma16816(D[2 * qp], A2, r0.hi, r1.hi);  // phi_lo G_hi (phi_lo G_lo is below 2^-27)
      const Split m0 = split2(im[0][0], im[0][1]), m1 = split2(im[1][0], im[1][1]);
      mma16816(D[2 * qp + 1], A, m0.hi, m1.hi);
      mma16816(D[2 * qp + 1], A, m0.lo, m1.lo);
      if (PHI3) mma16816(D[2 * qp + 1], A2, m0.hi, m1.hi);
    }
    if ((s % FLUSH) == FLUSH - 1 || s == nsteps - 1) {  // rows g (phi_hi, phi_lo parts) + g + 8 (phi_mid part) of source g
#pragma unroll
      for (int q = 0; q < 8; ++q)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          Ssum[q][e] += D[q][e] + D[q][2 + e];
          D[q][e] = D[q][2 + e] = 0.f;
        }
    }
  }
  // ---- U[b, i0 + bb, n = g, :, :]: source g; columns 2t, 2t + 1 of every tile ----
  if (!bin_valid) return;
  const float invJ = 1.0f / (float)J;
  cf* u = U + (((size_t)b * I + i0 + bb) * N + g) * N * N;
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int jc = 2 * t + e;  // column inside the tiles = lane group that produced it
#pragma unroll
    for (int qp = 0; qp < 4; ++qp) {
      const float re = Ssum[2 * qp][e] * invJ;
      const float im = Ssum[2 * qp + 1][e] * invJ;
      if (qp < 3 || jc < 4) {
        const int a = jc, c = (jc + qp + 1) & 7;
        u[a * N + c] = make_float2(re, im);
        u[c * N + a] = make_float2(re, -im);
      } else {
        const int a = jc, c = (jc + 4) & 7;
        u[a * N + a] = make_float2(re, 0.f);
        u[c * N + c] = make_float2(im, 0.f);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Four sources: the 4 + 2 * 6 = 16 real numbers of a frame's Hermitian products fill two column tiles (Re of the six
// pairs + |x_0|^2, |x_1|^2; Im of the six pairs + |x_2|^2, |x_3|^2), and ONE A operand holds the three-way weights of
// all four sources (rows 0-3 hi, 4-7 mid, 8-11 lo, 12-15 zero): 4 mma per (bin, 16 frames).  CTA = one tile of 16 bins,
// 8 warps; phase A: warp (source w & 3, frame half w >> 2); phase B: warp w takes bins w and w + 8.  Lane group g forms
// column g of both tiles: g < 4 the pair (g, g + 1 mod 4), g = 4, 5 the pair (g - 4, g - 2), g = 6, 7 the diagonals
// (g - 6, g - 4).  A source's sum is spread over rows n, n + 4, n + 8: lanes g = n and g = n + 4 combine at the end.
constexpr int N4 = 4;
constexpr int PHI_BIN4 = 16 * PHI_ROW + 16;  // 784 bytes per bin

template <int KS>
struct CovShape4 {
  static constexpr int KP = 16 * KS, JKS = KP + PADH;
  static constexpr int CHB = 2 * JCV * JKS * 2;
  static constexpr int XHB = 16 * N4 * 64;        // one frame group (8 frames): [bin][channel][64 bytes]
  static constexpr int XSB = 2 * XHB;
  // a 16-frame step is short at N = 4 (4 mma per bin): the rings must run further ahead than at N = 8 to cover the
  // latency of the bulk copies (two steps ahead measured 0.45 ms, slower than the FP32-pipe kernel)
  static constexpr int XSN = 5;                   // X stages
  static constexpr int VSN = 3;                   // V chunk slots per source
  static constexpr int X_BYTES = XSN * XSB;
  static constexpr int V_BYTES = N4 * VSN * CHB;
  static constexpr int PHI_BYTES = 2 * 16 * PHI_BIN4;
  static constexpr int NBAR = XSN + N4 * VSN;
  static constexpr int SMEM = X_BYTES + V_BYTES + PHI_BYTES + NBAR * 8 + 128;
};

template <int KS>
__global__ void __launch_bounds__(256, 2)
    kc_cov_mma4(const __grid_constant__ CUtensorMap tmX, const float* __restrict__ T, const __nv_bfloat16* __restrict__ Vs,
                cf* __restrict__ U, int I, int J, int K, int nchunk) {
  using S = CovShape4<KS>;
  constexpr int JKS = S::JKS, CHB = S::CHB, XSB = S::XSB, XHB = S::XHB, N = N4, XSN = S::XSN, VSN = S::VSN;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t smem_s = (uint32_t)__cvta_generic_to_shared(smem_raw);
  const uint32_t xring_s = smem_s;
  const uint32_t vring_s = smem_s + S::X_BYTES;                 // [source][VSN][CHB]
  const uint32_t phi_s = smem_s + S::X_BYTES + S::V_BYTES;      // [2][16 bins][PHI_BIN4]
  const uint32_t bars_s = phi_s + S::PHI_BYTES;
  auto xfull = [&](int st) { return bars_s + (uint32_t)(st * 8); };
  auto vfull = [&](int n, int st) { return bars_s + (uint32_t)((XSN + n * VSN + st) * 8); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int b = blockIdx.y, i0 = blockIdx.x * 16;
  const int nsteps = J >> 4;
  if (threadIdx.x == 0) {
    for (int e = 0; e < S::NBAR; ++e) mbar_init(bars_s + (uint32_t)(e * 8), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
  }
  // rows 12-15 of every bin's weight block stay zero for the whole kernel
  for (int e = threadIdx.x; e < 2 * 16 * 4 * (PHI_ROW / 4); e += 256) {
    const int w32 = e % (PHI_ROW / 4), row = 12 + (e / (PHI_ROW / 4)) % 4, bin = e / (4 * (PHI_ROW / 4));
    sts32(phi_s + (uint32_t)(bin * PHI_BIN4 + row * PHI_ROW + w32 * 4), 0u);
  }
  __syncthreads();

  auto x_request = [&](int s) {  // 16-frame stage s -> slot s % XSN
    const uint32_t bar = xfull(s % XSN), dst = xring_s + (uint32_t)((s % XSN) * XSB);
    mbar_expect_tx(bar, XSB);
    tma_load_3d(dst, &tmX, 32 * s, b * N, i0, bar);
    tma_load_3d(dst + XHB, &tmX, 32 * s + 16, b * N, i0, bar);
  };
  const unsigned char* vsrc = reinterpret_cast<const unsigned char*>(Vs);
  auto v_request = [&](int n, int c) {  // chunk c (32 frames) of source n -> slot c % VSN
    const uint32_t bar = vfull(n, c % VSN);
    mbar_expect_tx(bar, CHB);
    bulk_load(vring_s + (uint32_t)((n * VSN + (c % VSN)) * CHB), vsrc + (((size_t)b * N + n) * nchunk + c) * (size_t)CHB, CHB, bar);
  };
  if (warp == 0 && lane == 0)
    for (int s = 0; s < XSN && s < nsteps; ++s) x_request(s);
  if (warp < N && lane == 0)
    for (int c = 0; c < VSN && c < nchunk; ++c) v_request(warp, c);

  // ---- phase A role: source nA, frame half hA; T fragments of the tile's 16 bins ----
  const int nA = warp & 3, hA = warp >> 2;
  uint32_t Thi[KS][4], Tlo[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const float* tr = T + (((size_t)b * N + nA) * I + min(i0 + g + 8 * rr, I - 1)) * K;
#pragma unroll
      for (int nb = 0; nb < 2; ++nb) {
        const int k0 = ks * 16 + nb * 8 + 2 * t;
        const Split sp = split2((k0 < K) ? tr[k0] : 0.f, (k0 + 1 < K) ? tr[k0 + 1] : 0.f);
        Thi[ks][nb * 2 + rr] = sp.hi;
        Tlo[ks][nb * 2 + rr] = sp.lo;
      }
    }
  const int mid = lane >> 3, mrow = lane & 7;
  const uint32_t vlane = vring_s + (uint32_t)(nA * VSN * CHB) + (mid >> 1) * (JCV * JKS * 2) +
                         ((8 * hA + mrow) * JKS + (mid & 1) * 8) * 2;
  // phi store: bins g, g + 8; rows nA (hi), 4 + nA (mid), 8 + nA (lo); frames 8 hA + 2 t, + 1
  const uint32_t plane_st = phi_s + (uint32_t)(g * PHI_BIN4 + nA * PHI_ROW + (8 * hA + 2 * t) * 2);

  // ---- phase B role: bins warp and warp + 8 ----
  // A operand (16 x 16 bf16): matrices (rows 0-7, k 0-7), (rows 8-15, k 0-7), (rows 0-7, k 8-15), (rows 8-15, k 8-15)
  const uint32_t plane_ld = phi_s + (uint32_t)(warp * PHI_BIN4 + (((mid & 1) * 8 + mrow) * PHI_ROW) + (mid >> 1) * 16);
  const int ca = g < 4 ? g : (g < 6 ? g - 4 : g - 6);
  const int cb = g < 4 ? ((g + 1) & 3) : (g < 6 ? g - 2 : g - 4);
  const bool diag = g >= 6;
  const uint32_t xa_off = xring_s + (uint32_t)(warp * (N * 64) + ca * 64 + t * 16);
  const uint32_t xb_off = xring_s + (uint32_t)(warp * (N * 64) + cb * 64 + t * 16);
  constexpr int FLUSH = 2;  // see kc_cov_mma8
  float D[2][2][4], Ssum[2][2][2];
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int tl = 0; tl < 2; ++tl) {
#pragma unroll
      for (int c = 0; c < 4; ++c) D[q][tl][c] = 0.f;
      Ssum[q][tl][0] = Ssum[q][tl][1] = 0.f;
    }

#pragma unroll 1
  for (int s = 0; s < nsteps; ++s) {
    // ======== phase A: phi of (source nA, frames 8 hA ..) for the 16 bins ========
    {
      const int c = s >> 1;
      mbar_wait(vfull(nA, c % VSN), (c / VSN) & 1);
      const uint32_t vb = vlane + (c % VSN) * CHB + (s & 1) * (16 * JKS * 2);
      float R[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t bh0, bh1, bl0, bl1;
        ldsm_x4(bh0, bh1, bl0, bl1, vb + (ks * 16) * 2);
        mma16816(R, Tlo[ks], bl0, bl1);
        mma16816(R, Tlo[ks], bh0, bh1);
        mma16816(R, Thi[ks], bl0, bl1);
        mma16816(R, Thi[ks], bh0, bh1);
      }
      const Split3 p0 = split3(fast_rcp(R[0]), fast_rcp(R[1]));  // bin g
      const Split3 p1 = split3(fast_rcp(R[2]), fast_rcp(R[3]));  // bin g + 8
      const uint32_t pd = plane_st + (s & 1) * (16 * PHI_BIN4);
      sts32(pd, p0.hi);
      sts32(pd + 4 * PHI_ROW, p0.mid);
      sts32(pd + 8 * PHI_ROW, p0.lo);
      sts32(pd + 8 * PHI_BIN4, p1.hi);
      sts32(pd + 8 * PHI_BIN4 + 4 * PHI_ROW, p1.mid);
      sts32(pd + 8 * PHI_BIN4 + 8 * PHI_ROW, p1.lo);
    }
    __syncthreads();
    // behind the barrier every warp has left step s - 1: its X stage and (after an odd step) its V chunk are free
    if (lane == 0 && s >= 1) {
      if (warp == 0 && s - 1 + XSN < nsteps) x_request(s - 1 + XSN);
      if (warp < N && (s & 1) == 0) {  // chunk (s >> 1) - 1 was consumed in steps s - 2, s - 1: its slot is free
        const int cn = (s >> 1) - 1 + VSN;
        if (cn < nchunk) v_request(warp, cn);
      }
    }
    // ======== phase B: U[:, cols] += phi[:, frames] G[frames, cols] for bins warp, warp + 8 ========
    mbar_wait(xfull(s % XSN), (s / XSN) & 1);
    const uint32_t xst = (uint32_t)((s % XSN) * XSB);
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      uint32_t A[4];
      ldsm_x4(A[0], A[1], A[2], A[3], plane_ld + (s & 1) * (16 * PHI_BIN4) + q * (8 * PHI_BIN4));
      float re[2][2], im[2][2];  // [h][frame e]
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float4 xa = lds128(xa_off + xst + h * XHB + q * (8 * N * 64));
        const float4 xb = lds128(xb_off + xst + h * XHB + q * (8 * N * 64));
        // pair: x_a conj(x_c) = (ar cr + ai ci) + i (ai cr - ar ci); diagonal lanes: |x_a|^2 and |x_c|^2
        const float ux0 = diag ? xa.x : xb.x, uy0 = diag ? xa.y : xb.y;
        const float ux1 = diag ? xa.z : xb.z, uy1 = diag ? xa.w : xb.w;
        re[h][0] = fmaf(xa.x, ux0, xa.y * uy0);
        re[h][1] = fmaf(xa.z, ux1, xa.w * uy1);
        const float v10 = diag ? xb.x : xa.y, v20 = diag ? xb.y : -xa.x;
        const float v11 = diag ? xb.z : xa.w, v21 = diag ? xb.w : -xa.z;
        im[h][0] = fmaf(v10, xb.x, v20 * xb.y);
        im[h][1] = fmaf(v11, xb.z, v21 * xb.w);
      }
      const Split r0 = split2(re[0][0], re[0][1]), r1 = split2(re[1][0], re[1][1]);
      mma16816(D[q][0], A, r0.hi, r1.hi);
      mma16816(D[q][0], A, r0.lo, r1.lo);
      const Split m0 = split2(im[0][0], im[0][1]), m1 = split2(im[1][0], im[1][1]);
      mma16816(D[q][1], A, m0.hi, m1.hi);
      mma16816(D[q][1], A, m0.lo, m1.lo);
    }
    if ((s % FLUSH) == FLUSH - 1 || s == nsteps - 1) {
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int tl = 0; tl < 2; ++tl)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            Ssum[q][tl][e] += D[q][tl][e] + D[q][tl][2 + e];
            D[q][tl][e] = D[q][tl][2 + e] = 0.f;
          }
    }
  }
  // ---- lanes g = n hold rows n (hi) + n + 8 (lo), lanes g = n + 4 row n + 4 (mid): combine, then lane group n < 4
  //      writes U[b, bin, n, :, :]; columns 2t, 2t + 1 of both tiles ----
  const float invJ = 1.0f / (float)J;
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    float tot[2][2];
#pragma unroll
    for (int tl = 0; tl < 2; ++tl)
#pragma unroll
      for (int e = 0; e < 2; ++e)
        tot[tl][e] = (Ssum[q][tl][e] + __shfl_down_sync(0xffffffffu, Ssum[q][tl][e], 16)) * invJ;
    const int bin = i0 + warp + 8 * q;
    if (g >= 4 || bin >= I) continue;
    cf* u = U + (((size_t)b * I + bin) * N + g) * N * N;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int col = 2 * t + e;  // the lane group that produced the column
      const float re = tot[0][e], im = tot[1][e];
      if (col < 6) {
        const int a = col < 4 ? col : col - 4, c = col < 4 ? ((col + 1) & 3) : col - 2;
        u[a * N + c] = make_float2(re, im);
        u[c * N + a] = make_float2(re, -im);
      } else {
        const int a = col - 6, c = col - 4;
        u[a * N + a] = make_float2(re, 0.f);
        u[c * N + c] = make_float2(im, 0.f);
      }
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

// X as [I bins][B*N planes][2 J floats]: the plane axis sits inside the bin axis (its stride is the larger one), so a
// box of 8 frames x N planes x 16 bins lands in shared memory as [bin][channel][64 bytes]
int make_x_map_bin_major(CUtensorMap* tm, const cf* X, int B, int I, int J, int n_ch) {
  EncodeTiledFn enc = encode_fn();
  if (enc == nullptr || (reinterpret_cast<uintptr_t>(X) & 15) != 0) return 1;
  const cuuint64_t dims[3] = {(cuuint64_t)2 * J, (cuuint64_t)B * n_ch, (cuuint64_t)I};
  const cuuint64_t strides[2] = {(cuuint64_t)I * J * 8, (cuuint64_t)J * 8};
  const cuuint32_t box[3] = {16, (cuuint32_t)n_ch, 16};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<cf*>(X), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
             ? 0
             : 1;
}

template <int KS, bool PHI3>
int launch_cov_mma8(const ssb_config* c, const cf* X, const float* T, const __nv_bfloat16* Vs, cf* U, cudaStream_t st) {
  using S = CovShape8<KS>;
  const int B = c->n_batch, I = c->n_bins, J = c->n_frames, K = c->n_basis;
  CUtensorMap tm;
  SSB_REQUIRE(make_x_map_bin_major(&tm, X, B, I, J, N) == 0, "cov_mma: the bin-major tensor map of X cannot be encoded");
  static bool attr_dev[SSB_MAX_DEVICES] = {};
  bool& attr_set = attr_dev[ssb_current_device()];
  if (!attr_set) {
    SSB_CUDA(cudaFuncSetAttribute(kc_cov_mma8<KS, PHI3>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::SMEM));
    attr_set = true;
  }
  dim3 grid((I + 15) / 16, B);
  kc_cov_mma8<KS, PHI3><<<grid, 512, S::SMEM, st>>>(tm, T, Vs, U, I, J, K, (J + JCV - 1) / JCV);
  return ssb_check_launch("mma_phi_cov", st);
}

template <int KS>
int launch_cov_mma4(const ssb_config* c, const cf* X, const float* T, const __nv_bfloat16* Vs, cf* U, cudaStream_t st) {
  using S = CovShape4<KS>;
  const int B = c->n_batch, I = c->n_bins, J = c->n_frames, K = c->n_basis;
  CUtensorMap tm;
  SSB_REQUIRE(make_x_map_bin_major(&tm, X, B, I, J, N4) == 0, "cov_mma: the bin-major tensor map of X cannot be encoded");
  static bool attr_dev[SSB_MAX_DEVICES] = {};
  bool& attr_set = attr_dev[ssb_current_device()];
  if (!attr_set) {
    SSB_CUDA(cudaFuncSetAttribute(kc_cov_mma4<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::SMEM));
    attr_set = true;
  }
  dim3 grid((I + 15) / 16, B);
  kc_cov_mma4<KS><<<grid, 256, S::SMEM, st>>>(tm, T, Vs, U, I, J, K, (J + JCV - 1) / JCV);
  return ssb_check_launch("mma_phi_cov", st);
}

}  // namespace

// SSB_COV_MMA (read at every call): unset / 2 = tensor-core covariance at N = 8 (and N = 4 with SSB_COV_MMA4=1) for IP1 and
// IP2, 1 = IP1 only,
// 0 = never (kf_cov_coop on the FP32 pipe).  IP2's pairwise generalised eigenproblems amplify the rounding of the
// weights: with two-way operands BASELINE config 4 ended 1.1e-4 from the fp64 oracle (bound 1e-4; 7.3e-5 with
// kf_cov_coop); with the three-way weights, flushed accumulators and round-to-nearest splits everywhere it is 3.9e-5
// (profiles/r2_baseline_shapes_rn_split.txt), so IP2 runs here by default as well.
int ssb_cov_mma_supported(const ssb_config* c, const cf* X) {
  const char* e = getenv("SSB_COV_MMA");
  const int mode = e != nullptr ? atoi(e) : 2;
  if (mode == 0) return 0;
  if (c->spatial != SSB_SPATIAL_IP1 && mode < 2) return 0;
  if ((c->n_sources != 8 && c->n_sources != 4) || (c->n_frames % 16) != 0 || c->n_basis > 32) return 0;
  // N = 4 is opt-in (SSB_COV_MMA4=1): parity-green, but 0.45 ms against 0.35 ms for kf_cov_coop at I = 1025, J = 512,
  // B = 64 (gpurun_out/r2r_n4_*.json).  With 16 columns per frame the products, their bf16 splits and the diagonal-lane
  // selects cost as many issue slots (~210 per warp and 16-frame step) as the 64 FFMA per frame they replace; deeper
  // rings (5 X stages, 3 V slots) changed nothing, i.e. the kernel is issue-bound, not latency-bound.
  if (c->n_sources == 4 && !(getenv("SSB_COV_MMA4") != nullptr && atoi(getenv("SSB_COV_MMA4")) == 1)) return 0;
  static int map_ok = -1;  // whether the driver accepts a tensor map whose strides are not increasing
  if (map_ok < 0) {
    CUtensorMap tm;
    map_ok = make_x_map_bin_major(&tm, X, c->n_batch, c->n_bins, c->n_frames, 8) == 0 ? 1 : 0;
  }
  return map_ok;
}

int ssb_cov_mma(const ssb_config* c, const cf* X, const float* T, const void* Vs, cf* U, cudaStream_t st) {
  if (c->n_sources == 4)
    return c->n_basis <= 16 ? launch_cov_mma4<1>(c, X, T, (const __nv_bfloat16*)Vs, U, st)
                            : launch_cov_mma4<2>(c, X, T, (const __nv_bfloat16*)Vs, U, st);
  const bool phi3 = c->spatial != SSB_SPATIAL_IP1;
  if (c->n_basis <= 16)
    return phi3 ? launch_cov_mma8<1, true>(c, X, T, (const __nv_bfloat16*)Vs, U, st)
                : launch_cov_mma8<1, false>(c, X, T, (const __nv_bfloat16*)Vs, U, st);
  return phi3 ? launch_cov_mma8<2, true>(c, X, T, (const __nv_bfloat16*)Vs, U, st)
              : launch_cov_mma8<2, false>(c, X, T, (const __nv_bfloat16*)Vs, U, st);
}
