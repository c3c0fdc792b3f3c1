// Cooperative tensor-core kernels of the fused GaussILRMA path (source_algorithm="MM", domain=2) and of FastGaussMNMF.
//
//   kf_vsplit            V -> pre-split bf16 (hi, lo) chunks (only when V did not come from kf_activation_coop)
//   kf_basis_coop        T <- T sqrt( sum_j V P / R^2  /  sum_j V / R ),  P = |w_n^H x|^2,  R = T V
//                        (ssspy/bss/ilrma.py:1051-1128); also writes P (16 x 16 tiles) and the pre-split T
//   kf_activation_coop   V <- V sqrt( sum_i T P / R^2  /  sum_i T / R )   (ilrma.py:1130-1204); also the pre-split V
//   kf_update_ab         the same two updates with the elementwise factors given as arrays (ssspy/bss/mnmf.py:1351-1415)
//   kf_cov_coop          phi = 1 / (T V), U_n = mean_j phi x x^H for N = 4, 8 (ilrma.py:1494-1505)
//
// kf_basis (ssb_fused.cu) gives every (mixture, source) its own CTAs, so a bin's (channel x frame) slab is pulled
// from L2 once per source: N x the compulsory traffic, and at N = 8 the kernel sits on the L2 -> SM bandwidth
// (profiles/r1_bench_configs_e.jsonl: 1.96 ms against a 0.33 ms HBM floor).  Here one CTA owns BT tiles of 16 bins
// for ALL sources: warp (bt, n) updates source n of tile bt, the N warps of a tile share its X slab through a
// 3-stage cp.async ring in shared memory (each warp fetches "its" channel, a named barrier per tile hands the stage
// over) and every warp streams its own activation tile.  The small operands are pre-split once per update into bf16
// (hi, lo) in a [frame][basis] / [bin][basis] layout; both MMA operand orientations come out of that one layout with
// ldmatrix / ldmatrix.trans, so the hot loops have 4 LDSM instead of 16 LDS and no split arithmetic for T / V.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "ssb_fused.h"
#include "ssb_kernels.h"

namespace {

constexpr int PADH = 8;   // bf16 padding of a [frame][basis] row: 48-byte (KP=16) / 80-byte (KP=32) rows are
                          // conflict-free for ldmatrix
constexpr int JCV = 32;   // frames per staged V chunk (two 16-frame steps)
constexpr int XST = 3;    // stages of the shared X ring

template <int N>
struct CoopShape {
  static constexpr int BT = (8 / N) > 0 ? 8 / N : 1;  // 16-bin tiles per CTA
  static constexpr int NW = BT * N;                   // warps per CTA
};

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
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
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
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// the same with an L2 eviction policy (createpolicy): used for data that is read for the last time
__device__ __forceinline__ void cp_async16_hint(uint32_t dst, const void* src, uint64_t policy) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "l"(policy)
               : "memory");
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int NPEND>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(NPEND) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr)
               : "memory");
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr)
               : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
// keep a loop-invariant per-lane offset in a register (ptxas otherwise rematerialises it from %tid inside the loop)
__device__ __forceinline__ uint32_t pin(uint32_t v) {
  asm volatile("" : "+r"(v));
  return v;
}
template <typename T>
__device__ __forceinline__ T* pin_ptr(T* p) {
  asm volatile("" : "+l"(p));
  return p;
}
// named barrier of tile bt (ids 1..4 as immediates, so that ptxas reserves only the barriers in use)
template <int NT>
__device__ __forceinline__ void bar_sync_tile(int bt) {
  switch (bt) {
    case 0: asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); break;
    case 1: asm volatile("bar.sync 2, %0;" ::"n"(NT) : "memory"); break;
    case 2: asm volatile("bar.sync 3, %0;" ::"n"(NT) : "memory"); break;
    default: asm volatile("bar.sync 4, %0;" ::"n"(NT) : "memory"); break;
  }
}

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

// ------------------------------------------------------------------------------------------------
// kf_vsplit: V[bn, K, J] f32 -> Vs[bn][chunk][hi | lo][JCV frames][JKS] bf16 (zero beyond K / J).
// hi = bf16(rn), lo = bf16(rn) of the exact residual, as split2.
template <int KS>
__global__ void __launch_bounds__(128) kf_vsplit(const float* __restrict__ V, __nv_bfloat16* __restrict__ Vs, int J,
                                                 int K, int nchunk) {
  constexpr int KP = 16 * KS, JKS = KP + PADH, KG = KP / 4;
  constexpr int CHH = 2 * JCV * JKS;  // halfs per chunk
  __shared__ __align__(16) __nv_bfloat16 tile[CHH];
  const int chunk = blockIdx.x;
  const size_t bn = blockIdx.y;
  const int fr = threadIdx.x & 31, kg = threadIdx.x >> 5;
  const int j = chunk * JCV + fr;
  float v[KG];
#pragma unroll
  for (int e = 0; e < KG; ++e) {
    const int k = kg * KG + e;
    v[e] = (k < K && j < J) ? __ldg(V + (bn * K + k) * J + j) : 0.f;
  }
  __nv_bfloat16* hi = tile + fr * JKS + kg * KG;
  __nv_bfloat16* lo = hi + JCV * JKS;
#pragma unroll
  for (int e = 0; e < KG; e += 2) {
    const Split s = split2(v[e], v[e + 1]);
    *reinterpret_cast<uint32_t*>(hi + e) = s.hi;
    *reinterpret_cast<uint32_t*>(lo + e) = s.lo;
  }
  if (kg == 0) {  // the padding columns travel with the cp.async copies: keep them defined
#pragma unroll
    for (int e = 0; e < PADH; e += 2) {
      *reinterpret_cast<uint32_t*>(tile + fr * JKS + KP + e) = 0u;
      *reinterpret_cast<uint32_t*>(tile + JCV * JKS + fr * JKS + KP + e) = 0u;
    }
  }
  __syncthreads();
  uint4* dst = reinterpret_cast<uint4*>(Vs + (bn * nchunk + chunk) * (size_t)CHH);
  const uint4* src = reinterpret_cast<const uint4*>(tile);
  for (int c = threadIdx.x; c < CHH / 8; c += 128) dst[c] = src[c];
}

// ------------------------------------------------------------------------------------------------
// Fragment conventions (PTX m16n8k16): g = lane/4, t = lane%4;
//   C: c0,c1 = (row g, cols 2t,2t+1), c2,c3 = (row g+8, same cols)
//   A: a0 = (row g, k 2t..), a1 = (row g+8, k 2t..), a2 = (row g, k 2t+8..), a3 = (row g+8, k 2t+8..)
//   B: b0 = (k 2t.., n g), b1 = (k 2t+8.., n g)
// X stage layout: [tile][channel][16 rows][128 bytes = 16 frames], the two 64-byte halves of odd rows swapped so
// that the LDS.128 of a quarter warp (2 rows x 4 frame pairs) covers all 32 banks once.
template <int N, int KS>
__global__ void __launch_bounds__(CoopShape<N>::NW * 32)
    kf_basis_coop(const cf* __restrict__ X, const cf* __restrict__ W, float* __restrict__ T,
                  const __nv_bfloat16* __restrict__ Vs, float* __restrict__ Pout, __nv_bfloat16* __restrict__ Ts, int I,
                  int J, int K, int nchunk, int nchunk_i, int flooring, float eps) {
  constexpr int BT = CoopShape<N>::BT;
  constexpr int KP = 16 * KS, JKS = KP + PADH;
  constexpr int CHB = 2 * JCV * JKS * 2;  // bytes of one V chunk (hi + lo)
  constexpr int XTB = N * 2048;           // bytes of one tile's X stage
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t xs_s = (uint32_t)__cvta_generic_to_shared(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int bt = warp / N, n = warp - bt * N;
  const int b = blockIdx.y;
  const uint32_t vs_s = xs_s + XST * BT * XTB + warp * 2 * CHB;
  const int i0 = (blockIdx.x * BT + bt) * 16;
  if (i0 >= I) return;  // the N warps of a tile leave together; barriers are per tile
  const int row[2] = {i0 + g, i0 + g + 8};
  const bool rvalid[2] = {row[0] < I, row[1] < I};
  const int rowc[2] = {min(row[0], I - 1), min(row[1], I - 1)};
  const size_t bn = (size_t)b * N + n;

  uint32_t Thi[KS][4], Tlo[KS][4];
  float Told[KS][2][2][2];  // [ks][nb][rr][e]: basis ks*16 + nb*8 + 2t + e, row rr
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const float* tr = T + (bn * I + rowc[rr]) * K;
#pragma unroll
      for (int nb = 0; nb < 2; ++nb) {
        const int k0 = ks * 16 + nb * 8 + 2 * t;
        const float v0 = (k0 < K) ? tr[k0] : 0.f;
        const float v1 = (k0 + 1 < K) ? tr[k0 + 1] : 0.f;
        Told[ks][nb][rr][0] = v0;
        Told[ks][nb][rr][1] = v1;
        const Split s = split2(v0, v1);
        Thi[ks][nb * 2 + rr] = s.hi;
        Tlo[ks][nb * 2 + rr] = s.lo;
      }
    }
  cf w[2][N];
#pragma unroll
  for (int rr = 0; rr < 2; ++rr)
#pragma unroll
    for (int m = 0; m < N; ++m) w[rr][m] = W[(((size_t)b * I + rowc[rr]) * N + n) * N + m];

  float num[2 * KS][4], den[2 * KS][4];
#pragma unroll
  for (int q = 0; q < 2 * KS; ++q)
#pragma unroll
    for (int c = 0; c < 4; ++c) num[q][c] = den[q][c] = 0.f;

  // this warp fetches channel n of its tile: 16 rows x 128 bytes, 8 lanes per row (coalesced 128-byte segments)
  const cf* xsrc[4];
  uint32_t xdst[4];
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int c = it * 32 + lane, r = c >> 3, ch = c & 7;
    xsrc[it] = X + (bn * I + min(i0 + r, I - 1)) * (size_t)J + 2 * ch;
    xdst[it] = pin(xs_s + bt * XTB + (n * 16 + r) * 128 + ((ch ^ ((r & 1) << 2)) << 4));
  }
  // stages are requested in step order: the source pointers simply advance
  auto issue_x = [&](uint32_t slot_bytes) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      cp_async16(xdst[it] + slot_bytes, xsrc[it]);
      xsrc[it] += 16;
    }
  };
  const unsigned char* vsrc = reinterpret_cast<const unsigned char*>(Vs) + bn * (size_t)nchunk * CHB;
  const uint32_t vdst = pin(vs_s + lane * 16);
  vsrc += lane * 16;
  // chunks are requested in order as well
  auto issue_v = [&](int buf) {
#pragma unroll
    for (int c = 0; c < CHB / 16; c += 32) cp_async16(vdst + buf * CHB + c * 16, vsrc + c * 16);
    vsrc += CHB;
  };
  // per-lane shared-memory offsets of the X fragments and of the ldmatrix rows
  // (row g + 8rr, frames 8h + 2t..): the swizzle only flips the 64-byte half, so h selects between two lane bases
  const uint32_t xlane[2] = {pin(xs_s + bt * XTB + g * 128 + ((t ^ ((g & 1) << 2)) << 4)),
                             pin(xs_s + bt * XTB + g * 128 + (((4 + t) ^ ((g & 1) << 2)) << 4))};
  const int mid = lane >> 3, mrow = lane & 7;
  // GEMM1 (non-trans): matrices (hi k0-7, hi k8-15, lo k0-7, lo k8-15) of frames [.., +8)
  const uint32_t l1base = pin(vs_s + (mid >> 1) * (JCV * JKS * 2) + (mrow * JKS + (mid & 1) * 8) * 2);
  // GEMM2 (trans): matrices (hi frames 0-7, hi frames 8-15, lo 0-7, lo 8-15) of basis [.., +8)
  const uint32_t l2base = pin(vs_s + (mid >> 1) * (JCV * JKS * 2) + (((mid & 1) * 8 + mrow) * JKS) * 2);

  const int nsteps = J >> 4;
  static_assert(XST == 3, "the V chunk of steps 2c, 2c+1 is requested at step 2c - 2: XST - 1 must be 2");
  issue_v(0);
  issue_x(0);
  cp_async_commit();
  if (nsteps > 1) issue_x(BT * XTB);
  cp_async_commit();

  // rows past the last bin (last tile only) work on a clamped copy of row I-1: their accumulator rows are separate
  // and never stored, so the loop needs no masking beyond the P store
  // P is handed to kf_activation_coop in 16 x 16 tiles ([bn][bin tile][frame tile][16 bins][16 frames], 1 KB each):
  // a warp writes, and later reads, one contiguous kilobyte per step, and the eight warps of an activation CTA read
  // 8 KB contiguous, instead of sixteen 64-byte pieces 4 J bytes apart (DRAM page locality).  Rows past the last
  // bin fall into the padding of the last tile.
  const size_t ptile0 = ((bn * (size_t)((I + 15) >> 4) + (size_t)(i0 >> 4)) * (size_t)(J >> 4)) * 256;
  float* pout[2] = {pin_ptr(Pout + ptile0 + g * 16 + 2 * t), pin_ptr(Pout + ptile0 + (g + 8) * 16 + 2 * t)};
  uint32_t rd_slot = 0, wr_slot = 2 * (BT * XTB);  // byte offsets of the X stage read / refilled this step
  for (int s = 0; s < nsteps; ++s) {
    cp_async_wait<XST - 2>();
    bar_sync_tile<N * 32>(bt);
    if (s + 2 < nsteps) issue_x(wr_slot);
    if ((s & 1) == 0 && (s >> 1) + 1 < nchunk) issue_v(((s >> 1) + 1) & 1);
    cp_async_commit();

    const uint32_t voff = ((s >> 1) & 1) * CHB + (s & 1) * (16 * JKS * 2);
    const uint32_t vb1 = l1base + voff, vb2 = l2base + voff;
    const uint32_t xb0 = xlane[0] + rd_slot, xb1 = xlane[1] + rd_slot;
    rd_slot = rd_slot + BT * XTB == XST * BT * XTB ? 0 : rd_slot + BT * XTB;
    wr_slot = wr_slot + BT * XTB == XST * BT * XTB ? 0 : wr_slot + BT * XTB;
    // ---- GEMM1: R[16 bins x 16 frames] = T V ---------------------------------------------------------
    float R[2][4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int c = 0; c < 4; ++c) R[h][c] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t bh0, bh1, bl0, bl1;
        ldsm_x4(bh0, bh1, bl0, bl1, vb1 + (8 * h * JKS + ks * 16) * 2);
        mma_split(R[h], Thi[ks], Tlo[ks], bh0, bh1, bl0, bl1);
      }
    }
    // ---- elementwise: P = |w^H x|^2, A = P / R^2, B = 1 / R ------------------------------------------
    uint32_t Ahi[4], Alo[4], Bhi[4], Blo[4];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        float4 x[N];
#pragma unroll
        for (int m = 0; m < N; ++m) x[m] = lds128((h ? xb1 : xb0) + m * 2048 + rr * 1024);
        float p0, p1;
        power2<N>(x, w[rr], p0, p1);
        // the power spectrogram is kept for the activation update (same W => same P, ilrma.py:1169-1172)
        *reinterpret_cast<float2*>(pout[rr] + 8 * h) = make_float2(p0, p1);
        const float i0v = fast_rcp(R[h][rr * 2 + 0]);
        const float i1v = fast_rcp(R[h][rr * 2 + 1]);
        const Split sa = split2(p0 * i0v * i0v, p1 * i1v * i1v);
        const Split sb = split2(i0v, i1v);
        Ahi[h * 2 + rr] = sa.hi;
        Alo[h * 2 + rr] = sa.lo;
        Bhi[h * 2 + rr] = sb.hi;
        Blo[h * 2 + rr] = sb.lo;
      }
    pout[0] += 256;
    pout[1] += 256;
    // ---- GEMM2: num += A V^T, den += B V^T  (contraction over the 16 frames) ---------------------------
#pragma unroll
    for (int q = 0; q < 2 * KS; ++q) {
      uint32_t vh0, vh1, vl0, vl1;
      ldsm_x4_t(vh0, vh1, vl0, vl1, vb2 + q * 16);
      mma_split(num[q], Ahi, Alo, vh0, vh1, vl0, vl1);
      mma_split(den[q], Bhi, Blo, vh0, vh1, vl0, vl1);
    }
  }
  // ---- T <- floor(T * sqrt(num / den))      (ilrma.py:1125-1126, p = 2) --------------------------------
  // The new basis is also written pre-split (bf16 hi, lo; [bin][basis] chunks of 32 bins) for the activation
  // update of the same call (kf_activation_coop); rows >= I and columns >= K of Ts stay zero.
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
#pragma unroll
    for (int nb = 0; nb < 2; ++nb)
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        if (!rvalid[rr]) continue;
        const int q = ks * 2 + nb;
        const int k0 = ks * 16 + nb * 8 + 2 * t;
        float tn[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          tn[e] = 0.f;
          if (k0 + e < K) {
            const float ratio = num[q][rr * 2 + e] / den[q][rr * 2 + e];
            tn[e] = ssb_floor(sqrtf(ratio) * Told[ks][nb][rr][e], flooring, eps);
            T[(bn * I + row[rr]) * K + k0 + e] = tn[e];
          }
        }
        const Split sp = split2(tn[0], tn[1]);
        __nv_bfloat16* th = Ts + (bn * nchunk_i + (row[rr] >> 5)) * (size_t)(2 * JCV * JKS) + (row[rr] & 31) * JKS + k0;
        *reinterpret_cast<uint32_t*>(th) = sp.hi;
        *reinterpret_cast<uint32_t*>(th + JCV * JKS) = sp.lo;
      }
}

// ------------------------------------------------------------------------------------------------
// kf_activation_coop:  V <- V sqrt( sum_i T P / R^2  /  sum_i T / R ),  R = T V with the new T
// (ssspy/bss/ilrma.py:1130-1204).  CTA = (group of AW*16 frames, source, mixture); warp = 16 frames and owns its V
// entries completely (no partial sums, no atomics: deterministic).  All warps walk over the bins together: the
// pre-split basis chunks Ts (32 bins, written by kf_basis_coop) go through a CTA-wide double buffer, every warp
// streams its own 16-bin x 16-frame tiles of P (written by the basis kernel with the same W) through a private
// 3-stage cp.async ring.  Orientation is transposed w.r.t. the basis kernel (C rows = frames, C cols = bins), so the
// accumulator fragment of R^T = V^T T^T is the A operand of num^T += (P / R^2)^T T; both B operands come from the one
// [bin][basis] layout by ldmatrix / ldmatrix.trans.  The new V is also written pre-split (Vs) for kf_basis_coop /
// the covariance kernel of the following phases.
constexpr int AW = 8;     // warps per CTA
constexpr int PST = 6;    // stages of the per-warp P ring (PST - 1 tiles in flight per warp: the kernel is bound by
                          // the HBM latency per warp, not by issue)
constexpr int TD = PST / 2;   // T chunks issued ahead of use: a chunk must sit in a cp.async group no younger than
                             // the P tile of its first step
constexpr int TST = TD + 1;  // slots of the CTA-wide T chunk ring
constexpr int PRS = 80;   // bytes per P tile row: 16 frames fp32 + 16 pad => conflict-free fragment reads

__device__ __forceinline__ float lds32f(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

template <int KS, int ACW, int ACP, int MINB>
__global__ void __launch_bounds__(ACW * 32, MINB)
    kf_activation_coop(const float* __restrict__ P, const __nv_bfloat16* __restrict__ Ts, float* __restrict__ V,
                       __nv_bfloat16* __restrict__ Vs, int NS, int I, int J, int K, int nchunk_i, int nchunk_j,
                       int flooring, float eps) {
  constexpr int KP = 16 * KS, JKS = KP + PADH;
  constexpr int PST = ACP, TD = ACP / 2, TST = TD + 1, AW = ACW;  // shadow the defaults of the file scope
  constexpr int CHB = 2 * JCV * JKS * 2;  // bytes of one 32-bin T chunk (hi + lo)
  constexpr int PTB = 16 * PRS;           // bytes of one P tile stage
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t ts_s = (uint32_t)__cvta_generic_to_shared(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int n = blockIdx.y, b = blockIdx.z;
  const uint32_t pw_s = ts_s + TST * CHB + warp * (PST * PTB);
  const int j0 = (blockIdx.x * AW + warp) * 16;
  const bool warp_active = j0 < J;
  const size_t bn = (size_t)b * NS + n;
  float* Vb = V + bn * K * J;
  const float* Pb = P + bn * (size_t)((I + 15) >> 4) * (size_t)(J >> 4) * 256;  // tiled, see kf_basis_coop
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
        const Split sp = split2(v0, v1);
        Vhi[ks][nb * 2 + rr] = sp.hi;
        Vlo[ks][nb * 2 + rr] = sp.lo;
      }
  float num[2 * KS][4], den[2 * KS][4];
#pragma unroll
  for (int q = 0; q < 2 * KS; ++q)
#pragma unroll
    for (int c = 0; c < 4; ++c) num[q][c] = den[q][c] = 0.f;

  const unsigned char* tsrc = reinterpret_cast<const unsigned char*>(Ts) + bn * (size_t)nchunk_i * CHB;
  auto issue_t = [&](int chunk, int buf) {
    for (int c = threadIdx.x; c < CHB / 16; c += AW * 32)
      cp_async16(ts_s + buf * CHB + c * 16, tsrc + (size_t)chunk * CHB + c * 16);
  };
  // P tile of step s: 16 bins x 64 bytes, 4 lanes per row.  The P scratch is padded by 16 rows, so the (masked)
  // rows past the last bin need no clamping.
  const float* psrc[2];
  uint32_t pdst[2];
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int idx = it * 32 + lane, r = idx >> 2, part = idx & 3;
    psrc[it] = Pb + (size_t)(min(j0, J - 16) >> 4) * 256 + r * 16 + part * 4;
    pdst[it] = pin(pw_s + r * PRS + part * 16);
  }
  const size_t pstep = (size_t)(J >> 4) * 256;  // next bin tile of the same frame tile
  // tiles are requested in step order: the source pointers simply advance
  auto issue_p = [&](uint32_t slot_bytes) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      cp_async16(pdst[it] + slot_bytes, psrc[it]);
      psrc[it] += pstep;
    }
  };
  const int mid = lane >> 3, mrow = lane & 7;
  const uint32_t l1off = (mid >> 1) * (JCV * JKS * 2) + (mrow * JKS + (mid & 1) * 8) * 2;
  const uint32_t l2off = (mid >> 1) * (JCV * JKS * 2) + (((mid & 1) * 8 + mrow) * JKS) * 2;
  // fragment reads of P: (bin 8h + 2t + e, frame g + 8rr) = per-lane base + compile-time offset
  const uint32_t plane = pin(pw_s + (2 * t) * PRS + g * 4);
  const uint32_t l1base = pin(ts_s + l1off), l2base = pin(ts_s + l2off);
  const int nsteps = (I + 15) >> 4;
  const bool frames_full = fvalid[0] && fvalid[1];
  // group s carries the P tile of step s (+ the T chunk needed PST - 1 steps later); PST - 1 groups are in flight
#pragma unroll
  for (int q = 0; q < PST - 1; ++q) {
    if ((q & 1) == 0 && (q >> 1) < nchunk_i) issue_t(q >> 1, (q >> 1) % TST);
    if (warp_active && q < nsteps) issue_p(q * PTB);
    cp_async_commit();
  }
  uint32_t rd_slot = 0, wr_slot = (PST - 1) * PTB;  // byte offsets of the P ring slots read / refilled this step
  uint32_t t_slot = 0;                               // byte offset of the T chunk in use
  for (int s = 0; s < nsteps; ++s) {
    cp_async_wait<PST - 2>();
    if ((s & 1) == 0) {
      __syncthreads();  // chunk s/2 has landed for every thread; chunk s/2 - 1 is no longer read
      const int cn = (s >> 1) + TD;
      if (cn < nchunk_i) issue_t(cn, cn % TST);
    } else {
      __syncwarp();
    }
    if (warp_active && s + PST - 1 < nsteps) issue_p(wr_slot);
    cp_async_commit();
    const uint32_t tb1 = l1base + t_slot + (s & 1) * (16 * JKS * 2), tb2 = l2base + t_slot + (s & 1) * (16 * JKS * 2);
    const uint32_t pb = plane + rd_slot;
    rd_slot = rd_slot + PTB == PST * PTB ? 0 : rd_slot + PTB;
    wr_slot = wr_slot + PTB == PST * PTB ? 0 : wr_slot + PTB;
    if (s & 1) t_slot = t_slot + CHB == TST * CHB ? 0 : t_slot + CHB;
    if (!warp_active) continue;

    // ---- GEMM1: R^T[16 frames x 16 bins] = V^T T^T ---------------------------------------------------
    float R[2][4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {  // h: bins 8h .. 8h+7 of the step (n-tile)
#pragma unroll
      for (int c = 0; c < 4; ++c) R[h][c] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t bh0, bh1, bl0, bl1;
        ldsm_x4(bh0, bh1, bl0, bl1, tb1 + (8 * h * JKS + ks * 16) * 2);
        mma_split(R[h], Vhi[ks], Vlo[ks], bh0, bh1, bl0, bl1);
      }
    }
    // ---- elementwise at (frame rr, bin 8h + 2t + e) -------------------------------------------------------
    uint32_t Ahi[4], Alo[4], Bhi[4], Blo[4];
    const int lim = I - s * 16;  // valid bins in this step
    if (lim >= 16 && frames_full) {
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const float iv0 = fast_rcp(R[h][rr * 2]), iv1 = fast_rcp(R[h][rr * 2 + 1]);
          const float p0 = lds32f(pb + (8 * h) * PRS + 32 * rr), p1 = lds32f(pb + (8 * h + 1) * PRS + 32 * rr);
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
            const float p = lds32f(pb + (8 * h + e) * PRS + 32 * rr);
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
    // ---- GEMM2: num^T += A^T T, den^T += B^T T  (contraction over the 16 bins) ---------------------------
#pragma unroll
    for (int q = 0; q < 2 * KS; ++q) {
      uint32_t th0, th1, tl0, tl1;
      ldsm_x4_t(th0, th1, tl0, tl1, tb2 + q * 16);
      mma_split(num[q], Ahi, Alo, th0, th1, tl0, tl1);
      mma_split(den[q], Bhi, Blo, th0, th1, tl0, tl1);
    }
  }
  if (!warp_active) return;
  // ---- V <- floor(V * sqrt(num / den))      (ilrma.py:1201-1202, p = 2), plus the pre-split copy --------------
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
#pragma unroll
    for (int nb = 0; nb < 2; ++nb)
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        if (!fvalid[rr]) continue;
        const int q = ks * 2 + nb;
        const int k0 = ks * 16 + nb * 8 + 2 * t;
        float vn[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          vn[e] = 0.f;
          if (k0 + e < K) {
            const float ratio = num[q][rr * 2 + e] / den[q][rr * 2 + e];
            vn[e] = ssb_floor(sqrtf(ratio) * Vold[ks][nb][rr][e], flooring, eps);
            Vb[(size_t)(k0 + e) * J + fr[rr]] = vn[e];
          }
        }
        const Split sp = split2(vn[0], vn[1]);
        __nv_bfloat16* vh = Vs + (bn * nchunk_j + (fr[rr] >> 5)) * (size_t)(2 * JCV * JKS) + (fr[rr] & 31) * JKS + k0;
        *reinterpret_cast<uint32_t*>(vh) = sp.hi;
        *reinterpret_cast<uint32_t*>(vh + JCV * JKS) = sp.lo;
      }
}

template <int N, int KS>
int launch_coop(const ssb_config* c, const cf* X, const cf* W, float* T, float* V, float* P, __nv_bfloat16* Vs,
                __nv_bfloat16* Ts, int vs_valid, cudaStream_t st,
                int parts = 7) {  // parts: 1 = pre-split of V, 2 = basis kernel, 4 = activation kernel
  const int B = c->n_batch, I = c->n_bins, J = c->n_frames, K = c->n_basis;
  constexpr int BT = CoopShape<N>::BT, NW = CoopShape<N>::NW;
  constexpr int KP = 16 * KS, JKS = KP + PADH;
  constexpr int CHB = 2 * JCV * JKS * 2;
  const int nchunk = (J + JCV - 1) / JCV, nchunk_i = (I + JCV - 1) / JCV;
  if (!vs_valid && (parts & 1)) {
    dim3 gv(nchunk, B * N);
    kf_vsplit<KS><<<gv, 128, 0, st>>>(V, Vs, J, K, nchunk);
    if (ssb_check_launch("coop_vsplit", st)) return 1;
  }
  const size_t sm = (size_t)XST * BT * N * 2048 + (size_t)NW * 2 * CHB;
  // activation kernel shape: warps of 16 frames per CTA, stages of the P ring, CTAs per SM.  Default 8 / 6 / 2.
  // SSB_ACT_SHAPE=1 selects 4 warps x 7 CTAs = 28 warps per SM at 72 registers, which makes config 2 (4096 warp
  // tasks on 148 SMs) a single full wave instead of 1.73 waves of 16 warps: measured 0.106 vs 0.102 ms, i.e. the
  // kernel is bound by issued instructions, not by the wave shape.
  static int act_shape = -1;
  if (act_shape < 0) {
    const char* e = getenv("SSB_ACT_SHAPE");
    act_shape = e ? atoi(e) : 0;
  }
  constexpr int AW1 = 4, AP1 = 3, AB1 = 7;
  const size_t sm_act = (size_t)TST * CHB + (size_t)AW * PST * 16 * PRS;
  const size_t sm_act1 = (size_t)(AP1 / 2 + 1) * CHB + (size_t)AW1 * AP1 * 16 * PRS;
  static bool attr_dev[SSB_MAX_DEVICES] = {};  // function attributes are per device
  bool& attr_set = attr_dev[ssb_current_device()];
  if (!attr_set) {
    SSB_CUDA(cudaFuncSetAttribute(kf_basis_coop<N, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    SSB_CUDA(cudaFuncSetAttribute(kf_activation_coop<KS, AW, PST, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_act));
    SSB_CUDA(cudaFuncSetAttribute(kf_activation_coop<KS, AW1, AP1, AB1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_act1));
    attr_set = true;
  }
  dim3 grid((I + 16 * BT - 1) / (16 * BT), B);
  if (!(parts & 2)) {
    // the basis update is done by another kernel (ssb_tma.cu)
  } else {
    kf_basis_coop<N, KS><<<grid, NW * 32, sm, st>>>(X, W, T, Vs, P, Ts, I, J, K, nchunk, nchunk_i, c->flooring, c->eps);
    if (ssb_check_launch("coop_basis", st)) return 1;
  }
  if (!(parts & 4)) return 0;
  if (act_shape == 1 && KS == 1) {  // K > 16 spills at 72 registers
    dim3 ga((J + AW1 * 16 - 1) / (AW1 * 16), N, B);
    kf_activation_coop<KS, AW1, AP1, AB1><<<ga, AW1 * 32, sm_act1, st>>>(P, Ts, V, Vs, N, I, J, K, nchunk_i, nchunk,
                                                                         c->flooring, c->eps);
  } else {
    dim3 ga((J + AW * 16 - 1) / (AW * 16), N, B);
    kf_activation_coop<KS, AW, PST, 2><<<ga, AW * 32, sm_act, st>>>(P, Ts, V, Vs, N, I, J, K, nchunk_i, nchunk,
                                                                    c->flooring, c->eps);
  }
  return ssb_check_launch("coop_activation", st);
}

}  // namespace

static size_t coop_chunk_bytes(const ssb_config* c) {
  const size_t KP = c->n_basis <= 16 ? 16 : 32;
  return 2 * JCV * (KP + PADH) * sizeof(__nv_bfloat16);
}
static size_t coop_vs_bytes(const ssb_config* c) {
  const size_t nchunk = ((size_t)c->n_frames + JCV - 1) / JCV;
  return (((size_t)c->n_batch * c->n_sources * nchunk * coop_chunk_bytes(c)) + 255) & ~(size_t)255;
}
static size_t coop_ts_bytes(const ssb_config* c) {
  const size_t nchunk = ((size_t)c->n_bins + JCV - 1) / JCV;
  return (((size_t)c->n_batch * c->n_sources * nchunk * coop_chunk_bytes(c)) + 255) & ~(size_t)255;
}

// scratch: pre-split activation Vs followed by pre-split basis Ts (must be zero-initialised once)
size_t ssb_coop_ws_bytes(const ssb_config* c) { return coop_vs_bytes(c) + coop_ts_bytes(c); }

// MM source model for every source of every mixture: [V -> Vs unless vs_valid] basis (T, Ts, P), activation (V, Vs)
int ssb_coop_source(const ssb_config* c, const cf* X, const cf* W, float* T, float* V, float* P, void* ws,
                    int vs_valid, cudaStream_t st) {
  SSB_REQUIRE((c->n_frames % 16) == 0 && c->n_basis <= 32 && W != nullptr && ws != nullptr,
              "coop_source: unsupported configuration");
  __nv_bfloat16* Vs = (__nv_bfloat16*)ws;
  __nv_bfloat16* Ts = (__nv_bfloat16*)((char*)ws + coop_vs_bytes(c));
  if (c->n_basis <= 16) {
    SSB_DISPATCH_N(c->n_sources, return (launch_coop<NN, 1>(c, X, W, T, V, P, Vs, Ts, vs_valid, st)));
  } else {
    SSB_DISPATCH_N(c->n_sources, return (launch_coop<NN, 2>(c, X, W, T, V, P, Vs, Ts, vs_valid, st)));
  }
  return 0;
}

void* ssb_coop_vs(const ssb_config*, void* ws) { return ws; }
void* ssb_coop_ts(const ssb_config* c, void* ws) { return (char*)ws + coop_vs_bytes(c); }

int ssb_coop_vsplit(const ssb_config* c, const float* V, void* ws, int vs_valid, cudaStream_t st) {
  SSB_REQUIRE((c->n_frames % 16) == 0 && c->n_basis <= 32 && ws != nullptr, "coop_vsplit: unsupported configuration");
  __nv_bfloat16* Vs = (__nv_bfloat16*)ws;
  __nv_bfloat16* Ts = (__nv_bfloat16*)((char*)ws + coop_vs_bytes(c));
  float* Vm = const_cast<float*>(V);
  if (c->n_basis <= 16) {
    SSB_DISPATCH_N(c->n_sources, return (launch_coop<NN, 1>(c, nullptr, nullptr, nullptr, Vm, nullptr, Vs, Ts, vs_valid, st,
                                                            1)));
  } else {
    SSB_DISPATCH_N(c->n_sources, return (launch_coop<NN, 2>(c, nullptr, nullptr, nullptr, Vm, nullptr, Vs, Ts, vs_valid, st,
                                                            1)));
  }
  return 0;
}

int ssb_coop_activation(const ssb_config* c, float* V, float* P, void* ws, cudaStream_t st) {
  SSB_REQUIRE((c->n_frames % 16) == 0 && c->n_basis <= 32 && ws != nullptr, "coop_activation: unsupported configuration");
  __nv_bfloat16* Vs = (__nv_bfloat16*)ws;
  __nv_bfloat16* Ts = (__nv_bfloat16*)((char*)ws + coop_vs_bytes(c));
  if (c->n_basis <= 16) {
    SSB_DISPATCH_N(c->n_sources, return (launch_coop<NN, 1>(c, nullptr, nullptr, nullptr, V, P, Vs, Ts, 1, st, 4)));
  } else {
    SSB_DISPATCH_N(c->n_sources, return (launch_coop<NN, 2>(c, nullptr, nullptr, nullptr, V, P, Vs, Ts, 1, st, 4)));
  }
  return 0;
}

// ================================================================================================
// Multiplicative updates with the two elementwise factors given as arrays (FastGaussMNMF,
// ssspy/bss/mnmf.py:1351-1358, :1408-1415):
//   OUTER_ROWS = true  (basis):       T[i,k] <- floor(T sqrt(sum_j A[i,j] V[k,j] / sum_j Bm[i,j] V[k,j]))
//   OUTER_ROWS = false (activation):  V[k,j] <- floor(V sqrt(sum_i T[i,k] A[i,j] / sum_i T[i,k] Bm[i,j]))
// i.e. the second GEMM of kf_basis_coop / kf_activation_coop with its A operand streamed from memory instead of
// computed: warp = 16 "outer" indices (bins resp. frames) of one (mixture, source), all warps of the CTA walk the
// "inner" (reduction) index together; the pre-split operand chunks (Vs resp. Ts, 32 inner indices) go through a
// CTA-wide ring, every warp streams its own 16 x 16 tiles of A and Bm through a private cp.async ring, splits them
// to bf16 (hi, lo) and feeds mma.sync.  The result is also written pre-split (Ts resp. Vs) for the next kernel.
constexpr int ABST = 4;          // stages of the per-warp tile ring
constexpr int ABTD = ABST / 2;   // operand chunks issued ahead (see TD)
constexpr int ABTS = ABTD + 1;   // slots of the operand chunk ring

__device__ __forceinline__ float2 lds64f(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
  return v;
}

template <int KS, bool OUTER_ROWS>
__global__ void __launch_bounds__(AW * 32)
    kf_update_ab(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ Out,
                 const __nv_bfloat16* __restrict__ Opd, __nv_bfloat16* __restrict__ OutSplit, int I, int J, int K,
                 int nchunk_in, int nchunk_out, int flooring, float eps) {
  constexpr int KP = 16 * KS, JKS = KP + PADH;
  constexpr int CHB = 2 * JCV * JKS * 2;          // bytes of one 32-index operand chunk (hi + lo)
  constexpr int PITCH = OUTER_ROWS ? 96 : 80;     // tile row pitch: conflict-free LDS.64 resp. LDS.32 fragment reads
  constexpr int PTB = 16 * PITCH;                 // bytes of one tile
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t op_s = (uint32_t)__cvta_generic_to_shared(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const size_t bn = blockIdx.y;
  const int n_outer = OUTER_ROWS ? I : J, n_inner = OUTER_ROWS ? J : I;
  const uint32_t pw_s = op_s + ABTS * CHB + warp * (ABST * 2 * PTB);  // [stage][A | Bm][16 rows][PITCH]
  const int o0 = (blockIdx.x * AW + warp) * 16;
  const bool warp_active = o0 < n_outer;
  const int oc[2] = {min(o0 + g, n_outer - 1), min(o0 + g + 8, n_outer - 1)};
  const bool ovalid[2] = {o0 + g < n_outer, o0 + g + 8 < n_outer};
  const float* Ab = A + bn * (size_t)I * J;
  const float* Bb = Bm + bn * (size_t)I * J;
  float* Ob = Out + bn * (size_t)K * (OUTER_ROWS ? I : J);

  // current values of the entries this thread updates: [ks][nb][rr][e] = (outer rr, basis ks*16 + nb*8 + 2t + e)
  float old[KS][2][2][2];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
#pragma unroll
    for (int nb = 0; nb < 2; ++nb)
#pragma unroll
      for (int rr = 0; rr < 2; ++rr)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int k = ks * 16 + nb * 8 + 2 * t + e;
          old[ks][nb][rr][e] = (k < K) ? (OUTER_ROWS ? Ob[(size_t)oc[rr] * K + k] : Ob[(size_t)k * J + oc[rr]]) : 0.f;
        }
  float num[2 * KS][4], den[2 * KS][4];
#pragma unroll
  for (int q = 0; q < 2 * KS; ++q)
#pragma unroll
    for (int c = 0; c < 4; ++c) num[q][c] = den[q][c] = 0.f;

  const unsigned char* osrc = reinterpret_cast<const unsigned char*>(Opd) + bn * (size_t)nchunk_in * CHB;
  auto issue_op = [&](int chunk, int buf) {
    for (int c = threadIdx.x; c < CHB / 16; c += AW * 32)
      cp_async16(op_s + buf * CHB + c * 16, osrc + (size_t)chunk * CHB + c * 16);
  };
  // tile of step s: 16 rows x 64 bytes of A and of Bm, 4 lanes per row, 2 + 2 pieces per lane.
  //   OUTER_ROWS: rows = this warp's outer indices (clamped), columns = inner indices 16 s ..
  //   else      : rows = inner indices 16 s .. (clamped to the array), columns = this warp's outer indices
  const float* tsrc[2][2];
  uint32_t tdst[2];
  int trow[2];
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int idx = it * 32 + lane, r = idx >> 2, part = idx & 3;
    trow[it] = r;
    tdst[it] = pin(pw_s + r * PITCH + part * 16);
    if (OUTER_ROWS) {
      const size_t off = (size_t)min(o0 + r, I - 1) * J + part * 4;
      tsrc[it][0] = Ab + off;
      tsrc[it][1] = Bb + off;
    } else {
      const size_t off = (size_t)min(o0, J - 16) + part * 4;
      tsrc[it][0] = Ab + off;
      tsrc[it][1] = Bb + off;
    }
  }
  auto issue_tile = [&](int step, uint32_t slot_bytes) {
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const size_t adv = OUTER_ROWS ? (size_t)step * 16 : (size_t)min(step * 16 + trow[it], I - 1) * J;
      cp_async16(tdst[it] + slot_bytes, tsrc[it][0] + adv);
      cp_async16(tdst[it] + slot_bytes + PTB, tsrc[it][1] + adv);
    }
  };
  const int mid = lane >> 3, mrow = lane & 7;
  const uint32_t l2base = pin(op_s + (mid >> 1) * (JCV * JKS * 2) + (((mid & 1) * 8 + mrow) * JKS) * 2);
  const uint32_t flane = OUTER_ROWS ? pin(pw_s + g * PITCH + (2 * t) * 4) : pin(pw_s + (2 * t) * PITCH + g * 4);
  const int nsteps = (n_inner + 15) >> 4;
#pragma unroll
  for (int q = 0; q < ABST - 1; ++q) {
    if ((q & 1) == 0 && (q >> 1) < nchunk_in) issue_op(q >> 1, (q >> 1) % ABTS);
    if (warp_active && q < nsteps) issue_tile(q, q * (2 * PTB));
    cp_async_commit();
  }
  uint32_t rd_slot = 0, wr_slot = (ABST - 1) * (2 * PTB), o_slot = 0;
  for (int s = 0; s < nsteps; ++s) {
    cp_async_wait<ABST - 2>();
    if ((s & 1) == 0) {
      __syncthreads();
      const int cn = (s >> 1) + ABTD;
      if (cn < nchunk_in) issue_op(cn, cn % ABTS);
    } else {
      __syncwarp();
    }
    if (warp_active && s + ABST - 1 < nsteps) issue_tile(s + ABST - 1, wr_slot);
    cp_async_commit();
    const uint32_t ob = l2base + o_slot + (s & 1) * (16 * JKS * 2);
    const uint32_t fb = flane + rd_slot;
    rd_slot = rd_slot + 2 * PTB == ABST * 2 * PTB ? 0 : rd_slot + 2 * PTB;
    wr_slot = wr_slot + 2 * PTB == ABST * 2 * PTB ? 0 : wr_slot + 2 * PTB;
    if (s & 1) o_slot = o_slot + CHB == ABTS * CHB ? 0 : o_slot + CHB;
    if (!warp_active) continue;
    // A fragments: a0 = (outer g, inner 2t..), a1 = (outer g+8, inner 2t..), a2 = (outer g, inner 2t+8..), a3
    uint32_t Ahi[4], Alo[4], Bhi[4], Blo[4];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        float2 av, bv;
        if (OUTER_ROWS) {
          av = lds64f(fb + (8 * rr) * PITCH + (8 * h) * 4);
          bv = lds64f(fb + PTB + (8 * rr) * PITCH + (8 * h) * 4);
        } else {
          av = make_float2(lds32f(fb + (8 * h) * PITCH + 32 * rr), lds32f(fb + (8 * h + 1) * PITCH + 32 * rr));
          bv = make_float2(lds32f(fb + PTB + (8 * h) * PITCH + 32 * rr), lds32f(fb + PTB + (8 * h + 1) * PITCH + 32 * rr));
        }
        const Split sa = split2(av.x, av.y);
        const Split sb = split2(bv.x, bv.y);
        Ahi[h * 2 + rr] = sa.hi;
        Alo[h * 2 + rr] = sa.lo;
        Bhi[h * 2 + rr] = sb.hi;
        Blo[h * 2 + rr] = sb.lo;
      }
#pragma unroll
    for (int q = 0; q < 2 * KS; ++q) {
      uint32_t th0, th1, tl0, tl1;
      ldsm_x4_t(th0, th1, tl0, tl1, ob + q * 16);
      mma_split(num[q], Ahi, Alo, th0, th1, tl0, tl1);
      mma_split(den[q], Bhi, Blo, th0, th1, tl0, tl1);
    }
  }
  if (!warp_active) return;
#pragma unroll
  for (int ks = 0; ks < KS; ++ks)
#pragma unroll
    for (int nb = 0; nb < 2; ++nb)
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        if (!ovalid[rr]) continue;
        const int q = ks * 2 + nb;
        const int k0 = ks * 16 + nb * 8 + 2 * t;
        float vn[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          vn[e] = 0.f;
          if (k0 + e < K) {
            vn[e] = ssb_floor(old[ks][nb][rr][e] * sqrtf(num[q][rr * 2 + e] / den[q][rr * 2 + e]), flooring, eps);
            if (OUTER_ROWS) Ob[(size_t)oc[rr] * K + k0 + e] = vn[e];
            else Ob[(size_t)(k0 + e) * J + oc[rr]] = vn[e];
          }
        }
        const Split sp = split2(vn[0], vn[1]);
        __nv_bfloat16* oh = OutSplit + (bn * nchunk_out + (oc[rr] >> 5)) * (size_t)(2 * JCV * JKS) + (oc[rr] & 31) * JKS + k0;
        *reinterpret_cast<uint32_t*>(oh) = sp.hi;
        *reinterpret_cast<uint32_t*>(oh + JCV * JKS) = sp.lo;
      }
}

template <int KS>
int launch_update_ab(int which, const float* A, const float* Bm, float* T, float* V, __nv_bfloat16* Vs,
                     __nv_bfloat16* Ts, int BN, int I, int J, int K, int flooring, float eps, cudaStream_t st) {
  constexpr int KP = 16 * KS, JKS = KP + PADH, CHB = 2 * JCV * JKS * 2;
  const int nchunk_j = (J + JCV - 1) / JCV, nchunk_i = (I + JCV - 1) / JCV;
  const size_t sm_b = (size_t)ABTS * CHB + (size_t)AW * ABST * 2 * 16 * 96;
  const size_t sm_a = (size_t)ABTS * CHB + (size_t)AW * ABST * 2 * 16 * 80;
  static bool attr_dev[SSB_MAX_DEVICES] = {};  // function attributes are per device
  bool& attr_set = attr_dev[ssb_current_device()];
  if (!attr_set) {
    SSB_CUDA(cudaFuncSetAttribute(kf_update_ab<KS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_b));
    SSB_CUDA(cudaFuncSetAttribute(kf_update_ab<KS, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_a));
    attr_set = true;
  }
  if (which == 0) {
    dim3 gv(nchunk_j, BN);
    kf_vsplit<KS><<<gv, 128, 0, st>>>(V, Vs, J, K, nchunk_j);
    if (ssb_check_launch("coop_vsplit", st)) return 1;
    dim3 grid((I + AW * 16 - 1) / (AW * 16), BN);
    kf_update_ab<KS, true><<<grid, AW * 32, sm_b, st>>>(A, Bm, T, Vs, Ts, I, J, K, nchunk_j, nchunk_i, flooring, eps);
    return ssb_check_launch("coop_basis_ab", st);
  }
  dim3 grid((J + AW * 16 - 1) / (AW * 16), BN);
  kf_update_ab<KS, false><<<grid, AW * 32, sm_a, st>>>(A, Bm, V, Ts, Vs, I, J, K, nchunk_i, nchunk_j, flooring, eps);
  return ssb_check_launch("coop_activation_ab", st);
}

// ================================================================================================
// kf_mnmf_update<OUTER_ROWS>: the FastGaussMNMF multiplicative updates (ssspy/bss/mnmf.py:1344-1415) for four sources
// with their elementwise factors formed IN the kernel instead of read from G / H arrays (km_gh wrote 4.3 GB and read
// 6.6 GB per update, and Lambda = T V made another 2 GB round trip):
//   Lambda_n = T_n V_n                                   GEMM1 on the tensor pipe (as kf_basis_coop)
//   L_m = sum_n Lambda_n D[i,n,m], r_m = 1 / L_m, G_n = sum_m D[i,n,m] Z2_m r_m^2, H_n = sum_m D[i,n,m] r_m   per point
//   num_n += G_n Op_n^T, den_n += H_n Op_n^T             GEMM2 (as kf_update_ab), Op = pre-split V resp. T chunks
// Z2[b,m,i,j] = |q_m^H x|^2 is the only array streamed (written once per iteration by km_z2).  Warp = 16 outer indices
// (bins resp. frames) of one mixture x ALL four sources: the accumulator fragments of GEMM1 sit at the same (outer,
// inner) positions for every source, so the 4 x 4 mixing by D is thread-local.  K <= 16 (one k-step; 64 accumulator
// registers for the four sources).  OUTER_ROWS = true: basis update (outer = bins, inner = frames, D of the thread's
// two bins in registers); false: activation update (outer = frames, inner = bins, D of the 32 bins of a chunk staged in
// shared memory next to the operand chunks).
constexpr int MUW = 4;   // warps per CTA
constexpr int MUS = 3;   // slots of the operand chunk ring

template <bool OUTER_ROWS>
__global__ void __launch_bounds__(MUW * 32)
    kf_mnmf_update(const float* __restrict__ Z2, const float* __restrict__ Dm, float* Out,
                   const float* OuterSrc /* == Out: read at the start, written at the end */,
                   const __nv_bfloat16* __restrict__ Opd,
                   __nv_bfloat16* __restrict__ OutSplit, int I, int J, int K, int nchunk_in, int nchunk_out,
                   int flooring, float eps, const float* __restrict__ zscale) {
  constexpr int NS = 4, KP = 16, JKS = KP + PADH;
  constexpr int CHB = 2 * JCV * JKS * 2;   // bytes of one 32-index operand chunk (hi + lo) of one source
  constexpr int DCB = JCV * 16 * 4;        // bytes of the D block of one chunk (32 bins x 16 floats), activation only
  constexpr int SLOT = NS * CHB + (OUTER_ROWS ? 0 : DCB);
  constexpr int PITCH = OUTER_ROWS ? 96 : 80;  // Z2 tile row pitch: conflict-free LDS.64 resp. LDS.32 fragment reads
  constexpr int PTB = 16 * PITCH;              // bytes of one 16 x 16 tile
  constexpr int ZST = 2;                       // stages of the per-warp Z2 tile ring (four tiles, one per m, per stage)
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t op_s = (uint32_t)__cvta_generic_to_shared(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const size_t b = blockIdx.y;
  const int n_outer = OUTER_ROWS ? I : J, n_inner = OUTER_ROWS ? J : I;
  const int o0 = (blockIdx.x * MUW + warp) * 16;
  const bool warp_active = o0 < n_outer;
  const int oc[2] = {min(o0 + g, n_outer - 1), min(o0 + g + 8, n_outer - 1)};
  const bool ovalid[2] = {o0 + g < n_outer, o0 + g + 8 < n_outer};
  // outer-side A operand of GEMM1, per source: (outer row rr, basis nb * 8 + 2t, + 1), kept in shared memory
  // ([source][hi | lo][4][lane]: conflict-free 32-bit accesses) to leave the registers to the accumulators
  uint32_t* ofr = reinterpret_cast<uint32_t*>(smem_raw + MUS * SLOT) + warp * (NS * 2 * 4 * 32) + lane;
  // per-warp ring of Z2 tiles [stage][m][16 rows][PITCH]: cp.async one step ahead, so the loads of step s + 1 are in
  // flight while step s is computed (with direct global loads the kernel stalled on them: 40 % issue utilisation)
  const uint32_t zr_s = op_s + MUS * SLOT + MUW * (NS * 2 * 4 * 32 * 4) + warp * (ZST * NS * PTB);
#pragma unroll
  for (int n = 0; n < NS; ++n)
#pragma unroll
    for (int rr = 0; rr < 2; ++rr)
#pragma unroll
      for (int nb = 0; nb < 2; ++nb) {
        const int k0 = nb * 8 + 2 * t;
        float v0 = 0.f, v1 = 0.f;
        if (OUTER_ROWS) {  // T[b, n, bin, k]
          const float* tr = OuterSrc + ((b * NS + n) * I + oc[rr]) * (size_t)K;
          v0 = k0 < K ? tr[k0] : 0.f;
          v1 = k0 + 1 < K ? tr[k0 + 1] : 0.f;
        } else {           // V[b, n, k, frame]
          const float* vc = OuterSrc + (b * NS + n) * (size_t)K * J + oc[rr];
          v0 = k0 < K ? vc[(size_t)k0 * J] : 0.f;
          v1 = k0 + 1 < K ? vc[(size_t)(k0 + 1) * J] : 0.f;
        }
        const Split sp = split2(v0, v1);
        ofr[((n * 2 + 0) * 4 + nb * 2 + rr) * 32] = sp.hi;
        ofr[((n * 2 + 1) * 4 + nb * 2 + rr) * 32] = sp.lo;
      }
  // D of the thread's two bins (basis update): dreg[rr][n * 4 + m]
  float dreg[OUTER_ROWS ? 2 : 1][16];
  if (OUTER_ROWS) {
#pragma unroll
    for (int rr = 0; rr < 2; ++rr)
#pragma unroll
      for (int e = 0; e < 16; e += 4) {
        const float4 d4 = *reinterpret_cast<const float4*>(Dm + ((b * I + oc[rr]) * 16 + e));
        dreg[rr][e] = d4.x;
        dreg[rr][e + 1] = d4.y;
        dreg[rr][e + 2] = d4.z;
        dreg[rr][e + 3] = d4.w;
      }
  }
  float num[NS][2][4], den[NS][2][4];
#pragma unroll
  for (int n = 0; n < NS; ++n)
#pragma unroll
    for (int q = 0; q < 2; ++q)
#pragma unroll
      for (int c = 0; c < 4; ++c) num[n][q][c] = den[n][q][c] = 0.f;

  // operand chunks of the four sources (+ the D block of the chunk's 32 bins) -> ring slot
  const unsigned char* osrc = reinterpret_cast<const unsigned char*>(Opd);
  auto issue_op = [&](int chunk, int slot) {
    for (int c = threadIdx.x; c < NS * (CHB / 16); c += MUW * 32) {
      const int n = c / (CHB / 16), w = c - n * (CHB / 16);
      cp_async16(op_s + slot * SLOT + n * CHB + w * 16, osrc + ((b * NS + n) * (size_t)nchunk_in + chunk) * CHB + w * 16);
    }
    if (!OUTER_ROWS) {
      for (int c = threadIdx.x; c < DCB / 16; c += MUW * 32) {
        const int bin = min(chunk * JCV + (c >> 2), I - 1);
        cp_async16(op_s + slot * SLOT + NS * CHB + c * 16, Dm + ((b * I + bin) * 16 + (c & 3) * 4));
      }
    }
  };
  const int mid = lane >> 3, mrow = lane & 7;
  // GEMM1 (non-trans): matrices (hi k0-7, hi k8-15, lo k0-7, lo k8-15) of inner indices [.., +8)
  const uint32_t l1base = pin(op_s + (mid >> 1) * (JCV * JKS * 2) + (mrow * JKS + (mid & 1) * 8) * 2);
  // GEMM2 (trans): matrices (hi inner 0-7, hi inner 8-15, lo 0-7, lo 8-15) of basis [.., +8)
  const uint32_t l2base = pin(op_s + (mid >> 1) * (JCV * JKS * 2) + (((mid & 1) * 8 + mrow) * JKS) * 2);
  const size_t plane = (size_t)I * J;
  const float* z2b = Z2 + b * NS * plane;
  // Z2 written before the power normalisation of the diagonaliser (km_spatial): Z2_m / psi_m^2 is the current value
  float zs[NS];
#pragma unroll
  for (int m = 0; m < NS; ++m) zs[m] = zscale != nullptr ? zscale[b * NS + m] : 1.0f;
  const int nsteps = (n_inner + 15) >> 4;
  // tile of step s: 16 rows x 64 bytes per m, 4 lanes per row, 2 pieces per lane
  //   OUTER_ROWS: rows = this warp's outer bins (clamped), columns = inner frames 16 s ..
  //   else      : rows = inner bins 16 s .. (clamped), columns = this warp's outer frames
  const float* tsrc[2];
  uint32_t tdst[2];
  int trow[2];
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int idx = it * 32 + lane, r = idx >> 2, part = idx & 3;
    trow[it] = r;
    tdst[it] = zr_s + r * PITCH + part * 16;
    tsrc[it] = OUTER_ROWS ? z2b + (size_t)min(o0 + r, I - 1) * J + part * 4 : z2b + (size_t)min(o0, J - 16) + part * 4;
  }
  auto issue_tile = [&](int step) {
    const uint32_t st = (step & 1) * (NS * PTB);
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const size_t adv = OUTER_ROWS ? (size_t)step * 16 : (size_t)min(step * 16 + trow[it], I - 1) * J;
#pragma unroll
      for (int m = 0; m < NS; ++m) cp_async16(tdst[it] + st + m * PTB, tsrc[it] + m * plane + adv);
    }
  };
  const uint32_t zlane = OUTER_ROWS ? zr_s + g * PITCH + (2 * t) * 4 : zr_s + (2 * t) * PITCH + g * 4;
  issue_op(0, 0);
  if (warp_active) issue_tile(0);
  cp_async_commit();
  for (int s = 0; s < nsteps; ++s) {
    const int chunk = s >> 1, slot = chunk % MUS;
    cp_async_wait<0>();  // the group committed one step ago: tiles of step s and, on even steps, operand chunk `chunk`
    if ((s & 1) == 0) __syncthreads();  // the chunk was copied by all threads of the CTA
    else __syncwarp();                  // the tiles by all lanes of the warp
    // behind the barrier every lane has finished reading the tiles of step s - 1 (the stage refilled now) and every warp
    // the operand chunk two chunks back (the slot refilled now): group s + 1 flies while step s is computed
    if (warp_active && s + 1 < nsteps) issue_tile(s + 1);
    if ((s & 1) == 0 && chunk + 1 < nchunk_in) issue_op(chunk + 1, (chunk + 1) % MUS);
    cp_async_commit();
    if (!warp_active) continue;
    const uint32_t sb = slot * SLOT + (s & 1) * (16 * JKS * 2);
    uint32_t Ahi[NS][4], Alo[NS][4], Bhi[NS][4], Blo[NS][4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      // ---- Z2 of the half step from the ring: [m][rr] = (inner e = 0, 1) ----
      float2 z2v[NS][2];
      const uint32_t zb = zlane + (s & 1) * (NS * PTB);
#pragma unroll
      for (int m = 0; m < NS; ++m)
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          if (OUTER_ROWS) z2v[m][rr] = lds64f(zb + m * PTB + (8 * rr) * PITCH + (8 * h) * 4);
          else z2v[m][rr] = make_float2(lds32f(zb + m * PTB + (8 * h) * PITCH + 32 * rr),
                                        lds32f(zb + m * PTB + (8 * h + 1) * PITCH + 32 * rr));
        }
      // ---- GEMM1: Lambda_n[16 outer x 8 inner] ----
      float lam[NS][4];
#pragma unroll
      for (int n = 0; n < NS; ++n) {
        uint32_t oh[4], ol[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          oh[j] = ofr[((n * 2 + 0) * 4 + j) * 32];
          ol[j] = ofr[((n * 2 + 1) * 4 + j) * 32];
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) lam[n][c] = 0.f;
        uint32_t bh0, bh1, bl0, bl1;
        ldsm_x4(bh0, bh1, bl0, bl1, l1base + sb + n * CHB + (8 * h * JKS) * 2);
        mma_split(lam[n], oh, ol, bh0, bh1, bl0, bl1);
      }
      // ---- per point: L, r, G, H; split into the A fragments of GEMM2 ----
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        float G[NS][2], H[NS][2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          float d[16];
          bool valid = true;
          if (OUTER_ROWS) {
#pragma unroll
            for (int q = 0; q < 16; ++q) d[q] = dreg[OUTER_ROWS ? rr : 0][q];
          } else {
            const int ib = (s & 1) * 16 + 8 * h + 2 * t + e;  // bin inside the chunk
            valid = chunk * JCV + ib < I;
#pragma unroll
            for (int q = 0; q < 16; q += 4) {
              const float4 d4 = lds128(op_s + slot * SLOT + NS * CHB + ib * 64 + q * 4);
              d[q] = d4.x;
              d[q + 1] = d4.y;
              d[q + 2] = d4.z;
              d[q + 3] = d4.w;
            }
          }
          float r[NS], r2[NS];
#pragma unroll
          for (int m = 0; m < NS; ++m) {
            float l = 0.f;
#pragma unroll
            for (int n = 0; n < NS; ++n) l = fmaf(lam[n][rr * 2 + e], d[n * 4 + m], l);
            r[m] = fast_rcp(l);
            r2[m] = (e ? z2v[m][rr].y : z2v[m][rr].x) * zs[m] * r[m] * r[m];
          }
#pragma unroll
          for (int n = 0; n < NS; ++n) {
            float gg = 0.f, hh = 0.f;
#pragma unroll
            for (int m = 0; m < NS; ++m) {
              gg = fmaf(d[n * 4 + m], r2[m], gg);
              hh = fmaf(d[n * 4 + m], r[m], hh);
            }
            // inner indices past the end: the operand rows are zero, the factor must be finite (Lambda = 0 there)
            G[n][e] = valid ? gg : 0.f;
            H[n][e] = valid ? hh : 0.f;
          }
        }
#pragma unroll
        for (int n = 0; n < NS; ++n) {
          const Split sa = split2(G[n][0], G[n][1]);
          const Split sh = split2(H[n][0], H[n][1]);
          Ahi[n][h * 2 + rr] = sa.hi;
          Alo[n][h * 2 + rr] = sa.lo;
          Bhi[n][h * 2 + rr] = sh.hi;
          Blo[n][h * 2 + rr] = sh.lo;
        }
      }
    }
    // ---- GEMM2: num_n += G_n Op_n^T, den_n += H_n Op_n^T (contraction over the 16 inner indices) ----
#pragma unroll
    for (int n = 0; n < NS; ++n)
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        uint32_t th0, th1, tl0, tl1;
        ldsm_x4_t(th0, th1, tl0, tl1, l2base + sb + n * CHB + q * 16);
        mma_split(num[n][q], Ahi[n], Alo[n], th0, th1, tl0, tl1);
        mma_split(den[n][q], Bhi[n], Blo[n], th0, th1, tl0, tl1);
      }
  }
  cp_async_wait<0>();
  if (!warp_active) return;
  // ---- Out <- floor(Out sqrt(num / den)), also written pre-split for the next kernel ----
#pragma unroll
  for (int n = 0; n < NS; ++n) {
    const size_t bn = b * NS + n;
    float* Ob = Out + bn * (size_t)K * (OUTER_ROWS ? I : J);
#pragma unroll
    for (int nb = 0; nb < 2; ++nb)
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        if (!ovalid[rr]) continue;
        const int k0 = nb * 8 + 2 * t;
        float vn[2];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          vn[e] = 0.f;
          if (k0 + e < K) {
            float* op = OUTER_ROWS ? Ob + (size_t)oc[rr] * K + k0 + e : Ob + (size_t)(k0 + e) * J + oc[rr];
            vn[e] = ssb_floor(*op * sqrtf(num[n][nb][rr * 2 + e] / den[n][nb][rr * 2 + e]), flooring, eps);
            *op = vn[e];
          }
        }
        const Split sp = split2(vn[0], vn[1]);
        __nv_bfloat16* oh = OutSplit + (bn * nchunk_out + (oc[rr] >> 5)) * (size_t)(2 * JCV * JKS) + (oc[rr] & 31) * JKS + k0;
        *reinterpret_cast<uint32_t*>(oh) = sp.hi;
        *reinterpret_cast<uint32_t*>(oh + JCV * JKS) = sp.lo;
      }
  }
}

int launch_mnmf_update(int which, const float* Z2, const float* zscale, const float* Dm, float* T, float* V,
                       __nv_bfloat16* Vs, __nv_bfloat16* Ts, int B, int I, int J, int K, int flooring, float eps,
                       cudaStream_t st) {
  constexpr int JKS = 16 + PADH, CHB = 2 * JCV * JKS * 2;
  const int nchunk_j = (J + JCV - 1) / JCV, nchunk_i = (I + JCV - 1) / JCV;
  const size_t ofr = (size_t)MUW * 4 * 2 * 4 * 32 * 4;
  const size_t sm_b = (size_t)MUS * (4 * CHB) + ofr + (size_t)MUW * 2 * 4 * 16 * 96;
  const size_t sm_a = (size_t)MUS * (4 * CHB + JCV * 16 * 4) + ofr + (size_t)MUW * 2 * 4 * 16 * 80;
  static bool attr_dev[SSB_MAX_DEVICES] = {};  // function attributes are per device
  bool& attr_set = attr_dev[ssb_current_device()];
  if (!attr_set) {
    SSB_CUDA(cudaFuncSetAttribute(kf_mnmf_update<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_b));
    SSB_CUDA(cudaFuncSetAttribute(kf_mnmf_update<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm_a));
    attr_set = true;
  }
  if (which == 0) {
    dim3 gv(nchunk_j, B * 4);
    kf_vsplit<1><<<gv, 128, 0, st>>>(V, Vs, J, K, nchunk_j);
    if (ssb_check_launch("coop_vsplit", st)) return 1;
    dim3 grid((I + MUW * 16 - 1) / (MUW * 16), B);
    kf_mnmf_update<true><<<grid, MUW * 32, sm_b, st>>>(Z2, Dm, T, T, Vs, Ts, I, J, K, nchunk_j, nchunk_i, flooring, eps,
                                                        zscale);
    return ssb_check_launch("mnmf_basis_fused", st);
  }
  dim3 grid((J + MUW * 16 - 1) / (MUW * 16), B);
  kf_mnmf_update<false><<<grid, MUW * 32, sm_a, st>>>(Z2, Dm, V, V, Ts, Vs, I, J, K, nchunk_i, nchunk_j, flooring, eps,
                                                       zscale);
  return ssb_check_launch("mnmf_activation_fused", st);
}

// ================================================================================================
// Cooperative weighted covariance for N = 4 and N = 8 (GaussILRMA, p = 2):
//   phi = 1 / (T V)                                   (ssspy/bss/ilrma.py:1494-1498)
//   U[b,i,n,a,c] = (1/J) sum_j phi[n,i,j] x_a conj(x_c) (ilrma.py:1500-1505)
// kf_phi_cov (ssb_fused.cu) keeps every Hermitian entry of G sources in each thread: 2 N^2 registers per source, so
// at N = 8 only two sources share the products of a frame and the kernel runs 8 warps per SM at 255 registers.
// Here a CTA owns BT tiles of 16 bins; the X slab of a tile goes once through a shared cp.async ring and the work on
// it is split over WPT = (N / G) * N/2 warps: warp (sg, q) accumulates, for the G sources of group sg, the q-th
// Hamiltonian path of the complete graph on the N channels (Walecki's zigzag decomposition: K_N = N/2 paths of
// N - 1 edges, q, q+1, q-1, q+2, ... mod N) plus the two diagonal entries at the path's end points (q and q + N/2).
// Every warp therefore runs the SAME code on a per-warp permutation of the channels (only shared-memory offsets
// differ): N - 1 complex products shared by G sources, N accumulator registers per source instead of N^2, and phi
// from the pre-split V chunks (ldmatrix) like in kf_basis_coop.  One row group (8 bins) of the tile per pass over
// the frames, as in kf_phi_cov.
namespace {

template <int N, int G_, bool RSW_ = false>
struct CovCoop {
  static constexpr int G = G_;                 // sources per warp
  static constexpr int EQ = N / 2;             // Hamiltonian paths = warps per source group
  static constexpr bool RSW = RSW_;            // the two row groups (8 bins each) of a tile go to different warps
                                               // (one pass over the frames) instead of two passes of the same warp
  static constexpr int WPT = (N / G) * EQ * (RSW ? 2 : 1);  // warps per tile
  static constexpr int BT = (8 / WPT) > 0 ? 8 / WPT : 1;  // tiles per CTA
  static constexpr int NT = WPT * BT * 32;     // threads per CTA
  static_assert(N == 4 || N == 8, "N = 4, 8 only");
};

template <int N, int KS, int G_, bool RSW_>
__global__ void __launch_bounds__(CovCoop<N, G_, RSW_>::NT, RSW_ ? 2 : 0)
    kf_cov_coop(const cf* __restrict__ X, const float* __restrict__ T, const __nv_bfloat16* __restrict__ Vs,
                cf* __restrict__ U, int I, int J, int K, int nchunk) {
  using S = CovCoop<N, G_, RSW_>;
  constexpr int G = S::G, BT = S::BT, EQ = S::EQ;
  constexpr int KP = 16 * KS, JKS = KP + PADH;
  constexpr int CHB = 2 * JCV * JKS * 2;   // bytes of one source's V chunk
  constexpr int XTB = N * 2048;            // bytes of one tile's X stage
  constexpr int XSB = BT * XTB;            // bytes of one X stage
  constexpr int NT = S::NT;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t xs_s = (uint32_t)__cvta_generic_to_shared(smem_raw);
  const uint32_t vs_s = xs_s + XST * XSB;  // V ring: [2][N][CHB]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int bt = warp / S::WPT, p0 = warp - bt * S::WPT;
  const int rsw = S::RSW ? (p0 & 1) : 0, p = S::RSW ? (p0 >> 1) : p0;
  const int sg = p / EQ, q = p - sg * EQ;
  const int b = blockIdx.y;
  const int i0 = (blockIdx.x * BT + bt) * 16;
  const bool tile_active = i0 < I;
  const float invJ = 1.0f / (float)J;

  // channel order of this warp's Hamiltonian path: q, q+1, q-1, q+2, q-2, ... (mod N)
  int perm[N];
  uint32_t choff[N];
#pragma unroll
  for (int m = 0; m < N; ++m) {
    const int d = (m + 1) >> 1;
    perm[m] = (q + ((m & 1) ? d : N - d)) & (N - 1);
    choff[m] = pin((uint32_t)perm[m] * 2048u);
  }

  // cooperative loaders: X stage = BT*N*16 rows x 8 pieces of 16 bytes; V chunk = N*CHB bytes
  constexpr int XP = BT * N * 128 / NT;
  const cf* xsrc[XP];
  uint32_t xdst[XP];
#pragma unroll
  for (int it = 0; it < XP; ++it) {
    const int e = it * NT + tid, ch8 = e & 7, r = (e >> 3) & 15, m = (e >> 7) % N, tl = e / (128 * N);
    const int ib = min((blockIdx.x * BT + tl) * 16 + r, I - 1);
    xsrc[it] = X + (((size_t)b * N + m) * I + ib) * (size_t)J + 2 * ch8;
    xdst[it] = pin(xs_s + tl * XTB + (m * 16 + r) * 128 + ((ch8 ^ ((r & 1) << 2)) << 4));
  }
  auto issue_x = [&](uint32_t slot_bytes) {
#pragma unroll
    for (int it = 0; it < XP; ++it) {
      cp_async16(xdst[it] + slot_bytes, xsrc[it]);
      xsrc[it] += 16;
    }
  };
  constexpr int VP = N * CHB / 16 / NT;
  static_assert((N * CHB / 16) % NT == 0, "V chunk pieces must divide over the CTA");
  // piece e of a chunk: source e / (CHB/16), byte (e % (CHB/16)) * 16; source s of chunk c sits at ((b*N+s)*nchunk+c)*CHB
  const unsigned char* vp[VP];
  uint32_t vd[VP];
#pragma unroll
  for (int it = 0; it < VP; ++it) {
    const int e = it * NT + tid, s = e / (CHB / 16), off = (e % (CHB / 16)) * 16;
    vp[it] = reinterpret_cast<const unsigned char*>(Vs) + ((size_t)b * N + s) * nchunk * CHB + off;
    vd[it] = pin(vs_s + s * CHB + off);
  }
  auto issue_v = [&](int buf) {
#pragma unroll
    for (int it = 0; it < VP; ++it) {
      cp_async16(vd[it] + buf * (N * CHB), vp[it]);
      vp[it] += CHB;
    }
  };
  const int mid = lane >> 3, mrow = lane & 7;
  const uint32_t l1base = pin(vs_s + sg * G * CHB + (mid >> 1) * (JCV * JKS * 2) + (mrow * JKS + (mid & 1) * 8) * 2);
  const int nsteps = J >> 4;

#pragma unroll 1
  for (int pass = 0; pass < (S::RSW ? 1 : 2); ++pass) {
    const int rs = S::RSW ? rsw : pass;
    const int row = i0 + g + 8 * rs;
    // T fragments of the G sources (both row groups feed the MMA; only row group rs is used afterwards)
    uint32_t Thi[G][KS][4], Tlo[G][KS][4];
#pragma unroll
    for (int gs = 0; gs < G; ++gs)
#pragma unroll
      for (int ks = 0; ks < KS; ++ks)
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const float* tr = T + (((size_t)b * N + sg * G + gs) * I + min(i0 + g + 8 * rr, I - 1)) * K;
#pragma unroll
          for (int nb = 0; nb < 2; ++nb) {
            const int k0 = ks * 16 + nb * 8 + 2 * t;
            const Split sp = split2((k0 < K) ? tr[k0] : 0.f, (k0 + 1 < K) ? tr[k0 + 1] : 0.f);
            Thi[gs][ks][nb * 2 + rr] = sp.hi;
            Tlo[gs][ks][nb * 2 + rr] = sp.lo;
          }
        }
    float2 ao[G][N - 1], ad[G];  // path edges (x'_e conj x'_{e+1}); end-point powers (|x'_0|^2, |x'_{N-1}|^2)
#pragma unroll
    for (int gs = 0; gs < G; ++gs) {
#pragma unroll
      for (int e = 0; e < N - 1; ++e) ao[gs][e] = make_float2(0.f, 0.f);
      ad[gs] = make_float2(0.f, 0.f);
    }
    // one frame: the products of this warp's path, shared by its G sources
    auto accum = [&](const float (&xr)[N], const float (&xi)[N], const float (&phf)[G]) {
      float2 pp[G];
#pragma unroll
      for (int gs = 0; gs < G; ++gs) pp[gs] = make_float2(phf[gs], phf[gs]);
#pragma unroll
      for (int e = 0; e < N - 1; ++e) {
        const float2 pr = make_float2(fmaf(xr[e], xr[e + 1], xi[e] * xi[e + 1]), fmaf(xi[e], xr[e + 1], -(xr[e] * xi[e + 1])));
#pragma unroll
        for (int gs = 0; gs < G; ++gs) ao[gs][e] = __ffma2_rn(pp[gs], pr, ao[gs][e]);
      }
      const float2 pd = make_float2(fmaf(xr[0], xr[0], xi[0] * xi[0]), fmaf(xr[N - 1], xr[N - 1], xi[N - 1] * xi[N - 1]));
#pragma unroll
      for (int gs = 0; gs < G; ++gs) ad[gs] = __ffma2_rn(pp[gs], pd, ad[gs]);
    };
    const uint32_t xlane[2] = {pin(xs_s + bt * XTB + (g + 8 * rs) * 128 + ((t ^ ((g & 1) << 2)) << 4)),
                               pin(xs_s + bt * XTB + (g + 8 * rs) * 128 + (((4 + t) ^ ((g & 1) << 2)) << 4))};
    // restart the streams
    if (pass) {
#pragma unroll
      for (int it = 0; it < XP; ++it) xsrc[it] -= nsteps * 16;
#pragma unroll
      for (int it = 0; it < VP; ++it) vp[it] -= (size_t)nchunk * CHB;
    }
    __syncthreads();  // the previous pass no longer reads the rings
    issue_v(0);
    issue_x(0);
    cp_async_commit();
    if (nsteps > 1) issue_x(XSB);
    cp_async_commit();
    uint32_t rd_slot = 0, wr_slot = 2 * XSB;
#pragma unroll 1
    for (int s = 0; s < nsteps; ++s) {
      cp_async_wait<XST - 2>();
      __syncthreads();
      if (s + 2 < nsteps) issue_x(wr_slot);
      if ((s & 1) == 0 && (s >> 1) + 1 < nchunk) issue_v(((s >> 1) + 1) & 1);
      cp_async_commit();
      const uint32_t vb1 = l1base + ((s >> 1) & 1) * (N * CHB) + (s & 1) * (16 * JKS * 2);
      const uint32_t xb[2] = {xlane[0] + rd_slot, xlane[1] + rd_slot};
      rd_slot = rd_slot + XSB == XST * XSB ? 0 : rd_slot + XSB;
      wr_slot = wr_slot + XSB == XST * XSB ? 0 : wr_slot + XSB;
      if (!tile_active) continue;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float ph[2][G];
#pragma unroll
        for (int gs = 0; gs < G; ++gs) {
          float R[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int ks = 0; ks < KS; ++ks) {
            uint32_t bh0, bh1, bl0, bl1;
            ldsm_x4(bh0, bh1, bl0, bl1, vb1 + gs * CHB + (8 * h * JKS + ks * 16) * 2);
            mma_split(R, Thi[gs][ks], Tlo[gs][ks], bh0, bh1, bl0, bl1);
          }
          ph[0][gs] = fast_rcp(rs ? R[2] : R[0]);
          ph[1][gs] = fast_rcp(rs ? R[3] : R[1]);
        }
        float4 x[N];
#pragma unroll
        for (int m = 0; m < N; ++m) x[m] = lds128(xb[h] + choff[m]);
        float xr[N], xi[N];
#pragma unroll
        for (int m = 0; m < N; ++m) {
          xr[m] = x[m].x;
          xi[m] = x[m].y;
        }
        accum(xr, xi, ph[0]);
#pragma unroll
        for (int m = 0; m < N; ++m) {
          xr[m] = x[m].z;
          xi[m] = x[m].w;
        }
        accum(xr, xi, ph[1]);
      }
    }
    if (tile_active) {
      const bool wr = (t == 0) && (row < I);
#pragma unroll
      for (int gs = 0; gs < G; ++gs) {
        cf* u = U + (((size_t)b * I + min(row, I - 1)) * N + sg * G + gs) * N * N;
#pragma unroll
        for (int e = 0; e < N - 1; ++e) {
          float2 v = ao[gs][e];
          v.x += __shfl_xor_sync(SSB_FULL, v.x, 1);
          v.y += __shfl_xor_sync(SSB_FULL, v.y, 1);
          v.x += __shfl_xor_sync(SSB_FULL, v.x, 2);
          v.y += __shfl_xor_sync(SSB_FULL, v.y, 2);
          const int a = perm[e], c = perm[e + 1];
          if (wr) {
            u[a * N + c] = make_float2(v.x * invJ, v.y * invJ);
            u[c * N + a] = make_float2(v.x * invJ, -v.y * invJ);
          }
        }
        float2 v = ad[gs];
        v.x += __shfl_xor_sync(SSB_FULL, v.x, 1);
        v.y += __shfl_xor_sync(SSB_FULL, v.y, 1);
        v.x += __shfl_xor_sync(SSB_FULL, v.x, 2);
        v.y += __shfl_xor_sync(SSB_FULL, v.y, 2);
        if (wr) {
          u[perm[0] * N + perm[0]] = make_float2(v.x * invJ, 0.f);
          u[perm[N - 1] * N + perm[N - 1]] = make_float2(v.y * invJ, 0.f);
        }
      }
    }
  }
}

template <int N, int KS, int G_, bool RSW_ = false>
int launch_cov_coop(const ssb_config* c, const cf* X, const float* T, const __nv_bfloat16* Vs, cf* U, cudaStream_t st) {
  using S = CovCoop<N, G_, RSW_>;
  const int B = c->n_batch, I = c->n_bins, J = c->n_frames, K = c->n_basis;
  constexpr int KP = 16 * KS, JKS = KP + PADH, CHB = 2 * JCV * JKS * 2;
  const int nchunk = (J + JCV - 1) / JCV;
  const size_t sm = (size_t)XST * S::BT * N * 2048 + (size_t)2 * N * CHB;
  static bool attr_dev[SSB_MAX_DEVICES] = {};  // function attributes are per device
  bool& attr_set = attr_dev[ssb_current_device()];
  if (!attr_set) {
    SSB_CUDA(cudaFuncSetAttribute(kf_cov_coop<N, KS, G_, RSW_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    attr_set = true;
  }
  dim3 grid((I + 16 * S::BT - 1) / (16 * S::BT), B);
  kf_cov_coop<N, KS, G_, RSW_><<<grid, S::NT, sm, st>>>(X, T, Vs, U, I, J, K, nchunk);
  return ssb_check_launch("coop_phi_cov", st);
}

}  // namespace

// N = 8: two passes of each warp over the row groups (registers); N = 4: the row groups go to different warps (four
// warps per tile, one pass over X, two CTAs per SM): 0.41 -> 0.35 ms against kf_phi_cov.  The two-pass variant was
// slower than kf_phi_cov at N = 4 (0.47 ms: its second pass misses L2, profiles/r1_ncu_coop_summary.md).
// SSB_COOP_COV=0 falls back to kf_phi_cov.
int ssb_coop_cov_supported(const ssb_config* c) {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("SSB_COOP_COV");
    mode = e ? atoi(e) : 1;
  }
  const bool n_ok = mode != 0 && (c->n_sources == 8 || c->n_sources == 4);
  return n_ok && (c->n_frames % 16) == 0 && c->n_basis <= 32;
}

// weighted covariance of every source from the pre-split activation Vs held in ws (valid after ssb_coop_source)
int ssb_coop_cov(const ssb_config* c, const cf* X, const float* T, const void* ws, cf* U, cudaStream_t st) {
  SSB_REQUIRE(ssb_coop_cov_supported(c) && ws != nullptr, "coop_cov: unsupported configuration");
  const __nv_bfloat16* Vs = (const __nv_bfloat16*)ws;
  const bool k16 = c->n_basis <= 16;
  static int g2 = -1;  // SSB_COV_G: sources per warp at N = 8 (4: fewer product instructions, 2: twice the warps)
  if (g2 < 0) {
    const char* e = getenv("SSB_COV_G");
    g2 = (e && atoi(e) == 2) ? 1 : 0;
  }
  // N = 4: row groups over warps (4 warps per tile, one pass over X, 2 CTAs per SM)
  if (c->n_sources == 4)
    return k16 ? launch_cov_coop<4, 1, 4, true>(c, X, T, Vs, U, st) : launch_cov_coop<4, 2, 4, true>(c, X, T, Vs, U, st);
  if (g2) return k16 ? launch_cov_coop<8, 1, 2>(c, X, T, Vs, U, st) : launch_cov_coop<8, 2, 2>(c, X, T, Vs, U, st);
  return k16 ? launch_cov_coop<8, 1, 4>(c, X, T, Vs, U, st) : launch_cov_coop<8, 2, 4>(c, X, T, Vs, U, st);
}

// FastGaussMNMF source model with the factors formed in the kernel (four sources, K <= 16): which = 0 basis, 1 activation
int ssb_coop_mnmf_update(const ssb_config* c, int which, const float* Z2, const float* zscale, const float* Dm, float* T,
                         float* V, void* ws, cudaStream_t st) {
  SSB_REQUIRE(c->n_sources == 4 && (c->n_frames % 16) == 0 && c->n_basis <= 16 && ws != nullptr,
              "coop_mnmf_update: unsupported configuration");
  __nv_bfloat16* Vs = (__nv_bfloat16*)ws;
  __nv_bfloat16* Ts = (__nv_bfloat16*)((char*)ws + coop_vs_bytes(c));
  return launch_mnmf_update(which, Z2, zscale, Dm, T, V, Vs, Ts, c->n_batch, c->n_bins, c->n_frames, c->n_basis, c->flooring,
                            c->eps, st);
}

// FastGaussMNMF source model on the tensor pipe: which = 0 basis (reads V, writes T and the pre-split Ts),
// which = 1 activation (reads the Ts left by the basis call, writes V).  ws as for ssb_coop_source.
int ssb_coop_update_ab(const ssb_config* c, int which, const float* A, const float* Bm, float* T, float* V, void* ws,
                       cudaStream_t st) {
  SSB_REQUIRE((c->n_frames % 16) == 0 && c->n_basis <= 32 && ws != nullptr, "coop_update_ab: unsupported configuration");
  __nv_bfloat16* Vs = (__nv_bfloat16*)ws;
  __nv_bfloat16* Ts = (__nv_bfloat16*)((char*)ws + coop_vs_bytes(c));
  const int BN = c->n_batch * c->n_sources;
  if (c->n_basis <= 16)
    return launch_update_ab<1>(which, A, Bm, T, V, Vs, Ts, BN, c->n_bins, c->n_frames, c->n_basis, c->flooring, c->eps, st);
  return launch_update_ab<2>(which, A, Bm, T, V, Vs, Ts, BN, c->n_bins, c->n_frames, c->n_basis, c->flooring, c->eps, st);
}
