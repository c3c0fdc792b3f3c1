// Cooperative basis update of the fused GaussILRMA path (source_algorithm="MM", domain=2):
//   T <- T sqrt( sum_j V P / R^2  /  sum_j V / R ),  P = |w_n^H x|^2,  R = T V      (ssspy/bss/ilrma.py:1051-1128)
//
// kf_basis (ssb_fused.cu) gives every (mixture, source) its own CTAs, so a bin's (channel x frame) slab is pulled
// from L2 once per source: N x the compulsory traffic, and at N = 8 the kernel sits on the L2 -> SM bandwidth
// (profiles/r1_bench_configs_e.jsonl: 1.96 ms against a 0.33 ms HBM floor).  Here one CTA owns BT tiles of 16 bins
// for ALL sources: warp (bt, n) updates source n of tile bt, the N warps of a tile share its X slab through a
// 3-stage cp.async ring in shared memory (each warp fetches "its" channel, a named barrier per tile hands the stage
// over) and every warp streams its own activation tile.  V is pre-split once per call into bf16 (hi, lo) in the
// [frame][basis] layout (kf_vsplit); both MMA operand orientations come out of that one layout with
// ldmatrix / ldmatrix.trans, so the hot loop has 4 LDSM instead of 16 LDS and no split arithmetic for V.
#include <cuda_bf16.h>

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
  const uint32_t ua = __float_as_uint(a), ub = __float_as_uint(b);
  Split s;
  s.hi = __byte_perm(ua, ub, 0x7632);
  const float ra = a - __uint_as_float(ua & 0xffff0000u);
  const float rb = b - __uint_as_float(ub & 0xffff0000u);
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
// hi = truncated bf16, lo = bf16(rn) of the exact residual, as split2.
template <int KS>
__global__ void __launch_bounds__(128) kf_vsplit(const float* __restrict__ V, __nv_bfloat16* __restrict__ Vs, int J,
                                                 int K, int nchunk) {
  constexpr int KP = 16 * KS, JKS = KP + PADH, KG = KP / 4;
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
  __nv_bfloat16* hi = Vs + (bn * nchunk + chunk) * (size_t)(2 * JCV * JKS) + fr * JKS + kg * KG;
  __nv_bfloat16* lo = hi + JCV * JKS;
#pragma unroll
  for (int e = 0; e < KG; e += 2) {
    const Split s = split2(v[e], v[e + 1]);
    *reinterpret_cast<uint32_t*>(hi + e) = s.hi;
    *reinterpret_cast<uint32_t*>(lo + e) = s.lo;
  }
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
                  const __nv_bfloat16* __restrict__ Vs, float* __restrict__ Pout, int I, int J, int K, int nchunk,
                  int flooring, float eps) {
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
        const float v0 = (k0 < K && rvalid[rr]) ? tr[k0] : 0.f;
        const float v1 = (k0 + 1 < K && rvalid[rr]) ? tr[k0 + 1] : 0.f;
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
    xdst[it] = xs_s + bt * XTB + (n * 16 + r) * 128 + ((ch ^ ((r & 1) << 2)) << 4);
  }
  auto issue_x = [&](int step, int slot) {
#pragma unroll
    for (int it = 0; it < 4; ++it) cp_async16(xdst[it] + slot * (BT * XTB), xsrc[it] + step * 16);
  };
  const unsigned char* vsrc = reinterpret_cast<const unsigned char*>(Vs) + bn * (size_t)nchunk * CHB;
  auto issue_v = [&](int chunk, int buf) {
#pragma unroll
    for (int c = lane; c < CHB / 16; c += 32) cp_async16(vs_s + buf * CHB + c * 16, vsrc + (size_t)chunk * CHB + c * 16);
  };
  // per-lane shared-memory offsets of the X fragments and of the ldmatrix rows
  uint32_t xoff[2][2];
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) xoff[h][rr] = (g + 8 * rr) * 128 + (((4 * h + t) ^ ((g & 1) << 2)) << 4);
  const int mid = lane >> 3, mrow = lane & 7;
  // GEMM1 (non-trans): matrices (hi k0-7, hi k8-15, lo k0-7, lo k8-15) of frames [.., +8)
  const uint32_t l1off = (mid >> 1) * (JCV * JKS * 2) + (mrow * JKS + (mid & 1) * 8) * 2;
  // GEMM2 (trans): matrices (hi frames 0-7, hi frames 8-15, lo 0-7, lo 8-15) of basis [.., +8)
  const uint32_t l2off = (mid >> 1) * (JCV * JKS * 2) + (((mid & 1) * 8 + mrow) * JKS) * 2;

  const int nsteps = J >> 4;
  issue_v(0, 0);
  issue_x(0, 0);
  cp_async_commit();
  if (nsteps > 1) issue_x(1, 1);
  cp_async_commit();

  for (int s = 0; s < nsteps; ++s) {
    cp_async_wait<XST - 2>();
    bar_sync_tile<N * 32>(bt);
    if (s + 2 < nsteps) issue_x(s + 2, (s + 2) % XST);
    if ((s & 1) == 0 && (s >> 1) + 1 < nchunk) issue_v((s >> 1) + 1, ((s >> 1) + 1) & 1);
    cp_async_commit();

    const uint32_t vb = vs_s + ((s >> 1) & 1) * CHB + (s & 1) * (16 * JKS * 2);
    const uint32_t xb = xs_s + (s % XST) * (BT * XTB) + bt * XTB;
    // ---- GEMM1: R[16 bins x 16 frames] = T V ---------------------------------------------------------
    float R[2][4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int c = 0; c < 4; ++c) R[h][c] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t bh0, bh1, bl0, bl1;
        ldsm_x4(bh0, bh1, bl0, bl1, vb + l1off + (8 * h * JKS + ks * 16) * 2);
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
        for (int m = 0; m < N; ++m) x[m] = lds128(xb + m * 2048 + xoff[h][rr]);
        float p0, p1;
        power2<N>(x, w[rr], p0, p1);
        // the power spectrogram is kept for the activation update (same W => same P, ilrma.py:1169-1172)
        if (rvalid[rr])
          *reinterpret_cast<float2*>(Pout + (bn * I + row[rr]) * (size_t)J + s * 16 + 8 * h + 2 * t) = make_float2(p0, p1);
        const float i0v = rvalid[rr] ? fast_rcp(R[h][rr * 2 + 0]) : 0.f;
        const float i1v = rvalid[rr] ? fast_rcp(R[h][rr * 2 + 1]) : 0.f;
        const Split sa = split2(p0 * i0v * i0v, p1 * i1v * i1v);
        const Split sb = split2(i0v, i1v);
        Ahi[h * 2 + rr] = sa.hi;
        Alo[h * 2 + rr] = sa.lo;
        Bhi[h * 2 + rr] = sb.hi;
        Blo[h * 2 + rr] = sb.lo;
      }
    // ---- GEMM2: num += A V^T, den += B V^T  (contraction over the 16 frames) ---------------------------
#pragma unroll
    for (int q = 0; q < 2 * KS; ++q) {
      uint32_t vh0, vh1, vl0, vl1;
      ldsm_x4_t(vh0, vh1, vl0, vl1, vb + l2off + q * 16);
      mma_split(num[q], Ahi, Alo, vh0, vh1, vl0, vl1);
      mma_split(den[q], Bhi, Blo, vh0, vh1, vl0, vl1);
    }
  }
  // ---- T <- floor(T * sqrt(num / den))      (ilrma.py:1125-1126, p = 2) --------------------------------
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

template <int N, int KS>
int launch_coop(const ssb_config* c, const cf* X, const cf* W, float* T, const float* V, float* P,
                __nv_bfloat16* Vs, cudaStream_t st) {
  const int B = c->n_batch, I = c->n_bins, J = c->n_frames, K = c->n_basis;
  constexpr int BT = CoopShape<N>::BT, NW = CoopShape<N>::NW;
  constexpr int KP = 16 * KS, JKS = KP + PADH;
  const int nchunk = (J + JCV - 1) / JCV;
  dim3 gv(nchunk, B * N);
  kf_vsplit<KS><<<gv, 128, 0, st>>>(V, Vs, J, K, nchunk);
  if (ssb_check_launch("coop_vsplit", st)) return 1;
  const size_t sm = (size_t)XST * BT * N * 2048 + (size_t)NW * 2 * (2 * JCV * JKS * 2);
  static bool attr_set = false;
  if (!attr_set) {
    SSB_CUDA(cudaFuncSetAttribute(kf_basis_coop<N, KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    attr_set = true;
  }
  dim3 grid((I + 16 * BT - 1) / (16 * BT), B);
  kf_basis_coop<N, KS><<<grid, NW * 32, sm, st>>>(X, W, T, Vs, P, I, J, K, nchunk, c->flooring, c->eps);
  return ssb_check_launch("coop_basis", st);
}

}  // namespace

size_t ssb_coop_ws_bytes(const ssb_config* c) {
  const size_t KP = c->n_basis <= 16 ? 16 : 32;
  const size_t nchunk = ((size_t)c->n_frames + JCV - 1) / JCV;
  return (size_t)c->n_batch * c->n_sources * nchunk * (2 * JCV * (KP + PADH)) * sizeof(__nv_bfloat16);
}

// basis update for every source of every mixture; Vs is scratch of ssb_coop_ws_bytes() bytes
int ssb_coop_basis(const ssb_config* c, const cf* X, const cf* W, float* T, const float* V, float* P, void* Vs,
                   cudaStream_t st) {
  SSB_REQUIRE((c->n_frames % 16) == 0 && c->n_basis <= 32 && W != nullptr && Vs != nullptr,
              "coop_basis: unsupported configuration");
  if (c->n_basis <= 16) {
    SSB_DISPATCH_N(c->n_sources, return (launch_coop<NN, 1>(c, X, W, T, V, P, (__nv_bfloat16*)Vs, st)));
  } else {
    SSB_DISPATCH_N(c->n_sources, return (launch_coop<NN, 2>(c, X, W, T, V, P, (__nv_bfloat16*)Vs, st)));
  }
  return 0;
}
