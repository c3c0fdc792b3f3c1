// TMA-fed tile kernels of the GaussILRMA iteration (source_algorithm="MM", domain=2; N = 2, 4, 8).
//
//   kt_tile<N, KS, MODE, FS>   persistent CTA = one tile of 16 bins at a time x ALL sources x FS = 8 / N frame ranges:
//                          consumer warp (frame range fq, source n) + one producer warp
//     MODE_BASIS  T <- T sqrt(sum_j V P / R^2 / sum_j V / R), P = |w_n^H x|^2, R = T V   (ssspy/bss/ilrma.py:1051-1128)
//     MODE_COV    phi = 1 / (T V), U_n = mean_j phi x x^H                                  (ilrma.py:1494-1505), N = 2
//     MODE_FUSED  MODE_COV of iteration t, IP1 in fp64 (_update_spatial_model.py:63-76), then MODE_BASIS of iteration
//                 t + 1 with the new filter on the SAME tile (N = 2, inside ssb_run only)
//
// Data movement is Blackwell-native: the (channel x frame) slab of a tile is pulled by the producer warp with
// cp.async.bulk.tensor (TMA, one tensor map over X[B*N planes][I rows][2 J floats], box = 8 frames x 16 bins x N
// channels) into a 3-stage shared-memory ring guarded by mbarriers (full: complete_tx, empty: one arrive per source
// warp); the pre-split activation chunks travel as 1-D bulk copies into per-warp double buffers.  Consumer warps issue
// no global loads at all in their frame loops: ldmatrix + LDS.128 + mma.sync + the elementwise stage.  A box lands as
// [channel][16 rows][64 bytes]; a quarter warp's LDS.128 (2 rows x 4 frame pairs) covers 128 contiguous bytes, so the
// layout is bank-conflict free without a swizzle, and rows past the last bin are zero-filled by the TMA unit.
//
// Why the frames of a tile are split over warps (FS ranges): in MODE_FUSED the tile's slab is read twice, once for the
// covariance and once, after IP1, for the basis update.  With one warp per (tile, source) sweeping all frames, the
// slabs waiting between their two passes (148 SMs x 16 warps x 64 KB = 155 MB at N = 2, J = 512) exceed the 126 MB L2
// and the second pass goes back to HBM (round 1: 0.335 vs 0.345 ms, profiles/r1_fuse_iter_ab.jsonl).  With the frames
// split four ways a CTA finishes a tile's first pass after a quarter of the bytes per warp: 39 MB in flight, and the
// second pass is served by L2.  The per-range partial sums (U, then num / den) are combined through shared memory in a
// fixed order, so results do not depend on scheduling.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include <mutex>

#include "ssb_fused.h"
#include "ssb_kernels.h"

namespace {

constexpr int PADH = 8;   // bf16 padding of a [frame][basis] row (see ssb_coop.cu)
constexpr int JCV = 32;   // frames per pre-split V chunk
constexpr int XS = 3;     // stages of an X ring (16 frames each)
constexpr int MODE_BASIS = 0, MODE_COV = 1, MODE_FUSED = 2;

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
__device__ __forceinline__ uint32_t pin(uint32_t v) {
  asm volatile("" : "+r"(v));
  return v;
}
template <typename T>
__device__ __forceinline__ T* pin_ptr(T* p) {
  asm volatile("" : "+l"(p));
  return p;
}

// ---- mbarrier + TMA (PTX ISA 8.x, sm_90+) ------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
        : "r"(bar), "r"(parity), "r"(1000000)  // suspend-time hint (ns): sleep in hardware rather than spin in the loop
        : "memory");
  } while (!done);
}
// one box of the X tensor map: coordinates (float index in the row, bin, plane)
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_hint(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, "
      "%3, %4}], [%5], %6;"
      ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar), "l"(policy)
      : "memory");
}
// contiguous bytes (multiple of 16, 16-byte aligned on both sides)
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
template <int NT>
__device__ __forceinline__ void bar_sync_id(int id) {  // named barrier 1 + id over NT threads
  asm volatile("bar.sync %0, %1;" ::"r"(id + 1), "n"(NT) : "memory");
}
template <int NT>
__device__ __forceinline__ void bar_arrive_id(int id) {  // non-blocking arrival at named barrier 1 + id
  asm volatile("bar.arrive %0, %1;" ::"r"(id + 1), "n"(NT) : "memory");
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

template <int N, int KS, int FS_>
struct TileShape {
  static constexpr int FS = FS_;                 // frame ranges per tile
  static constexpr int TB = 8 / (N * FS);        // tiles per CTA (a "group" of consecutive bin tiles of one mixture)
  static constexpr int NCW = 8;                  // warps: (tile, frame range, source)
  static constexpr int NT = NCW * 32;
  static constexpr int KP = 16 * KS, JKS = KP + PADH;
  static constexpr int CHB = 2 * JCV * JKS * 2;  // bytes of one V chunk (hi + lo)
  static constexpr int XSB = N * 2048;           // bytes of one X stage: 16 frames x 16 bins x N channels
  static constexpr int NR = TB * FS;             // X rings (one per tile and frame range)
  static constexpr int X_BYTES = NR * XS * XSB;
  static constexpr int V_BYTES = NCW * 2 * CHB;
  static constexpr int NBAR = NR * XS + NCW * 2;
  static constexpr int U_FLOATS = 2 * FS * 16 * 4;    // MODE_COV / FUSED (N = 2), per tile: [source][fq][row][4]
  static constexpr int U_BYTES = TB * U_FLOATS * 4;
  static constexpr int NV = 2 * (2 * KS) * 4 + 2;     // values per lane combined over the frame ranges: num, den, qs
  static constexpr int C_FLOATS = FS > 1 ? N * (FS - 1) * NV * 32 : 0;  // per tile: [source][fq - 1][value][lane]
  static constexpr int C_BYTES = TB * C_FLOATS * 4;
  static constexpr int SMEM = X_BYTES + V_BYTES + U_BYTES + C_BYTES + NBAR * 8 + 128;
  static constexpr int TILE_THREADS = FS * N * 32;    // the warps of one tile (its combine barrier)
  static constexpr int RING_BAR0 = FS > 1 ? TB : 0;   // named barriers: [0, TB) combine (FS > 1 only), then XS per ring
  static_assert(N == 2 || N == 4 || N == 8, "N = 2, 4, 8");
  static_assert(N * FS * TB == 8 && RING_BAR0 + NR * XS <= 15, "eight warps; at most 15 named barriers");
};

// position of a warp in its sequence of 16-frame steps: tiles blockIdx.x, + gridDim.x, ...; passes; chunks of the
// warp's frame range; the two halves of a chunk
struct StepIter {
  int tile, pass, c, half;
};

// Fragment conventions as in ssb_coop.cu (PTX m16n8k16): g = lane / 4, t = lane % 4; C: (row g, cols 2t, 2t+1),
// (row g + 8, same); A: a0 (row g, k 2t..), a1 (row g+8, k 2t..), a2 (row g, k 2t+8..), a3; B: b0 (k 2t.., n g), b1.
//
// The kernel is persistent: CTA c works on the tiles c, c + gridDim.x, ... of the (mixture, bin tile) list.  There is
// no separate producer: lane 0 of the source-0 warp of a frame range issues the TMA loads of that range's X ring XS
// stages ahead (named barriers release the stage that is refilled: the other source warps only bar.arrive), and lane 0
// of every warp issues the bulk copies of its own V ring; both run ahead across pass and tile boundaries, so the rings never drain
// while a tile's partial sums are combined.  (A dedicated producer warp polling empty-mbarriers for 16 rings was
// measured 30 - 60 % slower than the cp.async kernels: one thread could not issue the ~220 waits / loads per tile fast
// enough, profiles/r2_tma_variants.md.)
template <int N, int KS, int MODE, int FS_>
__global__ void __launch_bounds__(TileShape<N, KS, FS_>::NT, KS == 1 ? 2 : 1)
    kt_tile(const __grid_constant__ CUtensorMap tmX, cf* Wrw, float* __restrict__ T,
            const __nv_bfloat16* __restrict__ Vs, float* __restrict__ Pout, __nv_bfloat16* __restrict__ Ts,
            cf* __restrict__ U, double* __restrict__ q, int B, int I, int J, int K, int nchunk, int nchunk_i,
            int flooring, float eps, int* __restrict__ status) {
  using S = TileShape<N, KS, FS_>;
  constexpr int FS = S::FS, TB = S::TB, JKS = S::JKS, CHB = S::CHB, XSB = S::XSB, NV = S::NV;
  static_assert(MODE == MODE_BASIS || N == 2, "the covariance modes are written for two sources");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const uint32_t smem_s = (uint32_t)__cvta_generic_to_shared(smem_raw);
  const uint32_t xring_s = smem_s;                          // [fq][stage][XSB]
  const uint32_t vring_s = smem_s + S::X_BYTES;             // [warp][2][CHB]
  float* ucomb0 = reinterpret_cast<float*>(smem_raw + S::X_BYTES + S::V_BYTES);
  float* sc0 = reinterpret_cast<float*>(smem_raw + S::X_BYTES + S::V_BYTES + S::U_BYTES);
  const uint32_t bars_s = smem_s + S::X_BYTES + S::V_BYTES + S::U_BYTES + S::C_BYTES;
  auto xfull = [&](int ring, int st) { return bars_s + (uint32_t)((ring * XS + st) * 8); };
  auto vfull = [&](int w, int st) { return bars_s + (uint32_t)((S::NR * XS + w * 2 + st) * 8); };

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nsteps = J >> 4;
  const int cq = (nchunk + FS - 1) / FS;  // chunks (32 frames) per frame range
  const int ntile_i = (I + 15) >> 4;
  const int gpm = (ntile_i + TB - 1) / TB;  // groups of TB consecutive bin tiles per mixture
  const int ntiles = gpm * B;               // number of groups; "tile" below is a group index, the warp's tile is fixed
  constexpr int NPASS = MODE == MODE_FUSED ? 2 : 1;

  if (threadIdx.x == 0) {
    for (int e = 0; e < S::NBAR; ++e) mbar_init(bars_s + (uint32_t)(e * 8), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmX) : "memory");
  }
  __syncthreads();

  // =================================== consumer warps ===================================
  const int g = lane >> 2, t = lane & 3;
  const int rg = warp / N, n = warp - rg * N;  // ring = (tile slot, frame range); source
  const int tb = rg / FS, fq = rg - tb * FS;
  // group index -> (mixture, first bin) of this warp's tile; the last group of a mixture may have empty tile slots
  auto tile_b = [&](int grp) { return grp / gpm; };
  auto tile_i0 = [&](int grp) { return ((grp - (grp / gpm) * gpm) * TB + tb) * 16; };
  auto next_group = [&](int grp) {  // the next group of this CTA in which the warp's tile slot holds bins
    grp += gridDim.x;
    while (TB > 1 && grp < ntiles && tile_i0(grp) >= I) grp += gridDim.x;
    return grp;
  };
  int first_group = blockIdx.x;
  if (TB > 1 && first_group < ntiles && tile_i0(first_group) >= I) first_group = next_group(first_group);
  const int c_lo = fq * cq, c_hi = min((fq + 1) * cq, nchunk);
  // per-lane shared-memory offsets: X fragment (row g + 8 rr, frames 8 h + 2 t ..) of channel m sits at
  //   stage + h * (N * 1024) + m * 1024 + (g + 8 rr) * 64 + t * 16
  const uint32_t xlane = pin(xring_s + (uint32_t)(rg * XS * XSB) + g * 64 + t * 16);
  const uint32_t vs_s = vring_s + (uint32_t)(warp * 2 * CHB);
  const int mid = lane >> 3, mrow = lane & 7;
  // GEMM1 (non-trans): matrices (hi k0-7, hi k8-15, lo k0-7, lo k8-15) of frames [.., +8)
  const uint32_t l1base = pin(vs_s + (mid >> 1) * (JCV * JKS * 2) + (mrow * JKS + (mid & 1) * 8) * 2);
  // GEMM2 (trans): matrices (hi frames 0-7, hi frames 8-15, lo 0-7, lo 8-15) of basis [.., +8)
  const uint32_t l2base = pin(vs_s + (mid >> 1) * (JCV * JKS * 2) + (((mid & 1) * 8 + mrow) * JKS) * 2);
  float* const ucomb = ucomb0 + tb * S::U_FLOATS;  // this tile's combine scratch
  float* const sc = sc0 + tb * S::C_FLOATS;
  const int rs = t & 1;  // the lane's bin of its pair (rows g, g + 8) in the per-bin algebra between the passes
  int xslot = 0, xphase = 0;  // X stage consumed next by this warp: slot in the ring and phase parity of its barrier
  int vk = 0;                 // V chunks consumed so far by this warp

  // ---- requests: X stages of this frame range (issued by lane 0 of its source-0 warp) and V chunks of this warp ----
  const bool range_active = c_lo < c_hi;
  const bool x_issuer = n == 0 && lane == 0;
  const uint64_t pol = l2_evict_first_policy();
  StepIter xit{first_group, 0, c_lo, 0};  // next X stage to request; every lane tracks it, one lane issues
  auto x_request = [&](int st) {  // into slot st (== requests so far % XS)
    if (xit.tile >= ntiles) return;
    if (x_issuer) {
      const int xb_ = tile_b(xit.tile), ti0 = tile_i0(xit.tile), s = 2 * xit.c + xit.half;
      const uint32_t dst = xring_s + (uint32_t)((rg * XS + st) * XSB), bar = xfull(rg, st);
      mbar_expect_tx(bar, XSB);
      if (MODE == MODE_FUSED && xit.pass == NPASS - 1) {
        // X read for the last time (two-pass mode only, where the first pass must stay in L2 for the second)
        tma_load_3d_hint(dst, &tmX, 32 * s, ti0, xb_ * N, bar, pol);
        tma_load_3d_hint(dst + N * 1024, &tmX, 32 * s + 16, ti0, xb_ * N, bar, pol);
      } else {
        tma_load_3d(dst, &tmX, 32 * s, ti0, xb_ * N, bar);
        tma_load_3d(dst + N * 1024, &tmX, 32 * s + 16, ti0, xb_ * N, bar);
      }
    }
    ++xit.half;
    if (xit.half == 2 || 2 * xit.c + xit.half >= nsteps) {
      xit.half = 0;
      if (++xit.c >= c_hi) {
        xit.c = c_lo;
        if (++xit.pass == NPASS) {
          xit.pass = 0;
          xit.tile = next_group(xit.tile);
        }
      }
    }
  };
  // End of step k: this warp is done reading X stage k (slot k % XS).  The non-issuing source warps only signal that
  // (bar.arrive, non-blocking); the issuing warp waits for them (bar.sync) and requests stage k + XS into the same
  // slot, so XS stages are in flight.  One named barrier per SLOT: a fast warp can run up to XS - 1 steps ahead of the
  // issuer, so its arrivals for steps k + 1, k + 2 must not count towards the barrier of step k; it cannot reach step
  // k + XS before that barrier has completed, because stage k + XS is only requested behind it.
  auto x_release_and_request = [&](int xst) {
    if (n != 0) {
      __syncwarp();
      bar_arrive_id<N * 32>(S::RING_BAR0 + rg * XS + xst);
    } else {
      bar_sync_id<N * 32>(S::RING_BAR0 + rg * XS + xst);
      x_request(xst);
    }
  };
  const unsigned char* vsrc = reinterpret_cast<const unsigned char*>(Vs);
  StepIter vit{first_group, 0, c_lo, 0};  // next V chunk to request
  int vreq = 0;
  auto v_request = [&]() {
    if (vit.tile >= ntiles) return;
    if (lane == 0) {
      const int st = vreq & 1;
      const uint32_t bar = vfull(warp, st);
      mbar_expect_tx(bar, CHB);
      bulk_load(vring_s + (uint32_t)((warp * 2 + st) * CHB),
                vsrc + (((size_t)tile_b(vit.tile) * N + n) * nchunk + vit.c) * (size_t)CHB, CHB, bar);
    }
    ++vreq;
    if (++vit.c >= c_hi) {
      vit.c = c_lo;
      if (++vit.pass == NPASS) {
        vit.pass = 0;
        vit.tile = next_group(vit.tile);
      }
    }
  };
  if (range_active) {
    v_request();
    v_request();
#pragma unroll
    for (int e = 0; e < XS; ++e) x_request(e);
  }

#pragma unroll 1
  for (int tile = first_group; tile < ntiles; tile = next_group(tile)) {
    const int b = tile_b(tile), i0 = tile_i0(tile);
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
    cf w[2][N];  // rows g, g + 8 of W: the filter of source n
    {  // pull the next tile's basis rows (and filters) towards the SM while this tile is processed
      const int nt = next_group(tile);
      if (nt < ntiles && t == 0) {
        const int nb_ = tile_b(nt), ni0 = tile_i0(nt);
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int r = min(ni0 + g + 8 * rr, I - 1);
          asm volatile("prefetch.global.L2 [%0];" ::"l"(T + (((size_t)nb_ * N + n) * I + r) * K));
          if (KS == 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(T + (((size_t)nb_ * N + n) * I + r) * K + 16));
          if (MODE != MODE_COV) asm volatile("prefetch.global.L2 [%0];" ::"l"(Wrw + ((size_t)nb_ * I + r) * N * N));
        }
      }
    }

    // ======================= pass 0: weighted covariance of source n (N = 2) =======================
    if constexpr (MODE != MODE_BASIS) {
      float ua[2][4];  // [rr]: U00, U11, Re U01, Im U01   (U_ac = sum phi x_a conj(x_c))
#pragma unroll
      for (int rr = 0; rr < 2; ++rr)
#pragma unroll
        for (int c = 0; c < 4; ++c) ua[rr][c] = 0.f;
#pragma unroll 1
      for (int c = c_lo; c < c_hi; ++c) {
        const int vst = vk & 1;
        mbar_wait(vfull(warp, vst), (vk >> 1) & 1);
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          const int s = 2 * c + half;
          if (s >= nsteps) break;
          const int xst = xslot;
          const uint32_t vb1 = l1base + vst * CHB + half * (16 * JKS * 2);
          const uint32_t xb = xlane + xst * XSB;
          float Rc[2][4];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) Rc[h][cc] = 0.f;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
              uint32_t bh0, bh1, bl0, bl1;
              ldsm_x4(bh0, bh1, bl0, bl1, vb1 + (8 * h * JKS + ks * 16) * 2);
              mma_split(Rc[h], Thi[ks], Tlo[ks], bh0, bh1, bl0, bl1);
            }
          }
          mbar_wait(xfull(rg, xst), xphase);  // behind the first GEMM, which only needs V
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float (&R)[4] = Rc[h];
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
              const float4 x0 = lds128(xb + h * (N * 1024) + rr * 512);
              const float4 x1 = lds128(xb + h * (N * 1024) + 1024 + rr * 512);
              const float f0 = fast_rcp(R[rr * 2 + 0]), f1 = fast_rcp(R[rr * 2 + 1]);  // no floor on R (ilrma.py:1494-1498)
              ua[rr][0] = fmaf(f0, fmaf(x0.x, x0.x, x0.y * x0.y), fmaf(f1, fmaf(x0.z, x0.z, x0.w * x0.w), ua[rr][0]));
              ua[rr][1] = fmaf(f0, fmaf(x1.x, x1.x, x1.y * x1.y), fmaf(f1, fmaf(x1.z, x1.z, x1.w * x1.w), ua[rr][1]));
              ua[rr][2] = fmaf(f0, fmaf(x0.x, x1.x, x0.y * x1.y), fmaf(f1, fmaf(x0.z, x1.z, x0.w * x1.w), ua[rr][2]));
              ua[rr][3] = fmaf(f0, fmaf(x0.y, x1.x, -x0.x * x1.y), fmaf(f1, fmaf(x0.w, x1.z, -x0.z * x1.w), ua[rr][3]));
            }
          }
          x_release_and_request(xst);
          if (++xslot == XS) {
            xslot = 0;
            xphase ^= 1;
          }
        }
        __syncwarp();
        v_request();  // this warp is done with the chunk: its buffer takes the chunk two ahead
        ++vk;
      }
      // partial sums of this frame range -> shared memory ([source][fq][row][4]); combined in fixed order
#pragma unroll
      for (int rr = 0; rr < 2; ++rr)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float v = ua[rr][c];
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          ua[rr][c] = v;
        }
      if constexpr (MODE == MODE_COV && FS == 1) {
        // one warp holds the complete sums of its source: rows g, g + 8 go straight to U[b, i, n, :, :]
        const float invJf = 1.0f / (float)J;
        if (t == 0) {
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            if (!rvalid[rr]) continue;
            cf* uo = U + (((size_t)b * I + row[rr]) * N + n) * 4;
            uo[0] = make_float2(ua[rr][0] * invJf, 0.f);
            uo[1] = make_float2(ua[rr][2] * invJf, ua[rr][3] * invJf);
            uo[2] = make_float2(ua[rr][2] * invJf, -ua[rr][3] * invJf);
            uo[3] = make_float2(ua[rr][1] * invJf, 0.f);
          }
        }
      } else {
      // The OLD filter of the lane's bin is read before the barrier: behind it the frame range 0 warps of both sources
      // rewrite W while others may still be loading
      cd wm[4];
      if constexpr (MODE == MODE_FUSED) {
        const cf* wold = Wrw + ((size_t)b * I + (rs ? rowc[1] : rowc[0])) * 4;
#pragma unroll
        for (int e = 0; e < 4; ++e) wm[e] = cf2cd(wold[e]);
      }
      if (t == 0) {
#pragma unroll
        for (int rr = 0; rr < 2; ++rr)
          *reinterpret_cast<float4*>(ucomb + ((n * FS + fq) * 16 + g + 8 * rr) * 4) =
              make_float4(ua[rr][0], ua[rr][1], ua[rr][2], ua[rr][3]);
      }
      bar_sync_id<S::TILE_THREADS>(tb);
      // both covariances of the lane's bin, summed over the frame ranges in fp64
      const int r16 = g + 8 * rs;
      const double invJ = 1.0 / (double)J;
      cd u[2][4];  // [source][u00, u01, u10, u11]
#pragma unroll
      for (int sn = 0; sn < 2; ++sn) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
        for (int f = 0; f < FS; ++f) {
          const float4 v = *reinterpret_cast<const float4*>(ucomb + ((sn * FS + f) * 16 + r16) * 4);
          a0 += (double)v.x;
          a1 += (double)v.y;
          a2 += (double)v.z;
          a3 += (double)v.w;
        }
        u[sn][0] = cd_make(a0 * invJ, 0.0);
        u[sn][3] = cd_make(a1 * invJ, 0.0);
        u[sn][1] = cd_make(a2 * invJ, a3 * invJ);
        u[sn][2] = cd_make(a2 * invJ, -a3 * invJ);
      }
      const bool my_valid = rs ? rvalid[1] : rvalid[0];
      if constexpr (MODE == MODE_COV) {
        // U[b, i, n, :, :] complex64 (the consumer is kf_ip1_n2 / kq_ip2)
        if (fq == 0 && t < 2 && my_valid) {
          cf* uo = U + (((size_t)b * I + (rs ? row[1] : row[0])) * N + n) * 4;
#pragma unroll
          for (int e = 0; e < 4; ++e) uo[e] = cd2cf(u[n][e]);
        }
        // ucomb is rewritten at the end of the next tile's pass: every warp must have read it by then
        bar_sync_id<S::TILE_THREADS>(tb);
      } else {
        // IP1, n = 0 then n = 1 with the updated row 0, all in fp64 (kf_ip1_n2); every lane of every frame range computes
        // the same values.  W is written back UNNORMALISED (kf_normalize runs after the activation update, see
        // ssb_fused_spatial_source in ssb_fused.cu for why that order is exact); P below uses the stored complex64 filter.
        cf* wmat = Wrw + ((size_t)b * I + (rs ? rowc[1] : rowc[0])) * 4;
        if (!my_valid) {  // rows past the last bin see a zero slab: keep their algebra finite (nothing is stored)
#pragma unroll
          for (int sn = 0; sn < 2; ++sn) {
            u[sn][0] = u[sn][3] = cd_make(1.0, 0.0);
            u[sn][1] = u[sn][2] = cd_make(0.0, 0.0);
          }
        }
#pragma unroll
        for (int sn = 0; sn < 2; ++sn) {
          const cd a00 = cd_add(cd_mul(wm[0], u[sn][0]), cd_mul(wm[1], u[sn][2]));
          const cd a01 = cd_add(cd_mul(wm[0], u[sn][1]), cd_mul(wm[1], u[sn][3]));
          const cd a10 = cd_add(cd_mul(wm[2], u[sn][0]), cd_mul(wm[3], u[sn][2]));
          const cd a11 = cd_add(cd_mul(wm[2], u[sn][1]), cd_mul(wm[3], u[sn][3]));
          const cd det = cd_sub(cd_mul(a00, a11), cd_mul(a01, a10));
          if (det.x == 0.0 && det.y == 0.0 && my_valid && fq == 0 && t < 2)
            atomicOr(status, SSB_STATUS_SINGULAR);  // np.linalg.solve raises (ssspy/linalg/_solve.py:15)
          const cd idet = cd_inv(det);
          const cd x0 = sn == 0 ? cd_mul(a11, idet) : cd_mul(cd_make(-a01.x, -a01.y), idet);
          const cd x1 = sn == 0 ? cd_mul(cd_make(-a10.x, -a10.y), idet) : cd_mul(a00, idet);
          const cd t0 = cd_add(cd_mul(u[sn][0], x0), cd_mul(u[sn][1], x1));
          const cd t1 = cd_add(cd_mul(u[sn][2], x0), cd_mul(u[sn][3], x1));
          const double qq = cd_mulc(t0, x0).x + cd_mulc(t1, x1).x;
          const double d = ssb_floor(sqrt(fmax(qq, 0.0)), flooring, (double)eps);
          wm[sn * 2 + 0] = cd_scale(cd_conj(x0), 1.0 / d);
          wm[sn * 2 + 1] = cd_scale(cd_conj(x1), 1.0 / d);
        }
        const cf wn0 = cd2cf(wm[n * 2 + 0]), wn1 = cd2cf(wm[n * 2 + 1]);
        if (fq == 0 && t < 2 && my_valid) {
          wmat[n * 2 + 0] = wn0;
          wmat[n * 2 + 1] = wn1;
        }
        // w[rr][m]: own bin from this lane, the other bin of the pair from the neighbour lane (t ^ 1)
        const float o0x = __shfl_xor_sync(0xffffffffu, wn0.x, 1), o0y = __shfl_xor_sync(0xffffffffu, wn0.y, 1);
        const float o1x = __shfl_xor_sync(0xffffffffu, wn1.x, 1), o1y = __shfl_xor_sync(0xffffffffu, wn1.y, 1);
        w[0][0] = rs ? make_float2(o0x, o0y) : wn0;
        w[0][1] = rs ? make_float2(o1x, o1y) : wn1;
        w[1][0] = rs ? wn0 : make_float2(o0x, o0y);
        w[1][1] = rs ? wn1 : make_float2(o1x, o1y);
      }
      }
    } else {
#pragma unroll
      for (int rr = 0; rr < 2; ++rr)
#pragma unroll
        for (int m = 0; m < N; ++m) w[rr][m] = Wrw[(((size_t)b * I + rowc[rr]) * N + n) * N + m];
    }

    // ======================= pass 1: basis update of source n with the filter w =======================
    if constexpr (MODE != MODE_COV) {
      float num[2 * KS][4], den[2 * KS][4];
#pragma unroll
      for (int qi = 0; qi < 2 * KS; ++qi)
#pragma unroll
        for (int c = 0; c < 4; ++c) num[qi][c] = den[qi][c] = 0.f;
      float qs[2] = {0.f, 0.f};
      // P is handed to kf_activation_coop in 16 x 16 tiles ([bn][bin tile][frame tile][16 bins][16 frames], ssb_coop.cu)
      const size_t ptile0 = ((bn * (size_t)ntile_i + (size_t)(i0 >> 4)) * (size_t)(J >> 4)) * 256;
      float* const pout0 = pin_ptr(Pout + ptile0 + g * 16 + 2 * t);
      float* const pout1 = pin_ptr(Pout + ptile0 + (g + 8) * 16 + 2 * t);
#pragma unroll 1
      for (int c = c_lo; c < c_hi; ++c) {
        const int vst = vk & 1;
        mbar_wait(vfull(warp, vst), (vk >> 1) & 1);
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          const int s = 2 * c + half;
          if (s >= nsteps) break;
          const int xst = xslot;
          const uint32_t voff = vst * CHB + half * (16 * JKS * 2);
          const uint32_t vb1 = l1base + voff, vb2 = l2base + voff;
          const uint32_t xb = xlane + xst * XSB;
          // ---- GEMM1: R[16 bins x 16 frames] = T V ----
          float R[2][4];
#pragma unroll
          for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) R[h][cc] = 0.f;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
              uint32_t bh0, bh1, bl0, bl1;
              ldsm_x4(bh0, bh1, bl0, bl1, vb1 + (8 * h * JKS + ks * 16) * 2);
              mma_split(R[h], Thi[ks], Tlo[ks], bh0, bh1, bl0, bl1);
            }
          }
          // ---- elementwise: P = |w^H x|^2, A = P / R^2, B = 1 / R ----
          mbar_wait(xfull(rg, xst), xphase);  // behind the first GEMM, which only needs V
          uint32_t Ahi[4], Alo[4], Bhi[4], Blo[4];
          float* const po[2] = {pout0 + (size_t)s * 256, pout1 + (size_t)s * 256};
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
              float4 x[N];
#pragma unroll
              for (int m = 0; m < N; ++m) x[m] = lds128(xb + h * (N * 1024) + m * 1024 + rr * 512);
              float p0, p1;
              power2<N>(x, w[rr], p0, p1);
              // the power spectrogram is kept for the activation update (same W => same P, ilrma.py:1169-1172)
              *reinterpret_cast<float2*>(po[rr] + 8 * h) = make_float2(p0, p1);
              if constexpr (MODE == MODE_FUSED) qs[rr] += p0 + p1;
              const float i0v = fast_rcp(R[h][rr * 2 + 0]);
              const float i1v = fast_rcp(R[h][rr * 2 + 1]);
              const Split sa = split2(p0 * i0v * i0v, p1 * i1v * i1v);
              const Split sb = split2(i0v, i1v);
              Ahi[h * 2 + rr] = sa.hi;
              Alo[h * 2 + rr] = sa.lo;
              Bhi[h * 2 + rr] = sb.hi;
              Blo[h * 2 + rr] = sb.lo;
            }
          x_release_and_request(xst);
          if (++xslot == XS) {
            xslot = 0;
            xphase ^= 1;
          }
          // ---- GEMM2: num += A V^T, den += B V^T (contraction over the 16 frames) ----
#pragma unroll
          for (int qi = 0; qi < 2 * KS; ++qi) {
            uint32_t vh0, vh1, vl0, vl1;
            ldsm_x4_t(vh0, vh1, vl0, vl1, vb2 + qi * 16);
            mma_split(num[qi], Ahi, Alo, vh0, vh1, vl0, vl1);
            mma_split(den[qi], Bhi, Blo, vh0, vh1, vl0, vl1);
          }
        }
        __syncwarp();
        v_request();
        ++vk;
      }
      // ---- combine the frame ranges (fixed order), then T <- floor(T sqrt(num / den))   (ilrma.py:1125-1126, p = 2) ----
      if constexpr (FS > 1) {
        if (fq > 0) {
          float* dst = sc + ((size_t)(n * (FS - 1) + fq - 1) * NV) * 32 + lane;
#pragma unroll
          for (int qi = 0; qi < 2 * KS; ++qi)
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
              dst[(qi * 4 + cc) * 32] = num[qi][cc];
              dst[((2 * KS + qi) * 4 + cc) * 32] = den[qi][cc];
            }
          dst[(NV - 2) * 32] = qs[0];
          dst[(NV - 1) * 32] = qs[1];
        }
        bar_sync_id<S::TILE_THREADS>(tb);
        if (fq == 0) {
#pragma unroll
          for (int f = 1; f < FS; ++f) {
            const float* src = sc + ((size_t)(n * (FS - 1) + f - 1) * NV) * 32 + lane;
#pragma unroll
            for (int qi = 0; qi < 2 * KS; ++qi)
#pragma unroll
              for (int cc = 0; cc < 4; ++cc) {
                num[qi][cc] += src[(qi * 4 + cc) * 32];
                den[qi][cc] += src[((2 * KS + qi) * 4 + cc) * 32];
              }
            qs[0] += src[(NV - 2) * 32];
            qs[1] += src[(NV - 1) * 32];
          }
        }
      }
      if (fq == 0) {
        if constexpr (MODE == MODE_FUSED) {
          // q[b, i, n] = mean_j |y|^2 with the unnormalised filter (the term kf_ip1_n2 gets from the unweighted covariance)
#pragma unroll
          for (int rr = 0; rr < 2; ++rr) {
            float v = qs[rr];
            v += __shfl_xor_sync(0xffffffffu, v, 1);
            v += __shfl_xor_sync(0xffffffffu, v, 2);
            if (t == 0 && rvalid[rr]) q[((size_t)b * I + row[rr]) * N + n] = (double)v / (double)J;
          }
        }
        // the new basis is also written pre-split (bf16 hi, lo; [bin][basis] chunks of 32 bins) for kf_activation_coop
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
          for (int nb = 0; nb < 2; ++nb)
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
              if (!rvalid[rr]) continue;
              const int qi = ks * 2 + nb;
              const int k0 = ks * 16 + nb * 8 + 2 * t;
              float tn[2];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                tn[e] = 0.f;
                if (k0 + e < K) {
                  const float ratio = num[qi][rr * 2 + e] / den[qi][rr * 2 + e];
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
      // the combine scratch is rewritten at the end of the next tile: the frame range 0 warps must have read it by then
      if constexpr (FS > 1) bar_sync_id<S::TILE_THREADS>(tb);
    }
  }
}

// ---- tensor maps -----------------------------------------------------------------------------------------------
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

// X[B*N planes][I bins][2 J floats]; box = 16 floats (8 frames) x 16 bins x N planes, no swizzle, zero fill
int make_x_map(CUtensorMap* tm, const cf* X, int B, int N, int I, int J) {
  EncodeTiledFn enc = encode_fn();
  SSB_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from this driver");
  SSB_REQUIRE((reinterpret_cast<uintptr_t>(X) & 15) == 0 && (J % 2) == 0, "TMA needs a 16-byte aligned X and even n_frames");
  const cuuint64_t dims[3] = {(cuuint64_t)2 * J, (cuuint64_t)I, (cuuint64_t)B * N};
  const cuuint64_t strides[2] = {(cuuint64_t)J * 8, (cuuint64_t)I * J * 8};  // bytes, dims 1 and 2
  const cuuint32_t box[3] = {16, 16, (cuuint32_t)N};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<cf*>(X), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SSB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) for X[%d,%d,%d,%d]", (int)r, B, N, I, J);
  return 0;
}

template <int N, int KS, int MODE, int FS>
int launch_tile(const ssb_config* c, const cf* X, cf* W, float* T, const __nv_bfloat16* Vs, float* P, __nv_bfloat16* Ts,
                cf* U, double* q, const char* name, cudaStream_t st) {
  using S = TileShape<N, KS, FS>;
  const int B = c->n_batch, I = c->n_bins, J = c->n_frames, K = c->n_basis;
  const int nchunk = (J + JCV - 1) / JCV, nchunk_i = (I + JCV - 1) / JCV;
  CUtensorMap tm;
  if (make_x_map(&tm, X, B, N, I, J)) return 1;
  static int ctas_dev[SSB_MAX_DEVICES] = {};  // resident CTAs of this kernel on the device (0: not configured yet)
  int& ctas = ctas_dev[ssb_current_device()];
  if (ctas == 0) {
    SSB_CUDA(cudaFuncSetAttribute(kt_tile<N, KS, MODE, FS>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::SMEM));
    int dev = 0, sms = 0, per_sm = 0;
    SSB_CUDA(cudaGetDevice(&dev));
    SSB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    SSB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kt_tile<N, KS, MODE, FS>, S::NT, S::SMEM));
    SSB_REQUIRE(per_sm >= 1, "%s: the kernel does not fit an SM (%d bytes of shared memory)", name, S::SMEM);
    ctas = sms * per_sm;
  }
  // FS > 1 (two-pass mode): persistent, one wave of CTAs walks the (mixture, tile group) list, so that at most one tile
  // per CTA waits in L2 for its second pass; FS == 1: one CTA per group, other streams' kernels fill the tail wave
  const long long ngroups = (long long)(((I + 15) / 16 + S::TB - 1) / S::TB) * B;
  const int grid = (int)((FS == 1 || ngroups < ctas) ? ngroups : ctas);
  kt_tile<N, KS, MODE, FS><<<grid, S::NT, S::SMEM, st>>>(tm, W, T, Vs, P, Ts, U, q, B, I, J, K, nchunk, nchunk_i, c->flooring,
                                                       c->eps, ssb_status_word());
  return ssb_check_launch(name, st);
}

}  // namespace

// Which TMA tile kernels run, as a bit mask: 1 = basis kernel, 2 = covariance kernel (N = 2), 4 = the fused covariance +
// IP1 + basis kernel inside ssb_run (N = 2).  SSB_TMA (read once) overrides it for A/B runs (0 = the cp.async kernels
// of ssb_coop.cu / ssb_fused.cu everywhere, 7 = every TMA kernel).  The default follows the B200 measurements of
// profiles/r2_tma_variants.md: at N = 8 (one 16 KB stage per step shared by eight warps) the TMA basis kernel is on par
// with the cp.async one and is used; at N = 2 / 4 (2 - 4 KB stages, one named barrier + one mbarrier wait per 16-frame
// step and warp pair) it is 15 - 20 % slower, so those sizes stay on cp.async until the stages get larger.
int ssb_tma_mask(const ssb_config* c) {
  const char* e = getenv("SSB_TMA");  // read at every call so that tests can toggle it
  if (e != nullptr && atoi(e) >= 0) return atoi(e);
  return c->n_sources == 8 ? 1 : 0;
}

int ssb_tma_supported(const ssb_config* c) {
  const int N = c->n_sources;
  return (N == 2 || N == 4 || N == 8) && (c->n_frames % 16) == 0 && c->n_basis <= 32;
}

#define SSB_TILE_DISPATCH(MODE_, ...)                                                      \
  do {                                                                                     \
    const bool k16 = c->n_basis <= 16;                                                     \
    switch (c->n_sources) {                                                                \
      case 2: return k16 ? launch_tile<2, 1, MODE_, 1>(__VA_ARGS__) : launch_tile<2, 2, MODE_, 1>(__VA_ARGS__); \
      case 4: return k16 ? launch_tile<4, 1, MODE_, 1>(__VA_ARGS__) : launch_tile<4, 2, MODE_, 1>(__VA_ARGS__); \
      case 8: return k16 ? launch_tile<8, 1, MODE_, 1>(__VA_ARGS__) : launch_tile<8, 2, MODE_, 1>(__VA_ARGS__); \
      default: break;                                                                      \
    }                                                                                      \
  } while (0)

// basis update of every source (reads the pre-split activation Vs, writes T, the pre-split basis Ts and P)
int ssb_tma_basis(const ssb_config* c, const cf* X, const cf* W, float* T, const void* Vs, float* P, void* Ts,
                  cudaStream_t st) {
  SSB_REQUIRE(ssb_tma_supported(c), "tma_basis: unsupported configuration");
  SSB_TILE_DISPATCH(MODE_BASIS, c, X, const_cast<cf*>(W), T, (const __nv_bfloat16*)Vs, P, (__nv_bfloat16*)Ts, nullptr,
                    nullptr, "tma_basis", st);
  return 1;
}

// N = 2: weighted covariances U[B,I,2,2,2] from the pre-split activation Vs
int ssb_tma_cov_n2(const ssb_config* c, const cf* X, float* T, const void* Vs, cf* U, cudaStream_t st) {
  SSB_REQUIRE(ssb_tma_supported(c) && c->n_sources == 2, "tma_cov: unsupported configuration");
  if (c->n_basis <= 16)
    return launch_tile<2, 1, MODE_COV, 1>(c, X, nullptr, T, (const __nv_bfloat16*)Vs, nullptr, nullptr, U, nullptr,
                                          "tma_phi_cov", st);
  return launch_tile<2, 2, MODE_COV, 1>(c, X, nullptr, T, (const __nv_bfloat16*)Vs, nullptr, nullptr, U, nullptr,
                                        "tma_phi_cov", st);
}

// N = 2 inside ssb_run: covariance + IP1 of iteration t (W rewritten unnormalised, q emitted) and the basis update of
// iteration t + 1 on the same tile
int ssb_tma_spatial_basis_n2(const ssb_config* c, const cf* X, cf* W, float* T, const void* Vs, float* P, void* Ts,
                             double* q, cudaStream_t st) {
  SSB_REQUIRE(ssb_tma_supported(c) && c->n_sources == 2 && q != nullptr, "tma_spatial_basis: unsupported configuration");
  if (c->n_basis <= 16)
    return launch_tile<2, 1, MODE_FUSED, 4>(c, X, W, T, (const __nv_bfloat16*)Vs, P, (__nv_bfloat16*)Ts, nullptr, q,
                                            "tma_cov_ip1_basis", st);
  return launch_tile<2, 2, MODE_FUSED, 4>(c, X, W, T, (const __nv_bfloat16*)Vs, P, (__nv_bfloat16*)Ts, nullptr, q,
                                          "tma_cov_ip1_basis", st);
}
