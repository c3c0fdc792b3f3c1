// Register-resident small dense complex linear algebra on lane groups.
//
// A group of GS = 4 (N <= 4) or 8 (N <= 8) consecutive lanes owns one N x N problem: lane r holds row
// r of the matrix in registers (fp64 complex), rows are exchanged with warp shuffles, and 32/GS
// problems run side by side in a warp.  No shared-memory round trips and no __syncwarp inside the
// elimination, which is what made the warp-per-matrix shared-memory Gauss-Jordan slow for N >= 3
// (profiles/r1_bench_configs_a.jsonl: update_by_ip1 1.05 ms at N = 8).
#pragma once
#include "ssb_common.cuh"

template <int N>
struct GroupShape {
  static constexpr int GS = N <= 4 ? 4 : 8;   // lanes per problem
  static constexpr int GW = 32 / GS;          // problems per warp
};

__device__ __forceinline__ cd shfl_cd(cd v, int src) {
  return make_double2(__shfl_sync(SSB_FULL, v.x, src), __shfl_sync(SSB_FULL, v.y, src));
}

template <int GS>
__device__ __forceinline__ double group_sum(double v) {
#pragma unroll
  for (int o = GS / 2; o > 0; o >>= 1) v += __shfl_xor_sync(SSB_FULL, v, o);
  return v;
}

// Gauss-Jordan with partial pivoting on a row-distributed system [A | RHS]:
//   a[c]   : row r of A (lane r of the group; rows r >= N must be passed as zeros)
//   rhs[q] : row r of the R right-hand sides
// On return lane r (r < N) holds component r of every solution column in rhs[q]
// (x = A^-1 RHS).  gbase = first lane of the group inside the warp, r = lane - gbase.
// logabs (optional) accumulates log|det A|.
template <int N, int R, int GS>
__device__ __forceinline__ void group_solve(cd (&a)[N], cd (&rhs)[R], int r, int gbase, double* logabs = nullptr,
                                            bool* singular = nullptr) {
  bool used = false;
  bool sing = false;
  int mysrc = r;  // which lane ends up holding solution component r
  double lad = 0.0;
#pragma unroll
  for (int p = 0; p < N; ++p) {
    // pivot: unused row with the largest |a[p]|
    double v = (used || r >= N) ? -1.0 : cd_abs2(a[p]);
    int l = r;
#pragma unroll
    for (int o = GS / 2; o > 0; o >>= 1) {
      const double ov = __shfl_xor_sync(SSB_FULL, v, o);
      const int ol = __shfl_xor_sync(SSB_FULL, l, o);
      if (ov > v || (ov == v && ol < l)) {
        v = ov;
        l = ol;
      }
    }
    const int L = gbase + l;
    sing = sing || (v == 0.0);  // the largest remaining pivot candidate is exactly zero (uniform over the group)
    const cd pv = shfl_cd(a[p], L);
    lad += 0.5 * log(cd_abs2(pv));
    const cd ipv = cd_inv(pv);
    const bool is_piv = (r == l);
    const cd f = a[p];
#pragma unroll
    for (int c = p + 1; c < N; ++c) {
      const cd pc = cd_mul(shfl_cd(a[c], L), ipv);
      a[c] = is_piv ? pc : cd_sub(a[c], cd_mul(f, pc));
    }
#pragma unroll
    for (int q = 0; q < R; ++q) {
      const cd pc = cd_mul(shfl_cd(rhs[q], L), ipv);
      rhs[q] = is_piv ? pc : cd_sub(rhs[q], cd_mul(f, pc));
    }
    if (is_piv) used = true;
    if (r == p) mysrc = l;
  }
  // lane `mysrc` holds component r: bring it home
#pragma unroll
  for (int q = 0; q < R; ++q) rhs[q] = shfl_cd(rhs[q], gbase + mysrc);
  if (logabs) *logabs = lad;
  if (singular) *singular = sing;
}
