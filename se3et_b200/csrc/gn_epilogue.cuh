// GroupNorm statistics accumulated inside a tensor-core epilogue (shared by gemm.cu and kpconv_fused.cu):
// every epilogue thread owns one output row and walks its columns in chunks of 32 (or 16); per-warp partial sums go
// to shared memory and each tile issues one set of fp64 atomics per (pair, group).
#pragma once
#include "common.cuh"

namespace se3et {

// Sums N per-lane values across the 32 lanes of a warp with N/2 + N/4 + ... shuffles (each exchange halves the
// number of values a lane still carries).  Afterwards lane l (with the low 5 - log2(N) bits clear) holds the
// total of value index l >> (5 - log2(N)) in v[0].
template <int N>
__device__ __forceinline__ void warp_multi_reduce(float (&v)[N], int lane) {
  static_assert(N >= 1 && N <= 32 && (N & (N - 1)) == 0, "N must be a power of two");
  int off = 16;
#pragma unroll
  for (int n = N; n > 1; n >>= 1, off >>= 1) {
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float send = upper ? v[i] : v[i + n / 2];
      const float keep = upper ? v[i + n / 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  for (; off > 0; off >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
}

// GroupNorm partial sums of one epilogue chunk: `vals` = the thread's 32 (or kCols) consecutive columns of its row.
// kGroups = groups inside the chunk (kCols / cpg, or 1 when cpg >= kCols).
template <int kCols, int kGroups>
__device__ __forceinline__ void gn_chunk_partials(const float (&vals)[kCols], bool row_ok, bool uniform, int lane,
                                                   float* warp_acc /* smem [64][2] of this warp */, int group0_local,
                                                   double* stats_row /* global, this row's pair, or null */,
                                                   int group0_global) {
  constexpr int kPer = kCols / kGroups;
  float s[kGroups], ss[kGroups];
#pragma unroll
  for (int g = 0; g < kGroups; ++g) {
    float a = 0.f, b = 0.f;
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      const float x = vals[g * kPer + j];
      a += x;
      b += x * x;
    }
    s[g] = row_ok ? a : 0.f;
    ss[g] = row_ok ? b : 0.f;
  }
  if (uniform) {
    warp_multi_reduce<kGroups>(s, lane);
    warp_multi_reduce<kGroups>(ss, lane);
    constexpr int kShift = kGroups == 32 ? 0 : kGroups == 16 ? 1 : kGroups == 8 ? 2 : kGroups == 4 ? 3 : kGroups == 2 ? 4 : 5;
    if ((lane & ((1 << kShift) - 1)) == 0) {
      const int g = lane >> kShift;
      warp_acc[2 * (group0_local + g)] += s[0];
      warp_acc[2 * (group0_local + g) + 1] += ss[0];
    }
  } else if (row_ok && stats_row) {
    // tile straddles a pair boundary (one tile per boundary): every row adds straight to its own pair
#pragma unroll
    for (int g = 0; g < kGroups; ++g) {
      atomicAdd(stats_row + 2 * (group0_global + g), (double)s[g]);
      atomicAdd(stats_row + 2 * (group0_global + g) + 1, (double)ss[g]);
    }
  }
}


// dispatch on channels-per-group: a power of two <= kCols, or a multiple of kCols (one group per chunk)
template <int kCols>
__device__ __forceinline__ void gn_accumulate_chunk(int cpg, const float (&v)[kCols], bool row_ok, bool uniform, int lane,
                                                    float* warp_acc, int g_loc, double* row_stats, int g_glob) {
#define SE3ET_GN_CASE(G) gn_chunk_partials<kCols, (G)>(v, row_ok, uniform, lane, warp_acc, g_loc, row_stats, g_glob)
  if (cpg >= kCols) SE3ET_GN_CASE(1);
  else if (cpg * 2 == kCols) SE3ET_GN_CASE(2);
  else if (cpg * 4 == kCols) SE3ET_GN_CASE(4);
  else if (cpg * 8 == kCols) SE3ET_GN_CASE(8);
  else if (cpg * 16 == kCols) SE3ET_GN_CASE(16);
  else if (kCols == 32) SE3ET_GN_CASE(kCols == 32 ? 32 : 1);
#undef SE3ET_GN_CASE
}

}  // namespace se3et
