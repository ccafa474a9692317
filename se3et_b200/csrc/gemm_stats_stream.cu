// Streaming statistics pass of Linear + GroupNorm blocks (UnaryBlockEPN / the ResNet shortcut, blocks_epn.py:639-665,
// 833-852): per pair and group the sum and the sum of squares of y = A W^T + b over the pair's rows -- y itself is
// never stored (the apply pass recomputes it, gemm_stream.cu).  The pass only has to read A once, so the kernel is the
// streaming GEMM of gemm_stream.cu with a different epilogue:
//
// Persistent CTAs (one per SM) walk contiguous ranges of 128 x 64 tiles, ROW TILES FASTEST for a fixed column tile:
//   warp 0     TMA producer, 6-stage ring of (A 128 x 64, W 64 x 64) K-blocks running ahead across tiles
//   warp 1     tcgen05.mma issuer, accumulators double-buffered in TMEM
//   warps 2-9  epilogue: warp w owns TMEM lanes 32 (w % 4) and 32 of the tile's 64 columns; every thread keeps fp32
//              running sums  sum y  and  sum y^2  of ITS row for the 32 columns across all tiles of the same pair
//              and column tile (64 registers).  Only when the pair or the column tile changes (or at the end) the warp
//              transposes-and-reduces them with 2 x 31 shuffles, folds the bias in, sums each group's columns and
//              issues one pair of fp64 atomics per group.  Per tile the epilogue is 2 tcgen05.ld + 64 FMAs per thread.
// Rows whose warp straddles a pair boundary take a per-row path (rare: at most once per pair and warp).
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc.cuh"

namespace se3et {

int make_tmap_bf16_2d(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows);

namespace gss {

constexpr int kBM = 128, kBN = 64, kBK = 64;
constexpr int kStages = 6;      // one CTA per SM (the running sums take 64 registers per epilogue thread): deep ring
constexpr int kEpWarps = 8;
constexpr int kThreads = (2 + kEpWarps) * 32;
constexpr int kABytes = kBM * 128, kBBytes = kBN * 128, kStageBytes = kABytes + kBBytes;
constexpr int kBarOff = kStages * kStageBytes;
constexpr int kSmem = kBarOff + 128 + 1024;

struct Args {
  int M, N, K;
  int m_tiles, n_tiles;
  const float* bias;      // nullable
  double* stats;          // [nseg, groups, 2]
  const int64_t* seg_off;
  int nseg, cpg, groups, rpp;
};

// lane j ends up with the sum over all lanes of v[j] (v is destroyed)
__device__ __forceinline__ float transpose_reduce32(float (&v)[32], int lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float keep = up ? v[i + o] : v[i];
      const float send = up ? v[i] : v[i + o];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}

// column totals (one column per lane) -> group totals -> fp64 atomics
__device__ __forceinline__ void flush_columns(const Args& a, int seg, int col, float sy, float syy, float rows,
                                              int lane) {
  const bool col_ok = col < a.N;
  double s = 0.0, q = 0.0;
  if (col_ok) {
    const double b = a.bias ? (double)__ldg(a.bias + col) : 0.0;
    s = (double)sy + (double)rows * b;
    q = (double)syy + 2.0 * b * (double)sy + (double)rows * b * b;
  }
  const int span = a.cpg < 32 ? a.cpg : 32;  // lanes per group inside this warp's 32 columns (power of two)
  for (int o = 1; o < span; o <<= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (col_ok && (lane & (span - 1)) == 0) {
    double* dst = a.stats + ((int64_t)seg * a.groups + col / a.cpg) * 2;
    atomicAdd(dst, s);
    atomicAdd(dst + 1, q);
  }
}

__global__ void __launch_bounds__(kThreads, 1)
gnstats_stream_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, Args args) {
  constexpr uint32_t kAcc = kBN;
  constexpr uint32_t kTmemAlloc = 2 * kAcc;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kBarOff);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;  // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (args.K + kBK - 1) / kBK;
  const int64_t W = (int64_t)args.m_tiles * args.n_tiles;
  const int64_t w_begin = W * blockIdx.x / gridDim.x, w_end = W * (blockIdx.x + 1) / gridDim.x;

  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&tma_a);
    tc::tma_prefetch_desc(&tma_b);
    for (int s = 0; s < kStages; ++s) {
      tc::mbar_init(&full_bar[s], 1);
      tc::mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(&tmem_full_bar[b], 1);
      tc::mbar_init(&tmem_empty_bar[b], kEpWarps);
    }
    tc::mbar_fence_init();
  }
  if (warp == 1) tc::tmem_alloc<kTmemAlloc>(tmem_ptr);
  tc::tcgen05_fence_before_sync();
  __syncthreads();
  tc::tcgen05_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t phase = 0;
      for (int64_t w = w_begin; w < w_end; ++w) {
        const int nt = (int)(w / args.m_tiles), mt = (int)(w - (int64_t)nt * args.m_tiles);
        for (int kb = 0; kb < nkb; ++kb) {
          tc::mbar_wait_long(&empty_bar[s], phase ^ 1);
          tc::mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
          uint8_t* a_dst = smem + s * kStageBytes;
          tc::tma_load_2d(a_dst, &tma_a, &full_bar[s], kb * kBK, mt * kBM);
          tc::tma_load_2d(a_dst + kABytes, &tma_b, &full_bar[s], kb * kBK, nt * kBN);
          if (++s == kStages) { s = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = tc::umma_idesc_bf16(kBM, kBN);
      int s = 0;
      uint32_t phase = 0, it = 0;
      for (int64_t w = w_begin; w < w_end; ++w, ++it) {
        const uint32_t buf = it & 1u, use = it >> 1;
        tc::mbar_wait_long(&tmem_empty_bar[buf], (use & 1u) ^ 1u);
        tc::tcgen05_fence_after_sync();
        for (int kb = 0; kb < nkb; ++kb) {
          tc::mbar_wait_long(&full_bar[s], phase);
          tc::tcgen05_fence_after_sync();
          const uint32_t a_addr = tc::smem_u32(smem + s * kStageBytes);
          const uint64_t a_desc = tc::umma_desc_sw128(a_addr);
          const uint64_t b_desc = tc::umma_desc_sw128(a_addr + kABytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            tc::umma_bf16(tmem_base + buf * kAcc, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc,
                          (kb | k) != 0);
          tc::umma_commit(&empty_bar[s]);
          if (++s == kStages) { s = 0; phase ^= 1; }
        }
        tc::umma_commit(&tmem_full_bar[buf]);
      }
    }
  } else {
    const int e = warp - 2;
    const int lg = warp & 3;   // TMEM lane group this warp may read
    const int ch = e >> 2;     // which 32 of the tile's 64 columns
    float sy[32], syy[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) { sy[j] = 0.f; syy[j] = 0.f; }
    float rows_acc = 0.f;           // rows this thread accumulated since the last flush
    int acc_seg = -1, acc_nt = -1;  // what the running sums belong to (-1: empty)
    int64_t seg_lo = 0, seg_hi = -1;
    int seg_cached = 0;
    auto flush = [&]() {
      if (acc_seg >= 0) {
        const float c1 = transpose_reduce32(sy, lane);
        const float c2 = transpose_reduce32(syy, lane);
        float r = rows_acc;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
        flush_columns(args, acc_seg, acc_nt * kBN + ch * 32 + lane, c1, c2, r, lane);
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) { sy[j] = 0.f; syy[j] = 0.f; }
      rows_acc = 0.f;
      acc_seg = -1;
    };
    uint32_t it = 0;
    for (int64_t w = w_begin; w < w_end; ++w, ++it) {
      const uint32_t buf = it & 1u, use = it >> 1;
      const int nt = (int)(w / args.m_tiles), mt = (int)(w - (int64_t)nt * args.m_tiles);
      const int64_t wfirst = (int64_t)mt * kBM + lg * 32;
      const int64_t row = wfirst + lane;
      const bool row_ok = row < args.M;
      const bool warp_ok = wfirst < args.M;
      bool uniform = true;
      if (warp_ok) {
        const int64_t wlast = min(wfirst + 31, (int64_t)args.M - 1);
        if (!(wfirst >= seg_lo && wlast < seg_hi)) {
          seg_cached = segment_of(args.seg_off, args.nseg, wfirst / args.rpp);
          seg_lo = args.seg_off[seg_cached] * args.rpp;
          seg_hi = args.seg_off[seg_cached + 1] * args.rpp;
        }
        uniform = wlast < seg_hi;
        if (acc_seg >= 0 && (!uniform || acc_seg != seg_cached || acc_nt != nt)) flush();
      }
      tc::mbar_wait_long(&tmem_full_bar[buf], use & 1u);
      tc::tcgen05_fence_after_sync();
      if (warp_ok) {
        uint32_t r0[16], r1[16];
        const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + buf * kAcc + (uint32_t)(ch * 32);
        tc::tmem_ld_32x32b_x16(taddr, r0);
        tc::tmem_ld_32x32b_x16(taddr + 16, r1);
        tc::tmem_ld_wait();
        if (uniform) {
          if (row_ok) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float a = __uint_as_float(r0[j]), b = __uint_as_float(r1[j]);
              sy[j] += a;
              syy[j] = fmaf(a, a, syy[j]);
              sy[16 + j] += b;
              syy[16 + j] = fmaf(b, b, syy[16 + j]);
            }
            rows_acc += 1.f;
          }
          acc_seg = seg_cached;
          acc_nt = nt;
        } else if (row_ok) {
          // the warp's rows straddle a pair boundary: this row alone, group by group
          const int row_seg = segment_of(args.seg_off, args.nseg, row / args.rpp);
          const int col0 = nt * kBN + ch * 32;
          const int span = args.cpg < 32 ? args.cpg : 32;
          for (int c0 = 0; c0 < 32 && col0 + c0 < args.N; c0 += span) {
            double s = 0.0, q = 0.0;
            for (int j = c0; j < c0 + span; ++j) {
              // static register indexing: select by unrolled compare
              float v = 0.f;
#pragma unroll
              for (int t = 0; t < 16; ++t) {
                if (j == t) v = __uint_as_float(r0[t]);
                if (j == 16 + t) v = __uint_as_float(r1[t]);
              }
              const double y = (double)v + (args.bias ? (double)__ldg(args.bias + col0 + j) : 0.0);
              s += y;
              q += y * y;
            }
            double* dst = args.stats + ((int64_t)row_seg * args.groups + (col0 + c0) / args.cpg) * 2;
            atomicAdd(dst, s);
            atomicAdd(dst + 1, q);
          }
        }
      }
      tc::tcgen05_fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tmem_empty_bar[buf]);
    }
    flush();
  }
  tc::tcgen05_fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<kTmemAlloc>(tmem_base);
}

}  // namespace gss

// ---------------------------------------------------------------------------------------------------------------
// Transposed form (round 2): D^T = W A^T, i.e. UMMA M = output columns (a 128-row W tile, zero-filled by TMA past N),
// UMMA N = 128 matrix rows.  The accumulator then has TMEM lanes = output columns and TMEM columns = rows, so the sum
// over rows that GroupNorm needs is a serial sum inside each epilogue thread: TWO running registers per column block
// instead of 64, no transposition, pair boundaries are a uniform split of the column loop, and up to 256 output columns
// are covered in ONE pass over A (the first kernel re-read A once per 64 output columns: 0.7-1.9 TB/s on the deep
// levels).  Both operands stay K-major exactly as they lie in memory.
//   warp 0     TMA producer: per K-block the A tile (128 rows) and NB W tiles (128 output columns each)
//   warp 1     tcgen05.mma issuer: NB accumulators (128 lanes x 128 columns), double-buffered: NB * 256 TMEM columns
//   warps 2-9  epilogue: lane quarter warp % 4 (32 output columns of every column block), row half (warp - 2) / 4
// ---------------------------------------------------------------------------------------------------------------
namespace gst {

constexpr int kBM = 128;   // matrix rows per tile (UMMA N)
constexpr int kBW = 128;   // output columns per W tile (UMMA M)
constexpr int kBK = 64;
constexpr int kEpWarps = 8;
constexpr int kThreads = (2 + kEpWarps) * 32;
constexpr int kTileBytes = 128 * 128;  // 128 rows x 64 bf16

template <int NB>
struct Cfg {
  static constexpr int kStages = NB == 1 ? 6 : 4;
  static constexpr int kStageBytes = (1 + NB) * kTileBytes;
  static constexpr int kBarOff = kStages * kStageBytes;
  static constexpr int kSmem = kBarOff + 128 + 1024;
};

struct Args {
  int M, N, K;
  int m_tiles, n_pass;    // row tiles, passes over A (one per NB * 128 output columns)
  const float* bias;      // nullable
  double* stats;          // [nseg, groups, 2]
  const int64_t* seg_off;
  int nseg, cpg, groups, rpp;
};

// one column per lane: bias folded in, columns of a group summed, fp64 atomics
__device__ __forceinline__ void flush_column(const Args& a, int seg, int col, double sy, double syy, double rows,
                                             int lane) {
  const bool col_ok = col < a.N;
  double s = 0.0, q = 0.0;
  if (col_ok) {
    const double b = a.bias ? (double)__ldg(a.bias + col) : 0.0;
    s = sy + rows * b;
    q = syy + 2.0 * b * sy + rows * b * b;
  }
  const int span = a.cpg < 32 ? a.cpg : 32;  // lanes per group inside this warp's 32 columns (power of two)
  for (int o = 1; o < span; o <<= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
  if (col_ok && (lane & (span - 1)) == 0) {
    double* dst = a.stats + ((int64_t)seg * a.groups + col / a.cpg) * 2;
    atomicAdd(dst, s);
    atomicAdd(dst + 1, q);
  }
}

template <int NB>
__global__ void __launch_bounds__(kThreads, 1)
gnstats_stream_t_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_w, Args args) {
  using C = Cfg<NB>;
  constexpr uint32_t kAcc = 128;                 // TMEM columns per accumulator
  constexpr uint32_t kTmemAlloc = NB * 2 * kAcc;  // 256 or 512
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::kBarOff);
  uint64_t* empty_bar = full_bar + C::kStages;
  uint64_t* tmem_full_bar = empty_bar + C::kStages;  // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (args.K + kBK - 1) / kBK;
  // work items (pass, row tile), row tiles fastest; contiguous ranges per CTA
  const int64_t W = (int64_t)args.m_tiles * args.n_pass;
  const int64_t w_begin = W * blockIdx.x / gridDim.x, w_end = W * (blockIdx.x + 1) / gridDim.x;

  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&tma_a);
    tc::tma_prefetch_desc(&tma_w);
    for (int s = 0; s < C::kStages; ++s) {
      tc::mbar_init(&full_bar[s], 1);
      tc::mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(&tmem_full_bar[b], 1);
      tc::mbar_init(&tmem_empty_bar[b], kEpWarps);
    }
    tc::mbar_fence_init();
  }
  if (warp == 1) tc::tmem_alloc<kTmemAlloc>(tmem_ptr);
  tc::tcgen05_fence_before_sync();
  __syncthreads();
  tc::tcgen05_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0;
      uint32_t phase = 0;
      for (int64_t w = w_begin; w < w_end; ++w) {
        const int np = (int)(w / args.m_tiles), mt = (int)(w - (int64_t)np * args.m_tiles);
        for (int kb = 0; kb < nkb; ++kb) {
          tc::mbar_wait_long(&empty_bar[s], phase ^ 1);
          tc::mbar_arrive_expect_tx(&full_bar[s], C::kStageBytes);
          uint8_t* dst = smem + s * C::kStageBytes;
          tc::tma_load_2d(dst, &tma_a, &full_bar[s], kb * kBK, mt * kBM);
#pragma unroll
          for (int nb = 0; nb < NB; ++nb)
            tc::tma_load_2d(dst + (1 + nb) * kTileBytes, &tma_w, &full_bar[s], kb * kBK, (np * NB + nb) * kBW);
          if (++s == C::kStages) { s = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = tc::umma_idesc_bf16(kBW, kBM);  // M = output columns, N = matrix rows
      int s = 0;
      uint32_t phase = 0, it = 0;
      for (int64_t w = w_begin; w < w_end; ++w, ++it) {
        const uint32_t buf = it & 1u, use = it >> 1;
        tc::mbar_wait_long(&tmem_empty_bar[buf], (use & 1u) ^ 1u);
        tc::tcgen05_fence_after_sync();
        for (int kb = 0; kb < nkb; ++kb) {
          tc::mbar_wait_long(&full_bar[s], phase);
          tc::tcgen05_fence_after_sync();
          const uint32_t a_addr = tc::smem_u32(smem + s * C::kStageBytes);
          const uint64_t rows_desc = tc::umma_desc_sw128(a_addr);
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) {
            const uint64_t w_desc = tc::umma_desc_sw128(a_addr + (1 + nb) * kTileBytes);
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              tc::umma_bf16(tmem_base + (buf * NB + nb) * kAcc, w_desc + (uint64_t)(k * 2), rows_desc + (uint64_t)(k * 2),
                            idesc, (kb | k) != 0);
          }
          tc::umma_commit(&empty_bar[s]);
          if (++s == C::kStages) { s = 0; phase ^= 1; }
        }
        tc::umma_commit(&tmem_full_bar[buf]);
      }
    }
  } else {
    const int e = warp - 2;
    const int lg = warp & 3;  // TMEM lane quarter this warp may read = 32 output columns of every column block
    const int hb = e >> 2;    // which 64 of the tile's 128 rows (TMEM columns)
    double sy[NB], syy[NB];   // running sums of this thread's output column(s) over the rows of (acc_seg, acc_np)
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) sy[nb] = syy[nb] = 0.0;
    double rows_acc = 0.0;
    int acc_seg = -1, acc_np = -1;
    int64_t seg_lo = 0, seg_hi = -1;  // row range of seg_cached
    int seg_cached = 0;
    auto flush = [&]() {
      if (acc_seg >= 0) {
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
          flush_column(args, acc_seg, (acc_np * NB + nb) * kBW + lg * 32 + lane, sy[nb], syy[nb], rows_acc, lane);
      }
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) sy[nb] = syy[nb] = 0.0;
      rows_acc = 0.0;
      acc_seg = -1;
    };
    uint32_t it = 0;
    for (int64_t w = w_begin; w < w_end; ++w, ++it) {
      const uint32_t buf = it & 1u, use = it >> 1;
      const int np = (int)(w / args.m_tiles), mt = (int)(w - (int64_t)np * args.m_tiles);
      const int64_t r_first = (int64_t)mt * kBM + hb * 64;          // this warp's rows of the tile: [r_first, r_end)
      const int64_t r_end = min(r_first + 64, (int64_t)args.M);
      tc::mbar_wait_long(&tmem_full_bar[buf], use & 1u);
      tc::tcgen05_fence_after_sync();
      // lane quarters past the last output column hold the zero rows TMA filled in: nothing to sum
      const bool active = (int64_t)np * NB * kBW + lg * 32 < args.N;
      if (active && r_first < r_end) {
        int64_t r = r_first;
        while (r < r_end) {
          if (!(r >= seg_lo && r < seg_hi)) {
            seg_cached = segment_of(args.seg_off, args.nseg, r / args.rpp);
            seg_lo = args.seg_off[seg_cached] * args.rpp;
            seg_hi = seg_cached == args.nseg - 1 ? (int64_t)args.M : args.seg_off[seg_cached + 1] * args.rpp;
          }
          if (acc_seg >= 0 && (acc_seg != seg_cached || acc_np != np)) flush();
          const int64_t r_stop = min(r_end, seg_hi);
          const int j0 = (int)(r - r_first), j1 = (int)(r_stop - r_first);
#pragma unroll
          for (int nb = 0; nb < NB; ++nb) {
            // the 64 accumulator columns (= rows of the matrix) of this column block -> registers (re-read per piece
            // when a pair boundary splits the 64 rows: rare)
            uint32_t v[64];
            const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (buf * NB + nb) * kAcc + (uint32_t)(hb * 64);
            tc::tmem_ld_32x32b_x32(taddr, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
            tc::tmem_ld_32x32b_x32(taddr + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
            tc::tmem_ld_wait();
            // four independent partial sums per statistic: 64 dependent adds would serialise on the FADD latency
            float s4[4] = {0.f, 0.f, 0.f, 0.f}, q4[4] = {0.f, 0.f, 0.f, 0.f};
            if (j0 == 0 && j1 == 64) {
#pragma unroll
              for (int j = 0; j < 64; ++j) {
                const float y = __uint_as_float(v[j]);
                s4[j & 3] += y;
                q4[j & 3] = fmaf(y, y, q4[j & 3]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 64; ++j) {
                const float y = (j >= j0 && j < j1) ? __uint_as_float(v[j]) : 0.f;
                s4[j & 3] += y;
                q4[j & 3] = fmaf(y, y, q4[j & 3]);
              }
            }
            sy[nb] += (double)((s4[0] + s4[1]) + (s4[2] + s4[3]));
            syy[nb] += (double)((q4[0] + q4[1]) + (q4[2] + q4[3]));
          }
          rows_acc += (double)(j1 - j0);
          acc_seg = seg_cached;
          acc_np = np;
          r = r_stop;
        }
      }
      tc::tcgen05_fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tmem_empty_bar[buf]);
    }
    flush();
  }
  tc::tcgen05_fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<kTmemAlloc>(tmem_base);
}

template <int NB>
static int launch(const void* a, int64_t lda, int64_t m, int64_t k, const void* w, int64_t ldw, int64_t n, Args args,
                  cudaStream_t st) {
  CUtensorMap ta, tw;
  int rc = make_tmap_bf16_2d(&ta, a, m, k, lda, kBM);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tw, w, n, k, ldw, kBW);
  if (rc) return rc;
  args.n_pass = (int)ceil_div(n, (int64_t)NB * kBW);
  SE3ET_ENSURE_SMEM(gnstats_stream_t_kernel<NB>, Cfg<NB>::kSmem);
  const int64_t work = (int64_t)args.m_tiles * args.n_pass;
  const unsigned grid = (unsigned)(work < kNumSMs ? work : kNumSMs);
  gnstats_stream_t_kernel<NB><<<grid, kThreads, Cfg<NB>::kSmem, st>>>(ta, tw, args);
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}

}  // namespace gst
}  // namespace se3et

using namespace se3et;

extern "C" int se3et_linear_gnstats_stream(const void* a, int64_t lda, int64_t m, int64_t k, const void* w_bf16,
                                           int64_t ldw, int64_t n, const float* bias, const int64_t* seg_offsets,
                                           int64_t nseg, int64_t groups, int64_t rows_per_point, double* stats,
                                           se3et_stream_t stream) {
  if (!a || !w_bf16 || !seg_offsets || !stats || m < 0 || n <= 0 || k <= 0 || nseg <= 0 || groups <= 0 || n % groups ||
      rows_per_point <= 0 || rows_per_point > INT32_MAX)
    return SE3ET_ERR_ARG;
  const int64_t cpg = n / groups;
  // a group is either a power-of-two run of columns inside one warp's 32 columns, or a multiple of 32 columns
  if (n % 32 != 0 || k % 8 != 0 || !(((cpg & (cpg - 1)) == 0 && cpg <= 32) || cpg % 32 == 0)) return SE3ET_ERR_UNSUPPORTED;
  if ((int64_t)ceil_div(m, gss::kBM) * ceil_div(n, gss::kBN) > INT32_MAX) return SE3ET_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SE3ET_CUDA_CHECK(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * nseg * groups, st));
  if (m == 0) return SE3ET_OK;
  static const bool transposed = [] {
    const char* e = getenv("SE3ET_STATS_TRANSPOSED");  // A/B switch: 0 = the first (rows = lanes) kernel
    return !(e && e[0] == '0');
  }();
  if (transposed) {
    gst::Args ta;
    ta.M = (int)m; ta.N = (int)n; ta.K = (int)k;
    ta.m_tiles = (int)ceil_div(m, gst::kBM);
    ta.n_pass = 0;
    ta.bias = bias; ta.stats = stats; ta.seg_off = seg_offsets;
    ta.nseg = (int)nseg; ta.cpg = (int)cpg; ta.groups = (int)groups; ta.rpp = (int)rows_per_point;
    return n > gst::kBW ? gst::launch<2>(a, lda, m, k, w_bf16, ldw, n, ta, st)
                        : gst::launch<1>(a, lda, m, k, w_bf16, ldw, n, ta, st);
  }
  CUtensorMap ta, tb;
  int rc = make_tmap_bf16_2d(&ta, a, m, k, lda, gss::kBM);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tb, w_bf16, n, k, ldw, gss::kBN);
  if (rc) return rc;
  gss::Args args;
  args.M = (int)m; args.N = (int)n; args.K = (int)k;
  args.m_tiles = (int)ceil_div(m, gss::kBM);
  args.n_tiles = (int)ceil_div(n, gss::kBN);
  args.bias = bias; args.stats = stats; args.seg_off = seg_offsets;
  args.nseg = (int)nseg; args.cpg = (int)cpg; args.groups = (int)groups; args.rpp = (int)rows_per_point;
  SE3ET_ENSURE_SMEM(gss::gnstats_stream_kernel, gss::kSmem);
  const int64_t work = (int64_t)args.m_tiles * args.n_tiles;
  const unsigned grid = (unsigned)(work < kNumSMs ? work : kNumSMs);
  gss::gnstats_stream_kernel<<<grid, gss::kThreads, gss::kSmem, st>>>(ta, tb, args);
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}
