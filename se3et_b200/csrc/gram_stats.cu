// GroupNorm statistics of y = A W^T + b without forming y: per pair the Gram matrix G = A^T A and the column sums
// s = 1^T A of the INPUT (one pass over A, K x K outputs), then
//     sum_rows y_j   = w_j . s + R b_j                    sum_rows y_j^2 = w_j^T G w_j + 2 b_j w_j . s + R b_j^2
// in fp64 per channel, summed over the channels of a group.  Replaces the statistics pass of se3et_gemm_bf16_gnstats
// (UnaryBlockEPN, blocks_epn.py:639-665) when the Linear widens (N >= 2 K): that pass computes all N outputs of every
// row only to square and sum them and is bound by its epilogue; this one is bound by reading A once.
//
// gram_kernel<K>: persistent CTAs over 128-row tiles that never straddle a pair (tiles are counted per pair; rows past
// the pair's end are zero-filled), a 4-stage cp.async ring, mma.sync m16n8k16 with both operands taken from the same
// shared-memory tile by ldmatrix.trans (the reduction runs over rows).  An all-ones A fragment yields the column sums.
#include <cuda_bf16.h>

#include "common.cuh"
#include "kpconv_mma.cuh"

namespace se3et {

using namespace kpm;

constexpr int kGramRows = 128;
constexpr int kGramThreads = 256;

template <int K>
struct GramCfg {
  static constexpr int kPitch = K * 2 + 16;                 // bytes; 16-byte pad: ldmatrix rows hit distinct banks
  static constexpr int kTileBytes = kGramRows * kPitch;
  static constexpr int kNP = K / 16;                        // column pairs of n-tiles = warps along N
  static constexpr int kRS = 8 / kNP;                       // row splits of a tile among the remaining warps
  static constexpr int kMT = K / 16;                        // m-tiles (all of them per warp)
  static constexpr int kKSteps = (kGramRows / 16) / kRS;    // 16-row reduction steps per warp and tile
  static constexpr int kEntries = (K + 1) * K;              // G rows 0..K-1, row K = column sums
  static constexpr int kStages = 4;                         // cp.async ring: three tiles in flight per CTA
  static constexpr int kSmem = kStages * kTileBytes + kEntries * 4;
  static constexpr int kCtasPerSM = K <= 32 ? 4 : (K <= 64 ? 2 : 1);
};

// global tile index -> (pair, first row, rows in the pair after it); pairs are ranges of points, rows = points * rpp
struct GramTile {
  int seg;
  int64_t row0, row_end;
};

__device__ __forceinline__ GramTile gram_locate(const int64_t* seg_off, int nseg, int rpp, int64_t tile) {
  GramTile t{nseg, 0, 0};
  int64_t first = 0;
  for (int s = 0; s < nseg; ++s) {
    const int64_t r0 = seg_off[s] * rpp, r1 = seg_off[s + 1] * rpp;
    const int64_t nt = (r1 - r0 + kGramRows - 1) / kGramRows;
    if (tile < first + nt) {
      t.seg = s;
      t.row0 = r0 + (tile - first) * kGramRows;
      t.row_end = r1;
      return t;
    }
    first += nt;
  }
  return t;
}
// the tile after t (seg = nseg past the last one); skips empty pairs
__device__ __forceinline__ void gram_advance(GramTile& t, const int64_t* seg_off, int nseg, int rpp) {
  if (t.seg >= nseg) return;
  t.row0 += kGramRows;
  while (t.row0 >= t.row_end) {
    if (++t.seg >= nseg) return;
    t.row0 = seg_off[t.seg] * rpp;
    t.row_end = seg_off[t.seg + 1] * rpp;
  }
}

template <int K>
__global__ void __launch_bounds__(kGramThreads) gram_kernel(const __nv_bfloat16* __restrict__ a, int64_t lda,
                                                            const int64_t* __restrict__ seg_off, int nseg, int rpp,
                                                            int64_t total_tiles, double* __restrict__ gram) {
  using C = GramCfg<K>;
  extern __shared__ __align__(16) uint8_t gsm[];
  float* red = reinterpret_cast<float*>(gsm + C::kStages * C::kTileBytes);  // [K + 1][K] CTA-level partial of the current pair
  const uint32_t tile_s = smem_addr(gsm);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int np = warp % C::kNP, rs = warp / C::kNP;
  const int64_t t_begin = total_tiles * blockIdx.x / gridDim.x, t_end = total_tiles * (blockIdx.x + 1) / gridDim.x;
  if (t_begin >= t_end) return;

  float acc[C::kMT][2][4];
  float ones_acc[2][4];
  auto zero_acc = [&]() {
#pragma unroll
    for (int m = 0; m < C::kMT; ++m)
#pragma unroll
      for (int n = 0; n < 2; ++n)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[m][n][i] = 0.f;
#pragma unroll
    for (int n = 0; n < 2; ++n)
#pragma unroll
      for (int i = 0; i < 4; ++i) ones_acc[n][i] = 0.f;
  };
  zero_acc();
  for (int i = threadIdx.x; i < C::kEntries; i += kGramThreads) red[i] = 0.f;

  auto load = [&](const GramTile& t, int buf) {  // always commits one group (empty past the last tile)
    constexpr int kChunks = K / 8;  // 16-byte chunks per row
    if (t.seg < nseg) {
      for (int i = threadIdx.x; i < kGramRows * kChunks; i += kGramThreads) {
        const int r = i / kChunks, c = i - r * kChunks;
        const int64_t row = t.row0 + r;
        const bool ok = row < t.row_end;
        cp_async_16(tile_s + buf * C::kTileBytes + r * C::kPitch + c * 16, a + (ok ? row : t.row0) * lda + c * 8,
                    ok ? 16 : 0);
      }
    }
    cp_async_commit();
  };
  // CTA partial of pair `seg` -> global fp64 (all threads; accumulators first meet in shared memory)
  auto flush = [&](int seg) {
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int m = 0; m < C::kMT; ++m)
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        const int col = np * 16 + n * 8 + 2 * q;
        atomicAdd(&red[(m * 16 + g) * K + col], acc[m][n][0]);
        atomicAdd(&red[(m * 16 + g) * K + col + 1], acc[m][n][1]);
        atomicAdd(&red[(m * 16 + g + 8) * K + col], acc[m][n][2]);
        atomicAdd(&red[(m * 16 + g + 8) * K + col + 1], acc[m][n][3]);
      }
    if (g == 0) {
#pragma unroll
      for (int n = 0; n < 2; ++n) {
        const int col = np * 16 + n * 8 + 2 * q;
        atomicAdd(&red[K * K + col], ones_acc[n][0]);
        atomicAdd(&red[K * K + col + 1], ones_acc[n][1]);
      }
    }
    __syncthreads();
    double* dst = gram + (int64_t)seg * C::kEntries;
    for (int i = threadIdx.x; i < C::kEntries; i += kGramThreads) {
      atomicAdd(dst + i, (double)red[i]);
      red[i] = 0.f;
    }
    zero_acc();
    __syncthreads();
  };

  // two iterators over this CTA's tiles: `ld` runs kStages - 1 tiles ahead of `cur`
  GramTile cur = gram_locate(seg_off, nseg, rpp, t_begin);
  GramTile ld = cur;
  const int64_t count = t_end - t_begin;
  for (int i = 0; i < C::kStages - 1; ++i) {
    if (i >= count) ld.seg = nseg;
    load(ld, i);
    gram_advance(ld, seg_off, nseg, rpp);
  }
  const uint32_t ones[4] = {0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u};  // bf16 1.0 pairs
  // ldmatrix.trans lane addressing inside a 16-row step: matrix i = lane / 8, row lane % 8
  const int li = lane >> 3, lr = lane & 7;
  const uint32_t a_lane = (uint32_t)(((li >> 1) * 8 + lr) * C::kPitch + (li & 1) * 16);   // k block i / 2, m block i % 2
  const uint32_t b_lane = (uint32_t)(((li & 1) * 8 + lr) * C::kPitch + (li >> 1) * 16);   // k block i % 2, n block i / 2
  int buf = 0;
  for (int64_t t = 0; t < count && cur.seg < nseg; ++t) {
    // refill the stage the previous iteration finished with (its trailing __syncthreads orders the overwrite)
    if (t + C::kStages - 1 >= count) ld.seg = nseg;
    load(ld, (buf + C::kStages - 1) % C::kStages);
    gram_advance(ld, seg_off, nseg, rpp);
    cp_async_wait<C::kStages - 1>();
    __syncthreads();
    const uint32_t base = tile_s + buf * C::kTileBytes;
#pragma unroll
    for (int ks = 0; ks < C::kKSteps; ++ks) {
      const uint32_t step = base + (uint32_t)((rs * C::kKSteps + ks) * 16 * C::kPitch);
      uint32_t b[4];
      ldmatrix_x4_trans(b, step + b_lane + np * 32);
      mma_16816(ones_acc[0], ones, b[0], b[1]);
      mma_16816(ones_acc[1], ones, b[2], b[3]);
#pragma unroll
      for (int m = 0; m < C::kMT; ++m) {
        uint32_t af[4];
        ldmatrix_x4_trans(af, step + a_lane + m * 32);
        mma_16816(acc[m][0], af, b[0], b[1]);
        mma_16816(acc[m][1], af, b[2], b[3]);
      }
    }
    __syncthreads();  // every warp is done with this stage before it is refilled
    const int seg = cur.seg;
    gram_advance(cur, seg_off, nseg, rpp);
    if (t + 1 == count || cur.seg != seg) flush(seg);
    buf = (buf + 1) % C::kStages;
  }
  cp_async_wait<0>();
}

// stats[seg][group] = {sum, sum sq} of y = A w^T + b over the pair's rows and the group's channels
__global__ void __launch_bounds__(128) gram_finalize_kernel(const double* __restrict__ gram, int K,
                                                            const __nv_bfloat16* __restrict__ w, int64_t ldw,
                                                            const float* __restrict__ bias,
                                                            const int64_t* __restrict__ seg_off, int rpp, int cpg,
                                                            int groups, double* __restrict__ stats) {
  const int g = blockIdx.x, seg = blockIdx.y;
  const double* G = gram + (int64_t)seg * (K + 1) * K;
  const double* s = G + (int64_t)K * K;
  const double rows = (double)(seg_off[seg + 1] - seg_off[seg]) * rpp;
  double sum = 0.0, sq = 0.0;
  // eight channels at a time: one pass over column k1 of G (symmetric: G[k2][k1], coalesced over k1) serves all of them
  __shared__ double wsh[8][128];  // fp64 once: the inner loop is bound by the fp64 pipe, conversions would double it
  for (int j0 = 0; j0 < cpg; j0 += 8) {
    const int nj = min(8, cpg - j0);
    __syncthreads();
    for (int i = threadIdx.x; i < 8 * K; i += blockDim.x) {
      const int jj = i / K, k = i - jj * K;
      wsh[jj][k] = jj < nj ? (double)__bfloat162float(w[(int64_t)(g * cpg + j0 + jj) * ldw + k]) : 0.0;
    }
    __syncthreads();
    for (int k1 = threadIdx.x; k1 < K; k1 += blockDim.x) {
      double t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      for (int k2 = 0; k2 < K; ++k2) {
        const double gv = G[(int64_t)k2 * K + k1];
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) t[jj] += gv * wsh[jj][k2];
      }
      const double s1 = s[k1];
#pragma unroll
      for (int jj = 0; jj < 8; ++jj) {
        const double bj = (bias && jj < nj) ? (double)bias[g * cpg + j0 + jj] : 0.0;
        const double w1 = wsh[jj][k1];
        sq += w1 * t[jj] + 2.0 * bj * w1 * s1;
        sum += w1 * s1;
      }
    }
    if (threadIdx.x == 0) {
      for (int jj = 0; jj < nj; ++jj) {
        const double bj = bias ? (double)bias[g * cpg + j0 + jj] : 0.0;
        sum += rows * bj;
        sq += rows * bj * bj;
      }
    }
  }
  __shared__ double sh[2][4];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, o);
    sq += __shfl_xor_sync(0xffffffffu, sq, o);
  }
  if ((threadIdx.x & 31) == 0) {
    sh[0][threadIdx.x >> 5] = sum;
    sh[1][threadIdx.x >> 5] = sq;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double* dst = stats + ((int64_t)seg * groups + g) * 2;
    dst[0] = sh[0][0] + sh[0][1] + sh[0][2] + sh[0][3];
    dst[1] = sh[1][0] + sh[1][1] + sh[1][2] + sh[1][3];
  }
}

template <int K>
static int launch_gram(const __nv_bfloat16* a, int64_t lda, const int64_t* seg_off, int nseg, int rpp,
                       int64_t total_tiles, double* gram, cudaStream_t st) {
  using C = GramCfg<K>;
  SE3ET_ENSURE_SMEM(gram_kernel<K>, C::kSmem);
  const int64_t cap = (int64_t)C::kCtasPerSM * kNumSMs;
  const int64_t grid = total_tiles < cap ? total_tiles : cap;
  gram_kernel<K><<<(unsigned)grid, kGramThreads, C::kSmem, st>>>(a, lda, seg_off, nseg, rpp, total_tiles, gram);
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}

}  // namespace se3et

using namespace se3et;

extern "C" int se3et_linear_gnstats_gram_workspace_bytes(int64_t k, int64_t nseg, size_t* bytes) {
  if (!bytes || k <= 0 || nseg <= 0) return SE3ET_ERR_ARG;
  *bytes = sizeof(double) * (size_t)nseg * (size_t)(k + 1) * (size_t)k;
  return SE3ET_OK;
}

extern "C" int se3et_linear_gnstats_gram(const void* a, int64_t lda, int64_t m, int64_t k, const void* w_bf16,
                                         int64_t ldw, int64_t n, const float* bias, const int64_t* seg_offsets,
                                         int64_t nseg, int64_t groups, int64_t rows_per_point, int64_t upper_tiles,
                                         void* workspace, size_t workspace_bytes, double* stats,
                                         se3et_stream_t stream) {
  if (!a || !w_bf16 || !seg_offsets || !stats || !workspace || m < 0 || n <= 0 || nseg <= 0 || groups <= 0 ||
      n % groups || rows_per_point <= 0 || rows_per_point > INT32_MAX || upper_tiles < 0)
    return SE3ET_ERR_ARG;
  if (k != 32 && k != 64 && k != 128) return SE3ET_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(a) & 15) || (lda % 8)) return SE3ET_ERR_ARG;
  const size_t need = sizeof(double) * (size_t)nseg * (size_t)(k + 1) * (size_t)k;
  if (workspace_bytes < need) return SE3ET_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  double* gram = static_cast<double*>(workspace);
  SE3ET_CUDA_CHECK(cudaMemsetAsync(gram, 0, need, st));
  // upper_tiles: an upper bound of sum over pairs of ceil(rows / 128) known on the host without a sync
  // (rows / 128 + nseg); tiles past the real count find no pair and are skipped
  const int64_t total_tiles = upper_tiles > 0 ? upper_tiles : m / kGramRows + nseg;
  int rc = SE3ET_OK;
  if (m > 0) {
    const __nv_bfloat16* ap = static_cast<const __nv_bfloat16*>(a);
    switch (k) {
      case 32: rc = launch_gram<32>(ap, lda, seg_offsets, (int)nseg, (int)rows_per_point, total_tiles, gram, st); break;
      case 64: rc = launch_gram<64>(ap, lda, seg_offsets, (int)nseg, (int)rows_per_point, total_tiles, gram, st); break;
      default: rc = launch_gram<128>(ap, lda, seg_offsets, (int)nseg, (int)rows_per_point, total_tiles, gram, st); break;
    }
    if (rc) return rc;
  }
  gram_finalize_kernel<<<dim3((unsigned)groups, (unsigned)nseg), 128, 0, st>>>(
      gram, (int)k, static_cast<const __nv_bfloat16*>(w_bf16), ldw, bias, seg_offsets, (int)rows_per_point,
      (int)(n / groups), (int)groups, stats);
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}

// One Gram pass for TWO Linears applied to the same input (the block input x feeds both the narrowing
// unary1 and the widening shortcut of a ResnetBottleneckBlockEPN, blocks_epn.py:833-852): G = A^T A and s = 1^T A give
// the GroupNorm statistics of every Linear on A, so x is read once for statistics instead of twice.
extern "C" int se3et_linear_gnstats_gram2(const void* a, int64_t lda, int64_t m, int64_t k, const void* w1_bf16,
                                          int64_t ldw1, int64_t n1, const float* bias1, int64_t groups1, double* stats1,
                                          const void* w2_bf16, int64_t ldw2, int64_t n2, const float* bias2,
                                          int64_t groups2, double* stats2, const int64_t* seg_offsets, int64_t nseg,
                                          int64_t rows_per_point, void* workspace, size_t workspace_bytes,
                                          se3et_stream_t stream) {
  if (!a || !w1_bf16 || !w2_bf16 || !seg_offsets || !stats1 || !stats2 || !workspace || m < 0 || n1 <= 0 || n2 <= 0 ||
      nseg <= 0 || groups1 <= 0 || groups2 <= 0 || n1 % groups1 || n2 % groups2 || rows_per_point <= 0 ||
      rows_per_point > INT32_MAX)
    return SE3ET_ERR_ARG;
  // the first Linear's statistics with the Gram pass, the second one from the same Gram matrix
  int rc = se3et_linear_gnstats_gram(a, lda, m, k, w1_bf16, ldw1, n1, bias1, seg_offsets, nseg, groups1, rows_per_point, 0,
                                     workspace, workspace_bytes, stats1, stream);
  if (rc) return rc;
  gram_finalize_kernel<<<dim3((unsigned)groups2, (unsigned)nseg), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const double*>(workspace), (int)k, static_cast<const __nv_bfloat16*>(w2_bf16), ldw2, bias2, seg_offsets,
      (int)rows_per_point, (int)(n2 / groups2), (int)groups2, stats2);
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}
