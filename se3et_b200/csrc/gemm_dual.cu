// out = LeakyReLU_slope( GroupNorm_1(A1 W1^T + b1) + GroupNorm_2(A2 W2^T + b2) )      bf16 in, bf16 out
//
// The tail of ResnetBottleneckBlockEPN (blocks_epn.py:833-852): unary2 (Linear + GroupNormEPN, no activation) plus the
// shortcut unary (same) and the final LeakyReLU, in ONE kernel.  Both Linears are recomputed here (their statistics come
// from se3et_gemm_bf16_gnstats passes), so neither pre-norm tensor nor the normalised shortcut reaches global memory.
//
// One 128 x BN output tile per CTA, two TMEM accumulators:
//   warp 0    TMA producer: the K-blocks of (A1, W1) then of (A2, W2) through one mbarrier ring
//   warp 1    TMEM allocator + tcgen05.mma issuer (first K1 blocks -> accumulator 0, the rest -> accumulator 1)
//   warps 2-5 epilogue: per-column (scale1, scale2, shift) table of the tile's pair, two tcgen05.ld per 16 columns, the
//             bf16 tile is staged in the (now idle) operand ring and leaves with row-contiguous 16-byte stores
#include <cuda_bf16.h>

#include "common.cuh"
#include "gemm_stream.cuh"
#include "tc.cuh"

namespace se3et {

int make_tmap_bf16_2d(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows);

constexpr int kDualBM = 128;
constexpr int kDualBK = 64;
constexpr int kDualMaxStages = 4;
constexpr int kDualThreads = 192;
constexpr int kDualEpCols = 16;

struct DualNorm {
  const double* stats;  // [nseg, groups, 2] {sum, sum sq} of the fp32 Linear output (bias included)
  const float* gamma;
  const float* beta;
  const float* bias;    // nullable
};

struct DualArgs {
  int M, N, K1, K2, stages;
  __nv_bfloat16* out;
  int64_t ldc;
  const int64_t* seg_off;
  int nseg, cpg, groups, rpp;
  DualNorm n1, n2;
  float eps, slope;
};

template <int BN>
struct DualSmem {
  static constexpr int kABytes = kDualBM * 128;
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTail = 128 + BN * 16;  // barriers | float4 column table
  static int total(int stages) { return stages * kStageBytes + kTail + 1024; }
};

// (scale, shift) of y -> (y + bias - mean) * rstd * gamma + beta for column c of pair seg
__device__ __forceinline__ float2 dual_affine(const DualNorm& n, int seg, int c, int groups, int cpg, double cnt,
                                              float eps) {
  const double* st = n.stats + ((int64_t)seg * groups + c / cpg) * 2;
  const double mean = st[0] / cnt;
  const double var = st[1] / cnt - mean * mean;
  const float sc = rsqrtf((float)fmax(var, 0.0) + eps) * __ldg(n.gamma + c);
  const float b = n.bias ? __ldg(n.bias + c) : 0.f;
  return make_float2(sc, __ldg(n.beta + c) + (b - (float)mean) * sc);
}

template <int BN>
__global__ void __launch_bounds__(kDualThreads, BN <= 64 ? 4 : 2)
gemm_dual_gnapply_kernel(const __grid_constant__ CUtensorMap tma_a1, const __grid_constant__ CUtensorMap tma_b1,
                         const __grid_constant__ CUtensorMap tma_a2, const __grid_constant__ CUtensorMap tma_b2,
                         DualArgs args) {
  using S = DualSmem<BN>;
  constexpr uint32_t kAccCols = BN < 32 ? 32 : BN;
  constexpr uint32_t kTmemAlloc = 2 * kAccCols;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int stages = args.stages;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + stages * S::kStageBytes);
  uint64_t* empty_bar = full_bar + kDualMaxStages;
  uint64_t* tmem_full_bar = empty_bar + kDualMaxStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);
  float4* col_tab = reinterpret_cast<float4*>(smem + stages * S::kStageBytes + 128);  // {scale1, scale2, shift, -}
  uint8_t* c_tile = smem;  // aliases the operand ring once every MMA has retired

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kDualBM, n0 = blockIdx.y * BN;
  const int nkb1 = (args.K1 + kDualBK - 1) / kDualBK, nkb2 = (args.K2 + kDualBK - 1) / kDualBK;
  const int nkb = nkb1 + nkb2;

  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&tma_a1);
    tc::tma_prefetch_desc(&tma_b1);
    tc::tma_prefetch_desc(&tma_a2);
    tc::tma_prefetch_desc(&tma_b2);
    for (int s = 0; s < stages; ++s) {
      tc::mbar_init(&full_bar[s], 1);
      tc::mbar_init(&empty_bar[s], 1);
    }
    tc::mbar_init(tmem_full_bar, 1);
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
      for (int kb = 0; kb < nkb; ++kb) {
        tc::mbar_wait_long(&empty_bar[s], phase ^ 1);
        tc::mbar_arrive_expect_tx(&full_bar[s], S::kStageBytes);
        uint8_t* a_dst = smem + s * S::kStageBytes;
        const bool first = kb < nkb1;
        const int kc = (first ? kb : kb - nkb1) * kDualBK;
        tc::tma_load_2d(a_dst, first ? &tma_a1 : &tma_a2, &full_bar[s], kc, m0);
        tc::tma_load_2d(a_dst + S::kABytes, first ? &tma_b1 : &tma_b2, &full_bar[s], kc, n0);
        if (++s == stages) { s = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = tc::umma_idesc_bf16(kDualBM, BN);
      int s = 0;
      uint32_t phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        tc::mbar_wait_long(&full_bar[s], phase);
        tc::tcgen05_fence_after_sync();
        const bool first = kb < nkb1;
        const int kbl = first ? kb : kb - nkb1;
        const uint32_t tmem_acc = tmem_base + (first ? 0u : kAccCols);
        const uint32_t a_addr = tc::smem_u32(smem + s * S::kStageBytes);
        const uint64_t a_desc = tc::umma_desc_sw128(a_addr);
        const uint64_t b_desc = tc::umma_desc_sw128(a_addr + S::kABytes);
#pragma unroll
        for (int k = 0; k < kDualBK / 16; ++k)
          tc::umma_bf16(tmem_acc, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (kbl | k) != 0);
        tc::umma_commit(&empty_bar[s]);
        if (++s == stages) { s = 0; phase ^= 1; }
      }
      tc::umma_commit(tmem_full_bar);
    }
  } else {
    const int ew = warp & 3;            // TMEM lanes [32 ew, 32 ew + 32)
    const int lane_base = ew * 32;
    const int et = ew * 32 + lane;      // epilogue thread 0..127 = tile row
    const int row = m0 + et;
    const bool row_ok = row < args.M;
    const int cpg = args.cpg;
    constexpr int kRowChunks = BN / 8;  // 16-byte chunks per staged row
    constexpr int kSwz = kRowChunks >= 8 ? 7 : kRowChunks - 1;
    // pair of this tile; tiles that straddle a pair boundary take the per-row path
    const int last = min(m0 + kDualBM, args.M) - 1;
    const int seg0 = segment_of(args.seg_off, args.nseg, m0 / args.rpp);
    const bool uniform = segment_of(args.seg_off, args.nseg, last / args.rpp) == seg0;
    const int row_seg = uniform ? seg0 : segment_of(args.seg_off, args.nseg, (row_ok ? row : last) / args.rpp);
    const double cnt = (double)(args.seg_off[row_seg + 1] - args.seg_off[row_seg]) * args.rpp * cpg;
    if (uniform) {
      for (int cc = et; cc < BN; cc += 128) {
        const float2 f1 = dual_affine(args.n1, seg0, n0 + cc, args.groups, cpg, cnt, args.eps);
        const float2 f2 = dual_affine(args.n2, seg0, n0 + cc, args.groups, cpg, cnt, args.eps);
        col_tab[cc] = make_float4(f1.x, f2.x, f1.y + f2.y, 0.f);
      }
    }
    tc::mbar_wait_long(tmem_full_bar, 0);
    tc::tcgen05_fence_after_sync();
    asm volatile("bar.sync 1, 128;" ::: "memory");  // column table visible; the ring is free to hold the output tile
    uint8_t* crow = c_tile + et * (BN * 2);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += kDualEpCols) {
      uint32_t r1[kDualEpCols], r2[kDualEpCols];
      tc::tmem_ld_32x32b_x16(tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)c0, r1);
      tc::tmem_ld_32x32b_x16(tmem_base + ((uint32_t)lane_base << 16) + kAccCols + (uint32_t)c0, r2);
      tc::tmem_ld_wait();
      float v[kDualEpCols];
#pragma unroll
      for (int j = 0; j < kDualEpCols; ++j) {
        float x;
        if (uniform) {
          const float4 tb = col_tab[c0 + j];
          x = fmaf(__uint_as_float(r1[j]), tb.x, fmaf(__uint_as_float(r2[j]), tb.y, tb.z));
        } else {
          const float2 f1 = dual_affine(args.n1, row_seg, n0 + c0 + j, args.groups, cpg, cnt, args.eps);
          const float2 f2 = dual_affine(args.n2, row_seg, n0 + c0 + j, args.groups, cpg, cnt, args.eps);
          x = fmaf(__uint_as_float(r1[j]), f1.x, fmaf(__uint_as_float(r2[j]), f2.x, f1.y + f2.y));
        }
        v[j] = fmaxf(x, x * args.slope);  // LeakyReLU for slope <= 1
      }
      const int cj = c0 / 8;
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        __nv_bfloat162 p0 = __floats2bfloat162_rn(v[8 * jj], v[8 * jj + 1]);
        __nv_bfloat162 p1 = __floats2bfloat162_rn(v[8 * jj + 2], v[8 * jj + 3]);
        __nv_bfloat162 p2 = __floats2bfloat162_rn(v[8 * jj + 4], v[8 * jj + 5]);
        __nv_bfloat162 p3 = __floats2bfloat162_rn(v[8 * jj + 6], v[8 * jj + 7]);
        uint4 u;
        u.x = *reinterpret_cast<uint32_t*>(&p0);
        u.y = *reinterpret_cast<uint32_t*>(&p1);
        u.z = *reinterpret_cast<uint32_t*>(&p2);
        u.w = *reinterpret_cast<uint32_t*>(&p3);
        *reinterpret_cast<uint4*>(crow + (((cj + jj) ^ (et & kSwz)) << 4)) = u;
      }
    }
    tc::tcgen05_fence_before_sync();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    for (int i = et; i < kDualBM * kRowChunks; i += 128) {
      const int rr = i / kRowChunks, j = i - rr * kRowChunks;
      if (m0 + rr < args.M)
        *reinterpret_cast<uint4*>(args.out + (int64_t)(m0 + rr) * args.ldc + n0 + j * 8) =
            *reinterpret_cast<const uint4*>(c_tile + rr * (BN * 2) + ((j ^ (rr & kSwz)) << 4));
    }
  }
  tc::tcgen05_fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<kTmemAlloc>(tmem_base);
}

template <int BN>
static int launch_dual(const CUtensorMap& ta1, const CUtensorMap& tb1, const CUtensorMap& ta2, const CUtensorMap& tb2,
                       DualArgs args, cudaStream_t st) {
  using S = DualSmem<BN>;
  const int nkb = (args.K1 + kDualBK - 1) / kDualBK + (args.K2 + kDualBK - 1) / kDualBK;
  args.stages = nkb < kDualMaxStages ? nkb : kDualMaxStages;
  while (args.stages > 1 && S::total(args.stages) > (BN <= 64 ? 55 : 72) * 1024) --args.stages;  // 4 / 3 CTAs per SM
  // the staged output tile must fit the ring it aliases
  while (args.stages * S::kStageBytes < kDualBM * BN * 2) ++args.stages;
  const int smem = S::total(args.stages);
  SE3ET_ENSURE_SMEM(gemm_dual_gnapply_kernel<BN>, smem);
  dim3 grid((unsigned)ceil_div(args.M, kDualBM), (unsigned)(args.N / BN), 1);
  gemm_dual_gnapply_kernel<BN><<<grid, kDualThreads, smem, st>>>(ta1, tb1, ta2, tb2, args);
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}

}  // namespace se3et

using namespace se3et;

extern "C" int se3et_gemm_bf16_gnapply_dual(const void* a1, int64_t lda1, const void* b1, int64_t ldb1, int64_t k1,
                                            const float* bias1, const double* stats1, const float* gamma1,
                                            const float* beta1, const void* a2, int64_t lda2, const void* b2,
                                            int64_t ldb2, int64_t k2, const float* bias2, const double* stats2,
                                            const float* gamma2, const float* beta2, int64_t m, int64_t n, float eps,
                                            float leaky_slope, void* out_bf16, int64_t ldc,
                                            const int64_t* seg_offsets, int64_t nseg, int64_t groups,
                                            int64_t rows_per_point, int tile_n, void* workspace,
                                            size_t workspace_bytes, se3et_stream_t stream) {
  if (m < 0 || n <= 0 || k1 <= 0 || k2 <= 0 || m > INT32_MAX || n > INT32_MAX || k1 > INT32_MAX || k2 > INT32_MAX)
    return SE3ET_ERR_ARG;
  if (!a1 || !b1 || !a2 || !b2 || !stats1 || !gamma1 || !beta1 || !stats2 || !gamma2 || !beta2 || !out_bf16 ||
      !seg_offsets || nseg <= 0 || groups <= 0 || n % groups || rows_per_point <= 0 || leaky_slope > 1.f)
    return SE3ET_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(out_bf16) & 15) || (ldc % 8)) return SE3ET_ERR_ARG;
  if (tile_n == 0 && m > 0 && workspace && gemm_stream_supported(n, k1, k2, ldc)) {
    const StreamNorm n1{stats1, gamma1, beta1, bias1}, n2{stats2, gamma2, beta2, bias2};
    return gemm_stream_gnapply(a1, lda1, b1, ldb1, k1, n1, a2, lda2, b2, ldb2, k2, n2, m, n, eps, leaky_slope, nullptr,
                               out_bf16, ldc, seg_offsets, nseg, groups, rows_per_point, workspace, workspace_bytes,
                               static_cast<cudaStream_t>(stream));
  }
  int bn = tile_n;
  // 64-wide tiles: 128 TMEM columns and ~53 KB per CTA, four CTAs (16 epilogue warps) per SM; measured a little faster
  // than 128-wide tiles (two CTAs per SM by TMEM), the second read of the A tiles hits L2
  if (bn == 0) bn = n % 64 == 0 ? 64 : 32;
  if ((bn != 32 && bn != 64 && bn != 128) || n % bn) return SE3ET_ERR_UNSUPPORTED;
  if (m == 0) return SE3ET_OK;
  CUtensorMap ta1, tb1, ta2, tb2;
  int rc = make_tmap_bf16_2d(&ta1, a1, m, k1, lda1, kDualBM);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tb1, b1, n, k1, ldb1, bn);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&ta2, a2, m, k2, lda2, kDualBM);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tb2, b2, n, k2, ldb2, bn);
  if (rc) return rc;
  DualArgs args;
  args.M = (int)m; args.N = (int)n; args.K1 = (int)k1; args.K2 = (int)k2; args.stages = 1;
  args.out = static_cast<__nv_bfloat16*>(out_bf16);
  args.ldc = ldc;
  args.seg_off = seg_offsets;
  args.nseg = (int)nseg; args.cpg = (int)(n / groups); args.groups = (int)groups; args.rpp = (int)rows_per_point;
  args.n1 = DualNorm{stats1, gamma1, beta1, bias1};
  args.n2 = DualNorm{stats2, gamma2, beta2, bias2};
  args.eps = eps;
  args.slope = leaky_slope;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (bn) {
    case 128: return launch_dual<128>(ta1, tb1, ta2, tb2, args, st);
    case 64: return launch_dual<64>(ta1, tb1, ta2, tb2, args, st);
    default: return launch_dual<32>(ta1, tb1, ta2, tb2, args, st);
  }
}
