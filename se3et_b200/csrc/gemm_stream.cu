// Streaming apply pass of Linear + GroupNorm blocks (UnaryBlockEPN and the tail of ResnetBottleneckBlockEPN,
// blocks_epn.py:639-665, 833-852):
//     out = LeakyReLU_slope( GN_1(A1 W1^T + b1) [+ GN_2(A2 W2^T + b2)] [+ resid] )          bf16 in, bf16 out
// These passes move far more bytes than they compute (K of 32..256 against N of 32..1024): the kernel is organised
// around keeping HBM busy, not the tensor core.
//
// Persistent CTAs (two per SM), each walks a contiguous range of 128 x 64 output tiles:
//   warp 0     TMA producer, 3-stage ring of (A 128 x 64, W 64 x 64) K-blocks, running ahead across tiles
//   warp 1     tcgen05.mma issuer; accumulators double-buffered in TMEM (tile t + 1 is multiplied while t drains)
//   warps 2-9  epilogue, EIGHT warps: warp w owns TMEM lanes 32 (w % 4) and 32 of the 64 columns.  Per-warp column
//              table {scale1, scale2, shift} (rebuilt only when the pair changes), two tcgen05.ld per 16 columns,
//              bf16 rows staged in a private 2 KB buffer and written as 64-byte row segments.  No CTA-wide barrier
//              after start-up: the warps only meet at the TMEM mbarriers.
#include <cuda_bf16.h>

#include "common.cuh"
#include "gemm_stream.cuh"
#include "tc.cuh"

namespace se3et {

int make_tmap_bf16_2d(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows);

constexpr int kSBM = 128, kSBN = 64, kSBK = 64;
constexpr int kSStages = 3;
constexpr int kSEpWarps = 8;
constexpr int kSThreads = (2 + kSEpWarps) * 32;
constexpr int kSABytes = kSBM * 128, kSBBytes = kSBN * 128, kSStageBytes = kSABytes + kSBBytes;
constexpr int kSStageOff = 0;
constexpr int kSStagingOff = kSStages * kSStageBytes;          // 8 warps x [32 rows][64 B]
constexpr int kSTableOff = kSStagingOff + kSEpWarps * 2048;     // 8 warps x [32 columns] float4
constexpr int kSBarOff = kSTableOff + kSEpWarps * 512;
constexpr int kSSmem = kSBarOff + 128 + 1024;

struct StreamArgs {
  int M, N, K1, K2;     // K2 = 0: one Linear
  int m_tiles, n_tiles;
  __nv_bfloat16* out;
  const __nv_bfloat16* resid;  // nullable, same shape / pitch as out
  int64_t ldc;
  const int64_t* seg_off;
  int nseg, cpg, groups, rpp;
  const float4* tab;    // [nseg][N] {scale1, shift, scale2, -} from gn_table_kernel; NULL = plain Linear (below)
  float slope;
  const float* plain_bias;  // plain mode: out = act(alpha * acc + bias), act = LeakyReLU_slope (slope 0 = ReLU, 1 = none)
  float plain_alpha;
};

__device__ __forceinline__ float2 stream_affine(const StreamNorm& n, int seg, int c, int groups, int cpg, double cnt,
                                                float eps) {
  const double* st = n.stats + ((int64_t)seg * groups + c / cpg) * 2;
  const double mean = st[0] / cnt;
  const double var = st[1] / cnt - mean * mean;
  const float sc = rsqrtf((float)fmax(var, 0.0) + eps) * __ldg(n.gamma + c);
  const float b = n.bias ? __ldg(n.bias + c) : 0.f;
  return make_float2(sc, __ldg(n.beta + c) + (b - (float)mean) * sc);
}

// tab[seg][c] = {scale1, shift1 + shift2, scale2, 0} (scale and shift adjacent: one 8-byte load on the single-Linear path): y -> (y + bias - mean) * rstd * gamma + beta folded per pair and
// column, so that the GEMM epilogue is one table load and one or two FMAs per element (no fp64 there)
__global__ void gn_table_kernel(StreamNorm n1, StreamNorm n2, int dual, const int64_t* __restrict__ seg_off, int rpp,
                                int N, int groups, float eps, float4* __restrict__ tab) {
  const int seg = blockIdx.y;
  const int cpg = N / groups;
  const double cnt = (double)(seg_off[seg + 1] - seg_off[seg]) * rpp * cpg;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < N; c += gridDim.x * blockDim.x) {
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (cnt > 0.0) {
      const float2 f1 = stream_affine(n1, seg, c, groups, cpg, cnt, eps);
      o.x = f1.x;
      o.y = f1.y;
      if (dual) {
        const float2 f2 = stream_affine(n2, seg, c, groups, cpg, cnt, eps);
        o.z = f2.x;
        o.y += f2.y;
      }
    }
    tab[(int64_t)seg * N + c] = o;
  }
}

__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

template <bool kDual>
__global__ void __launch_bounds__(kSThreads, 2)
gemm_stream_gnapply_kernel(const __grid_constant__ CUtensorMap tma_a1, const __grid_constant__ CUtensorMap tma_b1,
                           const __grid_constant__ CUtensorMap tma_a2, const __grid_constant__ CUtensorMap tma_b2,
                           StreamArgs args) {
  constexpr uint32_t kAcc = kDual ? 2 * kSBN : kSBN;  // TMEM columns per accumulator buffer
  constexpr uint32_t kTmemAlloc = 2 * kAcc;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kSBarOff);
  uint64_t* empty_bar = full_bar + kSStages;
  uint64_t* tmem_full_bar = empty_bar + kSStages;  // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;    // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb1 = (args.K1 + kSBK - 1) / kSBK, nkb2 = kDual ? (args.K2 + kSBK - 1) / kSBK : 0;
  const int nkb = nkb1 + nkb2;
  const int64_t W = (int64_t)args.m_tiles * args.n_tiles;
  const int64_t w_begin = W * blockIdx.x / gridDim.x, w_end = W * (blockIdx.x + 1) / gridDim.x;

  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&tma_a1);
    tc::tma_prefetch_desc(&tma_b1);
    if (kDual) {
      tc::tma_prefetch_desc(&tma_a2);
      tc::tma_prefetch_desc(&tma_b2);
    }
    for (int s = 0; s < kSStages; ++s) {
      tc::mbar_init(&full_bar[s], 1);
      tc::mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(&tmem_full_bar[b], 1);
      tc::mbar_init(&tmem_empty_bar[b], kSEpWarps);
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
        const int mt = (int)((uint32_t)w / (uint32_t)args.n_tiles), nt = (int)w - mt * args.n_tiles;
        for (int kb = 0; kb < nkb; ++kb) {
          tc::mbar_wait_long(&empty_bar[s], phase ^ 1);
          tc::mbar_arrive_expect_tx(&full_bar[s], kSStageBytes);
          uint8_t* a_dst = smem + kSStageOff + s * kSStageBytes;
          const bool first = kb < nkb1;
          const int kc = (first ? kb : kb - nkb1) * kSBK;
          tc::tma_load_2d(a_dst, (!kDual || first) ? &tma_a1 : &tma_a2, &full_bar[s], kc, mt * kSBM);
          tc::tma_load_2d(a_dst + kSABytes, (!kDual || first) ? &tma_b1 : &tma_b2, &full_bar[s], kc, nt * kSBN);
          if (++s == kSStages) { s = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = tc::umma_idesc_bf16(kSBM, kSBN);
      int s = 0;
      uint32_t phase = 0, it = 0;
      for (int64_t w = w_begin; w < w_end; ++w, ++it) {
        const uint32_t buf = it & 1u, use = it >> 1;
        tc::mbar_wait_long(&tmem_empty_bar[buf], (use & 1u) ^ 1u);  // the epilogue has drained this buffer
        tc::tcgen05_fence_after_sync();
        for (int kb = 0; kb < nkb; ++kb) {
          tc::mbar_wait_long(&full_bar[s], phase);
          tc::tcgen05_fence_after_sync();
          const bool first = kb < nkb1;
          const int kbl = first ? kb : kb - nkb1;
          const uint32_t tmem_acc = tmem_base + buf * kAcc + (first ? 0u : (uint32_t)kSBN);
          const uint32_t a_addr = tc::smem_u32(smem + kSStageOff + s * kSStageBytes);
          const uint64_t a_desc = tc::umma_desc_sw128(a_addr);
          const uint64_t b_desc = tc::umma_desc_sw128(a_addr + kSABytes);
#pragma unroll
          for (int k = 0; k < kSBK / 16; ++k)
            tc::umma_bf16(tmem_acc, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (kbl | k) != 0);
          tc::umma_commit(&empty_bar[s]);
          if (++s == kSStages) { s = 0; phase ^= 1; }
        }
        tc::umma_commit(&tmem_full_bar[buf]);
      }
    }
  } else {
    const int e = warp - 2;
    const int lg = warp & 3;      // TMEM lane group this warp may read
    const int ch = e >> 2;        // which 32 of the tile's 64 columns
    uint8_t* staging = smem + kSStagingOff + e * 2048;
    float4* table = reinterpret_cast<float4*>(smem + kSTableOff + e * 512);
    const uint32_t table_s = tc::smem_u32(table);
    int64_t seg_lo = 0, seg_hi = -1;  // rows of the cached pair
    int seg_cached = 0;
    uint32_t it = 0;
    for (int64_t w = w_begin; w < w_end; ++w, ++it) {
      const uint32_t buf = it & 1u, use = it >> 1;
      const int mt = (int)((uint32_t)w / (uint32_t)args.n_tiles), nt = (int)w - mt * args.n_tiles;
      const int64_t wfirst = (int64_t)mt * kSBM + lg * 32;
      const int64_t row = wfirst + lane;
      const bool row_ok = row < args.M;
      const int n0 = nt * kSBN + ch * 32;
      const bool warp_ok = wfirst < args.M && n0 < args.N;
      bool uniform = true;
      const float4* row_tab = args.tab;
      if (warp_ok) {
        const int64_t wlast = min(wfirst + 31, (int64_t)args.M - 1);
        if (!args.tab) {  // plain Linear: one segment, the table row is {alpha, bias, 0}
          seg_lo = 0;
          seg_hi = args.M;
        } else if (!(wfirst >= seg_lo && wlast < seg_hi)) {
          seg_cached = segment_of(args.seg_off, args.nseg, wfirst / args.rpp);
          seg_lo = args.seg_off[seg_cached] * args.rpp;
          seg_hi = args.seg_off[seg_cached + 1] * args.rpp;
        }
        uniform = wlast < seg_hi;
        if (uniform) {
          // this warp's 32 columns of the pair's table: one coalesced 512-byte load, broadcast from shared memory
          const float4 te = args.tab ? __ldg(args.tab + (int64_t)seg_cached * args.N + n0 + lane)
                                     : make_float4(args.plain_alpha,
                                                   args.plain_bias ? __ldg(args.plain_bias + n0 + lane) : 0.f, 0.f, 0.f);
          __syncwarp();
          table[lane] = te;
          __syncwarp();
        } else {  // the warp's rows straddle a pair boundary: per-row table rows straight from global memory
          const int row_seg = segment_of(args.seg_off, args.nseg, min(row, (int64_t)args.M - 1) / args.rpp);
          row_tab = args.tab + (int64_t)row_seg * args.N + n0;
        }
      }
      // residual: the warp's 32 rows x 64 bytes, loaded row-contiguously (four lanes per row) while the accumulators
      // are produced, then transposed through the staging buffer so that every thread holds its own row
      const bool has_resid = !kDual && args.resid != nullptr;
      uint4 res[4] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
      if (has_resid && warp_ok) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int i = q * 32 + lane;
          const int rr = i >> 2, j = i & 3;
          if (wfirst + rr < args.M)
            res[q] = __ldg(reinterpret_cast<const uint4*>(args.resid + (wfirst + rr) * args.ldc + n0 + j * 8));
        }
      }
      tc::mbar_wait_long(&tmem_full_bar[buf], use & 1u);
      tc::tcgen05_fence_after_sync();
      if (has_resid && warp_ok) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int i = q * 32 + lane;
          const int rr = i >> 2, j = i & 3;
          *reinterpret_cast<uint4*>(staging + rr * 64 + ((j ^ ((rr >> 1) & 3)) << 4)) = res[q];
        }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 4; ++c)
          res[c] = *reinterpret_cast<const uint4*>(staging + lane * 64 + ((c ^ ((lane >> 1) & 3)) << 4));
        __syncwarp();
      }
      if (warp_ok) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r1[16], r2[16];
          const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + buf * kAcc + (uint32_t)(ch * 32 + c * 16);
          tc::tmem_ld_32x32b_x16(taddr, r1);
          if (kDual) tc::tmem_ld_32x32b_x16(taddr + kSBN, r2);
          tc::tmem_ld_wait();
          float v[16];
          auto element = [&](int j, const float4 tb) {
            float x = kDual ? fmaf(__uint_as_float(r1[j]), tb.x, fmaf(__uint_as_float(r2[j]), tb.z, tb.y))
                            : fmaf(__uint_as_float(r1[j]), tb.x, tb.y);
            if (!kDual && has_resid) {
              const uint4 rq = res[c * 2 + (j >> 3)];
              const uint32_t rw = ((j >> 1) & 3) == 0 ? rq.x : (((j >> 1) & 3) == 1 ? rq.y : (((j >> 1) & 3) == 2 ? rq.z : rq.w));
              x += (j & 1) ? __uint_as_float(rw & 0xffff0000u) : __uint_as_float(rw << 16);
            }
            v[j] = fmaxf(x, x * args.slope);  // LeakyReLU for slope <= 1
          };
          // two code paths, not a select per element: with `uniform ? lds : ldg` inside the loop the compiler issued
          // BOTH loads (predicated, as 4-byte pieces) for every element -- four load instructions per output value
          if (uniform) {
#pragma unroll
            for (int j = 0; j < 16; ++j) element(j, lds_f4(table_s + (uint32_t)((c * 16 + j) * 16)));
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) element(j, __ldg(row_tab + c * 16 + j));
          }
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
            // staging row = lane (64 B), chunk (c * 2 + jj) swizzled by the row pair: conflict-free 16-byte stores
            *reinterpret_cast<uint4*>(staging + lane * 64 + (((c * 2 + jj) ^ ((lane >> 1) & 3)) << 4)) = u;
          }
        }
      }
      // the accumulator buffer may be overwritten by the MMA of tile it + 2
      tc::tcgen05_fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tmem_empty_bar[buf]);
      if (warp_ok) {
        // 32 rows x 64 bytes -> global memory, four lanes per row
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int i = q * 32 + lane;
          const int rr = i >> 2, j = i & 3;
          if (wfirst + rr < args.M)
            *reinterpret_cast<uint4*>(args.out + (wfirst + rr) * args.ldc + n0 + j * 8) =
                *reinterpret_cast<const uint4*>(staging + rr * 64 + ((j ^ ((rr >> 1) & 3)) << 4));
        }
      }
      __syncwarp();  // staging is rewritten by the next tile
    }
  }
  tc::tcgen05_fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<kTmemAlloc>(tmem_base);
}

static bool g_stream_enabled = true;
static bool g_stream_plain = true;

bool gemm_stream_supported(int64_t n, int64_t k1, int64_t k2, int64_t ldc) {
  // n = 32 (mod 64): the last column tile is half empty (TMA zero-fills the missing weight rows, its second four
  // epilogue warps idle) -- still far better than the one-tile-per-CTA kernel for the 64 -> 32 / 128 -> 32 unary blocks
  return g_stream_enabled && n % 32 == 0 && k1 % 8 == 0 && k2 % 8 == 0 && ldc % 8 == 0;
}

size_t gemm_stream_workspace_bytes(int64_t n, int64_t nseg) { return sizeof(float4) * (size_t)n * (size_t)nseg; }

int gemm_stream_gnapply(const void* a1, int64_t lda1, const void* b1, int64_t ldb1, int64_t k1, const StreamNorm& n1,
                        const void* a2, int64_t lda2, const void* b2, int64_t ldb2, int64_t k2, const StreamNorm& n2,
                        int64_t m, int64_t n, float eps, float slope, const void* resid, void* out, int64_t ldc,
                        const int64_t* seg_off, int64_t nseg, int64_t groups, int64_t rpp, void* workspace,
                        size_t workspace_bytes, cudaStream_t st) {
  if (!workspace || workspace_bytes < gemm_stream_workspace_bytes(n, nseg) ||
      (reinterpret_cast<uintptr_t>(workspace) & 15) || nseg > 65535)
    return SE3ET_ERR_WORKSPACE;
  if ((int64_t)ceil_div(m, kSBM) * ceil_div(n, kSBN) > INT32_MAX) return SE3ET_ERR_UNSUPPORTED;
  const bool dual = k2 > 0;
  CUtensorMap ta1, tb1, ta2, tb2;
  int rc = make_tmap_bf16_2d(&ta1, a1, m, k1, lda1, kSBM);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tb1, b1, n, k1, ldb1, kSBN);
  if (rc) return rc;
  ta2 = ta1;
  tb2 = tb1;
  if (dual) {
    rc = make_tmap_bf16_2d(&ta2, a2, m, k2, lda2, kSBM);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&tb2, b2, n, k2, ldb2, kSBN);
    if (rc) return rc;
  }
  StreamArgs args;
  args.M = (int)m; args.N = (int)n; args.K1 = (int)k1; args.K2 = (int)k2;
  args.m_tiles = (int)ceil_div(m, kSBM);
  args.n_tiles = (int)ceil_div(n, kSBN);
  args.out = static_cast<__nv_bfloat16*>(out);
  args.resid = static_cast<const __nv_bfloat16*>(resid);
  args.ldc = ldc;
  args.seg_off = seg_off;
  args.nseg = (int)nseg; args.cpg = (int)(n / groups); args.groups = (int)groups; args.rpp = (int)rpp;
  args.tab = static_cast<const float4*>(workspace);
  args.slope = slope;
  args.plain_bias = nullptr;
  args.plain_alpha = 1.f;
  gn_table_kernel<<<dim3((unsigned)ceil_div(n, 256), (unsigned)nseg), 256, 0, st>>>(
      n1, n2, dual ? 1 : 0, seg_off, (int)rpp, (int)n, (int)groups, eps, static_cast<float4*>(workspace));
  SE3ET_LAUNCH_CHECK();
  SE3ET_ENSURE_SMEM(gemm_stream_gnapply_kernel<false>, kSSmem);
  SE3ET_ENSURE_SMEM(gemm_stream_gnapply_kernel<true>, kSSmem);
  const int64_t work = (int64_t)args.m_tiles * args.n_tiles;
  const unsigned grid = (unsigned)(work < 2 * kNumSMs ? work : 2 * kNumSMs);
  if (dual)
    gemm_stream_gnapply_kernel<true><<<grid, kSThreads, kSSmem, st>>>(ta1, tb1, ta2, tb2, args);
  else
    gemm_stream_gnapply_kernel<false><<<grid, kSThreads, kSSmem, st>>>(ta1, tb1, ta2, tb2, args);
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}

// out_bf16 = act(alpha * A B^T + bias) through the streaming kernel (act: slope 0 = ReLU, 1 = none)
int gemm_stream_plain(const void* a, int64_t lda, const void* b, int64_t ldb, int64_t m, int64_t n, int64_t k,
                      const float* bias, float alpha, float slope, void* out, int64_t ldc, cudaStream_t st) {
  if ((int64_t)ceil_div(m, kSBM) * (n / kSBN) > INT32_MAX) return SE3ET_ERR_UNSUPPORTED;
  CUtensorMap ta, tb;
  int rc = make_tmap_bf16_2d(&ta, a, m, k, lda, kSBM);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tb, b, n, k, ldb, kSBN);
  if (rc) return rc;
  StreamArgs args;
  args.M = (int)m; args.N = (int)n; args.K1 = (int)k; args.K2 = 0;
  args.m_tiles = (int)ceil_div(m, kSBM);
  args.n_tiles = (int)ceil_div(n, kSBN);
  args.out = static_cast<__nv_bfloat16*>(out);
  args.resid = nullptr;
  args.ldc = ldc;
  args.seg_off = nullptr;
  args.nseg = 1; args.cpg = 1; args.groups = 1; args.rpp = 1;
  args.tab = nullptr;
  args.slope = slope;
  args.plain_bias = bias;
  args.plain_alpha = alpha;
  SE3ET_ENSURE_SMEM(gemm_stream_gnapply_kernel<false>, kSSmem);
  const int64_t work = (int64_t)args.m_tiles * args.n_tiles;
  const unsigned grid = (unsigned)(work < 2 * kNumSMs ? work : 2 * kNumSMs);
  gemm_stream_gnapply_kernel<false><<<grid, kSThreads, kSSmem, st>>>(ta, tb, ta, tb, args);
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}

bool gemm_stream_plain_enabled() { return g_stream_plain; }

}  // namespace se3et

extern "C" int se3et_gemm_set_stream_plain(int on) {
  se3et::g_stream_plain = on != 0;
  return SE3ET_OK;
}

// A/B switch for measurements: 0 routes se3et_gemm_bf16_gnapply(_dual) back to the one-tile-per-CTA kernels
extern "C" int se3et_gemm_set_stream_apply(int on) {
  se3et::g_stream_enabled = on != 0;
  return SE3ET_OK;
}
