// C[M,N] = alpha * A[M,K] * B[N,K]^T (+ bias) (ReLU)   bf16 operands, fp32 accumulation in TMEM.
//
// The one GEMM behind every nn.Linear of the path (UnaryBlockEPN, decoder, in/out_proj, q/k/v,
// AttentionOutput) and the KPConvInterSO3 contraction (A = gathered/weighted features, B = W_flat).
// Warp-specialised tcgen05 kernel, one 128 x BN output tile per CTA:
//   warp 0   TMA producer  (A and B tiles, 128-byte swizzle, 4-stage mbarrier ring)
//   warp 1   TMEM allocator + single-thread tcgen05.mma issuer (UMMA 128 x BN x 16)
//   warps 2-5 epilogue: tcgen05.ld 32 lanes x 32 columns -> bias / alpha / ReLU -> fp32 and/or bf16 rows
// Ragged M and K are handled by TMA out-of-bounds zero fill; N must be a multiple of BN (16..256).
// Batched problems (grid.z) address the flattened [batch*rows, K] tensor maps by row offset.
#include <cuda_bf16.h>

#include "common.cuh"
#include "gn_epilogue.cuh"
#include "tc.cuh"

namespace se3et {

constexpr int kGemmBM = 128;
constexpr int kGemmBK = 64;  // 64 bf16 = 128 bytes = one swizzle atom
constexpr int kGemmStages = 4;
constexpr int kGemmThreads = 192;

struct GemmEpilogue {
  float* out_f32;            // nullable, [M, ldc]
  __nv_bfloat16* out_bf16;   // nullable, [M, ldc]
  const float* bias;         // nullable, [N]
  int64_t ldc;               // row stride of the outputs (elements)
  int64_t c_batch_stride;    // elements between batches of C
  float alpha;
  int act;                   // 0 none, 1 ReLU
  int transposed;            // 1: out_f32[c_off + col * ldc + row] (row-contiguous columns), fp32 only
  int n_valid;               // transposed mode: only columns < n_valid are stored
  // optional fused GroupNorm statistics of the stored fp32 values (se3et_gemm_bf16_gnstats):
  double* gn_stats = nullptr; // [nseg, groups, 2] {sum, sum of squares}, zeroed by the host before the launch
  const int64_t* gn_seg_off = nullptr; // [nseg + 1] point offsets of the pairs
  int gn_nseg = 0;
  int gn_cpg = 1;                // channels per group: a power of two <= 32, or a multiple of 32
  int gn_groups = 0;
  int gn_rpp = 1;                // rows per point (6 for equivariant features)
};

struct GemmShape {
  int M, N, K;               // per batch (M = largest group when a group table is used)
  int64_t a_batch_rows;      // rows between batches in the flattened A map
  int64_t b_batch_rows;      // rows between batches in the flattened B map (0 = shared B)
  const int64_t* groups;     // optional device table, 6 int64 per blockIdx.z: {a_row0, b_row0, m_rows, c_off, ldc, -}
};

template <int BN>
struct GemmSmem {
  static constexpr int kABytes = kGemmBM * 128;
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = kGemmStages * kStageBytes;
  static constexpr int kGnOffset = kBarOffset + 128;      // 4 epilogue warps x 64 groups x {sum, sumsq} floats
  static constexpr int kTotal = kGnOffset + 4 * 64 * 2 * 4 + 1024;  // + barriers + GN scratch + alignment slack
};


template <int BN>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tma_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, GemmShape shape,
                GemmEpilogue ep) {
  using S = GemmSmem<BN>;
  constexpr uint32_t kTmemCols = BN < 32 ? 32 : BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);
  uint64_t* empty_bar = full_bar + kGemmStages;
  uint64_t* tmem_full_bar = empty_bar + kGemmStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * kGemmBM, n0 = blockIdx.y * BN, z = blockIdx.z;
  const int num_kb = (shape.K + kGemmBK - 1) / kGemmBK;
  int64_t a_row0 = (int64_t)z * shape.a_batch_rows, b_row0 = (int64_t)z * shape.b_batch_rows;
  int m_rows = shape.M;
  int64_t c_base = (int64_t)z * ep.c_batch_stride;
  int64_t ldc = ep.ldc;
  if (shape.groups) {
    a_row0 = shape.groups[6 * z + 0];
    b_row0 = shape.groups[6 * z + 1];
    m_rows = (int)shape.groups[6 * z + 2];
    c_base = shape.groups[6 * z + 3];
    ldc = shape.groups[6 * z + 4];
    if (m0 >= m_rows) return;  // uniform for the whole CTA, before any barrier / TMEM allocation
  }

  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&tma_a);
    tc::tma_prefetch_desc(&tma_b);
    for (int s = 0; s < kGemmStages; ++s) {
      tc::mbar_init(&full_bar[s], 1);
      tc::mbar_init(&empty_bar[s], 1);
    }
    tc::mbar_init(tmem_full_bar, 1);
    tc::mbar_fence_init();
  }
  if (warp == 1) tc::tmem_alloc<kTmemCols>(tmem_ptr);
  tc::tcgen05_fence_before_sync();
  __syncthreads();
  tc::tcgen05_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    if (lane == 0) {
      const int a_row = (int)a_row0 + m0;
      const int b_row = (int)b_row0 + n0;
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kGemmStages;
        const uint32_t phase = (kb / kGemmStages) & 1;
        tc::mbar_wait(&empty_bar[s], phase ^ 1);
        tc::mbar_arrive_expect_tx(&full_bar[s], S::kStageBytes);
        uint8_t* a_dst = smem + s * S::kStageBytes;
        tc::tma_load_2d(a_dst, &tma_a, &full_bar[s], kb * kGemmBK, a_row);
        tc::tma_load_2d(a_dst + S::kABytes, &tma_b, &full_bar[s], kb * kGemmBK, b_row);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = tc::umma_idesc_bf16(kGemmBM, BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kGemmStages;
        const uint32_t phase = (kb / kGemmStages) & 1;
        tc::mbar_wait(&full_bar[s], phase);
        tc::tcgen05_fence_after_sync();
        const uint32_t a_addr = tc::smem_u32(smem + s * S::kStageBytes);
        const uint64_t a_desc = tc::umma_desc_sw128(a_addr);
        const uint64_t b_desc = tc::umma_desc_sw128(a_addr + S::kABytes);
#pragma unroll
        for (int k = 0; k < kGemmBK / 16; ++k) {
          // +32 bytes per UMMA_K step inside the 128-byte swizzle atom (encoded >> 4)
          tc::umma_bf16(tmem_base, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
        }
        tc::umma_commit(&empty_bar[s]);  // frees the smem stage once these MMAs retire
      }
      tc::umma_commit(tmem_full_bar);
    }
  } else {
    // epilogue: warp w may only touch TMEM lanes [32*(w%4), 32*(w%4)+32)
    const int lane_base = (warp & 3) * 32;
    const int row = m0 + lane_base + lane;
    tc::mbar_wait(tmem_full_bar, 0);
    tc::tcgen05_fence_after_sync();
    const bool row_ok = row < m_rows;
    const int64_t c_off = c_base + (int64_t)row * ldc + n0;
    // fused GroupNorm statistics: per-warp accumulators in smem, one set of fp64 atomics per tile
    float* gn_acc = reinterpret_cast<float*>(smem + S::kGnOffset);
    float* warp_acc = gn_acc + (warp & 3) * 128;
    bool gn_uniform = false;
    int gn_seg = 0;
    double* gn_row_stats = nullptr;
    if (ep.gn_stats) {
#pragma unroll
      for (int i = 0; i < 4; ++i) warp_acc[lane * 4 + i] = 0.f;
      const int last = min(m0 + kGemmBM, m_rows) - 1;
      gn_seg = segment_of(ep.gn_seg_off, ep.gn_nseg, m0 / ep.gn_rpp);
      gn_uniform = segment_of(ep.gn_seg_off, ep.gn_nseg, last / ep.gn_rpp) == gn_seg;
      if (!gn_uniform && row_ok)
        gn_row_stats = ep.gn_stats +
                       (int64_t)segment_of(ep.gn_seg_off, ep.gn_nseg, row / ep.gn_rpp) * ep.gn_groups * 2;
      __syncwarp();
    }
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      if (BN >= 32) {
        tc::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)c0, r);
      } else {
        // BN == 16: only 16 valid columns were written; load 32 (allocated) and ignore the rest
        tc::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)lane_base << 16), r);
      }
      tc::tmem_ld_wait();
      constexpr int kCols = BN >= 32 ? 32 : BN;
      float v[kCols];
#pragma unroll
      for (int j = 0; j < kCols; ++j) {
        float x = __uint_as_float(r[j]) * ep.alpha;
        if (ep.bias) x += __ldg(ep.bias + n0 + c0 + j);
        if (ep.act == 1) x = fmaxf(x, 0.f);
        v[j] = x;
      }
      if (ep.gn_stats) {
        const int cpg = ep.gn_cpg;
        const int g_glob = (n0 + c0) / cpg, g_loc = g_glob - n0 / cpg;
gn_accumulate_chunk<kCols>(cpg, v, row_ok, gn_uniform, lane, warp_acc, g_loc, gn_row_stats, g_glob);
      }
      if (row_ok && ep.transposed) {
#pragma unroll
        for (int j = 0; j < kCols; ++j)
          if (n0 + c0 + j < ep.n_valid) ep.out_f32[c_base + (int64_t)(n0 + c0 + j) * ldc + row] = v[j];
      } else if (row_ok) {
        if (ep.out_f32) {
          float4* dst = reinterpret_cast<float4*>(ep.out_f32 + c_off + c0);
#pragma unroll
          for (int j = 0; j < kCols / 4; ++j) dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        }
        if (ep.out_bf16) {
          uint4* dst = reinterpret_cast<uint4*>(ep.out_bf16 + c_off + c0);
#pragma unroll
          for (int j = 0; j < kCols / 8; ++j) {
            __nv_bfloat162 p0 = __floats2bfloat162_rn(v[8 * j], v[8 * j + 1]);
            __nv_bfloat162 p1 = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
            __nv_bfloat162 p2 = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]);
            __nv_bfloat162 p3 = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
            uint4 u;
            u.x = *reinterpret_cast<uint32_t*>(&p0);
            u.y = *reinterpret_cast<uint32_t*>(&p1);
            u.z = *reinterpret_cast<uint32_t*>(&p2);
            u.w = *reinterpret_cast<uint32_t*>(&p3);
            dst[j] = u;
          }
        }
      }
    }
    if (ep.gn_stats) {
      asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps only
      if (gn_uniform) {
        const int e = (warp & 3) * 32 + lane;
        const int ngr2 = 2 * ((n0 + BN - 1) / ep.gn_cpg - n0 / ep.gn_cpg + 1);
        for (int i = e; i < ngr2; i += 128) {
          const float t = gn_acc[i] + gn_acc[128 + i] + gn_acc[256 + i] + gn_acc[384 + i];
          atomicAdd(ep.gn_stats + ((int64_t)gn_seg * ep.gn_groups + n0 / ep.gn_cpg) * 2 + i, (double)t);
        }
      }
    }
  }
  tc::tcgen05_fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<kTmemCols>(tmem_base);
}

// ---- host side ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D bf16 row-major [rows, cols] (row pitch `ld` elements) -> tensor map with a [box_rows x 64] box, 128B swizzle
int make_tmap_bf16_2d(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_last_error("cuTensorMapEncodeTiled unavailable", cudaErrorNotSupported); return SE3ET_ERR_CUDA; }
  if ((ld * 2) % 16 != 0 || (reinterpret_cast<uintptr_t>(base) & 15) != 0) return SE3ET_ERR_ARG;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)kGemmBK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_last_error("cuTensorMapEncodeTiled", cudaErrorInvalidValue); return SE3ET_ERR_CUDA; }
  return SE3ET_OK;
}

template <int BN>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmShape& shape, const GemmEpilogue& ep,
                       int batch, cudaStream_t st) {
  using S = GemmSmem<BN>;
  static bool configured = false;
  if (!configured) {
    SE3ET_CUDA_CHECK(cudaFuncSetAttribute(gemm_tma_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::kTotal));
    configured = true;
  }
  dim3 grid((unsigned)ceil_div(shape.M, kGemmBM), (unsigned)(shape.N / BN), (unsigned)batch);
  gemm_tma_kernel<BN><<<grid, kGemmThreads, S::kTotal, st>>>(ta, tb, shape, ep);
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}

int pick_bn(int n) {
  for (int bn : {256, 128, 64, 32, 16})
    if (n % bn == 0) return bn;
  return 0;
}

// a: [batch*a_batch_rows (or M), K] bf16 row-major with pitch lda; b: [.., K] bf16 with pitch ldb
int gemm_bf16(const __nv_bfloat16* a, int64_t lda, const __nv_bfloat16* b, int64_t ldb, int M, int N, int K, int batch,
              int64_t a_batch_rows, int64_t b_batch_rows, const int64_t* groups, int64_t a_rows_total,
              int64_t b_rows_total, const GemmEpilogue& ep, cudaStream_t st) {
  if (M <= 0 || batch <= 0) return SE3ET_OK;
  if (N <= 0 || K <= 0 || !a || !b) return SE3ET_ERR_ARG;
  const int bn = pick_bn(N);
  if (!bn) return SE3ET_ERR_UNSUPPORTED;
  if (ep.transposed && (!ep.out_f32 || ep.out_bf16)) return SE3ET_ERR_ARG;
  if (!ep.transposed && ep.out_f32 &&
      ((reinterpret_cast<uintptr_t>(ep.out_f32) & 15) || (ep.ldc % 4) || (ep.c_batch_stride % 4) || groups))
    return SE3ET_ERR_ARG;
  if (ep.out_bf16 && ((reinterpret_cast<uintptr_t>(ep.out_bf16) & 15) || (ep.ldc % 8) || (ep.c_batch_stride % 8)))
    return SE3ET_ERR_ARG;
  CUtensorMap ta, tb;
  int64_t a_rows = batch > 1 ? (int64_t)(batch - 1) * a_batch_rows + M : M;
  int64_t b_rows = (batch > 1 && b_batch_rows > 0) ? (int64_t)(batch - 1) * b_batch_rows + N : N;
  if (a_rows_total > 0) a_rows = a_rows_total;  // true extents: rows beyond them are zero-filled by TMA
  if (b_rows_total > 0) b_rows = b_rows_total;
  if (groups && (a_rows_total <= 0 || b_rows_total <= 0)) return SE3ET_ERR_ARG;
  int rc = make_tmap_bf16_2d(&ta, a, a_rows, K, lda, kGemmBM);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tb, b, b_rows, K, ldb, bn);
  if (rc) return rc;
  GemmShape shape{M, N, K, a_batch_rows, b_batch_rows, groups};
  switch (bn) {
    case 256: return launch_gemm<256>(ta, tb, shape, ep, batch, st);
    case 128: return launch_gemm<128>(ta, tb, shape, ep, batch, st);
    case 64: return launch_gemm<64>(ta, tb, shape, ep, batch, st);
    case 32: return launch_gemm<32>(ta, tb, shape, ep, batch, st);
    default: return launch_gemm<16>(ta, tb, shape, ep, batch, st);
  }
}

}  // namespace se3et

using namespace se3et;

extern "C" int se3et_gemm_grouped_bf16(const void* a, int64_t lda, int64_t a_rows_total, const void* b, int64_t ldb,
                                       int64_t b_rows_total, const int64_t* groups, int64_t num_groups, int64_t max_m,
                                       int64_t n, int64_t n_valid, int64_t k, float alpha, float* out_f32, int64_t ldc,
                                       int transposed, se3et_stream_t stream) {
  if (max_m < 0 || n <= 0 || k <= 0 || num_groups < 0 || !groups || !out_f32 || max_m > INT32_MAX) return SE3ET_ERR_ARG;
  if (num_groups == 0 || max_m == 0) return SE3ET_OK;
  if (num_groups > 65535) return SE3ET_ERR_UNSUPPORTED;
  GemmEpilogue ep;
  ep.out_f32 = out_f32;
  ep.out_bf16 = nullptr;
  ep.bias = nullptr;
  ep.ldc = ldc;
  ep.c_batch_stride = 0;
  ep.alpha = alpha;
  ep.act = 0;
  ep.transposed = transposed ? 1 : 0;
  ep.n_valid = (int)(n_valid > 0 ? n_valid : n);
  if (!transposed) return SE3ET_ERR_UNSUPPORTED;
  return gemm_bf16(static_cast<const __nv_bfloat16*>(a), lda, static_cast<const __nv_bfloat16*>(b), ldb, (int)max_m,
                   (int)n, (int)k, (int)num_groups, 0, 0, groups, a_rows_total, b_rows_total, ep,
                   static_cast<cudaStream_t>(stream));
}

extern "C" int se3et_gemm_bf16(const void* a, int64_t lda, const void* b, int64_t ldb, int64_t m, int64_t n, int64_t k,
                               int64_t batch, int64_t a_batch_rows, int64_t b_batch_rows, const float* bias,
                               float alpha, int act, float* out_f32, void* out_bf16, int64_t ldc,
                               int64_t c_batch_stride, se3et_stream_t stream) {
  if (m < 0 || n <= 0 || k <= 0 || batch <= 0 || m > INT32_MAX || n > INT32_MAX || k > INT32_MAX) return SE3ET_ERR_ARG;
  if (!out_f32 && !out_bf16) return SE3ET_ERR_ARG;
  GemmEpilogue ep;
  ep.out_f32 = out_f32;
  ep.out_bf16 = static_cast<__nv_bfloat16*>(out_bf16);
  ep.bias = bias;
  ep.ldc = ldc;
  ep.c_batch_stride = c_batch_stride;
  ep.alpha = alpha;
  ep.act = act;
  ep.transposed = 0;
  ep.n_valid = (int)n;
  if (batch > 65535) return SE3ET_ERR_UNSUPPORTED;
  return gemm_bf16(static_cast<const __nv_bfloat16*>(a), lda, static_cast<const __nv_bfloat16*>(b), ldb, (int)m, (int)n,
                   (int)k, (int)batch, a_batch_rows, b_batch_rows, nullptr, 0, 0, ep, static_cast<cudaStream_t>(stream));
}

extern "C" int se3et_gemm_bf16_gnstats(const void* a, int64_t lda, const void* b, int64_t ldb, int64_t m, int64_t n,
                                       int64_t k, const float* bias, float* out_f32, int64_t ldc, double* stats,
                                       const int64_t* seg_offsets, int64_t nseg, int64_t groups,
                                       int64_t rows_per_point, se3et_stream_t stream) {
  if (m < 0 || n <= 0 || k <= 0 || m > INT32_MAX || n > INT32_MAX || k > INT32_MAX) return SE3ET_ERR_ARG;
  if (!out_f32 || !stats || !seg_offsets || nseg <= 0 || groups <= 0 || n % groups || rows_per_point <= 0)
    return SE3ET_ERR_ARG;
  const int bn = pick_bn((int)n);
  if (!bn) return SE3ET_ERR_UNSUPPORTED;
  const int64_t cpg = n / groups;
  const int chunk = bn >= 32 ? 32 : bn;
  const bool pow2 = (cpg & (cpg - 1)) == 0;
  // the epilogue reduces groups inside 32-column chunks: cpg | chunk, or chunk | cpg; <= 64 groups per tile
  if (!((pow2 && cpg <= chunk) || cpg % chunk == 0) || bn / cpg > 64) return SE3ET_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SE3ET_CUDA_CHECK(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * nseg * groups, st));
  if (m == 0) return SE3ET_OK;
  GemmEpilogue ep;
  ep.out_f32 = out_f32;
  ep.out_bf16 = nullptr;
  ep.bias = bias;
  ep.ldc = ldc;
  ep.c_batch_stride = 0;
  ep.alpha = 1.f;
  ep.act = 0;
  ep.transposed = 0;
  ep.n_valid = (int)n;
  ep.gn_stats = stats;
  ep.gn_seg_off = seg_offsets;
  ep.gn_nseg = (int)nseg;
  ep.gn_cpg = (int)cpg;
  ep.gn_groups = (int)groups;
  ep.gn_rpp = (int)rows_per_point;
  return gemm_bf16(static_cast<const __nv_bfloat16*>(a), lda, static_cast<const __nv_bfloat16*>(b), ldb, (int)m, (int)n,
                   (int)k, 1, 0, 0, nullptr, 0, 0, ep, st);
}
