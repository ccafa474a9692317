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
#include "gemm_stream.cuh"
#include "gn_epilogue.cuh"
#include "tc.cuh"

namespace se3et {

constexpr int kGemmBM = 128;
constexpr int kGemmBK = 64;  // 64 bf16 = 128 bytes = one swizzle atom
constexpr int kGemmMaxStages = 4;
constexpr int kGemmThreads = 192;
constexpr int kEpCols = 16;  // epilogue chunk: columns per tcgen05.ld

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
  // GroupNorm over (pair, channel group); a row r belongs to point r / gn_rpp, pairs are gn_seg_off ranges of points
  const int64_t* gn_seg_off = nullptr;
  int gn_nseg = 0;
  int gn_cpg = 1;            // channels per group: a power of two <= 16, or a multiple of 16
  int gn_groups = 0;
  int gn_rpp = 1;            // rows per point (6 for equivariant features)
  // (a) accumulate the statistics of the fp32 values (se3et_gemm_bf16_gnstats): [nseg, groups, 2] {sum, sum sq}
  double* gn_stats = nullptr;
  // (b) apply: out_bf16 = LeakyReLU_slope((v - mean) * rstd * gamma + beta [+ resid])   (se3et_gemm_bf16_gnapply)
  const double* norm_stats = nullptr;
  const float* norm_gamma = nullptr;
  const float* norm_beta = nullptr;
  const __nv_bfloat16* norm_resid = nullptr;  // nullable, [M, ldc]
  float norm_eps = 1e-5f;
  float norm_slope = 1.f;
};

struct GemmShape {
  int M, N, K;               // per batch (M = largest group when a group table is used)
  int64_t a_batch_rows;      // rows between batches in the flattened A map
  int64_t b_batch_rows;      // rows between batches in the flattened B map (0 = shared B)
  const int64_t* groups;     // optional device table, 6 int64 per blockIdx.z: {a_row0, b_row0, m_rows, c_off, ldc, -}
  int stages;                // smem ring depth, 1..kGemmMaxStages (fewer stages -> more CTAs per SM for short K)
  int tiles_per_cta;         // consecutive 128-row tiles one CTA walks (loads / MMA / epilogue overlap across tiles)
};

template <int BN>
struct GemmSmem {
  static constexpr int kABytes = kGemmBM * 128;
  static constexpr int kBBytes = BN * 128;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTail = 128 + 4 * 64 * 2 * 4 + 64 * 8;  // barriers, GN accumulators, (mean, rstd) table
  static constexpr int kCBytes = kGemmBM * BN * 2;             // bf16 output tile staged for coalesced stores
  __host__ __device__ static int bar_offset(int stages) { return stages * kStageBytes; }
  // tail: 128 B barriers | 2 KB GN accumulators | 512 B (mean, rstd) | 2 KB per-column (scale, shift) | pad
  __host__ __device__ static int c_offset(int stages) { return stages * kStageBytes + 5120; }
  static int total(int stages, bool c_tile) { return stages * kStageBytes + 5120 + (c_tile ? kCBytes : 0) + 1024; }
};

// kMulti = true : strips of row tiles per CTA, double-buffered accumulators (long K: the ring hides the load latency)
// kMulti = false: one tile per CTA, <= 112 registers so that three CTAs share an SM (K of one or two blocks: the
//                 epilogue dominates and more resident epilogue warps win)
template <int BN, bool kMulti>
__global__ void __launch_bounds__(kGemmThreads, (kMulti ? (BN <= 128 ? 2 : 1) : (BN <= 128 ? 3 : 2)))
gemm_tma_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, GemmShape shape,
                GemmEpilogue ep) {
  using S = GemmSmem<BN>;
  constexpr uint32_t kTmemCols = BN < 32 ? 32 : BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int stages = shape.stages;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::bar_offset(stages));
  uint64_t* empty_bar = full_bar + kGemmMaxStages;
  uint64_t* tmem_full_bar = empty_bar + kGemmMaxStages;   // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;            // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
  float* gn_acc = reinterpret_cast<float*>(smem + S::bar_offset(stages) + 128);  // [4 warps][64 groups][2]
  float2* gn_tab = reinterpret_cast<float2*>(gn_acc + 4 * 128);                  // [64 groups] {mean, rstd}
  float2* col_tab = reinterpret_cast<float2*>(gn_tab + 64);                       // [BN] apply mode: {scale, shift}
  uint8_t* c_tile = smem + S::c_offset(stages);  // apply mode only: [128 rows][BN bf16], 16-byte chunks swizzled

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * BN, z = blockIdx.z;
  const int tile0 = blockIdx.x * shape.tiles_per_cta;
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
    if (tile0 * kGemmBM >= m_rows) return;  // uniform for the whole CTA, before any barrier / TMEM allocation
  }
  // tiles this CTA owns: [tile0, tile0 + ntiles)
  const int ntiles = kMulti ? min(shape.tiles_per_cta, (m_rows - tile0 * kGemmBM + kGemmBM - 1) / kGemmBM) : 1;
  // two accumulator buffers when they fit TMEM next to a co-resident CTA (epilogue of tile t under the MMA of t + 1)
  constexpr int kAccBufs = (kMulti && BN <= 128) ? 2 : 1;
  constexpr uint32_t kTmemAlloc = kTmemCols * kAccBufs;

  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&tma_a);
    tc::tma_prefetch_desc(&tma_b);
    for (int s = 0; s < stages; ++s) {
      tc::mbar_init(&full_bar[s], 1);
      tc::mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(&tmem_full_bar[b], 1);
      tc::mbar_init(&tmem_empty_bar[b], 4);
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
      const int b_row = (int)b_row0 + n0;
      int s = 0;
      uint32_t phase = 0;
      for (int t = 0; t < ntiles; ++t) {
        const int a_row = (int)a_row0 + (tile0 + t) * kGemmBM;
        for (int kb = 0; kb < num_kb; ++kb) {
          tc::mbar_wait_long(&empty_bar[s], phase ^ 1);
          tc::mbar_arrive_expect_tx(&full_bar[s], S::kStageBytes);
          uint8_t* a_dst = smem + s * S::kStageBytes;
          tc::tma_load_2d(a_dst, &tma_a, &full_bar[s], kb * kGemmBK, a_row);
          tc::tma_load_2d(a_dst + S::kABytes, &tma_b, &full_bar[s], kb * kGemmBK, b_row);
          if (++s == stages) { s = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = tc::umma_idesc_bf16(kGemmBM, BN);
      int s = 0;
      uint32_t phase = 0;
      for (int t = 0; t < ntiles; ++t) {
        const int buf = t % kAccBufs;
        const uint32_t use = (uint32_t)(t / kAccBufs);
        tc::mbar_wait_long(&tmem_empty_bar[buf], (use & 1) ^ 1);  // the epilogue has drained this accumulator buffer
        tc::tcgen05_fence_after_sync();
        const uint32_t tmem_acc = tmem_base + (uint32_t)(buf * kTmemCols);
        for (int kb = 0; kb < num_kb; ++kb) {
          tc::mbar_wait_long(&full_bar[s], phase);
          tc::tcgen05_fence_after_sync();
          const uint32_t a_addr = tc::smem_u32(smem + s * S::kStageBytes);
          const uint64_t a_desc = tc::umma_desc_sw128(a_addr);
          const uint64_t b_desc = tc::umma_desc_sw128(a_addr + S::kABytes);
#pragma unroll
          for (int k = 0; k < kGemmBK / 16; ++k) {
            // +32 bytes per UMMA_K step inside the 128-byte swizzle atom (encoded >> 4)
            tc::umma_bf16(tmem_acc, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          }
          tc::umma_commit(&empty_bar[s]);  // frees the smem stage once these MMAs retire
          if (++s == stages) { s = 0; phase ^= 1; }
        }
        tc::umma_commit(&tmem_full_bar[buf]);
      }
    }
  } else {
    // epilogue: warp w may only touch TMEM lanes [32*(w%4), 32*(w%4)+32)
    const int ew = warp & 3;
    const int lane_base = ew * 32;
    const int et = ew * 32 + lane;      // epilogue thread 0..127
    const bool use_gn = ep.gn_stats != nullptr || ep.norm_stats != nullptr;
    float* warp_acc = gn_acc + ew * 128;
    const int cpg = ep.gn_cpg;
    const int g_tile0 = n0 / cpg;
    constexpr int kRowChunks = BN / 8;  // 16-byte chunks per tile row
    constexpr int kSwz = kRowChunks >= 8 ? 7 : kRowChunks - 1;  // swizzle stays inside the row
    for (int t = 0; t < ntiles; ++t) {
      const int m0 = (tile0 + t) * kGemmBM;
      const int buf = t % kAccBufs;
      const uint32_t use = (uint32_t)(t / kAccBufs);
      const uint32_t tmem_acc = tmem_base + (uint32_t)(buf * kTmemCols);
      const int row = m0 + lane_base + lane;
      const bool row_ok = row < m_rows;
      const int64_t c_off = c_base + (int64_t)row * ldc + n0;
      bool gn_uniform = false;
      int gn_seg = 0, row_seg = 0;
      double* gn_row_stats = nullptr;
      // ---- before the accumulators are ready: pair lookup, per-column table of this tile, residual prefetch
      if (use_gn) {
        const int last = min(m0 + kGemmBM, m_rows) - 1;
        gn_seg = segment_of(ep.gn_seg_off, ep.gn_nseg, m0 / ep.gn_rpp);
        gn_uniform = segment_of(ep.gn_seg_off, ep.gn_nseg, last / ep.gn_rpp) == gn_seg;
        row_seg = gn_uniform ? gn_seg : segment_of(ep.gn_seg_off, ep.gn_nseg, (row_ok ? row : last) / ep.gn_rpp);
        if (ep.gn_stats) {
#pragma unroll
          for (int i = 0; i < 4; ++i) warp_acc[lane * 4 + i] = 0.f;
          if (!gn_uniform && row_ok) gn_row_stats = ep.gn_stats + (int64_t)row_seg * ep.gn_groups * 2;
        }
        if (ep.norm_stats && gn_uniform) {
          // per-column affine map of this tile: x_norm = acc * scale + shift (bias, mean, rstd, gamma, beta folded)
          const double cnt = (double)(ep.gn_seg_off[gn_seg + 1] - ep.gn_seg_off[gn_seg]) * ep.gn_rpp * cpg;
          for (int cc = et; cc < BN; cc += 128) {
            const int c = n0 + cc;
            const double* st = ep.norm_stats + ((int64_t)gn_seg * ep.gn_groups + c / cpg) * 2;
            const double mean = st[0] / cnt;
            const double var = st[1] / cnt - mean * mean;
            const float sc = rsqrtf((float)fmax(var, 0.0) + ep.norm_eps) * ep.norm_gamma[c];
            const float b = ep.bias ? ep.bias[c] : 0.f;
            col_tab[cc] = make_float2(sc, ep.norm_beta[c] + (b - (float)mean) * sc);
          }
        }
        __syncwarp();
      }
      // apply mode: the output tile is staged in shared memory so that global loads (residual) and stores are
      // row-contiguous.  chunk (row, j) of 16 bytes lives at row * BN * 2 + ((j ^ (row & kSwz)) << 4).
      if (ep.norm_stats && ep.norm_resid) {
        for (int i = et; i < kGemmBM * kRowChunks; i += 128) {
          const int rr = i / kRowChunks, j = i - rr * kRowChunks;
          if (m0 + rr < m_rows) {
            const uint32_t dst = tc::smem_u32(c_tile) + rr * (BN * 2) + ((j ^ (rr & kSwz)) << 4);
            const __nv_bfloat16* src = ep.norm_resid + c_base + (int64_t)(m0 + rr) * ldc + n0 + j * 8;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      }
      tc::mbar_wait_long(&tmem_full_bar[buf], use & 1);
      tc::tcgen05_fence_after_sync();
      if (ep.norm_stats) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        asm volatile("bar.sync 1, 128;" ::: "memory");  // residual tile and column table visible to the 4 warps
      }
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += kEpCols) {
        uint32_t r[kEpCols];
        tc::tmem_ld_32x32b_x16(tmem_acc + ((uint32_t)lane_base << 16) + (uint32_t)c0, r);
        // residual row segment of the apply mode (this thread's own row of the staged tile)
        uint4 res[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
        const int trow = lane_base + lane;
        uint8_t* crow = c_tile + trow * (BN * 2);
        const int cj = c0 / 8;
        if (ep.norm_stats && ep.norm_resid) {
          res[0] = *reinterpret_cast<const uint4*>(crow + (((cj) ^ (trow & kSwz)) << 4));
          res[1] = *reinterpret_cast<const uint4*>(crow + (((cj + 1) ^ (trow & kSwz)) << 4));
        }
        tc::tmem_ld_wait();
        if (ep.norm_stats) {
          const uint32_t rw[8] = {res[0].x, res[0].y, res[0].z, res[0].w, res[1].x, res[1].y, res[1].z, res[1].w};
          float v[kEpCols];
#pragma unroll
          for (int j = 0; j < kEpCols; ++j) {
            float x;
            if (gn_uniform) {
              const float2 tb = col_tab[c0 + j];
              x = fmaf(__uint_as_float(r[j]), tb.x, tb.y);
            } else {  // tile straddles a pair boundary: per-row statistics straight from global memory
              const int c = n0 + c0 + j;
              const double cnt = (double)(ep.gn_seg_off[row_seg + 1] - ep.gn_seg_off[row_seg]) * ep.gn_rpp * cpg;
              const double* st = ep.norm_stats + ((int64_t)row_seg * ep.gn_groups + c / cpg) * 2;
              const double mean = st[0] / cnt;
              const double var = st[1] / cnt - mean * mean;
              const float sc = rsqrtf((float)fmax(var, 0.0) + ep.norm_eps) * __ldg(ep.norm_gamma + c);
              const float b = ep.bias ? __ldg(ep.bias + c) : 0.f;
              x = (__uint_as_float(r[j]) + b - (float)mean) * sc + __ldg(ep.norm_beta + c);
            }
            const uint32_t w = rw[j >> 1];
            x += (j & 1) ? __uint_as_float(w & 0xffff0000u) : __uint_as_float(w << 16);
            v[j] = fmaxf(x, x * ep.norm_slope);  // LeakyReLU for slope <= 1
          }
          // result back into the staged tile (same chunks the residual came from)
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
            *reinterpret_cast<uint4*>(crow + (((cj + jj) ^ (trow & kSwz)) << 4)) = u;
          }
          continue;
        }
        float v[kEpCols];
#pragma unroll
        for (int j = 0; j < kEpCols; ++j) {
          float x = __uint_as_float(r[j]) * ep.alpha;
          if (ep.bias) x += __ldg(ep.bias + n0 + c0 + j);
          if (ep.act == 1) x = fmaxf(x, 0.f);
          v[j] = x;
        }
        if (ep.gn_stats) {
          const int g_glob = (n0 + c0) / cpg, g_loc = g_glob - g_tile0;
          gn_accumulate_chunk<kEpCols>(cpg, v, row_ok, gn_uniform, lane, warp_acc, g_loc, gn_row_stats, g_glob);
        }
        if (row_ok && ep.transposed) {
#pragma unroll
          for (int j = 0; j < kEpCols; ++j)
            if (n0 + c0 + j < ep.n_valid) ep.out_f32[c_base + (int64_t)(n0 + c0 + j) * ldc + row] = v[j];
        } else if (row_ok) {
          if (ep.out_f32) {
            float4* dst = reinterpret_cast<float4*>(ep.out_f32 + c_off + c0);
#pragma unroll
            for (int j = 0; j < kEpCols / 4; ++j)
              dst[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
          if (ep.out_bf16) {
            uint4* dst = reinterpret_cast<uint4*>(ep.out_bf16 + c_off + c0);
#pragma unroll
            for (int j = 0; j < kEpCols / 8; ++j) {
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
      // the accumulator buffer may be overwritten by the MMA of tile t + kAccBufs
      tc::tcgen05_fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tmem_empty_bar[buf]);
      if (ep.norm_stats) {
        // staged tile -> global memory, row-contiguous 16-byte stores
        asm volatile("bar.sync 1, 128;" ::: "memory");
        for (int i = et; i < kGemmBM * kRowChunks; i += 128) {
          const int rr = i / kRowChunks, j = i - rr * kRowChunks;
          if (m0 + rr < m_rows)
            *reinterpret_cast<uint4*>(ep.out_bf16 + c_base + (int64_t)(m0 + rr) * ldc + n0 + j * 8) =
                *reinterpret_cast<const uint4*>(c_tile + rr * (BN * 2) + ((j ^ (rr & kSwz)) << 4));
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");  // before the next tile's residual lands in the staging tile
      }
      if (ep.gn_stats) {
        asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps only
        if (gn_uniform) {
          const int ngr2 = 2 * ((n0 + BN - 1) / cpg - g_tile0 + 1);
          for (int i = et; i < ngr2; i += 128) {
            const float tt = gn_acc[i] + gn_acc[128 + i] + gn_acc[256 + i] + gn_acc[384 + i];
            atomicAdd(ep.gn_stats + ((int64_t)gn_seg * ep.gn_groups + g_tile0) * 2 + i, (double)tt);
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");  // accumulators are re-zeroed by the next tile
      }
    }
  }
  tc::tcgen05_fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<kTmemAlloc>(tmem_base);
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

template <int BN, bool kMulti>
static int launch_gemm_variant(const CUtensorMap& ta, const CUtensorMap& tb, GemmShape shape, const GemmEpilogue& ep,
                               int batch, int max_smem, cudaStream_t st) {
  using S = GemmSmem<BN>;
  const int num_kb = (shape.K + kGemmBK - 1) / kGemmBK;
  const bool c_tile = ep.norm_stats != nullptr;
  // ring depth: as deep as the co-resident CTAs allow; the strip variant keeps the ring running across its tiles
  shape.stages = kMulti ? kGemmMaxStages : (num_kb < kGemmMaxStages ? num_kb : kGemmMaxStages);
  while (shape.stages > (kMulti ? 2 : 1) && S::total(shape.stages, c_tile) > max_smem) --shape.stages;
  const int smem = S::total(shape.stages, c_tile);
  SE3ET_ENSURE_SMEM((gemm_tma_kernel<BN, kMulti>), smem);
  const int64_t m_tiles = ceil_div(shape.M, kGemmBM);
  const int64_t n_tiles = shape.N / BN;
  int64_t tpc = 1;
  if (kMulti && !shape.groups) {
    // strip of consecutive row tiles per CTA: enough CTAs for ~4 waves of 2 per SM, at most 16 tiles each
    tpc = (m_tiles * n_tiles * batch) / ((int64_t)kNumSMs * 8);
    tpc = tpc < 1 ? 1 : (tpc > 16 ? 16 : tpc);
  }
  shape.tiles_per_cta = (int)tpc;
  dim3 grid((unsigned)ceil_div(m_tiles, tpc), (unsigned)n_tiles, (unsigned)batch);
  gemm_tma_kernel<BN, kMulti><<<grid, kGemmThreads, smem, st>>>(ta, tb, shape, ep);
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}

static bool g_grouped_small_cta = true;
// widest output tile of the plain / statistics GEMMs: 128 columns keep two CTAs (eight epilogue warps) per SM with
// double-buffered accumulators; 256-wide tiles are epilogue-bound at the K of this path (566 vs 553 pairs/s)
static int g_plain_bn_cap = 128;

template <int BN>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmShape& shape, const GemmEpilogue& ep,
                       int batch, cudaStream_t st) {
  const int num_kb = (shape.K + kGemmBK - 1) / kGemmBK;
  // short K: three (BN <= 128) or two single-tile CTAs per SM.  Grouped problems with narrow outputs (the positional
  // score term: one small problem per query point, N = 32) are bound by CTAs in flight, not by the ring depth: same variant
  if ((num_kb <= 2 && !shape.groups) || (shape.groups && BN <= 64 && g_grouped_small_cta))
    return launch_gemm_variant<BN, false>(ta, tb, shape, ep, batch, (BN <= 128 ? 75 : 113) * 1024, st);
  return launch_gemm_variant<BN, true>(ta, tb, shape, ep, batch, (BN <= 128 ? 113 : 227) * 1024, st);
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
  int bn = pick_bn(N);
  if (!bn) return SE3ET_ERR_UNSUPPORTED;
  if (ep.norm_stats && bn > 128) bn = 128;  // apply mode stages its output tile: 3 CTAs per SM
  if (bn > g_plain_bn_cap && !groups) bn = g_plain_bn_cap;
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
  GemmShape shape{M, N, K, a_batch_rows, b_batch_rows, groups, kGemmMaxStages, 1};
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

extern "C" int se3et_gemm_set_plain_tile_cap(int bn) {
  if (bn != 64 && bn != 128 && bn != 256) return SE3ET_ERR_ARG;
  g_plain_bn_cap = bn;
  return SE3ET_OK;
}

extern "C" int se3et_gemm_set_grouped_small_cta(int on) {
  g_grouped_small_cta = on != 0;
  return SE3ET_OK;
}

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
  // large bf16-only Linears (q/k/v, FFN expand, in_proj): streaming kernel, eight epilogue warps per CTA
  if (gemm_stream_plain_enabled() && batch == 1 && !out_f32 && out_bf16 && m >= 4096 && n % 64 == 0 && k % 8 == 0 &&
      ldc % 8 == 0 && !(reinterpret_cast<uintptr_t>(out_bf16) & 15) && (act == 0 || act == 1) && m > 0)
    return gemm_stream_plain(a, lda, b, ldb, m, n, k, bias, alpha, act == 1 ? 0.f : 1.f, out_bf16, ldc,
                             static_cast<cudaStream_t>(stream));
  return gemm_bf16(static_cast<const __nv_bfloat16*>(a), lda, static_cast<const __nv_bfloat16*>(b), ldb, (int)m, (int)n,
                   (int)k, (int)batch, a_batch_rows, b_batch_rows, nullptr, 0, 0, ep, static_cast<cudaStream_t>(stream));
}

// shapes the GroupNorm epilogues cover: groups inside 16-column chunks (cpg | 16) or chunks inside groups (16 | cpg),
// at most 64 groups per output tile
static bool gn_epilogue_ok(int64_t n, int64_t groups, int bn) {
  if (groups <= 0 || n % groups) return false;
  const int64_t cpg = n / groups;
  const bool pow2 = (cpg & (cpg - 1)) == 0;
  return ((pow2 && cpg <= kEpCols) || cpg % kEpCols == 0) && bn / cpg <= 64;
}

extern "C" int se3et_gemm_bf16_gnstats(const void* a, int64_t lda, const void* b, int64_t ldb, int64_t m, int64_t n,
                                       int64_t k, const float* bias, float* out_f32, int64_t ldc, double* stats,
                                       const int64_t* seg_offsets, int64_t nseg, int64_t groups,
                                       int64_t rows_per_point, se3et_stream_t stream) {
  if (m < 0 || n <= 0 || k <= 0 || m > INT32_MAX || n > INT32_MAX || k > INT32_MAX) return SE3ET_ERR_ARG;
  if (!stats || !seg_offsets || nseg <= 0 || groups <= 0 || n % groups || rows_per_point <= 0) return SE3ET_ERR_ARG;
  const int bn = pick_bn((int)n);
  if (!bn) return SE3ET_ERR_UNSUPPORTED;
  if (!gn_epilogue_ok(n, groups, bn)) return SE3ET_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SE3ET_CUDA_CHECK(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * nseg * groups, st));
  if (m == 0) return SE3ET_OK;
  GemmEpilogue ep;
  ep.out_f32 = out_f32;
  ep.out_bf16 = nullptr;
  ep.bias = bias;
  ep.ldc = out_f32 ? ldc : n;
  ep.c_batch_stride = 0;
  ep.alpha = 1.f;
  ep.act = 0;
  ep.transposed = 0;
  ep.n_valid = (int)n;
  ep.gn_stats = stats;
  ep.gn_seg_off = seg_offsets;
  ep.gn_nseg = (int)nseg;
  ep.gn_cpg = (int)(n / groups);
  ep.gn_groups = (int)groups;
  ep.gn_rpp = (int)rows_per_point;
  return gemm_bf16(static_cast<const __nv_bfloat16*>(a), lda, static_cast<const __nv_bfloat16*>(b), ldb, (int)m, (int)n,
                   (int)k, 1, 0, 0, nullptr, 0, 0, ep, st);
}

extern "C" int se3et_gemm_bf16_gnapply(const void* a, int64_t lda, const void* b, int64_t ldb, int64_t m, int64_t n,
                                       int64_t k, const float* bias, const double* stats, const float* gamma,
                                       const float* beta, float eps, float leaky_slope, const void* resid_bf16,
                                       void* out_bf16, int64_t ldc, const int64_t* seg_offsets, int64_t nseg,
                                       int64_t groups, int64_t rows_per_point, void* workspace,
                                       size_t workspace_bytes, se3et_stream_t stream) {
  if (m < 0 || n <= 0 || k <= 0 || m > INT32_MAX || n > INT32_MAX || k > INT32_MAX) return SE3ET_ERR_ARG;
  if (!stats || !gamma || !beta || !out_bf16 || !seg_offsets || nseg <= 0 || groups <= 0 || n % groups ||
      rows_per_point <= 0)
    return SE3ET_ERR_ARG;
  int bn = pick_bn((int)n);
  if (!bn) return SE3ET_ERR_UNSUPPORTED;
  if (bn > 128) bn = 128;
  if (!gn_epilogue_ok(n, groups, bn)) return SE3ET_ERR_UNSUPPORTED;
  if (resid_bf16 && (reinterpret_cast<uintptr_t>(resid_bf16) & 15)) return SE3ET_ERR_ARG;
  if (leaky_slope > 1.f) return SE3ET_ERR_ARG;
  if (m == 0) return SE3ET_OK;
  if (workspace && gemm_stream_supported(n, k, 0, ldc) && !(reinterpret_cast<uintptr_t>(out_bf16) & 15)) {
    const StreamNorm n1{stats, gamma, beta, bias};
    return gemm_stream_gnapply(a, lda, b, ldb, k, n1, nullptr, 0, nullptr, 0, 0, n1, m, n, eps, leaky_slope, resid_bf16,
                               out_bf16, ldc, seg_offsets, nseg, groups, rows_per_point, workspace, workspace_bytes,
                               static_cast<cudaStream_t>(stream));
  }
  GemmEpilogue ep;
  ep.out_f32 = nullptr;
  ep.out_bf16 = static_cast<__nv_bfloat16*>(out_bf16);
  ep.bias = bias;
  ep.ldc = ldc;
  ep.c_batch_stride = 0;
  ep.alpha = 1.f;
  ep.act = 0;
  ep.transposed = 0;
  ep.n_valid = (int)n;
  ep.gn_seg_off = seg_offsets;
  ep.gn_nseg = (int)nseg;
  ep.gn_cpg = (int)(n / groups);
  ep.gn_groups = (int)groups;
  ep.gn_rpp = (int)rows_per_point;
  ep.norm_stats = stats;
  ep.norm_gamma = gamma;
  ep.norm_beta = beta;
  ep.norm_resid = static_cast<const __nv_bfloat16*>(resid_bf16);
  ep.norm_eps = eps;
  ep.norm_slope = leaky_slope;
  return gemm_bf16(static_cast<const __nv_bfloat16*>(a), lda, static_cast<const __nv_bfloat16*>(b), ldb, (int)m, (int)n,
                   (int)k, 1, 0, 0, nullptr, 0, 0, ep, static_cast<cudaStream_t>(stream));
}
