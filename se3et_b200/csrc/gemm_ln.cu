// Linear + residual + LayerNorm in one kernel for the transformer's 256-wide hidden state:
//     out = LayerNorm( resid[row / resid_div] + A W^T + b )           A bf16 [M, K], W bf16 [256, K], out bf16 [M, 256]
// (rpe_transformer.py:163-175 / vanilla_transformer.py:905-913: `linear` + dropout(identity) + `norm(hidden + input)`;
//  output_layer.py:17-22: `squeeze` + `norm(input + hidden)`).  The two-kernel form wrote the fp32 Linear output (134 MB for
// 130k rows) and read it back in add_layernorm; here a row's 256 outputs never leave the SM.
//
// Persistent CTAs (one per SM) walk contiguous ranges of 128-row tiles; the accumulator is ONE 128 x 256 tile (UMMA N = 256),
// double-buffered: all 512 TMEM columns.
//   warp 0     TMA producer, 4-stage ring of (A 128 x 64, W 256 x 64) K-blocks running ahead across tiles
//   warp 1     tcgen05.mma issuer
//   warps 2-9  epilogue: thread = one row (TMEM lane) x 128 of the 256 columns (two warps share a lane quarter).  The 128
//              values stay in registers: + bias + residual, row mean and centred variance (two-pass, as nn.LayerNorm; the
//              two half-row partials meet through shared memory and a 64-thread named barrier), normalise, bf16 out.
//              Measured and rejected: residual in / result out through a per-warp staging buffer (whole 256-byte row
//              segments per instruction instead of 32 half-used sectors) with a 3-stage ring: 87 -> 122 us for 130k rows.
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc.cuh"

namespace se3et {

int make_tmap_bf16_2d(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows);

namespace gln {

constexpr int kBM = 128, kN = 256, kBK = 64;
constexpr int kStages = 4;
constexpr int kEpWarps = 8;
constexpr int kThreads = (2 + kEpWarps) * 32;
constexpr int kABytes = kBM * 128, kWBytes = kN * 128, kStageBytes = kABytes + kWBytes;
constexpr int kXchgOff = kStages * kStageBytes;          // [2 rounds][8 warps][32 lanes] float
constexpr int kBarOff = kXchgOff + 2 * kEpWarps * 32 * 4;
constexpr int kSmem = kBarOff + 128 + 1024;

struct Args {
  int M, K, m_tiles;
  const float* bias;              // [256], nullable
  const __nv_bfloat16* resid;     // [ceil(M / resid_div), 256], nullable
  int resid_div;
  const float* gamma;
  const float* beta;
  float eps;
  __nv_bfloat16* out;             // [M, 256]
};

__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}

__global__ void __launch_bounds__(kThreads, 1)
linear_add_layernorm_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_w, Args args) {
  constexpr uint32_t kAcc = kN;
  constexpr uint32_t kTmemAlloc = 2 * kAcc;  // 512
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* xchg = reinterpret_cast<float*>(smem + kXchgOff);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kBarOff);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;  // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;   // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nkb = (args.K + kBK - 1) / kBK;
  const int64_t W = args.m_tiles;
  const int64_t w_begin = W * blockIdx.x / gridDim.x, w_end = W * (blockIdx.x + 1) / gridDim.x;

  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&tma_a);
    tc::tma_prefetch_desc(&tma_w);
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
        for (int kb = 0; kb < nkb; ++kb) {
          tc::mbar_wait_long(&empty_bar[s], phase ^ 1);
          tc::mbar_arrive_expect_tx(&full_bar[s], kStageBytes);
          uint8_t* dst = smem + s * kStageBytes;
          tc::tma_load_2d(dst, &tma_a, &full_bar[s], kb * kBK, (int)w * kBM);
          tc::tma_load_2d(dst + kABytes, &tma_w, &full_bar[s], kb * kBK, 0);
          if (++s == kStages) { s = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = tc::umma_idesc_bf16(kBM, kN);
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
          const uint64_t w_desc = tc::umma_desc_sw128(a_addr + kABytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            tc::umma_bf16(tmem_base + buf * kAcc, a_desc + (uint64_t)(k * 2), w_desc + (uint64_t)(k * 2), idesc,
                          (kb | k) != 0);
          tc::umma_commit(&empty_bar[s]);
          if (++s == kStages) { s = 0; phase ^= 1; }
        }
        tc::umma_commit(&tmem_full_bar[buf]);
      }
    }
  } else {
    const int e = warp - 2;
    const int lg = warp & 3;   // TMEM lane quarter this warp may read
    const int ch = e >> 2;     // which 128 of the 256 columns
    const int partner = ((ch ^ 1) << 2) | lg;   // index (0..7) of the warp with the other half of the same rows
    const int me = (ch << 2) | lg;
    const int col0 = ch * 128;
    uint32_t it = 0;
    for (int64_t w = w_begin; w < w_end; ++w, ++it) {
      const uint32_t buf = it & 1u, use = it >> 1;
      const int64_t row = w * kBM + lg * 32 + lane;
      const bool row_ok = row < args.M;
      tc::mbar_wait_long(&tmem_full_bar[buf], use & 1u);
      tc::tcgen05_fence_after_sync();
      float v[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t r[32];
        tc::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(lg * 32) << 16) + buf * kAcc + (uint32_t)(col0 + c * 32), r);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) v[c * 32 + j] = __uint_as_float(r[j]);
      }
      // the accumulator buffer may be overwritten by the MMA of tile it + 2
      tc::tcgen05_fence_before_sync();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&tmem_empty_bar[buf]);
      // + bias + residual
      if (args.bias) {
#pragma unroll
        for (int j = 0; j < 128; j += 4) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(args.bias + col0 + j));
          v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
        }
      }
      if (args.resid && row_ok) {
        const __nv_bfloat16* rp = args.resid + (row / args.resid_div) * kN + col0;
#pragma unroll
        for (int j = 0; j < 128; j += 8) {
          const uint4 q = __ldg(reinterpret_cast<const uint4*>(rp + j));
          v[j] += bf_lo(q.x); v[j + 1] += bf_hi(q.x); v[j + 2] += bf_lo(q.y); v[j + 3] += bf_hi(q.y);
          v[j + 4] += bf_lo(q.z); v[j + 5] += bf_hi(q.z); v[j + 6] += bf_lo(q.w); v[j + 7] += bf_hi(q.w);
        }
      }
      // row mean, then the centred sum of squares (two-pass, like nn.LayerNorm); the half rows meet in shared memory
      float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 128; ++j) s4[j & 3] += v[j];
      const float s_mine = (s4[0] + s4[1]) + (s4[2] + s4[3]);
      xchg[me * 32 + lane] = s_mine;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + lg) : "memory");
      const float mean = (s_mine + xchg[partner * 32 + lane]) * (1.f / kN);
      float q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 128; ++j) {
        const float d = v[j] - mean;
        q4[j & 3] = fmaf(d, d, q4[j & 3]);
      }
      const float q_mine = (q4[0] + q4[1]) + (q4[2] + q4[3]);
      xchg[kEpWarps * 32 + me * 32 + lane] = q_mine;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + lg) : "memory");
      const float rstd = rsqrtf((q_mine + xchg[kEpWarps * 32 + partner * 32 + lane]) * (1.f / kN) + args.eps);
      if (row_ok) {
        __nv_bfloat16* op = args.out + row * kN + col0;
#pragma unroll
        for (int j = 0; j < 128; j += 8) {
          const float4 g0 = __ldg(reinterpret_cast<const float4*>(args.gamma + col0 + j));
          const float4 g1 = __ldg(reinterpret_cast<const float4*>(args.gamma + col0 + j + 4));
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(args.beta + col0 + j));
          const float4 b1 = __ldg(reinterpret_cast<const float4*>(args.beta + col0 + j + 4));
          uint4 o;
          o.x = pack2((v[j] - mean) * rstd * g0.x + b0.x, (v[j + 1] - mean) * rstd * g0.y + b0.y);
          o.y = pack2((v[j + 2] - mean) * rstd * g0.z + b0.z, (v[j + 3] - mean) * rstd * g0.w + b0.w);
          o.z = pack2((v[j + 4] - mean) * rstd * g1.x + b1.x, (v[j + 5] - mean) * rstd * g1.y + b1.y);
          o.w = pack2((v[j + 6] - mean) * rstd * g1.z + b1.z, (v[j + 7] - mean) * rstd * g1.w + b1.w);
          *reinterpret_cast<uint4*>(op + j) = o;
        }
      }
      // the exchange slots are rewritten by the next tile: both warps of the pair must have read them
      asm volatile("bar.sync %0, 64;" ::"r"(1 + lg) : "memory");
    }
  }
  tc::tcgen05_fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<kTmemAlloc>(tmem_base);
}

}  // namespace gln
}  // namespace se3et

using namespace se3et;

extern "C" int se3et_linear_add_layernorm(const void* a_bf16, int64_t lda, const void* w_bf16, int64_t ldw, int64_t m,
                                          int64_t n, int64_t k, const float* bias, const void* resid_bf16,
                                          int64_t resid_div, const float* gamma, const float* beta, float eps,
                                          void* out_bf16, se3et_stream_t stream) {
  if (m < 0 || n <= 0 || k <= 0 || resid_div <= 0) return SE3ET_ERR_ARG;
  if (n != gln::kN || k % 8 != 0) return SE3ET_ERR_UNSUPPORTED;  // the host uses se3et_gemm_bf16 + se3et_add_layernorm
  if (m == 0) return SE3ET_OK;
  if (!a_bf16 || !w_bf16 || !gamma || !beta || !out_bf16) return SE3ET_ERR_ARG;
  if ((reinterpret_cast<uintptr_t>(out_bf16) | reinterpret_cast<uintptr_t>(resid_bf16) | reinterpret_cast<uintptr_t>(bias) |
       reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15)
    return SE3ET_ERR_UNSUPPORTED;
  if (ceil_div(m, gln::kBM) > INT32_MAX / gln::kBM) return SE3ET_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUtensorMap ta, tw;
  int rc = make_tmap_bf16_2d(&ta, a_bf16, m, k, lda, gln::kBM);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&tw, w_bf16, n, k, ldw, gln::kN);
  if (rc) return rc;
  gln::Args args;
  args.M = (int)m; args.K = (int)k; args.m_tiles = (int)ceil_div(m, gln::kBM);
  args.bias = bias; args.resid = static_cast<const __nv_bfloat16*>(resid_bf16); args.resid_div = (int)resid_div;
  args.gamma = gamma; args.beta = beta; args.eps = eps; args.out = static_cast<__nv_bfloat16*>(out_bf16);
  SE3ET_ENSURE_SMEM(gln::linear_add_layernorm_kernel, gln::kSmem);
  const unsigned grid = (unsigned)(args.m_tiles < kNumSMs ? args.m_tiles : kNumSMs);
  gln::linear_add_layernorm_kernel<<<grid, gln::kThreads, gln::kSmem, st>>>(ta, tw, args);
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}
