// Superpoint-transformer kernels other than attention (sm_100a):
//   geo_embed_indices  pair distances + 3-NN triplet angles              (geotransformer.py:69-99)
//   geo_embed_project  sinusoidal embedding generated on the fly as the A operand of a tcgen05 GEMM with
//                      four TMEM accumulators (d, a_0, a_1, a_2), epilogue  d + max_k a_k + bias
//                      -> never materialises the (N, N, 3, C) tensor    (geotransformer.py:101-115,
//                                                                         positional_embedding.py:18-34)
//   add_layernorm      LayerNorm(x + residual), residual optionally broadcast over anchors
//                                                      (rpe_transformer.py:161-163, vanilla_transformer.py:908-911,
//                                                       output_layer.py:16-22)
//   l2_normalize_rows  F.normalize(p=2, dim=1) of model.py:156-157
// Clouds are stored flat: cloud b owns points [cu[b], cu[b+1]) and embedding rows [eoff[b], eoff[b] + n_b^2).
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc.cuh"

namespace se3et {

__device__ __forceinline__ uint32_t pack2_bf16(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}

// ---------------------------------------------------------------------------------------------
// geo_embed_indices: one CTA per (cloud, anchor point n)
// ---------------------------------------------------------------------------------------------
constexpr int kEmbThreads = 128;

__global__ void __launch_bounds__(kEmbThreads) geo_embed_indices_kernel(const float* __restrict__ pts,
                                                                         const int64_t* __restrict__ cu, int nclouds,
                                                                         const int64_t* __restrict__ eoff,
                                                                         float inv_sigma_d, float factor_a,
                                                                         float4* __restrict__ out) {
  extern __shared__ float sh[];  // [3 * n_b] points, [n_b] squared distances
  __shared__ unsigned long long sh_red[kEmbThreads / 32];
  __shared__ unsigned long long sh_sel[4];
  const int64_t gi = blockIdx.x;
  const int b = segment_of(cu, nclouds, gi);
  const int64_t start = cu[b];
  const int nb = (int)(cu[b + 1] - start);
  const int n = (int)(gi - start);
  float* px = sh;
  float* d2 = sh + 3 * nb;
  for (int i = threadIdx.x; i < 3 * nb; i += kEmbThreads) px[i] = pts[3 * start + i];
  __syncthreads();
  const float qx = px[3 * n], qy = px[3 * n + 1], qz = px[3 * n + 2];
  for (int m = threadIdx.x; m < nb; m += kEmbThreads) {
    const float dx = px[3 * m] - qx, dy = px[3 * m + 1] - qy, dz = px[3 * m + 2] - qz;
    d2[m] = dx * dx + dy * dy + dz * dz;
  }
  __syncthreads();
  // the k+1 = 4 nearest by (distance, index); the first one (the point itself) is dropped (geotransformer.py:86)
  unsigned long long last = 0;
  for (int round = 0; round < 4; ++round) {
    unsigned long long best = ~0ull;
    for (int m = threadIdx.x; m < nb; m += kEmbThreads) {
      const unsigned long long key = ((unsigned long long)__float_as_uint(d2[m]) << 32) | (unsigned)m;
      if ((round == 0 || key > last) && key < best) best = key;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other < best ? other : best;
    }
    if ((threadIdx.x & 31) == 0) sh_red[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long v = sh_red[0];
      for (int w = 1; w < kEmbThreads / 32; ++w) v = sh_red[w] < v ? sh_red[w] : v;
      sh_sel[round] = v;
    }
    __syncthreads();
    last = sh_sel[round];
  }
  float rx[3], ry[3], rz[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const unsigned long long key = sh_sel[k + 1];
    int j = (key == ~0ull) ? n : (int)(key & 0xffffffffu);  // clouds with < 4 points: degenerate reference vector
    rx[k] = px[3 * j] - qx; ry[k] = px[3 * j + 1] - qy; rz[k] = px[3 * j + 2] - qz;
  }
  float4* row = out + eoff[b] + (int64_t)n * nb;
  for (int m = threadIdx.x; m < nb; m += kEmbThreads) {
    const float ax = px[3 * m] - qx, ay = px[3 * m + 1] - qy, az = px[3 * m + 2] - qz;
    float ang[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float cx = ry[k] * az - rz[k] * ay, cy = rz[k] * ax - rx[k] * az, cz = rx[k] * ay - ry[k] * ax;
      const float s = sqrtf(cx * cx + cy * cy + cz * cz);
      float c = rx[k] * ax + ry[k] * ay + rz[k] * az;
      // torch.sum starts from +0, so the reference never sees cos = -0 (atan2(0, -0) would be pi on the diagonal)
      c = (c == 0.f) ? 0.f : c;
      ang[k] = atan2f(s, c) * factor_a;
    }
    row[m] = make_float4(sqrtf(d2[m]) * inv_sigma_d, ang[0], ang[1], ang[2]);
  }
}

// ---------------------------------------------------------------------------------------------
// geo_embed_project: E[r, :] = W_d emb(x_d) + max_k W_a emb(x_ak) + (b_d + b_a)      r = flat (cloud, n, m) row
// 128 rows per CTA, BN output columns per CTA (grid.y), K = C in blocks of 64.
//   warp 0: TMA of the W_d / W_a tiles      warp 1: TMEM alloc + tcgen05.mma issue (4 accumulators)
//   warps 2-5: generate sin/cos (A operand, written straight into the 128B-swizzled smem tile), then epilogue
// ---------------------------------------------------------------------------------------------
constexpr int kEpStages = 2;
constexpr int kEpThreads = 320;  // TMA warp, MMA warp, 8 producer / epilogue warps (2 per TMEM lane quadrant)
__constant__ float c_div_term[512];  // exp(-2j ln(1e4) / C), j < C/2

template <int BN>
struct EmbedSmem {
  static constexpr int kATile = 128 * 128;          // one index type: 128 rows x 64 bf16
  static constexpr int kABytes = 4 * kATile;        // d, a0, a1, a2
  static constexpr int kBTile = BN * 128;
  static constexpr int kBBytes = 2 * kBTile;        // W_d, W_a
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kBarOffset = kEpStages * kStageBytes;
  static constexpr int kTotal = kBarOffset + 128 + 1024;
};

__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

template <int BN>
__global__ void __launch_bounds__(kEpThreads, 1)
geo_embed_project_kernel(const __grid_constant__ CUtensorMap tma_wd, const __grid_constant__ CUtensorMap tma_wa,
                         const float4* __restrict__ idx, int64_t rows, int C, const float* __restrict__ bias_sum,
                         __nv_bfloat16* __restrict__ out) {
  using S = EmbedSmem<BN>;
  constexpr uint32_t kTmemCols = 4 * BN;  // 512 or 256
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + S::kBarOffset);
  uint64_t* empty_bar = full_bar + kEpStages;
  uint64_t* tmem_full_bar = empty_bar + kEpStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r0 = (int64_t)blockIdx.x * 128;
  const int n0 = blockIdx.y * BN;
  const int num_kb = C / 64;

  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&tma_wd);
    tc::tma_prefetch_desc(&tma_wa);
    for (int s = 0; s < kEpStages; ++s) {
      tc::mbar_init(&full_bar[s], 1 + 256);  // TMA thread (expect_tx) + 256 producer threads
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
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kEpStages;
        const uint32_t phase = (kb / kEpStages) & 1;
        tc::mbar_wait_long(&empty_bar[s], phase ^ 1);
        tc::mbar_arrive_expect_tx(&full_bar[s], S::kBBytes);
        uint8_t* b_dst = smem + s * S::kStageBytes + S::kABytes;
        tc::tma_load_2d(b_dst, &tma_wd, &full_bar[s], kb * 64, n0);
        tc::tma_load_2d(b_dst + S::kBTile, &tma_wa, &full_bar[s], kb * 64, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = tc::umma_idesc_bf16(128, BN);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kEpStages;
        const uint32_t phase = (kb / kEpStages) & 1;
        tc::mbar_wait_long(&full_bar[s], phase);
        tc::tcgen05_fence_after_sync();
        const uint32_t a_addr = tc::smem_u32(smem + s * S::kStageBytes);
        const uint32_t b_addr = a_addr + S::kABytes;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const uint64_t a_desc = tc::umma_desc_sw128(a_addr + t * S::kATile);
          const uint64_t b_desc = tc::umma_desc_sw128(b_addr + (t == 0 ? 0 : S::kBTile));
#pragma unroll
          for (int k = 0; k < 4; ++k)
            tc::umma_bf16(tmem_base + (uint32_t)(t * BN), a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc,
                          (kb | k) != 0);
        }
        tc::umma_commit(&empty_bar[s]);
      }
      tc::umma_commit(tmem_full_bar);
    }
  } else {
    const int lane_base = (warp & 3) * 32;
    const int trow = lane_base + lane;  // tile row == TMEM lane
    const int half = (warp - 2) >> 2;   // the two warps of a lane quadrant split the sinusoid chunks / output columns
    const int64_t row = r0 + trow;
    float x[4] = {0.f, 0.f, 0.f, 0.f};
    if (row < rows) {
      const float4 v = idx[row];
      x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
    }
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % kEpStages;
      const uint32_t phase = (kb / kEpStages) & 1;
      tc::mbar_wait_long(&empty_bar[s], phase ^ 1);
      uint8_t* a_base = smem + s * S::kStageBytes;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
#pragma unroll
        for (int j = 4 * half; j < 4 * half + 4; ++j) {  // 16-byte chunk j: frequencies kb*32 + 4j .. +3, (sin, cos)
          uint32_t w[4];
#pragma unroll
          for (int f = 0; f < 4; ++f) {
            const float arg = x[t] * c_div_term[kb * 32 + j * 4 + f];
            // range reduction to [-pi, pi] then the SFU approximations (abs error ~1e-6, far below bf16)
            const float kf = rintf(arg * 0.15915494309189535f);
            float r = fmaf(-kf, 6.2831854820251465f, arg);
            r = fmaf(-kf, -1.7484555e-7f, r);
            float sv, cv;
            __sincosf(r, &sv, &cv);
            w[f] = pack2_bf16(sv, cv);
          }
          *reinterpret_cast<uint4*>(a_base + t * S::kATile + tc::sw128_offset(trow, j)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      tc::fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core's async proxy
      tc::mbar_arrive(&full_bar[s]);
    }
    // epilogue
    tc::mbar_wait_long(tmem_full_bar, 0);
    tc::tcgen05_fence_after_sync();
#pragma unroll 1
    for (int c0 = half * (BN / 2); c0 < (half + 1) * (BN / 2); c0 += 16) {
      uint32_t d[16], a0[16], a1[16], a2[16];
      const uint32_t t0 = tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)c0;
      tmem_ld_32x32b_x16(t0, d);
      tmem_ld_32x32b_x16(t0 + BN, a0);
      tmem_ld_32x32b_x16(t0 + 2 * BN, a1);
      tmem_ld_32x32b_x16(t0 + 3 * BN, a2);
      tc::tmem_ld_wait();
      if (row < rows) {
        uint32_t packed[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float v[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int col = 2 * j + e;
            const float am = fmaxf(fmaxf(__uint_as_float(a0[col]), __uint_as_float(a1[col])), __uint_as_float(a2[col]));
            v[e] = __uint_as_float(d[col]) + am + __ldg(bias_sum + n0 + c0 + col);
          }
          packed[j] = pack2_bf16(v[0], v[1]);
        }
        uint4* dst = reinterpret_cast<uint4*>(out + row * C + n0 + c0);
        dst[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        dst[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
      }
    }
  }
  tc::tcgen05_fence_before_sync();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc<kTmemCols>(tmem_base);
}

int make_tmap_bf16_2d(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows);

template <int BN>
static int launch_embed(const CUtensorMap& td, const CUtensorMap& ta, const float4* idx, int64_t rows, int C,
                        const float* bias_sum, __nv_bfloat16* out, cudaStream_t st) {
  using S = EmbedSmem<BN>;
  SE3ET_ENSURE_SMEM(geo_embed_project_kernel<BN>, S::kTotal);
  dim3 grid((unsigned)ceil_div(rows, 128), (unsigned)(C / BN));
  geo_embed_project_kernel<BN><<<grid, kEpThreads, S::kTotal, st>>>(td, ta, idx, rows, C, bias_sum, out);
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}

// ---------------------------------------------------------------------------------------------
// geo_embed_lookup: the projected sinusoidal embedding of a scalar u, W emb(u) + b, is a smooth function R -> R^C that
// depends on the weights only, so it is tabulated once per weight version (step 1/512: the embedding's highest
// frequency is 1 rad per unit, the tabulation error stays below the bf16 rounding of the result) and the
// (sum n^2, C) embedding is assembled from four 16-byte-per-lane table reads per row:
//   out[row] = table_d[round(512 d)] + max_k table_a[round(512 a_k)]
// One row per C / 8 lanes; tables are L2 resident.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void bf16x8_to_f32(const uint4& v, float (&f)[8]) {
  f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
  f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
  f[4] = __uint_as_float(v.z << 16); f[5] = __uint_as_float(v.z & 0xffff0000u);
  f[6] = __uint_as_float(v.w << 16); f[7] = __uint_as_float(v.w & 0xffff0000u);
}

__global__ void __launch_bounds__(256) geo_embed_lookup_kernel(const float4* __restrict__ idx, int64_t rows, int C,
                                                                const __nv_bfloat16* __restrict__ table_d, int nd,
                                                                const __nv_bfloat16* __restrict__ table_a, int na,
                                                                float inv_step, __nv_bfloat16* __restrict__ out) {
  const int lpr = C >> 3;  // lanes per row (16 bytes each); C is a power-of-two multiple of 64 up to 256 -> 8..32
  const int lane = threadIdx.x & 31;
  const int sub = lane / lpr, c8 = (lane - sub * lpr) * 8, rpw = 32 / lpr;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = warp0 * rpw + sub; row < rows; row += nwarps * rpw) {
    const float4 v = __ldg(idx + row);
    const int id = min(max(__float2int_rn(v.x * inv_step), 0), nd - 1);
    const int i0 = min(max(__float2int_rn(v.y * inv_step), 0), na - 1);
    const int i1 = min(max(__float2int_rn(v.z * inv_step), 0), na - 1);
    const int i2 = min(max(__float2int_rn(v.w * inv_step), 0), na - 1);
    const uint4 qd = __ldg(reinterpret_cast<const uint4*>(table_d + (int64_t)id * C + c8));
    const uint4 q0 = __ldg(reinterpret_cast<const uint4*>(table_a + (int64_t)i0 * C + c8));
    const uint4 q1 = __ldg(reinterpret_cast<const uint4*>(table_a + (int64_t)i1 * C + c8));
    const uint4 q2 = __ldg(reinterpret_cast<const uint4*>(table_a + (int64_t)i2 * C + c8));
    // packed bf16 arithmetic: the maximum of bf16 values is exact, and the bf16 sum of two bf16 values (one rounding of
    // the exact sum) equals the fp32 sum rounded to bf16 -- same bits as the unpacked form, a third of the instructions
    const uint32_t dd[4] = {qd.x, qd.y, qd.z, qd.w}, x0[4] = {q0.x, q0.y, q0.z, q0.w};
    const uint32_t x1[4] = {q1.x, q1.y, q1.z, q1.w}, x2[4] = {q2.x, q2.y, q2.z, q2.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_bfloat162 m = __hmax2(__hmax2(*reinterpret_cast<const __nv_bfloat162*>(&x0[j]),
                                               *reinterpret_cast<const __nv_bfloat162*>(&x1[j])),
                                       *reinterpret_cast<const __nv_bfloat162*>(&x2[j]));
      const __nv_bfloat162 r = __hadd2(*reinterpret_cast<const __nv_bfloat162*>(&dd[j]), m);
      o[j] = *reinterpret_cast<const uint32_t*>(&r);
    }
    __stcs(reinterpret_cast<uint4*>(out + row * C + c8), make_uint4(o[0], o[1], o[2], o[3]));
  }
}

// ---------------------------------------------------------------------------------------------
// add_layernorm: one warp per row
// ---------------------------------------------------------------------------------------------
template <int kPerLane>
__global__ void __launch_bounds__(256) add_layernorm_kernel(const float* __restrict__ x,
                                                             const __nv_bfloat16* __restrict__ resid, int resid_div,
                                                             int64_t rows, int C, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float eps,
                                                             float* __restrict__ out_f32,
                                                             __nv_bfloat16* __restrict__ out_bf16) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float v[kPerLane];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < kPerLane; ++j) {
    const int c = lane + 32 * j;
    float t = 0.f;
    if (c < C) {
      t = x[row * C + c];
      if (resid) t += __bfloat162float(resid[(row / resid_div) * C + c]);
    }
    v[j] = t;
    s += t;
  }
  const float mean = warp_sum(s) / C;
  float ss = 0.f;
#pragma unroll
  for (int j = 0; j < kPerLane; ++j) {
    const int c = lane + 32 * j;
    const float dlt = c < C ? v[j] - mean : 0.f;
    ss += dlt * dlt;
  }
  const float rstd = rsqrtf(warp_sum(ss) / C + eps);
#pragma unroll
  for (int j = 0; j < kPerLane; ++j) {
    const int c = lane + 32 * j;
    if (c < C) {
      const float o = (v[j] - mean) * rstd * gamma[c] + beta[c];
      if (out_f32) out_f32[row * C + c] = o;
      if (out_bf16) out_bf16[row * C + c] = __float2bfloat16(o);
    }
  }
}

// C a multiple of 256 (the transformer's hidden width): a lane owns 8 consecutive columns per 256-column chunk, so x is read
// as two 16-byte loads, the residual as one and the bf16 result written as one (the scalar form above moves 4 / 2 bytes
// per instruction: 2.8 TB/s on 130k x 256 rows)
template <int kChunks>
__global__ void __launch_bounds__(256) add_layernorm_vec_kernel(const float* __restrict__ x,
                                                                 const __nv_bfloat16* __restrict__ resid, int resid_div,
                                                                 int64_t rows, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, float eps,
                                                                 float* __restrict__ out_f32,
                                                                 __nv_bfloat16* __restrict__ out_bf16) {
  constexpr int C = 256 * kChunks;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float v[kChunks][8];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < kChunks; ++k) {
    const int c = 256 * k + 8 * lane;
    const float4 a = __ldcs(reinterpret_cast<const float4*>(x + row * C + c));
    const float4 b = __ldcs(reinterpret_cast<const float4*>(x + row * C + c + 4));
    v[k][0] = a.x; v[k][1] = a.y; v[k][2] = a.z; v[k][3] = a.w;
    v[k][4] = b.x; v[k][5] = b.y; v[k][6] = b.z; v[k][7] = b.w;
    if (resid) {
      float r[8];
      bf16x8_to_f32(__ldg(reinterpret_cast<const uint4*>(resid + (row / resid_div) * C + c)), r);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[k][j] += r[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[k][j];
  }
  const float mean = warp_sum(s) / C;
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < kChunks; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float dlt = v[k][j] - mean;
      ss += dlt * dlt;
    }
  const float rstd = rsqrtf(warp_sum(ss) / C + eps);
#pragma unroll
  for (int k = 0; k < kChunks; ++k) {
    const int c = 256 * k + 8 * lane;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
    const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = (v[k][j] - mean) * rstd * g[j] + bb[j];
    if (out_f32) {
      *reinterpret_cast<float4*>(out_f32 + row * C + c) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4*>(out_f32 + row * C + c + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
    if (out_bf16)
      *reinterpret_cast<uint4*>(out_bf16 + row * C + c) =
          make_uint4(pack2_bf16(o[0], o[1]), pack2_bf16(o[2], o[3]), pack2_bf16(o[4], o[5]), pack2_bf16(o[6], o[7]));
  }
}

__global__ void __launch_bounds__(256) l2_normalize_rows_kernel(const float* __restrict__ x, int64_t rows, int C,
                                                                 float eps, float* __restrict__ out) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float ss = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float v = x[row * C + c];
    ss += v * v;
  }
  const float inv = 1.f / fmaxf(sqrtf(warp_sum(ss)), eps);  // F.normalize: x / max(||x||, eps)
  for (int c = lane; c < C; c += 32) out[row * C + c] = x[row * C + c] * inv;
}

}  // namespace se3et

using namespace se3et;

extern "C" int se3et_geo_embed_indices(const float* points, const int64_t* cloud_offsets, int64_t nclouds,
                                       int64_t total_points, int64_t max_cloud, const int64_t* emb_offsets,
                                       float sigma_d, float sigma_a, int64_t angle_k, float* out_idx4,
                                       se3et_stream_t stream) {
  if (nclouds <= 0 || total_points < 0 || max_cloud < 0 || !(sigma_d > 0.f) || !(sigma_a > 0.f)) return SE3ET_ERR_ARG;
  if (angle_k != 3) return SE3ET_ERR_UNSUPPORTED;
  if (total_points == 0) return SE3ET_OK;
  if (!points || !cloud_offsets || !emb_offsets || !out_idx4) return SE3ET_ERR_ARG;
  const size_t smem = sizeof(float) * 4 * (size_t)max_cloud;
  if (smem > 200 * 1024) return SE3ET_ERR_UNSUPPORTED;
  if (smem > 48 * 1024) SE3ET_ENSURE_SMEM(geo_embed_indices_kernel, smem);
  geo_embed_indices_kernel<<<(unsigned)total_points, kEmbThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      points, cloud_offsets, (int)nclouds, emb_offsets, 1.f / sigma_d, 180.f / (sigma_a * 3.14159265358979323846f),
      reinterpret_cast<float4*>(out_idx4));
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}

extern "C" int se3et_geo_embed_project(const float* idx4, int64_t rows, int64_t channels, const void* w_d_bf16,
                                       const void* w_a_bf16, const float* bias_sum, void* out_bf16,
                                       se3et_stream_t stream) {
  if (rows < 0 || channels < 64 || channels > 1024 || channels % 64) return SE3ET_ERR_ARG;
  if (rows == 0) return SE3ET_OK;
  if (!idx4 || !w_d_bf16 || !w_a_bf16 || !bias_sum || !out_bf16) return SE3ET_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int C = (int)channels;
  static int div_for = 0;
  if (div_for != C) {
    float h[512];
    for (int j = 0; j < C / 2; ++j) h[j] = expf((float)(2 * j) * (-logf(10000.0f) / (float)C));
    SE3ET_CUDA_CHECK(cudaMemcpyToSymbolAsync(c_div_term, h, sizeof(float) * (C / 2), 0, cudaMemcpyHostToDevice, st));
    SE3ET_CUDA_CHECK(cudaStreamSynchronize(st));  // `h` is a stack buffer; happens once per channel count
    div_for = C;
  }
  const int bn = C % 128 == 0 ? 128 : 64;
  CUtensorMap td, ta;
  int rc = make_tmap_bf16_2d(&td, w_d_bf16, C, C, C, bn);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(&ta, w_a_bf16, C, C, C, bn);
  if (rc) return rc;
  const float4* idx = reinterpret_cast<const float4*>(idx4);
  auto* out = static_cast<__nv_bfloat16*>(out_bf16);
  return bn == 128 ? launch_embed<128>(td, ta, idx, rows, C, bias_sum, out, st)
                   : launch_embed<64>(td, ta, idx, rows, C, bias_sum, out, st);
}

extern "C" int se3et_geo_embed_lookup(const float* idx4, int64_t rows, int64_t channels, const void* table_d_bf16,
                                      int64_t nd, const void* table_a_bf16, int64_t na, float step, void* out_bf16,
                                      se3et_stream_t stream) {
  if (rows < 0 || nd <= 0 || na <= 0 || !(step > 0.f)) return SE3ET_ERR_ARG;
  if (channels != 64 && channels != 128 && channels != 256) return SE3ET_ERR_UNSUPPORTED;
  if (rows == 0) return SE3ET_OK;
  if (!idx4 || !table_d_bf16 || !table_a_bf16 || !out_bf16) return SE3ET_ERR_ARG;
  int64_t blocks = ceil_div(rows * (channels / 8), 256);
  if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;
  geo_embed_lookup_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(idx4), rows, (int)channels, static_cast<const __nv_bfloat16*>(table_d_bf16),
      (int)nd, static_cast<const __nv_bfloat16*>(table_a_bf16), (int)na, 1.f / step,
      static_cast<__nv_bfloat16*>(out_bf16));
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}

extern "C" int se3et_add_layernorm(const float* x, const void* resid_bf16, int64_t resid_div, int64_t rows,
                                   int64_t channels, const float* gamma, const float* beta, float eps, float* out_f32,
                                   void* out_bf16, se3et_stream_t stream) {
  if (rows < 0 || channels <= 0 || channels > 1024 || resid_div <= 0) return SE3ET_ERR_ARG;
  if (rows == 0) return SE3ET_OK;
  if (!x || !gamma || !beta || (!out_f32 && !out_bf16)) return SE3ET_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const auto* r = static_cast<const __nv_bfloat16*>(resid_bf16);
  auto* ob = static_cast<__nv_bfloat16*>(out_bf16);
  const unsigned blocks = (unsigned)ceil_div(rows, 8);
  const int C = (int)channels;
  const bool aligned = !((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(r) | reinterpret_cast<uintptr_t>(gamma) |
                          reinterpret_cast<uintptr_t>(beta) | reinterpret_cast<uintptr_t>(out_f32) |
                          reinterpret_cast<uintptr_t>(ob)) & 15);
  if (aligned && (C == 256 || C == 512)) {
    if (C == 256) add_layernorm_vec_kernel<1><<<blocks, 256, 0, st>>>(x, r, (int)resid_div, rows, gamma, beta, eps, out_f32, ob);
    else add_layernorm_vec_kernel<2><<<blocks, 256, 0, st>>>(x, r, (int)resid_div, rows, gamma, beta, eps, out_f32, ob);
    SE3ET_LAUNCH_CHECK();
    return SE3ET_OK;
  }
  if (C <= 64) add_layernorm_kernel<2><<<blocks, 256, 0, st>>>(x, r, (int)resid_div, rows, C, gamma, beta, eps, out_f32, ob);
  else if (C <= 128) add_layernorm_kernel<4><<<blocks, 256, 0, st>>>(x, r, (int)resid_div, rows, C, gamma, beta, eps, out_f32, ob);
  else if (C <= 256) add_layernorm_kernel<8><<<blocks, 256, 0, st>>>(x, r, (int)resid_div, rows, C, gamma, beta, eps, out_f32, ob);
  else if (C <= 512) add_layernorm_kernel<16><<<blocks, 256, 0, st>>>(x, r, (int)resid_div, rows, C, gamma, beta, eps, out_f32, ob);
  else add_layernorm_kernel<32><<<blocks, 256, 0, st>>>(x, r, (int)resid_div, rows, C, gamma, beta, eps, out_f32, ob);
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}

extern "C" int se3et_l2_normalize_rows(const float* x, int64_t rows, int64_t channels, float eps, float* out,
                                       se3et_stream_t stream) {
  if (rows < 0 || channels <= 0) return SE3ET_ERR_ARG;
  if (rows == 0) return SE3ET_OK;
  if (!x || !out) return SE3ET_ERR_ARG;
  l2_normalize_rows_kernel<<<(unsigned)ceil_div(rows, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      x, rows, (int)channels, eps, out);
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}
