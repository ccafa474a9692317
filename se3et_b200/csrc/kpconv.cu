// KPConvInterSO3 operand builder on tensor cores (blocks_epn.py:334-390, 454-506): see kpconv_tables.cuh for the
// algebra and kpconv_mma.cuh for the warp-level producer.
//
//   kpconv_gather_mma_kernel   stand-alone form: one warp per query point, A' rows written to global memory for
//                              se3et_gemm_bf16 (used when Cin % 16 == 0; the scalar kernel in e2pn.cu covers the rest)
#include <cuda_bf16.h>

#include "common.cuh"
#include "kpconv_mma.cuh"

namespace se3et {

using namespace kpm;

constexpr int kGatherWarps = 8;

__constant__ int8_t c_basis_target[16][6] = {
#define SE3ET_BT(row) {(int8_t)basis_target(row, 0), (int8_t)basis_target(row, 1), (int8_t)basis_target(row, 2), \
                       (int8_t)basis_target(row, 3), (int8_t)basis_target(row, 4), (int8_t)basis_target(row, 5)}
    SE3ET_BT(0), SE3ET_BT(1), SE3ET_BT(2), SE3ET_BT(3), SE3ET_BT(4), SE3ET_BT(5), SE3ET_BT(6), SE3ET_BT(7),
    SE3ET_BT(8), SE3ET_BT(9), SE3ET_BT(10), SE3ET_BT(11), SE3ET_BT(12), SE3ET_BT(13), SE3ET_BT(14), SE3ET_BT(15)
#undef SE3ET_BT
};
__constant__ uint32_t c_ridx_cols[6] = {ridx_col_packed(0), ridx_col_packed(1), ridx_col_packed(2),
                                        ridx_col_packed(3), ridx_col_packed(4), ridx_col_packed(5)};

template <int KS>
struct GatherSmem {
  static constexpr int kNP = KS * 16;                 // neighbour rows (padded to the MMA k-step)
  static constexpr int kWRowBytes = (kNP + 8) * 2;    // W16 row pitch: 112 B / 144 B -> conflict-free ldmatrix
  static constexpr int kXStage = kNP * kXRowBytes;
  static constexpr int kPerWarp = 2 * kXStage + 16 * kWRowBytes + 64 * 4;  // 2 gather stages, W16, neighbour ids
  static constexpr int kShared = 48 * 4 + 16 * 6 + 6 * 4 + 8;              // kernel points, target table, ridx
  static constexpr int kTotal = kGatherWarps * kPerWarp + 512;
};

template <int KS>
__global__ void __launch_bounds__(kGatherWarps * 32, 1)
kpconv_gather_mma_kernel(const float* __restrict__ q_pts, const float* __restrict__ s_pts,
                         const int64_t* __restrict__ idx, int H, int64_t nq, int64_t ns,
                         const __nv_bfloat16* __restrict__ x, int cin, const float* __restrict__ kernel_points,
                         float inv_extent, __nv_bfloat16* __restrict__ out, int kpad) {
  using S = GatherSmem<KS>;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  __shared__ float sh_kp[48];
  __shared__ int8_t sh_target[16][6];
  __shared__ uint32_t sh_ridx[6];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < 45) sh_kp[threadIdx.x] = kernel_points[threadIdx.x];
  if (threadIdx.x < 96) sh_target[threadIdx.x / 6][threadIdx.x % 6] = c_basis_target[threadIdx.x / 6][threadIdx.x % 6];
  if (threadIdx.x < 6) sh_ridx[threadIdx.x] = c_ridx_cols[threadIdx.x];
  uint8_t* base = smem_raw + warp * S::kPerWarp;
  uint8_t* xs = base;                                   // [2][kNP][208]
  uint8_t* w16 = base + 2 * S::kXStage;                 // [16][kWRowBytes]
  int* nbr = reinterpret_cast<int*>(w16 + 16 * S::kWRowBytes);  // [64]
  // rows past H are never written by the gather: zero both stages once (their weights are zero as well)
  for (int i = lane; i < 2 * S::kXStage / 16; i += 32) reinterpret_cast<uint4*>(xs)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  const LaneTargets T = make_lane_targets(lane, sh_target, sh_ridx);
  const int g = lane >> 2, q = lane & 3;
  const int nchunks = cin / kChunk;
  const int row_elems = kA * cin;

  for (int64_t p = (int64_t)blockIdx.x * kGatherWarps + warp; p < nq; p += (int64_t)gridDim.x * kGatherWarps) {
    // ---- neighbours of p and their 16 basis weights (bf16, the A operand) -------------------------------------
    const float qx = q_pts[3 * p], qy = q_pts[3 * p + 1], qz = q_pts[3 * p + 2];
#pragma unroll
    for (int it = 0; it < S::kNP / 32 + (S::kNP % 32 ? 1 : 0); ++it) {
      const int n = it * 32 + lane;
      if (n < S::kNP) {
        int64_t j = n < H ? idx[p * H + n] : -1;
        const bool valid = j >= 0 && j < ns;
        if (!valid) j = 0;
        float row[16];
        basis_weights(s_pts[3 * j] - qx, s_pts[3 * j + 1] - qy, s_pts[3 * j + 2] - qz, sh_kp, inv_extent,
                      valid && ns > 0, row);
        nbr[n] = valid ? (int)j : -1;
#pragma unroll
        for (int r = 0; r < 16; ++r)
          *reinterpret_cast<__nv_bfloat16*>(w16 + r * S::kWRowBytes + n * 2) = __float2bfloat16(row[r]);
      }
    }
    __syncwarp();
    uint32_t afrag[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
      ldmatrix_x4(afrag[ks], smem_addr(w16 + ((lane & 7) + ((lane >> 3) & 1) * 8) * S::kWRowBytes +
                                       (ks * 16 + (lane >> 4) * 8) * 2));

    auto issue_gather = [&](int chunk, int stage) {
      uint8_t* dst = xs + stage * S::kXStage;
      for (int i = lane; i < H * kPieces; i += 32) {
        const int n = i / kPieces, part = i - n * kPieces;
        const int j = nbr[n];
        const __nv_bfloat16* src = x + (int64_t)(j < 0 ? 0 : j) * row_elems + (part >> 1) * cin + chunk * kChunk +
                                   (part & 1) * 8;
        cp_async_16(smem_addr(dst + n * kXRowBytes + part * 16), src, j < 0 ? 0 : 16);
      }
      cp_async_commit();
    };

    issue_gather(0, 0);
    __nv_bfloat16* out_p = out + p * kA * (int64_t)kpad;
    for (int chunk = 0; chunk < nchunks; ++chunk) {
      const int stage = chunk & 1;
      if (chunk + 1 < nchunks) {
        issue_gather(chunk + 1, stage ^ 1);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncwarp();
      const uint8_t* xsb = xs + stage * S::kXStage;
#pragma unroll
      for (int a = 0; a < kA; ++a) {
        float d[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          uint32_t b[4];
          ldmatrix_x4_trans(b, smem_addr(xsb + (ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * kXRowBytes + a * 32 +
                                         (lane >> 4) * 16));
          mma_16816(d[0], afrag[ks], b[0], b[1]);
          mma_16816(d[1], afrag[ks], b[2], b[3]);
        }
        // every accumulator element is one A' entry: row (p, r), column (kc, ridx[a][r], chunk * 16 + c)
        const int col0 = chunk * kChunk + 2 * q;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const uint32_t ap = (T.ridx[h][t] >> (3 * a)) & 7u;
            __nv_bfloat16* dst = out_p + T.r[h][t] * (int64_t)kpad + (T.kc[h][t] * kA + ap) * cin + col0;
            *reinterpret_cast<uint32_t*>(dst) = pack2(d[0][2 * h], d[0][2 * h + 1]);
            *reinterpret_cast<uint32_t*>(dst + 8) = pack2(d[1][2 * h], d[1][2 * h + 1]);
          }
        }
        if (T.centre) {
#pragma unroll
          for (int r = 2; r < kA; ++r) {
            __nv_bfloat16* dst = out_p + r * (int64_t)kpad + (5 * kA + ridx_tab(a, r)) * cin + col0;
            *reinterpret_cast<uint32_t*>(dst) = pack2(d[0][2], d[0][3]);
            *reinterpret_cast<uint32_t*>(dst + 8) = pack2(d[1][2], d[1][3]);
          }
        }
      }
      __syncwarp();  // all lanes done with this stage before the next gather overwrites it
    }
    (void)g;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// First backbone layer (SimpleBlockEPN on the lifted input, Cin = 1): the whole KPConvInterSO3.forward on CUDA cores.
// K = 36 only, so no tensor-core tile is worth building: one warp per query point,
//   D[row16][a] = sum_n W16[row16][n] x[idx[n]][a]                                  (96 values per point)
//   out[r][d]   = sum_{kc, a} D[basis_row(r, kc)][a] * W[kc][ridx[a][r]][0][d]       (lanes own output channels)
// plus the per-pair GroupNorm sums of the output.  Every CTA owns a contiguous range of points.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kC1Warps = 8;
constexpr int kC1MaxH = 40;  // neighbour columns (static shared memory budget)
__constant__ int8_t c_basis_row[6][6] = {
#define SE3ET_BR(r) {(int8_t)basis_row(r, 0), (int8_t)basis_row(r, 1), (int8_t)basis_row(r, 2), (int8_t)basis_row(r, 3), \
                     (int8_t)basis_row(r, 4), (int8_t)basis_row(r, 5)}
    SE3ET_BR(0), SE3ET_BR(1), SE3ET_BR(2), SE3ET_BR(3), SE3ET_BR(4), SE3ET_BR(5)
#undef SE3ET_BR
};
__constant__ int8_t c_ridx_tab[6][6] = {
#define SE3ET_RI(a) {(int8_t)ridx_tab(a, 0), (int8_t)ridx_tab(a, 1), (int8_t)ridx_tab(a, 2), (int8_t)ridx_tab(a, 3), \
                     (int8_t)ridx_tab(a, 4), (int8_t)ridx_tab(a, 5)}
    SE3ET_RI(0), SE3ET_RI(1), SE3ET_RI(2), SE3ET_RI(3), SE3ET_RI(4), SE3ET_RI(5)
#undef SE3ET_RI
};

// kLifted: the input is the same for all six anchors (LiftBlockEPN output, x[ns] one value per support point).  Then
// D does not depend on the anchor and out[r][d] = sum_kc D[basis_row(r, kc)] * (sum_a W[kc][a][0][d]): 16 products and
// 6 weight rows per point instead of 96 and 36.
template <int CPL, bool kLifted>  // output channels per lane: cout = 32 * CPL
__global__ void __launch_bounds__(kC1Warps * 32)
kpconv_cin1_kernel(const float* __restrict__ q_pts, const float* __restrict__ s_pts, const int64_t* __restrict__ idx,
                   int H, int64_t nq, int64_t ns, const __nv_bfloat16* __restrict__ x,
                   const float* __restrict__ w /* [36][cout] */, const float* __restrict__ kernel_points,
                   float inv_extent, float* __restrict__ out, double* __restrict__ stats,
                   const int64_t* __restrict__ seg_off, int nseg, int cpg) {
  constexpr int COUT = 32 * CPL;
  __shared__ float sh_w[36 * COUT];
  __shared__ float sh_kp[48];
  __shared__ float sh_w16[kC1Warps][kC1MaxH][17];  // [warp][neighbour][basis row], padded
  __shared__ __align__(16) float sh_x[kC1Warps][kC1MaxH][8];  // anchors 0-2 | pad | anchors 3-5 | pad
  __shared__ float sh_d[kC1Warps][96];
  __shared__ float sh_stat[2 * COUT];          // per output channel: sum, sum of squares of the current pair
  __shared__ uint8_t sh_dsel[kA][36];          // [r][kc * 6 + a] -> index into D: basis_row(r, kc) * 6 + a
  __shared__ uint8_t sh_wsel[kA][36];          // [r][kc * 6 + a] -> weight row kc * 6 + ridx[a][r]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < kA * 36) {
    const int r = threadIdx.x / 36, t = threadIdx.x % 36, kc = t / kA, a = t % kA;
    sh_dsel[r][t] = (uint8_t)(c_basis_row[r][kc] * kA + a);
    sh_wsel[r][t] = (uint8_t)(kc * kA + c_ridx_tab[a][r]);
  }
  if (kLifted) {  // rows 0-5 of sh_w: weights summed over the anchor slot
    for (int i = threadIdx.x; i < kA * COUT; i += blockDim.x) {
      const int kc = i / COUT, d = i - kc * COUT;
      float t = 0.f;
      for (int a = 0; a < kA; ++a) t += w[(kc * kA + a) * COUT + d];
      sh_w[i] = t;
    }
  } else {
    for (int i = threadIdx.x; i < 36 * COUT; i += blockDim.x) sh_w[i] = w[i];
  }
  if (threadIdx.x < 45) sh_kp[threadIdx.x] = kernel_points[threadIdx.x];
  __syncthreads();
  const int64_t per_cta = (nq + gridDim.x - 1) / gridDim.x;
  const int64_t p0 = (int64_t)blockIdx.x * per_cta, p1 = min(nq, p0 + per_cta);
  if (p0 >= p1) return;
  const int G = COUT / cpg;
  int seg = stats ? segment_of(seg_off, nseg, p0) : 0;
  int64_t pa = p0;
  while (pa < p1) {
    int64_t pb = p1;
    if (stats) {
      while (seg + 1 < nseg && seg_off[seg + 1] <= pa) ++seg;
      pb = min(p1, seg == nseg - 1 ? nq : seg_off[seg + 1]);
      for (int i = threadIdx.x; i < 2 * COUT; i += blockDim.x) sh_stat[i] = 0.f;
      __syncthreads();
    }
    float st_s[CPL], st_ss[CPL];
#pragma unroll
    for (int c = 0; c < CPL; ++c) st_s[c] = st_ss[c] = 0.f;
    for (int64_t p = pa + warp; p < pb; p += kC1Warps) {
      const float qx = q_pts[3 * p], qy = q_pts[3 * p + 1], qz = q_pts[3 * p + 2];
      for (int n = lane; n < H; n += 32) {
        int64_t j = idx[p * H + n];
        const bool valid = j >= 0 && j < ns;
        if (!valid) j = 0;
        float row[16];
        basis_weights(s_pts[3 * j] - qx, s_pts[3 * j + 1] - qy, s_pts[3 * j + 2] - qz, sh_kp, inv_extent, valid, row);
        if (kLifted) {
          const float f = valid ? __bfloat162float(x[j]) : 0.f;
#pragma unroll
          for (int r = 0; r < 16; ++r) sh_w16[warp][n][r] = row[r] * f;
          continue;
        }
#pragma unroll
        for (int r = 0; r < 16; ++r) sh_w16[warp][n][r] = row[r];
#pragma unroll
        for (int a = 0; a < kA; ++a) sh_x[warp][n][a + (a >= 3)] = valid ? __bfloat162float(x[j * kA + a]) : 0.f;
      }
      __syncwarp();
      if (kLifted) {
        // 16 basis products: lane = (basis row, neighbour parity)
        const int r16 = lane & 15, par = lane >> 4;
        float a0 = 0.f;
        for (int n = par; n < H; n += 2) a0 += sh_w16[warp][n][r16];
        a0 += __shfl_xor_sync(0xffffffffu, a0, 16);
        if (par == 0) sh_d[warp][r16] = a0;
        __syncwarp();
#pragma unroll 1
        for (int r = 0; r < kA; ++r) {
          float o[CPL];
#pragma unroll
          for (int c = 0; c < CPL; ++c) o[c] = 0.f;
#pragma unroll
          for (int kc = 0; kc < kA; ++kc) {
            const float dv = sh_d[warp][c_basis_row[r][kc]];
#pragma unroll
            for (int c = 0; c < CPL; ++c) o[c] = fmaf(dv, sh_w[kc * COUT + lane + 32 * c], o[c]);
          }
          float* dst = out + (p * kA + r) * COUT + lane;
#pragma unroll
          for (int c = 0; c < CPL; ++c) {
            dst[32 * c] = o[c];
            st_s[c] += o[c];
            st_ss[c] = fmaf(o[c], o[c], st_ss[c]);
          }
        }
        __syncwarp();
        continue;
      }
      {  // 96 (basis row, anchor) products: lane = (basis row, anchor half), three anchors each
        const int r16 = lane & 15, ah = lane >> 4;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
        for (int n = 0; n < H; ++n) {
          const float wv = sh_w16[warp][n][r16];
          const float4 xv = *reinterpret_cast<const float4*>(&sh_x[warp][n][4 * ah]);
          a0 = fmaf(wv, xv.x, a0);
          a1 = fmaf(wv, xv.y, a1);
          a2 = fmaf(wv, xv.z, a2);
        }
        sh_d[warp][r16 * kA + 3 * ah] = a0;
        sh_d[warp][r16 * kA + 3 * ah + 1] = a1;
        sh_d[warp][r16 * kA + 3 * ah + 2] = a2;
      }
      __syncwarp();
#pragma unroll 1
      for (int r = 0; r < kA; ++r) {
        float o[CPL];
#pragma unroll
        for (int c = 0; c < CPL; ++c) o[c] = 0.f;
#pragma unroll 4
        for (int t = 0; t < 36; ++t) {  // t = kc * 6 + a
          const float dv = sh_d[warp][sh_dsel[r][t]];
          const float* wr = sh_w + sh_wsel[r][t] * COUT + lane;
#pragma unroll
          for (int c = 0; c < CPL; ++c) o[c] = fmaf(dv, wr[32 * c], o[c]);
        }
        float* dst = out + (p * kA + r) * COUT + lane;
#pragma unroll
        for (int c = 0; c < CPL; ++c) {
          dst[32 * c] = o[c];
          st_s[c] += o[c];
          st_ss[c] = fmaf(o[c], o[c], st_ss[c]);
        }
      }
      __syncwarp();
    }
    if (stats) {
#pragma unroll
      for (int c = 0; c < CPL; ++c) {
        atomicAdd(&sh_stat[2 * (lane + 32 * c)], st_s[c]);
        atomicAdd(&sh_stat[2 * (lane + 32 * c) + 1], st_ss[c]);
      }
      __syncthreads();
      for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) {
        const int g = i >> 1, which = i & 1;
        float t = 0.f;
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) t += sh_stat[2 * c + which];
        atomicAdd(stats + ((int64_t)seg * G + g) * 2 + which, (double)t);
      }
      __syncthreads();
    }
    pa = pb;
  }
}

template <int KS>
static int launch_gather_mma(const float* q_pts, const float* s_pts, const int64_t* neighbors, int64_t nq, int64_t ns,
                             int h, const __nv_bfloat16* x, int cin, const float* kp, float inv_extent,
                             __nv_bfloat16* out, int kpad, cudaStream_t st) {
  using S = GatherSmem<KS>;
  SE3ET_ENSURE_SMEM(kpconv_gather_mma_kernel<KS>, S::kTotal);
  int64_t blocks = ceil_div(nq, kGatherWarps);
  if (blocks > (int64_t)kNumSMs * 4) blocks = (int64_t)kNumSMs * 4;
  kpconv_gather_mma_kernel<KS><<<(unsigned)blocks, kGatherWarps * 32, S::kTotal, st>>>(
      q_pts, s_pts, neighbors, h, nq, ns, x, cin, kp, inv_extent, out, kpad);
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}

// called by se3et_kpconv_gather (e2pn.cu) when cin % 16 == 0 and 36 * cin == kpad
int kpconv_gather_mma(const float* q_pts, const float* s_pts, const int64_t* neighbors, int64_t nq, int64_t ns, int h,
                      const __nv_bfloat16* x, int cin, const float* kp, float inv_extent, __nv_bfloat16* out, int kpad,
                      cudaStream_t st) {
  if (h <= 48)
    return launch_gather_mma<3>(q_pts, s_pts, neighbors, nq, ns, h, x, cin, kp, inv_extent, out, kpad, st);
  return launch_gather_mma<4>(q_pts, s_pts, neighbors, nq, ns, h, x, cin, kp, inv_extent, out, kpad, st);
}

}  // namespace se3et

extern "C" int se3et_kpconv_cin1(const float* q_pts, const float* s_pts, const int64_t* neighbors, int64_t nq, int64_t ns,
                                 int64_t h, const void* x_bf16, const float* w_36xcout, int64_t cout,
                                 const float* kernel_points_15x3, float kp_extent, float* out_f32, double* stats,
                                 const int64_t* seg_offsets, int64_t nseg, int64_t groups, int lifted,
                                 se3et_stream_t stream) {
  using namespace se3et;
  if (nq < 0 || ns <= 0 || h <= 0 || cout <= 0 || !(kp_extent > 0.f)) return SE3ET_ERR_ARG;
  if (cout % 32 != 0 || cout > 64 || h > kC1MaxH) return SE3ET_ERR_UNSUPPORTED;  // static shared memory budget
  if (!q_pts || !s_pts || !neighbors || !x_bf16 || !w_36xcout || !kernel_points_15x3 || !out_f32) return SE3ET_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int cpg = 1;
  if (stats) {
    if (!seg_offsets || nseg <= 0 || groups <= 0 || cout % groups) return SE3ET_ERR_ARG;
    cpg = (int)(cout / groups);
    SE3ET_CUDA_CHECK(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * nseg * groups, st));
  }
  if (nq == 0) return SE3ET_OK;
  int64_t blocks = ceil_div(nq, kC1Warps * 4);
  if (blocks > (int64_t)kNumSMs * 4) blocks = (int64_t)kNumSMs * 4;
  const auto* x = static_cast<const __nv_bfloat16*>(x_bf16);
#define SE3ET_C1(CPL, LIFT)                                                                                          \
  kpconv_cin1_kernel<CPL, LIFT><<<(unsigned)blocks, kC1Warps * 32, 0, st>>>(                                          \
      q_pts, s_pts, neighbors, (int)h, nq, ns, x, w_36xcout, kernel_points_15x3, 1.f / kp_extent, out_f32, stats,     \
      seg_offsets, (int)nseg, cpg)
  if (cout == 32) {
    if (lifted) SE3ET_C1(1, true); else SE3ET_C1(1, false);
  } else {
    if (lifted) SE3ET_C1(2, true); else SE3ET_C1(2, false);
  }
#undef SE3ET_C1
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}
