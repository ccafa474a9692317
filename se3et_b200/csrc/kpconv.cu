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

template <int KS>
static int launch_gather_mma(const float* q_pts, const float* s_pts, const int64_t* neighbors, int64_t nq, int64_t ns,
                             int h, const __nv_bfloat16* x, int cin, const float* kp, float inv_extent,
                             __nv_bfloat16* out, int kpad, cudaStream_t st) {
  using S = GatherSmem<KS>;
  static bool configured = false;
  if (!configured) {
    SE3ET_CUDA_CHECK(cudaFuncSetAttribute(kpconv_gather_mma_kernel<KS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          S::kTotal));
    configured = true;
  }
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
