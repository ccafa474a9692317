// First backbone layer (SimpleBlockEPN on the LiftBlockEPN output, Cin = 1, the same value for the six anchors):
// KPConvInterSO3.forward (blocks_epn.py:454-546 with 334-390) as
//   D[p][beta]   = sum_n W16[p][beta][n] * f[idx[p][n]]                       (16 basis products per point)
//   out[p][r][d] = sum_kc D[p][basis_row(r, kc)] * (sum_a W[kc][a][0][d])     (a [points x 16] x [16 x 6 Cout] product)
// Round-1 kernel (kpconv.cu, kpconv_cin1_kernel<., true>): one warp per point, a lane per neighbour (38 of 64 lane slots
// busy), D reduced through shared memory, the second product on CUDA cores: 1.7 ms per 32 stacked pairs.  Here:
//   * a THREAD per point walks its neighbours (ids of the NEXT tile staged by cp.async, support points as one 16-byte
//     {x, y, z, f} load from a packed copy), the 16 accumulators stay in registers: no idle lanes, no reduction;
//   * the second product runs on mma.sync (m16n8k16, bf16 operands split hi + lo, three products: fp32-grade result),
//     warp w owning a quarter of the output columns for all 128 points of the tile, so that the per-pair GroupNorm
//     sums of its columns stay in registers until the pair (or the CTA's range) ends.
#include <cuda_bf16.h>

#include "common.cuh"
#include "kpconv_mma.cuh"
#include "kpconv_tables.cuh"

namespace se3et {
namespace lift {
using namespace kpm;

constexpr int kThreads = 128;
constexpr int kTile = 128;  // points per tile: one per thread
constexpr int kMaxH = 48;

__constant__ int8_t c_basis_row[6][6] = {
#define SE3ET_BR(r) {(int8_t)basis_row(r, 0), (int8_t)basis_row(r, 1), (int8_t)basis_row(r, 2), (int8_t)basis_row(r, 3), \
                     (int8_t)basis_row(r, 4), (int8_t)basis_row(r, 5)}
    SE3ET_BR(0), SE3ET_BR(1), SE3ET_BR(2), SE3ET_BR(3), SE3ET_BR(4), SE3ET_BR(5)
#undef SE3ET_BR
};

// {x, y, z, f}: one 16-byte gather per neighbour instead of three 4-byte ones and a 2-byte one
__global__ void __launch_bounds__(256) pack_support_kernel(const float* __restrict__ s_pts,
                                                            const __nv_bfloat16* __restrict__ x, int64_t ns,
                                                            float4* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ns) out[i] = make_float4(s_pts[3 * i], s_pts[3 * i + 1], s_pts[3 * i + 2], __bfloat162float(x[i]));
}

__device__ __forceinline__ float sqrt_approx(float v) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}

// influence weights of one neighbour (blocks_epn.py:341-353, 'linear'), folded into the 16 basis rows and scaled by f
__device__ __forceinline__ void accumulate(float dx, float dy, float dz, float f, const float4* __restrict__ kp,
                                           float inv_extent, float (&D)[16]) {
  float w[kKP];
#pragma unroll
  for (int k = 0; k < kKP; ++k) {
    const float4 kk = kp[k];
    const float ex = dx - kk.x, ey = dy - kk.y, ez = dz - kk.z;
    w[k] = fmaxf(0.f, 1.f - sqrt_approx(ex * ex + ey * ey + ez * ez) * inv_extent);
  }
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kKP; ++k)
      if (basis_mask(r) & (1u << k)) s += w[k];
    D[r] = fmaf(s, f, D[r]);
  }
}

__device__ __forceinline__ uint32_t bf16_bits(float v) {
  return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v));
}
__device__ __forceinline__ float bf16_value(uint32_t bits) { return __uint_as_float(bits << 16); }

template <int COUT, bool kOutBf16>
__global__ void __launch_bounds__(kThreads, 3)
kpconv_lift_kernel(const float* __restrict__ q_pts, const float4* __restrict__ s4, const int64_t* __restrict__ idx, int H,
                   int64_t nq, int64_t ns, const float* __restrict__ w /* [36][COUT] */,
                   const float* __restrict__ kernel_points, float inv_extent, void* __restrict__ out_v,
                   double* __restrict__ stats, const int64_t* __restrict__ seg_off, int nseg, int cpg) {
  constexpr int NC = 6 * COUT;  // output columns (r, d)
  constexpr int NT = NC / 8;    // n-tiles of 8 columns
  constexpr int NTW = NT / 4;   // per warp
  extern __shared__ __align__(128) uint8_t smem[];
  // rows of 32 bytes = 16 bf16 (k = basis row), the two 16-byte halves swapped on rows with bit 2 set: conflict-free
  // ldmatrix and 16-byte row stores
  uint8_t* sh_b = smem;                                      // [hi | lo][NC rows]: Bm[beta][(r, d)], row = column (r, d)
  uint8_t* sh_d = sh_b + 2 * NC * 32;                        // [hi | lo][kTile rows]: D[p][beta]
  float4* sh_kp = reinterpret_cast<float4*>(sh_d + 2 * kTile * 32);
  float* sh_stat = reinterpret_cast<float*>(sh_kp + 16);     // [COUT][sum, sum sq]
  int* sh_idx = reinterpret_cast<int*>(sh_stat + 2 * COUT);  // [2][kTile][HS]: this tile's ids, the next tile's in flight
  const int HS = H | 1;  // odd row pitch: a thread per row reads without bank conflicts
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  if (tid < kKP)
    sh_kp[tid] = make_float4(kernel_points[3 * tid], kernel_points[3 * tid + 1], kernel_points[3 * tid + 2], 0.f);
  float* wsum = reinterpret_cast<float*>(sh_idx);  // [6][COUT] weights summed over the anchor slot
  for (int i = tid; i < kKC * COUT; i += kThreads) {
    const int kc = i / COUT, d = i - kc * COUT;
    float t = 0.f;
    for (int a = 0; a < kA; ++a) t += w[(kc * kA + a) * COUT + d];
    wsum[i] = t;
  }
  __syncthreads();
  for (int e = tid; e < NC * 16; e += kThreads) {
    const int col = e >> 4, beta = e & 15, r = col / COUT, d = col - r * COUT;
    float v = 0.f;
#pragma unroll
    for (int kc = 0; kc < kKC; ++kc)
      if (c_basis_row[r][kc] == beta) v += wsum[kc * COUT + d];
    const uint32_t hi = bf16_bits(v), lo = bf16_bits(v - bf16_value(hi));
    const int off = col * 32 + ((((beta >> 3) ^ (col >> 2)) & 1) << 4) + ((beta & 7) << 1);
    *reinterpret_cast<uint16_t*>(sh_b + off) = (uint16_t)hi;
    *reinterpret_cast<uint16_t*>(sh_b + NC * 32 + off) = (uint16_t)lo;
  }
  __syncthreads();

  const int64_t ntiles = (nq + kTile - 1) / kTile;
  const int64_t tiles_per_cta = (ntiles + gridDim.x - 1) / gridDim.x;
  const int64_t p0 = (int64_t)blockIdx.x * tiles_per_cta * kTile;
  const int64_t p1 = min(nq, p0 + tiles_per_cta * kTile);
  if (p0 >= p1) return;

  float st_s[NTW][2], st_ss[NTW][2];
#pragma unroll
  for (int i = 0; i < NTW; ++i) st_s[i][0] = st_s[i][1] = st_ss[i][0] = st_ss[i][1] = 0.f;
  const uint32_t sh_b_s = smem_addr(sh_b), sh_d_s = smem_addr(sh_d);
  // ldmatrix row addresses.  A (D rows = points): matrices (rows 0-7, k 0-7), (rows 8-15, k 0-7), (rows 0-7, k 8-15),
  // (rows 8-15, k 8-15); B (rows = output columns): (hi, k 0-7), (hi, k 8-15), (lo, k 0-7), (lo, k 8-15)
  const int a_row = (lane & 7) + ((lane >> 3) & 1) * 8, a_kh = lane >> 4;
  const int b_row = lane & 7, b_kh = (lane >> 3) & 1, b_lo = lane >> 4;

  // ---- tiles: consecutive 128-point chunks of the CTA's range that never straddle a pair -----------------------------
  struct Tile {
    int64_t t0;
    int np, seg;
  };
  auto tile_at = [&](int64_t t0, int seg_hint) {
    Tile t;
    t.t0 = t0;
    t.seg = seg_hint;
    t.np = 0;
    if (t0 >= p1) return t;
    int64_t pb = p1;
    if (stats) {
      while (t.seg + 1 < nseg && seg_off[t.seg + 1] <= t0) ++t.seg;
      pb = min(p1, t.seg == nseg - 1 ? nq : seg_off[t.seg + 1]);
    }
    t.np = (int)min((int64_t)kTile, pb - t0);
    return t;
  };
  // neighbour ids of a tile -> shared memory, asynchronously (4-byte cp.async of the LOW words of the int64 ids: every
  // valid id is < ns < 2^31 and the shadow id is ns; negative ids have the top bit set and compare >= ns as unsigned)
  const uint32_t sh_idx_s = smem_addr(sh_idx);
  auto fetch_ids = [&](const Tile& t, int buf) {
    const uint32_t dst0 = sh_idx_s + (uint32_t)(buf * kTile * HS * 4);
    for (int r = warp; r < t.np; r += kThreads / 32) {
      const int64_t* src = idx + (t.t0 + r) * H;
      for (int n = lane; n < H; n += 32)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst0 + (uint32_t)((r * HS + n) * 4)), "l"(src + n)
                     : "memory");
    }
    cp_async_commit();
  };

  Tile cur = tile_at(p0, stats ? segment_of(seg_off, nseg, p0) : 0);
  fetch_ids(cur, 0);
  for (int it = 0; cur.np > 0; ++it) {
    const Tile nxt = tile_at(cur.t0 + cur.np, cur.seg);
    if (nxt.np > 0) fetch_ids(nxt, (it + 1) & 1); else cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const int64_t t0 = cur.t0;
    const int np = cur.np;
    const uint32_t* ids = reinterpret_cast<const uint32_t*>(sh_idx) + (it & 1) * kTile * HS;
    {
      // ---- 2. the 16 basis products, a thread per point ----------------------------------------------------------
      float D[16];
#pragma unroll
      for (int r = 0; r < 16; ++r) D[r] = 0.f;
      if (tid < np) {
        const int64_t p = t0 + tid;
        const float qx = q_pts[3 * p], qy = q_pts[3 * p + 1], qz = q_pts[3 * p + 2];
        const uint32_t* my = ids + tid * HS;
        // software pipeline: the gathers of the next two neighbours are in flight while one is evaluated
        const uint32_t nsu = (uint32_t)ns;
        auto id_at = [&](int n) {  // -1: shadow neighbour or past the row
          const uint32_t u = n < H ? my[n] : 0xffffffffu;
          return u < nsu ? (int)u : -1;
        };
        int j0 = id_at(0), j1 = id_at(1);
        float4 sp0 = __ldg(s4 + max(j0, 0)), sp1 = __ldg(s4 + max(j1, 0));  // shadows: a valid address, weight zero
#pragma unroll 1
        for (int n = 0; n < H; n += 2) {
          const int j2 = id_at(n + 2), j3 = id_at(n + 3);
          const float4 sp2 = __ldg(s4 + max(j2, 0)), sp3 = __ldg(s4 + max(j3, 0));
          accumulate(sp0.x - qx, sp0.y - qy, sp0.z - qz, j0 < 0 ? 0.f : sp0.w, sh_kp, inv_extent, D);
          accumulate(sp1.x - qx, sp1.y - qy, sp1.z - qz, j1 < 0 ? 0.f : sp1.w, sh_kp, inv_extent, D);
          j0 = j2; j1 = j3; sp0 = sp2; sp1 = sp3;
        }
      }
      // ---- 3. D -> bf16 hi + lo rows (rows past np are zero) -----------------------------------------------------
      {
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t h0 = bf16_bits(D[2 * i]), h1 = bf16_bits(D[2 * i + 1]);
          hi[i] = h0 | (h1 << 16);
          lo[i] = pack2(D[2 * i] - bf16_value(h0), D[2 * i + 1] - bf16_value(h1));
        }
        const int sw = (tid >> 2) & 1;
        uint4* dh = reinterpret_cast<uint4*>(sh_d + tid * 32);
        uint4* dl = reinterpret_cast<uint4*>(sh_d + kTile * 32 + tid * 32);
        dh[sw] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        dh[sw ^ 1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
        dl[sw] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        dl[sw ^ 1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
      }
      __syncthreads();
      // ---- 4. out[128 x NC] = D[128 x 16] Bm[16 x NC]: warp -> n-tiles [warp * NTW, (warp + 1) * NTW) -------------
      uint32_t ah[8][4], al[8][4];
#pragma unroll
      for (int mt = 0; mt < 8; ++mt) {
        const int row = mt * 16 + a_row;
        const uint32_t off = (uint32_t)(row * 32 + (((a_kh ^ (row >> 2)) & 1) << 4));
        ldmatrix_x4(ah[mt], sh_d_s + off);
        ldmatrix_x4(al[mt], sh_d_s + kTile * 32 + off);
      }
#pragma unroll
      for (int i = 0; i < NTW; ++i) {
        const int nt = warp * NTW + i;
        uint32_t b[4];
        {
          const int row = nt * 8 + b_row;
          ldmatrix_x4(b, sh_b_s + (uint32_t)(b_lo * NC * 32 + row * 32 + (((b_kh ^ (row >> 2)) & 1) << 4)));
        }
        const int col = nt * 8 + 2 * (lane & 3);
#pragma unroll
        for (int mt = 0; mt < 8; ++mt) {
          if (mt * 16 >= np) break;
          float c[4] = {0.f, 0.f, 0.f, 0.f};
          mma_16816(c, al[mt], b[0], b[1]);
          mma_16816(c, ah[mt], b[2], b[3]);
          mma_16816(c, ah[mt], b[0], b[1]);
          st_s[i][0] += c[0] + c[2];
          st_s[i][1] += c[1] + c[3];
          st_ss[i][0] = fmaf(c[0], c[0], fmaf(c[2], c[2], st_ss[i][0]));
          st_ss[i][1] = fmaf(c[1], c[1], fmaf(c[3], c[3], st_ss[i][1]));
          const int r0 = mt * 16 + (lane >> 2);
          const int64_t o0 = (t0 + r0) * NC + col;
          if (kOutBf16) {
            __nv_bfloat16* out = static_cast<__nv_bfloat16*>(out_v);
            if (r0 < np) *reinterpret_cast<uint32_t*>(out + o0) = pack2(c[0], c[1]);
            if (r0 + 8 < np) *reinterpret_cast<uint32_t*>(out + o0 + 8 * NC) = pack2(c[2], c[3]);
          } else {
            float* out = static_cast<float*>(out_v);
            if (r0 < np) *reinterpret_cast<float2*>(out + o0) = make_float2(c[0], c[1]);
            if (r0 + 8 < np) *reinterpret_cast<float2*>(out + o0 + 8 * NC) = make_float2(c[2], c[3]);
          }
        }
      }
      __syncthreads();  // sh_d and this tile's id buffer are rewritten two tiles from now / by the next tile
    }
    if (stats && (nxt.np == 0 || nxt.seg != cur.seg)) {
      const int seg = cur.seg;
      // per-pair GroupNorm sums: lanes (registers) -> columns (shared memory) -> groups (fp64 atomics)
      for (int i = tid; i < 2 * COUT; i += kThreads) sh_stat[i] = 0.f;
      __syncthreads();
#pragma unroll
      for (int i = 0; i < NTW; ++i) {
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          float a = st_s[i][c], b = st_ss[i][c];
#pragma unroll
          for (int o = 4; o < 32; o <<= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            b += __shfl_xor_sync(0xffffffffu, b, o);
          }
          if (lane < 4) {
            const int col = (warp * NTW + i) * 8 + 2 * lane + c;
            const int d = col % COUT;
            atomicAdd(&sh_stat[2 * d], a);
            atomicAdd(&sh_stat[2 * d + 1], b);
          }
          st_s[i][c] = st_ss[i][c] = 0.f;
        }
      }
      __syncthreads();
      const int G = COUT / cpg;
      for (int i = tid; i < 2 * G; i += kThreads) {
        const int g = i >> 1, which = i & 1;
        float t = 0.f;
        for (int c = g * cpg; c < (g + 1) * cpg; ++c) t += sh_stat[2 * c + which];
        atomicAdd(stats + ((int64_t)seg * G + g) * 2 + which, (double)t);
      }
      __syncthreads();
    }
    cur = nxt;
  }
}

template <int COUT, bool kOutBf16>
static int launch(const float* q_pts, const float4* s4, const int64_t* idx, int h, int64_t nq, int64_t ns, const float* w,
                  const float* kp, float inv_extent, void* out, double* stats, const int64_t* seg_off, int nseg, int cpg,
                  cudaStream_t st) {
  const int hs = h | 1;
  const size_t smem = (size_t)2 * 6 * COUT * 32 + 2 * kTile * 32 + 16 * sizeof(float4) + 2 * COUT * sizeof(float) +
                      (size_t)2 * kTile * hs * sizeof(int);   // two id buffers
  SE3ET_ENSURE_SMEM((kpconv_lift_kernel<COUT, kOutBf16>), smem);
  int64_t blocks = ceil_div(nq, kTile);
  if (blocks > (int64_t)kNumSMs * 3) blocks = (int64_t)kNumSMs * 3;
  kpconv_lift_kernel<COUT, kOutBf16><<<(unsigned)blocks, kThreads, smem, st>>>(
      q_pts, s4, idx, h, nq, ns, w, kp, inv_extent, out, stats, seg_off, nseg, cpg);
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}

}  // namespace lift
}  // namespace se3et

extern "C" int64_t se3et_kpconv_lift_workspace_bytes(int64_t ns) { return ns > 0 ? ns * 16 + 256 : 256; }

extern "C" int se3et_kpconv_lift(const float* q_pts, const float* s_pts, const int64_t* neighbors, int64_t nq, int64_t ns,
                                 int64_t h, const void* f_bf16, const float* w_36xcout, int64_t cout,
                                 const float* kernel_points_15x3, float kp_extent, void* out, int out_bf16,
                                 double* stats, const int64_t* seg_offsets, int64_t nseg, int64_t groups,
                                 void* workspace, int64_t workspace_bytes, se3et_stream_t stream) {
  using namespace se3et;
  if (nq < 0 || ns <= 0 || h <= 0 || cout <= 0 || !(kp_extent > 0.f)) return SE3ET_ERR_ARG;
  if ((cout != 32 && cout != 64) || h > lift::kMaxH || ns >= ((int64_t)1 << 31)) return SE3ET_ERR_UNSUPPORTED;
  if (!q_pts || !s_pts || !neighbors || !f_bf16 || !w_36xcout || !kernel_points_15x3 || !out || !workspace)
    return SE3ET_ERR_ARG;
  if (workspace_bytes < se3et_kpconv_lift_workspace_bytes(ns)) return SE3ET_ERR_WORKSPACE;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int cpg = 1;
  if (stats) {
    if (!seg_offsets || nseg <= 0 || groups <= 0 || cout % groups) return SE3ET_ERR_ARG;
    cpg = (int)(cout / groups);
    SE3ET_CUDA_CHECK(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * nseg * groups, st));
  }
  if (nq == 0) return SE3ET_OK;
  float4* s4 = reinterpret_cast<float4*>(align_up(reinterpret_cast<size_t>(workspace), 256));
  lift::pack_support_kernel<<<(unsigned)ceil_div(ns, 256), 256, 0, st>>>(
      s_pts, static_cast<const __nv_bfloat16*>(f_bf16), ns, s4);
  SE3ET_LAUNCH_CHECK();
  const float inv = 1.f / kp_extent;
#define SE3ET_LIFT(C, B)                                                                                              \
  return lift::launch<C, B>(q_pts, s4, neighbors, (int)h, nq, ns, w_36xcout, kernel_points_15x3, inv, out, stats,    \
                            seg_offsets, (int)nseg, cpg, st)
  if (cout == 32) {
    if (out_bf16) SE3ET_LIFT(32, true);
    SE3ET_LIFT(32, false);
  }
  if (out_bf16) SE3ET_LIFT(64, true);
  SE3ET_LIFT(64, false);
#undef SE3ET_LIFT
}
