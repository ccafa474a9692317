// KPConvInterSO3.forward (blocks_epn.py:454-546 with feat_gather_by_perm :334-390) with the UMMA rows = QUERY POINTS.
//
//   out[p][r][d] = sum_{kc, a, c} B[p][beta(r, kc)][a][c] * W[kc][ridx[a][r]][c][d]
//   B[p][beta][a][c] = sum_n W16[p][beta][n] * x[idx[p][n]][a][c]            (16-row basis, kpconv_tables.cuh)
//
// The products B are stored ONCE; the (r, kc) fan-out that kpconv_fused.cu pays for with 2.25 copies of every product
// is done by descriptor selection: a step = (16-channel chunk, input anchor a) produces the operand
// [128 points][16 basis rows][16 channels] (64 KB, 4 sub-tiles of [128 rows x 128 B], 128-byte swizzle, double
// buffered) and the MMA warp issues, for each of the 36 (r, kc), one K = 16 tcgen05.mma whose A descriptor points at
// basis slab beta(r, kc) and whose B descriptor points at W[kc][ridx[a][r]][chunk] -- host pre-arranged in step order
// (e2pn.py:KPConvInterSO3._w_rows) so a TMA ring streams it -- into the TMEM accumulator of output anchor r
// (6 x BN fp32 columns).
//
// Persistent CTA, tile = 16 * PPW points:
//   warps 0-15  producers, PPW points each (row = i * 16 + warp).  Per tile: the 16 x H basis weights of each point as
//               mma.sync A fragments, kept for all steps of the tile in TENSOR MEMORY columns the accumulators leave free
//               (tcgen05.st once, tcgen05.ld per step; the first PPW - TP points stay in registers).  Per (step,
//               point): cp.async gather of the neighbour rows x[idx[n]][a][chunk] (32 B sectors, 3-stage ring per
//               warp), mma.sync W16 . X, ONE stmatrix.x4 of the 16 x 16 product into the operand tile.
//   warp 16     tcgen05.mma issuer.
//   warp 17     TMA producer of the weight K-blocks (4 (r, kc) slices of 16 channels each).
//   epilogue    all 16 producer warps (TMEM lane quadrant warp % 4, column group warp / 4): fp32 rows to global memory.
#include <cuda_bf16.h>

#include "common.cuh"
#include "kpconv_mma.cuh"
#include "tc.cuh"

namespace se3et {
namespace rows {

using namespace kpm;

constexpr int kProdWarps = 16;
constexpr int kThreads = (kProdWarps + 2) * 32;
constexpr int kSubTile = 128 * 128;      // [128 rows][4 basis rows x 16 channels] bf16
constexpr int kBBuf = 4 * kSubTile;      // operand of one step
constexpr int kXRow = 32;                // bytes per gathered row (16 channels)
#ifndef SE3ET_ROWS_XS32
#define SE3ET_ROWS_XS32 3   // measured on B200: a fourth gather stage changes nothing (1.538 vs 1.532 ms at level 0)
#endif
#ifndef SE3ET_ROWS_FRAG_PREFETCH
#define SE3ET_ROWS_FRAG_PREFETCH 1
#endif
constexpr int kWMaxStages = 12;
constexpr int kW16Row = 112;             // scratch row pitch (bytes): 48 neighbours + 8 pad, conflict-free ldmatrix

__device__ __forceinline__ void mma_1688(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ void ldmatrix_x2(uint32_t& r0, uint32_t& r1, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void cp_async_16_if(uint32_t dst, const void* src, bool pred) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %2, 0;\n"
      "@p cp.async.cg.shared.global [%0], [%1], 16;\n"
      "}\n" ::"r"(dst),
      "l"(src), "r"((uint32_t)pred)
      : "memory");
}
// TMEM as a register stash: thread t of warp w <-> lane 32 (w % 4) + t, consecutive columns <-> consecutive registers
__device__ __forceinline__ void tmem_st2(uint32_t taddr, uint32_t a, uint32_t b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
               "r"(r[2]), "r"(r[3])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t& a, uint32_t& b) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}

// basis_weights (kpconv_mma.cuh) with the approximate square root (MUFU.SQRT, <= 2 ulp): the result is rounded to bf16
// right after, so the two agree except on bf16 rounding ties; invalid slots give zero rows
__device__ __forceinline__ float sqrt_approx(float v) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void basis_weights_fast(float dx, float dy, float dz, const float4* __restrict__ kp,
                                                   float inv_extent, bool valid, float (&row)[16]) {
  float w[kKP];
#pragma unroll
  for (int k = 0; k < kKP; ++k) {
    const float4 kk = kp[k];  // one 16-byte shared-memory load per kernel point
    const float ex = dx - kk.x, ey = dy - kk.y, ez = dz - kk.z;
    w[k] = valid ? fmaxf(0.f, 1.f - sqrt_approx(ex * ex + ey * ey + ez * ez) * inv_extent) : 0.f;
  }
#pragma unroll
  for (int r = 0; r < 16; ++r) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < kKP; ++k)
      if (basis_mask(r) & (1u << k)) s += w[k];
    row[r] = s;
  }
}

// ---- thread-block cluster helpers (CL > 1: the CTAs of a cluster share one tile of points, each accumulates its own
// BN output columns and produces 1 / CL of the operand rows, which it copies to its peers through distributed shared
// memory) ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t num_clusters_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t smem_addr_cta, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr_cta), "r"(rank));
  return r;
}
// local shared memory -> a peer CTA's shared memory, completing on the PEER's mbarrier (both cluster addresses)
__device__ __forceinline__ void dsmem_bulk_copy(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes, uint32_t mbar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   dst_cluster),
               "r"(src_cta), "r"(bytes), "r"(mbar_cluster)
               : "memory");
}
// TMA tile -> the same shared-memory offset of every CTA in `mask`, completing bytes on each CTA's own mbarrier
__device__ __forceinline__ void tma_load_2d_multicast(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                      uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], "
      "[%2], %5;" ::"r"(tc::smem_u32(smem_dst)),
      "l"(m), "r"(tc::smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
// arrive on the same-offset mbarrier of every CTA in `mask` when all previously issued MMAs of this thread completed
__device__ __forceinline__ void umma_commit_multicast(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   tc::smem_u32(bar)),
               "h"(mask)
               : "memory");
}

// (r, kc) -> basis row, t = r * 6 + kc
struct BetaTab {
  int8_t v[36];
};
__host__ __device__ constexpr BetaTab make_beta_tab() {
  BetaTab t{};
  for (int r = 0; r < kA; ++r)
    for (int kc = 0; kc < kKC; ++kc) t.v[r * 6 + kc] = (int8_t)basis_row(r, kc);
  return t;
}

struct Args {
  const float* q_pts;
  const float* s_pts;
  const int64_t* idx;
  const __nv_bfloat16* x;
  const float* kernel_points;
  float* out;  // fp32 [nq * 6, cout]
  __nv_bfloat16* out_bf16;  // or bf16, same shape (exactly one of the two is set)
  int64_t nq, ns;
  int H;
  int cin, cout;
  int wstages;
  float inv_extent;
};

template <int BN, int KH>
struct Smem {
  static constexpr int kHR = 8 * KH;
  static constexpr int kXStage = kHR * kXRow;
  static constexpr int kBOff = 0;
  static constexpr int kXOff = 2 * kBBuf;
  // gather ring stages per producer warp: the 32-column kernels have the shared memory for a fourth one
  static constexpr int kXS = BN == 32 ? SE3ET_ROWS_XS32 : 3;
  static constexpr int kXBytes = kProdWarps * kXS * kXStage;
  static constexpr int kBarOff = kXOff + kXBytes;                       // 512 B of barriers
  static constexpr int kKpOff = kBarOff + 512;                          // 15 x float4
  static constexpr int kWOff = (kKpOff + 256 + 1023) / 1024 * 1024;
  static constexpr int kWStage = (BN * 128 + 1023) / 1024 * 1024;
  static int max_wstages() {
    const int n = (227 * 1024 - 1024 - kWOff) / kWStage;
    return n > kWMaxStages ? kWMaxStages : n;
  }
  static int total(int wstages) { return kWOff + wstages * kWStage + 1024; }
};

// BN: output columns per CTA; KH: neighbour columns / 8 rounded up (4, 5, 6); PPW: points per producer warp (of this
// CTA); TP: points per warp whose basis fragments live in tensor memory; CL: CTAs per cluster sharing a tile (1: none);
// WM: CTAs per cluster that work on DIFFERENT tiles but consume the same weight stream: every CTA loads 1 / WM of each
// weight K-block and TMA-multicasts it to the cluster (L2 reads and bytes in flight per CTA divided by WM)
template <int BN, int KH, int PPW, int TP, int CL, int WM>
__global__ void __launch_bounds__(kThreads + 32, 1)
kpconv_rows_kernel(const __grid_constant__ CUtensorMap tma_w, Args args) {
  using S = Smem<BN, KH>;
  constexpr int FR = 2 * KH;              // fragment registers per point
  constexpr int RP = PPW - TP;            // points kept in registers
  constexpr int kStash = TP * FR;         // stash columns per producer warp
  constexpr int kAccCols = 6 * BN;
  static_assert(kAccCols + 4 * kStash <= 512, "tensor memory over-subscribed");
  constexpr int XS = S::kXS, PF = XS - 1;   // ring stages, items the gather runs ahead
  static_assert(kW16Row * 16 <= XS * S::kXStage, "W16 scratch must fit the gather ring");
  static_assert((3 * PPW) % XS == 0 && PPW >= PF, "ring stage of item (step, i) must be a compile-time constant");
  constexpr int kTilePts = 16 * PPW * CL;
  static_assert(kTilePts <= 128 && (CL == 1 || kTilePts == 128), "a cluster shares one full 128-row tile");
  static_assert(CL == 1 || WM == 1, "one kind of cluster at a time");
  static_assert((BN / WM) % 8 == 0, "weight slices are whole swizzle atoms");
  constexpr bool kCluster = CL > 1 || WM > 1;
  constexpr int kHR = S::kHR;
  constexpr int kXStage = S::kXStage;
  constexpr BetaTab kBeta = make_beta_tab();

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* b_tile = smem + S::kBOff;
  uint8_t* w_tile = smem + S::kWOff;
  uint64_t* b_full = reinterpret_cast<uint64_t*>(smem + S::kBarOff);  // [2]
  uint64_t* b_empty = b_full + 2;                                     // [2]
  uint64_t* tmem_full = b_full + 4;
  uint64_t* tmem_empty = b_full + 5;
  uint64_t* p_done = b_full + 6;                                      // [2] (CL > 1: this CTA's rows are stored)
  uint64_t* w_full = b_full + 8;
  uint64_t* w_empty = w_full + kWMaxStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(w_empty + kWMaxStages);
  float4* sh_kp = reinterpret_cast<float4*>(smem + S::kKpOff);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CL > 1 ? cluster_ctarank() : 0u;
  const int n0 = (CL > 1 ? (int)rank : (int)blockIdx.y) * BN;
  const int64_t tile0 = CL > 1 ? (int64_t)cluster_id_x() : (int64_t)blockIdx.x;
  const int64_t tile_stride = CL > 1 ? (int64_t)num_clusters_x() : (int64_t)gridDim.x;
  const int64_t ntiles = (args.nq + kTilePts - 1) / kTilePts;
  // WM > 1: the CTAs of a cluster advance through the weight stream in lockstep, so all of them run the same number of
  // rounds; a CTA whose tile index passes the end works on a dummy tile (no valid point)
  const int64_t nrounds = (ntiles + tile_stride - 1) / tile_stride;
  const uint32_t wrank = WM > 1 ? cluster_ctarank() : 0u;
  const int nsteps = args.cin / kChunk * kA;  // (chunk, a), a fastest
  const int wstages = args.wstages;

  if (threadIdx.x < kKP)
    sh_kp[threadIdx.x] = make_float4(args.kernel_points[3 * threadIdx.x], args.kernel_points[3 * threadIdx.x + 1],
                                     args.kernel_points[3 * threadIdx.x + 2], 0.f);
  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&tma_w);
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&b_full[s], CL > 1 ? 1 : kProdWarps);
      tc::mbar_init(&b_empty[s], CL);
      tc::mbar_init(&p_done[s], kProdWarps);
    }
    for (int s = 0; s < wstages; ++s) {
      tc::mbar_init(&w_full[s], 1);
      tc::mbar_init(&w_empty[s], WM);
    }
    tc::mbar_init(tmem_full, 1);
    tc::mbar_init(tmem_empty, kProdWarps);
    tc::mbar_fence_init();
  }
  if (kTilePts < 128) {
    // rows kTilePts .. 127 of the operand are never produced: keep them finite
    for (int i = threadIdx.x; i < 2 * kBBuf / 16; i += blockDim.x) reinterpret_cast<uint4*>(b_tile)[i] = make_uint4(0, 0, 0, 0);
  }
  if (warp == kProdWarps) tc::tmem_alloc<512>(tmem_ptr);
  tc::fence_proxy_async_smem();
  tc::tcgen05_fence_before_sync();
  if (kCluster) cluster_sync_all(); else __syncthreads();   // peers' mbarriers are initialised before anything remote
  tc::tcgen05_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp < kProdWarps) {
    // =========================================== producers ===================================================
    uint8_t* xs = smem + S::kXOff + warp * XS * kXStage;
    const uint32_t xs_s = smem_addr(xs);
    const int H = args.H;
    const uint32_t row_chunks = (uint32_t)(kA * args.cin / 8);  // 16-byte units per support row
    // ldmatrix lane addressing inside a 16-neighbour k-step (see kpconv_fused.cu)
    const int ld_row = (lane & 7) + ((lane >> 3) & 1) * 8, ld_half = lane >> 4;
    uint32_t ld_off[(KH + 1) / 2];
#pragma unroll
    for (int ks = 0; ks < KH / 2; ++ks) {
      const int n = ks * 16 + ld_row;
      ld_off[ks] = n * kXRow + ((ld_half ^ ((n >> 2) & 1)) << 4);
    }
    if (KH & 1) {  // 8-neighbour tail: matrices (rows, channels 0-7), (rows, channels 8-15) from lanes 0-7, 8-15
      const int n = (KH - 1) * 8 + (lane & 7), half = (lane >> 3) & 1;
      ld_off[KH / 2] = n * kXRow + ((half ^ ((n >> 2) & 1)) << 4);
    }
    // gather pieces of this lane: piece i = lane + 32 u -> neighbour i / 2, 16-byte half i % 2
    constexpr int NU = (2 * kHR + 31) / 32;
    uint32_t gdst[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const int i = lane + 32 * u, n = i >> 1;
      gdst[u] = xs_s + n * kXRow + (((i & 1) ^ ((n >> 2) & 1)) << 4);
    }
    // stmatrix.x4 row address of this lane: matrix lane / 8 = (basis rows 0-7 | 8-15) x (channels 0-7 | 8-15), basis
    // row beta -> sub-tile beta / 4, slot beta % 4; odd sub-tiles hold the channel halves swapped (conflict-free
    // stores; the step-ordered weights carry the same swap)
    uint32_t st_off;
    {
      const int beta = (lane & 7) + 8 * ((lane >> 3) & 1), half = lane >> 4;
      const int sub = beta >> 2, slot = beta & 3;
      const int chunk = slot * 2 + (half ^ (sub & 1));
      st_off = smem_addr(b_tile) + sub * kSubTile + ((int)rank * PPW * 16 + warp) * 128 + ((chunk ^ (warp & 7)) << 4);
    }
    const uint32_t stash = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(kAccCols + (warp >> 2) * kStash);

    uint32_t afrag[RP > 0 ? RP : 1][FR];
    uint32_t goff[PPW][NU];  // 16-byte units relative to x (+ the piece's half)

    // ---- per-tile setup: basis weights of a point as mma.sync A fragments (through the W16 scratch at the ring's
    // head).  Neighbour ids are loaded two points ahead and neighbour coordinates one point ahead of the arithmetic.
    struct Ids { int j0, j1; };
    struct Rel { float x0, y0, z0, x1, y1, z1; };
    auto load_ids = [&](int64_t tile, int i) -> Ids {
      const int64_t p = tile * kTilePts + ((int)rank * PPW + i) * 16 + warp;
      Ids r{-1, -1};
      if (p < args.nq) {
        const int64_t* row = args.idx + p * H;
        if (lane < H) { const int64_t j = row[lane]; r.j0 = j < args.ns ? (int)j : -1; }
        if (KH > 4 && 32 + lane < H) { const int64_t j = row[32 + lane]; r.j1 = j < args.ns ? (int)j : -1; }
      }
      return r;
    };
    auto load_rel = [&](int64_t tile, int i, Ids id) -> Rel {
      int64_t p = tile * kTilePts + ((int)rank * PPW + i) * 16 + warp;
      if (p >= args.nq) p = 0;
      const float qx = args.q_pts[3 * p], qy = args.q_pts[3 * p + 1], qz = args.q_pts[3 * p + 2];
      Rel r{0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (id.j0 >= 0) {
        const float* s = args.s_pts + 3 * (int64_t)id.j0;
        r.x0 = s[0] - qx; r.y0 = s[1] - qy; r.z0 = s[2] - qz;
      }
      if (KH > 4 && id.j1 >= 0) {
        const float* s = args.s_pts + 3 * (int64_t)id.j1;
        r.x1 = s[0] - qx; r.y1 = s[1] - qy; r.z1 = s[2] - qz;
      }
      return r;
    };
    auto build = [&](Ids id, Rel rel, uint32_t (&fr)[FR]) {
      // every column below 8 KH is written (zeros for shadow / padding slots): no separate clearing pass
      {
        float row[16];
        basis_weights_fast(rel.x0, rel.y0, rel.z0, sh_kp, args.inv_extent, id.j0 >= 0, row);
        uint8_t* dst = xs + lane * 2;
#pragma unroll
        for (int r = 0; r < 16; ++r) *reinterpret_cast<__nv_bfloat16*>(dst + r * kW16Row) = __float2bfloat16(row[r]);
      }
      if (KH > 4 && 32 + lane < kHR) {
        float row[16];
        basis_weights_fast(rel.x1, rel.y1, rel.z1, sh_kp, args.inv_extent, id.j1 >= 0, row);
        uint8_t* dst = xs + (32 + lane) * 2;
#pragma unroll
        for (int r = 0; r < 16; ++r) *reinterpret_cast<__nv_bfloat16*>(dst + r * kW16Row) = __float2bfloat16(row[r]);
      }
      __syncwarp();
#pragma unroll
      for (int ks = 0; ks < KH / 2; ++ks) {
        uint32_t t4[4];
        ldmatrix_x4(t4, xs_s + ld_row * kW16Row + (ks * 16 + ld_half * 8) * 2);
        fr[4 * ks] = t4[0]; fr[4 * ks + 1] = t4[1]; fr[4 * ks + 2] = t4[2]; fr[4 * ks + 3] = t4[3];
      }
      if (KH & 1) ldmatrix_x2(fr[FR - 2], fr[FR - 1], xs_s + (lane & 15) * kW16Row + (KH - 1) * 8 * 2);
      __syncwarp();
    };
    auto setup = [&](int64_t tile) {
      // gather offsets: the neighbour of piece u is lane / 2 + 16 u
#pragma unroll
      for (int i = 0; i < PPW; ++i) {
        const int64_t p = tile * kTilePts + ((int)rank * PPW + i) * 16 + warp;
        const bool pvalid = p < args.nq;
        const int64_t pc = pvalid ? p : 0;
        int64_t j0 = (pvalid && lane < H) ? args.idx[pc * H + lane] : -1;
        int64_t j1 = (pvalid && 32 + lane < H) ? args.idx[pc * H + 32 + lane] : -1;
        if (j0 >= args.ns) j0 = -1;
        if (j1 >= args.ns) j1 = -1;
        // shadow / padding slots carry weight zero: they re-read the point's first neighbour (same sectors as a valid
        // piece of the same instruction: no extra traffic, and no hot spot on one row for the whole grid)
        const int jfb = max(__shfl_sync(0xffffffffu, (int)j0, 0), 0);
#pragma unroll
        for (int u = 0; u < NU; ++u) {
          const int src = (lane >> 1) + 16 * (u & 1);
          const int jj = __shfl_sync(0xffffffffu, (int)(u < 2 ? j0 : j1), src);
          goff[i][u] = (uint32_t)(jj >= 0 ? jj : jfb) * row_chunks + (uint32_t)(lane & 1);
        }
      }
      Ids id_a = load_ids(tile, 0);
      Rel rel_a = load_rel(tile, 0, id_a);
      Ids id_b = load_ids(tile, 1);
#pragma unroll
      for (int i = 0; i < RP; ++i) {
        const Rel rel_b = i + 1 < PPW ? load_rel(tile, i + 1, id_b) : Rel{0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        const Ids id_c = i + 2 < PPW ? load_ids(tile, i + 2) : Ids{-1, -1};
        build(id_a, rel_a, afrag[i]);
        id_a = id_b; rel_a = rel_b; id_b = id_c;
      }
      if (TP > 0) {
#pragma unroll 1
        for (int i = RP; i < PPW; ++i) {
          Rel rel_b{0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          Ids id_c{-1, -1};
          if (i + 1 < PPW) rel_b = load_rel(tile, i + 1, id_b);
          if (i + 2 < PPW) id_c = load_ids(tile, i + 2);
          uint32_t fr[FR];
          build(id_a, rel_a, fr);
          id_a = id_b; rel_a = rel_b; id_b = id_c;
          const uint32_t ta = stash + (uint32_t)((i - RP) * FR);
#pragma unroll
          for (int f = 0; f + 4 <= FR; f += 4) tmem_st4(ta + f, fr + f);
          if (FR & 2) tmem_st2(ta + FR - 2, fr[FR - 2], fr[FR - 1]);
        }
        tmem_st_wait();
      }
      __syncwarp();
    };

    // 64-bit source base of a step: x + (a * cin + chunk * 16) elements
    auto step_base = [&](int step) -> const uint8_t* {
      const int chunk = step / kA, a = step - chunk * kA;
      return reinterpret_cast<const uint8_t*>(args.x + a * args.cin + chunk * kChunk);
    };
    auto issue = [&](const uint8_t* base, int i, int stage) {
#pragma unroll
      for (int u = 0; u < NU; ++u) {
#ifndef SE3ET_ROWS_NOGATHER
        if (u * 32 + 31 < 2 * kHR)
          cp_async_16(gdst[u] + stage * kXStage, base + (uint64_t)goff[i][u] * 16u, 16);
        else  // partial instruction: predicated, no branch
          cp_async_16_if(gdst[u] + stage * kXStage, base + (uint64_t)goff[i][u] * 16u, lane + 32 * u < 2 * kHR);
#endif
      }
    };

    uint32_t frag_a[FR], frag_b[FR];
    auto load_stash = [&](int i, uint32_t (&dst)[FR]) {
      const uint32_t ta = stash + (uint32_t)((i - RP) * FR);
      if (FR >= 8) {
        tmem_ld8(ta, dst);
        if (FR == 12) tmem_ld4(ta + 8, dst + 8);
        if (FR == 10) tmem_ld2(ta + 8, dst[8], dst[9]);
      } else {
        tmem_ld4(ta, dst);
        tmem_ld4(ta + 4, dst + 4);
      }
    };

    uint32_t gstep = 0;   // steps produced so far (all tiles)
    // rotated tile loop (one call site per phase): production(t), setup(t + 1), epilogue(t)
    for (int64_t tile = tile0 - tile_stride, titer = -1;; tile += tile_stride, ++titer) {
      const bool have = titer >= 0;
      const bool have_next = WM > 1 ? titer + 1 < nrounds : tile + tile_stride < ntiles;
      if (have) {
      if (SE3ET_ROWS_FRAG_PREFETCH && RP == 0) load_stash(0, frag_a);   // point 0's fragments (later ones: one item ahead)
      // items (step, point) in order; the ring runs PF items ahead
      {
        const uint8_t* b0 = step_base(0);
#pragma unroll
        for (int j = 0; j < PF; ++j) {
          issue(b0, j, j);
          cp_async_commit();
        }
      }
      // three steps per trip: the ring stage of item (step, i) is then a compile-time constant (nsteps = 6 chunks)
#pragma unroll 1
      for (int step3 = 0; step3 < nsteps; step3 += 3) {
#pragma unroll
      for (int s3 = 0; s3 < 3; ++s3, ++gstep) {
        const int step = step3 + s3;
        const uint32_t buf = gstep & 1u;
        const uint8_t* base = step_base(step);
        const uint8_t* base_next = step + 1 < nsteps ? step_base(step + 1) : nullptr;
        const uint32_t st_base = st_off + buf * kBBuf;
#pragma unroll
        for (int i = 0; i < PPW; ++i) {
          {  // prefetch the item PF ahead
            const int i2 = (i + PF) % PPW;
            const uint8_t* b2 = i + PF < PPW ? base : base_next;
            const int stage = (s3 * PPW + i) % XS;
            const int st2 = (stage + PF) % XS;
            if (b2) issue(b2, i2, st2);
            cp_async_commit();
          }
          // basis fragments of point i: registers, or the tensor-memory stash -- loaded one item ahead into the other
          // of two register sets (PPW is even, so the parity of i names the set in every step)
          uint32_t (&fr)[FR] = (i & 1) ? frag_b : frag_a;
          uint32_t (&fr_next)[FR] = (i & 1) ? frag_a : frag_b;
          if (i < RP) {
#pragma unroll
            for (int f = 0; f < FR; ++f) fr[f] = afrag[i < RP ? i : 0][f];
          } else if (!SE3ET_ROWS_FRAG_PREFETCH) {
            load_stash(i, fr);
          }
          if (i >= RP) tc::tmem_ld_wait();
          if (SE3ET_ROWS_FRAG_PREFETCH && (i + 1) % PPW >= RP) load_stash((i + 1) % PPW, fr_next);
          cp_async_wait<PF>();
          __syncwarp();
          if (i == 0) {
            // the MMAs of the step before last have consumed this operand buffer
            tc::mbar_wait_long(&b_empty[buf], ((gstep >> 1) & 1u) ^ 1u);
          }
          float d[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
          const uint32_t xsb = xs_s + ((s3 * PPW + i) % XS) * kXStage;
#ifndef SE3ET_ROWS_NOLDSM
#pragma unroll
          for (int ks = 0; ks < KH / 2; ++ks) {
            uint32_t b[4];
            ldmatrix_x4_trans(b, xsb + ld_off[ks]);
            const uint32_t a4[4] = {fr[4 * ks], fr[4 * ks + 1], fr[4 * ks + 2], fr[4 * ks + 3]};
            mma_16816(d[0], a4, b[0], b[1]);
            mma_16816(d[1], a4, b[2], b[3]);
          }
          if (KH & 1) {
            uint32_t b0, b1;
            ldmatrix_x2_trans(b0, b1, xsb + ld_off[KH / 2]);
            mma_1688(d[0], fr[FR - 2], fr[FR - 1], b0);
            mma_1688(d[1], fr[FR - 2], fr[FR - 1], b1);
          }
#else
          d[0][0] = __uint_as_float(fr[0] ^ xsb); d[1][1] = __uint_as_float(fr[FR - 1]);
#endif
#ifndef SE3ET_ROWS_NOSTSM
          stmatrix_x4(st_base + i * (16 * 128), pack2(d[0][0], d[0][1]), pack2(d[0][2], d[0][3]), pack2(d[1][0], d[1][1]),
                      pack2(d[1][2], d[1][3]));
#else
          if (d[0][0] == 1.2345f && d[1][1] == 3.f && d[0][3] == 7.f) st_shared_b32(st_base, pack2(d[0][2], d[1][0]));
#endif
          __syncwarp();  // every lane is done with this ring stage
        }
        tc::fence_proxy_async_smem();  // generic-proxy stores -> visible to the async proxy (tensor core, bulk copies)
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(CL > 1 ? &p_done[buf] : &b_full[buf]);
      }
      }
        cp_async_wait<0>();
        if (TP > 0) tc::tmem_ld_wait();   // the last prefetched fragments: the stash is rewritten by the next tile's setup
        __syncwarp();
      }
      // the next tile's basis weights are built under this tile's last MMAs
      if (have_next) setup(tile + tile_stride);
      if (have) {
        // ---- epilogue: TMEM lane quadrant warp % 4 (rows = points), column group warp / 4 -------------------
        tc::mbar_wait_long(tmem_full, (uint32_t)titer & 1u);
        tc::tcgen05_fence_after_sync();
        const int row = (warp & 3) * 32 + lane;
        const int64_t p = tile * kTilePts + row;
        const bool row_ok = row < kTilePts && p < args.nq;
        constexpr int kColsPerWarp = kAccCols / 4;
        const int c_begin = (warp >> 2) * kColsPerWarp;
#pragma unroll 1
        for (int c0 = c_begin; c0 < c_begin + kColsPerWarp; c0 += 16) {
          uint32_t rr[16];
          tc::tmem_ld_32x32b_x16(tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)c0, rr);
          tc::tmem_ld_wait();
          if (row_ok) {
            const int r = c0 / BN, dcol = c0 - r * BN;
            const int64_t o = (p * kA + r) * args.cout + n0 + dcol;
            if (args.out_bf16) {
              uint32_t pk[8];
#pragma unroll
              for (int jj = 0; jj < 8; ++jj)
                pk[jj] = kpm::pack2(__uint_as_float(rr[2 * jj]), __uint_as_float(rr[2 * jj + 1]));
              uint4* dst = reinterpret_cast<uint4*>(args.out_bf16 + o);
              dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            } else {
              float4* dst = reinterpret_cast<float4*>(args.out + o);
#pragma unroll
              for (int jj = 0; jj < 4; ++jj)
                dst[jj] = make_float4(__uint_as_float(rr[4 * jj]), __uint_as_float(rr[4 * jj + 1]),
                                      __uint_as_float(rr[4 * jj + 2]), __uint_as_float(rr[4 * jj + 3]));
            }
          }
        }
        tc::tcgen05_fence_before_sync();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(tmem_empty);
      }
      if (!have_next) break;
    }
  } else if (warp == kProdWarps) {
    // =========================================== MMA issuer ==================================================
    if (lane == 0) {
      constexpr uint32_t idesc = tc::umma_idesc_bf16(128, BN);
      uint32_t gstep = 0, titer = 0, ws = 0, wphase = 0;
      const uint32_t b_s = tc::smem_u32(b_tile), w_s = tc::smem_u32(w_tile);
      for (int64_t tile = tile0; WM > 1 ? (int64_t)titer < nrounds : tile < ntiles; tile += tile_stride, ++titer) {
        tc::mbar_wait_long(tmem_empty, (titer & 1) ^ 1);  // the epilogue has drained the previous tile's accumulators
        tc::tcgen05_fence_after_sync();
        for (int step = 0; step < nsteps; ++step, ++gstep) {
          const uint32_t buf = gstep & 1u;
          tc::mbar_wait_long(&b_full[buf], (gstep >> 1) & 1u);
          tc::tcgen05_fence_after_sync();
          const uint32_t bb = b_s + buf * kBBuf;
#pragma unroll
          for (int kb = 0; kb < 9; ++kb) {
            tc::mbar_wait_long(&w_full[ws], wphase);
            tc::tcgen05_fence_after_sync();
            const uint64_t b_desc = tc::umma_desc_sw128(w_s + ws * S::kWStage);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int t = kb * 4 + q;          // r * 6 + kc
              const int beta = kBeta.v[t];
              const int r = t / 6, kc = t - r * 6;
              const uint64_t a_desc = tc::umma_desc_sw128(bb + (beta >> 2) * kSubTile) + (uint64_t)((beta & 3) * 2);
#ifndef SE3ET_ROWS_NOMMA
              tc::umma_bf16(tmem_base + (uint32_t)(r * BN), a_desc, b_desc + (uint64_t)(q * 2), idesc,
                            (step | kc) != 0);
#endif
            }
            if (WM > 1) umma_commit_multicast(&w_empty[ws], (uint16_t)((1u << WM) - 1u));
            else tc::umma_commit(&w_empty[ws]);
            if (++ws == (uint32_t)wstages) { ws = 0; wphase ^= 1; }
          }
          // every CTA of the cluster writes rows into this buffer: all of them learn that it is free again
          if (CL > 1) umma_commit_multicast(&b_empty[buf], (uint16_t)((1u << CL) - 1u));
          else tc::umma_commit(&b_empty[buf]);
        }
        tc::umma_commit(tmem_full);
      }
    }
  } else if (warp == kProdWarps + 1) {
    // =========================================== weight TMA ==================================================
    if (lane == 0) {
      uint32_t ws = 0, wphase = 0;
      int64_t round = 0;
      for (int64_t tile = tile0; WM > 1 ? round < nrounds : tile < ntiles; tile += tile_stride, ++round) {
        for (int kbg = 0; kbg < nsteps * 9; ++kbg) {
          tc::mbar_wait_long(&w_empty[ws], wphase ^ 1);     // WM > 1: every CTA of the cluster has consumed the stage
          tc::mbar_arrive_expect_tx(&w_full[ws], BN * 128);
          if (WM > 1) {
            constexpr int kSliceRows = BN / WM;
            tma_load_2d_multicast(w_tile + ws * S::kWStage + wrank * (kSliceRows * 128), &tma_w, &w_full[ws], kbg * 64,
                                  n0 + (int)wrank * kSliceRows, (uint16_t)((1u << WM) - 1u));
          } else {
            tc::tma_load_2d(w_tile + ws * S::kWStage, &tma_w, &w_full[ws], kbg * 64, n0);
          }
          if (++ws == (uint32_t)wstages) { ws = 0; wphase ^= 1; }
        }
      }
    }
  } else if (CL > 1) {
    // ============================== operand exchange (CL > 1): this CTA's rows -> the peers ==========================
    if (lane == 0) {
      constexpr uint32_t kBlk = PPW * 16 * 128;                 // this CTA's rows of one sub-tile
      const uint32_t b_s = tc::smem_u32(b_tile) + rank * kBlk;
      uint32_t gstep = 0;
      for (int64_t tile = tile0; tile < ntiles; tile += tile_stride) {
        for (int step = 0; step < nsteps; ++step, ++gstep) {
          const uint32_t buf = gstep & 1u;
          tc::mbar_wait_long(&p_done[buf], (gstep >> 1) & 1u);  // the 16 producer warps stored (and fenced) their rows
          // the one arrival of this CTA's b_full phase, expecting the peers' rows
          tc::mbar_arrive_expect_tx(&b_full[buf], (CL - 1) * 4 * kBlk);
          const uint32_t bar = tc::smem_u32(&b_full[buf]);
#pragma unroll
          for (int d = 1; d < CL; ++d) {
            const uint32_t peer = (rank + d) % CL;
            const uint32_t bar_peer = map_to_cta(bar, peer);
#pragma unroll
            for (int sub = 0; sub < 4; ++sub) {
              const uint32_t src = b_s + buf * kBBuf + sub * kSubTile;
              dsmem_bulk_copy(map_to_cta(src, peer), src, kBlk, bar_peer);
            }
          }
        }
      }
    }
  }
  tc::tcgen05_fence_before_sync();
  if (kCluster) cluster_sync_all(); else __syncthreads();   // no CTA leaves while a peer may still write to it
  if (warp == kProdWarps) tc::tmem_dealloc<512>(tmem_base);
}

}  // namespace rows

int make_tmap_bf16_2d(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows);  // gemm.cu

namespace rows {

template <int BN, int KH, int PPW, int TP, int CL, int WM>
static int launch(const void* w, Args args, cudaStream_t st) {
  using S = Smem<BN, KH>;
  args.wstages = S::max_wstages();
  if (args.wstages < 2) return SE3ET_ERR_UNSUPPORTED;
  const int smem = S::total(args.wstages);
  CUtensorMap tw;
  int rc = make_tmap_bf16_2d(&tw, w, args.cout, 216 * (int64_t)args.cin, 216 * (int64_t)args.cin, BN / WM);
  if (rc) return rc;
  auto kernel = kpconv_rows_kernel<BN, KH, PPW, TP, CL, WM>;
  SE3ET_ENSURE_SMEM(kernel, smem);
  const int64_t ntiles = ceil_div(args.nq, 16 * PPW * CL);
  if (CL == 1 && WM == 1) {
    dim3 grid((unsigned)(ntiles < kNumSMs ? ntiles : kNumSMs), (unsigned)(args.cout / BN));
    kernel<<<grid, kThreads + 32, smem, st>>>(tw, args);
    SE3ET_LAUNCH_CHECK();
    return SE3ET_OK;
  }
  constexpr int kC = CL * WM;   // cluster size (one of the two is 1)
  cudaLaunchConfig_t cfg = {};
  if (CL > 1) {
    // one cluster of CL CTAs per tile of 128 points; CTA rank = column block of BN outputs (cout == CL * BN)
    const int64_t max_clusters = kNumSMs / CL;
    cfg.gridDim = dim3((unsigned)((ntiles < max_clusters ? ntiles : max_clusters) * CL), 1, 1);
  } else {
    // clusters of WM CTAs on consecutive tiles, same column block (blockIdx.y)
    const int64_t want = ceil_div(ntiles, WM) * WM, cap = kNumSMs / WM * WM;
    cfg.gridDim = dim3((unsigned)(want < cap ? want : cap), (unsigned)(args.cout / BN), 1);
  }
  cfg.blockDim = dim3(kThreads + 32, 1, 1);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kC;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SE3ET_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kernel, tw, args));
  return SE3ET_OK;
}

}  // namespace rows
}  // namespace se3et

using namespace se3et;

extern "C" int se3et_kpconv_rows_layout(int32_t* src_slot_6x36, int32_t* flip_36) {
  for (int a = 0; a < kA; ++a)
    for (int r = 0; r < kA; ++r)
      for (int kc = 0; kc < kKC; ++kc) src_slot_6x36[a * 36 + r * 6 + kc] = kc * 6 + ridx_tab(a, r);
  for (int r = 0; r < kA; ++r)
    for (int kc = 0; kc < kKC; ++kc) flip_36[r * 6 + kc] = (basis_row(r, kc) >> 2) & 1;
  return SE3ET_OK;
}

extern "C" int se3et_kpconv_rows(const float* q_pts, const float* s_pts, const int64_t* neighbors, int64_t nq,
                                 int64_t ns, int64_t h, const void* x_bf16, int64_t cin, const void* w_rows_bf16,
                                 int64_t cout, const float* kernel_points_15x3, float kp_extent, void* out,
                                 int out_bf16, se3et_stream_t stream) {
  if (nq < 0 || ns <= 0 || h <= 0 || cin <= 0 || cout <= 0 || !(kp_extent > 0.f)) return SE3ET_ERR_ARG;
  if (h > 48 || cin % kpm::kChunk != 0 || cout % 32 != 0) return SE3ET_ERR_UNSUPPORTED;
  if (ns * kA * cin / 8 >= ((int64_t)1 << 32)) return SE3ET_ERR_UNSUPPORTED;  // 32-bit gather offsets (16-byte units)
  if (!q_pts || !s_pts || !neighbors || !x_bf16 || !w_rows_bf16 || !kernel_points_15x3 || !out) return SE3ET_ERR_ARG;
  if (nq == 0) return SE3ET_OK;
  const int bn = cout % 64 == 0 ? 64 : 32;
  const int kh = h <= 32 ? 4 : (h <= 40 ? 5 : 6);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  rows::Args a;
  a.q_pts = q_pts; a.s_pts = s_pts; a.idx = neighbors; a.x = static_cast<const __nv_bfloat16*>(x_bf16);
  a.kernel_points = kernel_points_15x3; a.nq = nq;
  a.out = out_bf16 ? nullptr : static_cast<float*>(out);
  a.out_bf16 = out_bf16 ? static_cast<__nv_bfloat16*>(out) : nullptr;
  a.ns = ns; a.H = (int)h;
  a.cin = (int)cin; a.cout = (int)cout; a.inv_extent = 1.f / kp_extent; a.wstages = 0;
  const void* w = w_rows_bf16;
#ifndef SE3ET_ROWS_WM
#define SE3ET_ROWS_WM 1   // measured on B200: 2 and 4 are 2x SLOWER (lockstep on a 4-stage ring: any skew between the CTAs stalls all)
#endif
  constexpr int WM = SE3ET_ROWS_WM;   // weight-multicast cluster size of the 64-column kernels
  // cout 128 / 256: clusters of 2 / 4 CTAs share a tile of 128 points (each CTA 64 output columns and 1 / CL of the
  // operand rows, exchanged through distributed shared memory).  Other widths: one CTA per (tile, column block); the
  // 64-column kernels run in clusters of WM CTAs that share the weight stream by TMA multicast.
#ifdef SE3ET_ROWS_DEV
  if (kh != 5) return SE3ET_ERR_UNSUPPORTED;
  if (cout == 128) return rows::launch<64, 5, 4, 3, 2, 1>(w, a, st);
  if (cout == 256) return rows::launch<64, 5, 2, 2, 4, 1>(w, a, st);
  return bn == 32 ? rows::launch<32, 5, 8, 8, 1, 1>(w, a, st) : rows::launch<64, 5, 6, 3, 1, WM>(w, a, st);
#else
  if (cout == 128 || cout == 256) {
    if (cout == 128) {
      switch (kh) {
        case 4: return rows::launch<64, 4, 4, 4, 2, 1>(w, a, st);
        case 5: return rows::launch<64, 5, 4, 3, 2, 1>(w, a, st);
        default: return rows::launch<64, 6, 4, 2, 2, 1>(w, a, st);
      }
    }
    switch (kh) {
      case 4: return rows::launch<64, 4, 2, 2, 4, 1>(w, a, st);
      case 5: return rows::launch<64, 5, 2, 2, 4, 1>(w, a, st);
      default: return rows::launch<64, 6, 2, 2, 4, 1>(w, a, st);
    }
  }
  if (bn == 32) {
    switch (kh) {
      case 4: return rows::launch<32, 4, 8, 8, 1, 1>(w, a, st);
      case 5: return rows::launch<32, 5, 8, 8, 1, 1>(w, a, st);
      default: return rows::launch<32, 6, 8, 6, 1, 1>(w, a, st);
    }
  }
  switch (kh) {
    case 4: return rows::launch<64, 4, 6, 4, 1, WM>(w, a, st);
    case 5: return rows::launch<64, 5, 6, 3, 1, WM>(w, a, st);
    default: return rows::launch<64, 6, 6, 2, 1, WM>(w, a, st);
  }
#endif
}
