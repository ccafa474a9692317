// Fused KPConvInterSO3.forward (blocks_epn.py:454-546 with feat_gather_by_perm :334-390): neighbour gather, kernel
// point influence, rotate-by-permute and the (kernel points x anchors x Cin) -> Cout contraction in ONE kernel; the
// gathered operand A' (kpconv_tables.cuh) lives only in shared memory.
//
// Persistent CTA, tile = 16 query points = 96 operand rows (p, r) of a UMMA M = 128 tile:
//   warps 0-15 producers, one point each.  Per point: the 16 x H basis weight matrix W16 (bf16 A fragments kept in
//              registers for the whole tile).  Per (16-channel chunk, input anchor): cp.async gather of the
//              neighbour rows x[idx[n]][a][chunk] (32 B sectors, 3-stage ring per warp), mma.sync m16n8k16
//              W16 . X, and every accumulator element is stored once (twice .. six times for its (r, kc) copies) as
//              bf16 into the 128-byte-swizzled K-major operand tile the tensor core reads.
//   warp 16    tcgen05.mma issuer: per chunk 9 K-blocks of 64 (36 (kc, a') slots x 16 channels) against the weight
//              K-blocks, fp32 accumulation in TMEM across all chunks of the tile.
//   warp 17    TMA producer of the weight K-blocks (3-stage mbarrier ring).
//   warps 0-3  epilogue after their production: tcgen05.ld, fp32 rows to global memory, GroupNorm statistics.
// The operand tile is single-buffered (9 x 12 KB): the MMA of chunk i and the production of chunk i + 1 alternate,
// while gathers and the next tile's weights run ahead.
#include <cuda_bf16.h>

#include "common.cuh"
#include "gn_epilogue.cuh"
#include "kpconv_mma.cuh"
#include "tc.cuh"

namespace se3et {

using namespace kpm;

constexpr int kFPts = 16;                     // points per tile
constexpr int kFRows = kFPts * kA;            // 96 operand rows
constexpr int kFProdWarps = 16;               // one point each
constexpr int kFThreads = (kFProdWarps + 2) * 32;
constexpr int kFKBlocks = 9;                  // 36 slots x 16 channels = 576 = 9 x 64
constexpr int kFKBlockBytes = kFRows * 128;   // 12288, a multiple of the 1024-byte swizzle period
constexpr int kFStages = 3;                   // gather ring per warp, item = (chunk, anchor)
constexpr int kFXRow = 32;                    // bytes per gathered row: 16 channels bf16; the two 16-byte halves of row n
                                              // are swapped when (n / 4) is odd -> conflict-free ldmatrix
constexpr int kFWMaxStages = 18;             // weight K-block ring: as many stages as fit (two chunks at most)
constexpr int kFKS = 3;                       // neighbour k-steps of 16 (H <= 48)
constexpr int kFW16Row = (kFKS * 16 + 8) * 2; // 112 bytes (one neighbour half per ring stage)
// Neighbour columns beyond 48 (KITTI-calibrated limits, utils/data.py:212-252) are processed as NH = 2 halves of <= 48:
// a ring stage holds one half, the products of both halves accumulate in registers before the operand stores, and
// the second half's basis-weight fragments are parked in tensor-memory columns the accumulators leave free.
__host__ __device__ constexpr int w16_pitch(int nh, int hr) { return nh == 1 ? kFW16Row : (2 * hr + 8) * 2; }

__constant__ int8_t c_f_basis_target[16][6] = {
#define SE3ET_BT(row) {(int8_t)basis_target(row, 0), (int8_t)basis_target(row, 1), (int8_t)basis_target(row, 2), \
                       (int8_t)basis_target(row, 3), (int8_t)basis_target(row, 4), (int8_t)basis_target(row, 5)}
    SE3ET_BT(0), SE3ET_BT(1), SE3ET_BT(2), SE3ET_BT(3), SE3ET_BT(4), SE3ET_BT(5), SE3ET_BT(6), SE3ET_BT(7),
    SE3ET_BT(8), SE3ET_BT(9), SE3ET_BT(10), SE3ET_BT(11), SE3ET_BT(12), SE3ET_BT(13), SE3ET_BT(14), SE3ET_BT(15)
#undef SE3ET_BT
};
__constant__ uint32_t c_f_ridx_cols[6] = {ridx_col_packed(0), ridx_col_packed(1), ridx_col_packed(2),
                                          ridx_col_packed(3), ridx_col_packed(4), ridx_col_packed(5)};

struct FusedArgs {
  const float* q_pts;
  const float* s_pts;
  const int64_t* idx;
  const __nv_bfloat16* x;
  const float* kernel_points;
  float* out;                 // fp32 [nq * 6, cout]
  __nv_bfloat16* out_bf16;    // or bf16, same shape (exactly one of the two is set)
  int64_t nq, ns;
  int H, HR;                  // neighbour columns; rows per ring stage (one half, rounded up to 8)
  int cin, cout;
  int wstages;                // weight ring stages (<= kFWMaxStages)
  float inv_extent;
  // GroupNorm statistics of the output (optional)
  double* gn_stats;
  const int64_t* gn_seg_off;
  int gn_nseg, gn_cpg, gn_groups;
};

template <int BN>
struct FusedSmem {
  static constexpr int kAOff = 0;
  static constexpr int kABytes = kFKBlocks * kFKBlockBytes;          // 110592; the last K-block's unused rows
                                                                      // 96..127 alias the first 4 KB after it
  static constexpr int kBarOff = kAOff + kABytes + 4096;              // mbarriers + tmem pointer (512 B)
  static constexpr int kZeroOff = kBarOff + 512;                      // 16 zero bytes
  static constexpr int kGnOff = kZeroOff + 16;                        // 4 warps x 128 floats
  static constexpr int kXOff = kGnOff + 4 * 128 * 4;                  // gather rings [warps][stages][HR][48]
  static constexpr int kWStage = (BN * 128 + 1023) / 1024 * 1024;
  __host__ __device__ static int x_bytes(int hr) { return kFProdWarps * kFStages * hr * kFXRow; }
  __host__ __device__ static int w_off(int hr) { return (kXOff + x_bytes(hr) + 1023) / 1024 * 1024; }
  static int total(int hr, int wstages) { return w_off(hr) + wstages * kWStage + 1024; }
  static int max_wstages(int hr) {
    const int n = (227 * 1024 - 1024 - w_off(hr)) / kWStage;
    return n > kFWMaxStages ? kFWMaxStages : n;
  }
};

template <int BN, int NH>
__global__ void __launch_bounds__(kFThreads, 1)
kpconv_fused_kernel(const __grid_constant__ CUtensorMap tma_w, FusedArgs args) {
  using S = FusedSmem<BN>;
  constexpr uint32_t kAccCols = BN < 32 ? 32 : BN;
  // NH = 2: 12 stash columns per producer warp (4 warps per lane quadrant) after the accumulators
  constexpr uint32_t kTmemCols = NH == 1 ? kAccCols : (kAccCols + 48 <= 64 ? 64 : (kAccCols + 48 <= 128 ? 128 : 256));
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_tile = smem + S::kAOff;
  uint8_t* w_tile = smem + S::w_off(args.HR);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + S::kBarOff);
  uint64_t* a_empty = a_full + 1;
  uint64_t* tmem_full = a_full + 2;
  uint64_t* tmem_empty = a_full + 3;
  uint64_t* w_full = a_full + 4;
  uint64_t* w_empty = w_full + kFWMaxStages;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(w_empty + kFWMaxStages);
  const int wstages = args.wstages;
  uint8_t* zero16 = smem + S::kZeroOff;
  float* gn_acc = reinterpret_cast<float*>(smem + S::kGnOff);
  __shared__ float sh_kp[48];
  __shared__ int8_t sh_target[16][6];
  __shared__ uint32_t sh_ridx[6];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * BN;
  const int64_t ntiles = (args.nq + kFPts - 1) / kFPts;
  const int nchunks = args.cin / kChunk;
  const int xstage = args.HR * kFXRow;

  if (threadIdx.x < 45) sh_kp[threadIdx.x] = args.kernel_points[threadIdx.x];
  if (threadIdx.x < 96) sh_target[threadIdx.x / 6][threadIdx.x % 6] = c_f_basis_target[threadIdx.x / 6][threadIdx.x % 6];
  if (threadIdx.x < 6) sh_ridx[threadIdx.x] = c_f_ridx_cols[threadIdx.x];
  if (threadIdx.x < 4) reinterpret_cast<uint32_t*>(zero16)[threadIdx.x] = 0u;
  if (threadIdx.x == 0) {
    tc::tma_prefetch_desc(&tma_w);
    tc::mbar_init(a_full, kFProdWarps);
    tc::mbar_init(a_empty, 1);
    for (int s = 0; s < wstages; ++s) {
      tc::mbar_init(&w_full[s], 1);
      tc::mbar_init(&w_empty[s], 1);
    }
    tc::mbar_init(tmem_full, 1);
    tc::mbar_init(tmem_empty, 4);
    tc::mbar_fence_init();
  }
  if (warp < kFProdWarps) {
    // gather rows H..HR-1 are never written: zero the whole ring once
    uint8_t* xs = smem + S::kXOff + warp * kFStages * xstage;
    for (int i = lane; i < kFStages * xstage / 16; i += 32) reinterpret_cast<uint4*>(xs)[i] = make_uint4(0, 0, 0, 0);
  }
  if (warp == kFProdWarps) tc::tmem_alloc<kTmemCols>(tmem_ptr);
  tc::tcgen05_fence_before_sync();
  __syncthreads();
  tc::tcgen05_fence_after_sync();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp < kFProdWarps) {
    // =========================================== producers ===================================================
    uint8_t* xs = smem + S::kXOff + warp * kFStages * xstage;
    const uint32_t xs_s = smem_addr(xs);
    const uint32_t a_tile_s = smem_addr(a_tile);
    const uint32_t zero_s = smem_addr(zero16);
    // stmatrix address rows of this lane.  x4: matrix lane / 8 = (basis rows 0-7 | 8-15) x (channels 0-7 | 8-15), row
    // lane % 8; each basis row has two (r, kc) targets.  x2 (centre copies): matrices = basis rows 8-15 x both channel
    // halves, addresses from lanes 0-15.
    const int ar4 = (lane & 7) + 8 * ((lane >> 3) & 1), nt4 = lane >> 4;
    uint32_t r4[2], kc4[2], rw4[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int v = sh_target[ar4][t];
      r4[t] = (uint32_t)(v >> 3);
      kc4[t] = (uint32_t)(v & 7) * kA;
      rw4[t] = sh_ridx[v >> 3];
    }
    const bool centre = (lane >> 2) == 7;  // accumulator row g + 8 == 15
    const int q = lane & 3;
    const int H = args.H, HR = args.HR;
    const int cin = args.cin;
    // ldmatrix row of this lane inside a k-step and its 16-byte half
    const int ld_row = (lane & 7) + ((lane >> 3) & 1) * 8, ld_half = lane >> 4;
    // this lane's three gather pieces of an item: neighbour n = i / 2, 16-byte half i % 2, i = lane + 32 u
    uint32_t gdst[3];
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      const int i = lane + 32 * u;
      const int n = i >> 1;
      gdst[u] = xs_s + n * kFXRow + (((i & 1) ^ ((n >> 2) & 1)) << 4);
    }
    // per-tile state (registers): A fragments of this warp's point's basis weights, gather source pointers
    uint32_t afrag[kFKS][4];
    const __nv_bfloat16* gsrc[NH][3];
    uint32_t gvalid = 0;  // bit 3 * half + u: the piece reads a real neighbour (else zero fill)
    const uint32_t stash = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + kAccCols + (uint32_t)((warp >> 2) * 12);
    const int wpitch = w16_pitch(NH, HR);

    auto setup = [&](int64_t tile) {
      const int64_t p = tile * kFPts + warp;
      const bool pvalid = p < args.nq;
      const int64_t pc = pvalid ? p : 0;
      const float qx = args.q_pts[3 * pc], qy = args.q_pts[3 * pc + 1], qz = args.q_pts[3 * pc + 2];
      gvalid = 0;
#pragma unroll
      for (int hf = 0; hf < NH; ++hf)
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          const int i = lane + 32 * u, n = hf * HR + (i >> 1);
          int64_t j = (pvalid && (i >> 1) < HR && n < H) ? args.idx[pc * H + n] : -1;
          const bool valid = j >= 0 && j < args.ns;
          gsrc[hf][u] = args.x + (valid ? j : 0) * (int64_t)(kA * cin) + (i & 1) * 8;
          gvalid |= (valid ? 1u : 0u) << (3 * hf + u);
        }
      // W16 scratch aliases the head of the gather ring: [16][wpitch]
      for (int i = lane; i < 16 * wpitch / 16; i += 32) reinterpret_cast<uint4*>(xs)[i] = make_uint4(0, 0, 0, 0);
      __syncwarp();
#pragma unroll 1
      for (int n = lane; n < H; n += 32) {
        int64_t j = pvalid ? args.idx[pc * H + n] : -1;
        const bool valid = j >= 0 && j < args.ns;
        if (valid) {  // shadow / padding neighbours keep their zero weights
          float row[16];
          basis_weights(args.s_pts[3 * j] - qx, args.s_pts[3 * j + 1] - qy, args.s_pts[3 * j + 2] - qz, sh_kp,
                        args.inv_extent, true, row);
          // column of neighbour n: half n / HR, position n % HR inside it
          const int col = NH == 1 ? n : (n >= HR ? HR + (n - HR) : n);
          uint8_t* dst = xs + col * 2;
#pragma unroll
          for (int r = 0; r < 16; ++r) *reinterpret_cast<__nv_bfloat16*>(dst + r * wpitch) = __float2bfloat16(row[r]);
        }
      }
      __syncwarp();
      if (NH == 2) {  // second half first: its fragments go to tensor memory
        uint32_t t4[kFKS][4];
#pragma unroll
        for (int ks = 0; ks < kFKS; ++ks) {
          if (ks * 16 < HR) ldmatrix_x4(t4[ks], xs_s + ld_row * wpitch + (HR + ks * 16 + ld_half * 8) * 2);
          else { t4[ks][0] = t4[ks][1] = t4[ks][2] = t4[ks][3] = 0u; }
          asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stash + ks * 4), "r"(t4[ks][0]),
                       "r"(t4[ks][1]), "r"(t4[ks][2]), "r"(t4[ks][3]) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
#pragma unroll
      for (int ks = 0; ks < kFKS; ++ks) {
        if (NH == 1 || ks * 16 < HR) ldmatrix_x4(afrag[ks], xs_s + ld_row * wpitch + (ks * 16 + ld_half * 8) * 2);
        else { afrag[ks][0] = afrag[ks][1] = afrag[ks][2] = afrag[ks][3] = 0u; }
      }
      __syncwarp();
      // gather rows >= H must read as zero again
      for (int i = lane; i < 16 * wpitch / 16; i += 32) reinterpret_cast<uint4*>(xs)[i] = make_uint4(0, 0, 0, 0);
      __syncwarp();
    };

    // item = (chunk, input anchor a, neighbour half): idx = a * NH + half inside a chunk, ring stage idx % 3
    auto issue = [&](int chunk, int idx, int stage) {
      const int a = idx / NH, hf = idx % NH;
      const int off = a * cin + chunk * kChunk;
#pragma unroll
      for (int u = 0; u < 3; ++u) {
        if (lane + 32 * u < 2 * HR)
          cp_async_16(gdst[u] + stage * xstage, gsrc[hf][u] + off, (gvalid >> (3 * hf + u)) & 1u ? 16 : 0);
      }
    };
    // the ring runs two items ahead (6 * NH items per chunk, a multiple of the three stages)
    auto prologue = [&]() {
      issue(0, 0, 0);
      cp_async_commit();
      issue(0, 1, 1);
      cp_async_commit();
    };

    uint32_t gc = 0;       // chunks produced so far (all tiles): parity of the operand-tile barriers
    uint32_t titer = 0;    // tiles done by this CTA
    if ((int64_t)blockIdx.x < ntiles) {
      setup(blockIdx.x);
      prologue();
    }
    const uint32_t mbase = warp * kA;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++titer) {
      for (int chunk = 0; chunk < nchunks; ++chunk) {
#pragma unroll
        for (int a = 0; a < kA; ++a) {
          float d[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
          for (int hf = 0; hf < NH; ++hf) {
            constexpr int kItems = kA * NH;
            const int idx = a * NH + hf;
            {  // prefetch the item two ahead (possibly of the next chunk)
              const int idx2 = (idx + 2) % kItems;
              const int chunk2 = chunk + (idx + 2) / kItems;
              if (chunk2 < nchunks) issue(chunk2, idx2, idx2 % kFStages);
              cp_async_commit();
            }
            uint32_t bfrag[kFKS][4];
            if (hf == 1) {
#pragma unroll
              for (int ks = 0; ks < kFKS; ++ks)
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(bfrag[ks][0]), "=r"(bfrag[ks][1]), "=r"(bfrag[ks][2]), "=r"(bfrag[ks][3])
                             : "r"(stash + ks * 4)
                             : "memory");
            }
            cp_async_wait<2>();
            __syncwarp();
            if (idx == 0) {
              tc::mbar_wait_long(a_empty, (gc & 1) ^ 1);  // the MMA of the previous chunk has consumed the operand tile
            }
            if (hf == 1) tc::tmem_ld_wait();
            const uint32_t xsb = xs_s + (idx % kFStages) * xstage;
#pragma unroll
            for (int ks = 0; ks < kFKS; ++ks) {
              const int n = ks * 16 + ld_row;
              uint32_t b[4];
              ldmatrix_x4_trans(b, n < HR ? xsb + n * kFXRow + ((ld_half ^ ((n >> 2) & 1)) << 4) : zero_s);
              if (hf == 0) {
                mma_16816(d[0], afrag[ks], b[0], b[1]);
                mma_16816(d[1], afrag[ks], b[2], b[3]);
              } else {
                mma_16816(d[0], bfrag[ks], b[0], b[1]);
                mma_16816(d[1], bfrag[ks], b[2], b[3]);
              }
            }
            if (hf + 1 < NH) __syncwarp();  // every lane is done with this ring stage
          }
          // accumulator rows g / g + 8 = basis rows; each is copied to its (r, kc) targets: operand row
          // m = warp * 6 + r, slot j = kc * 6 + ridx[a][r], K-block j / 4, 16-byte chunk (j % 4) * 2 + nt swizzled by
          // (m % 8).  One stmatrix.x4 writes all 16 basis rows x 16 channels to their t-th targets.
          const uint32_t p0 = pack2(d[0][0], d[0][1]), p1 = pack2(d[0][2], d[0][3]);
          const uint32_t p2 = pack2(d[1][0], d[1][1]), p3 = pack2(d[1][2], d[1][3]);
#pragma unroll
          for (int t = 0; t < 2; ++t) {
            const uint32_t j = kc4[t] + ((rw4[t] >> (3 * a)) & 7u);
            const uint32_t m = mbase + r4[t];
            stmatrix_x4(a_tile_s + (j >> 2) * kFKBlockBytes + m * 128 + ((((j & 3) * 2 + nt4) ^ (m & 7)) << 4), p0, p1,
                        p2, p3);
          }
          if (centre) {  // the centre row (basis row 15, held by lanes 28-31) has four more targets r = 2..5
#pragma unroll
            for (int r = 2; r < kA; ++r) {
              const uint32_t j = 5 * kA + ridx_tab(a, r);
              const uint32_t m = mbase + r;
              const uint32_t addr = a_tile_s + (j >> 2) * kFKBlockBytes + m * 128 + q * 4 +
                                    ((((j & 3) * 2) ^ (m & 7)) << 4);
              st_shared_b32(addr, p1);
              st_shared_b32(addr ^ 16u, p3);
            }
          }
          __syncwarp();  // every lane is done with this ring stage and its stores are issued
          if (a == kA - 1) {
            tc::fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async proxy
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(a_full);
            ++gc;
          }
        }
      }
      cp_async_wait<0>();
      __syncwarp();
      // the next tile's weights and first gathers run under this tile's last MMA and epilogue
      if (tile + gridDim.x < ntiles) {
        setup(tile + gridDim.x);
        prologue();
      }

      // ---- epilogue (warps 0-3 own TMEM lanes 32 w .. 32 w + 31 = operand rows) ------------------------------
      if (warp < 4) {
        tc::mbar_wait_long(tmem_full, titer & 1);
        tc::tcgen05_fence_after_sync();
        const int m = warp * 32 + lane;
        const int64_t grow = tile * kFRows + m;
        const bool row_ok = m < kFRows && grow < args.nq * kA;
        float* warp_acc = gn_acc + warp * 128;
        bool gn_uniform = false;
        int gn_seg = 0;
        double* gn_row_stats = nullptr;
        if (args.gn_stats) {
#pragma unroll
          for (int i = 0; i < 4; ++i) warp_acc[lane * 4 + i] = 0.f;
          const int64_t p_first = tile * kFPts;
          const int64_t p_last = min(p_first + kFPts, args.nq) - 1;
          gn_seg = segment_of(args.gn_seg_off, args.gn_nseg, p_first);
          gn_uniform = segment_of(args.gn_seg_off, args.gn_nseg, p_last) == gn_seg;
          if (!gn_uniform && row_ok)
            gn_row_stats = args.gn_stats +
                           (int64_t)segment_of(args.gn_seg_off, args.gn_nseg, grow / kA) * args.gn_groups * 2;
          __syncwarp();
        }
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
          uint32_t rr[32];
          tc::tmem_ld_32x32b_x32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(BN >= 32 ? c0 : 0), rr);
          tc::tmem_ld_wait();
          constexpr int kCols = BN >= 32 ? 32 : BN;
          float v[kCols];
#pragma unroll
          for (int jj = 0; jj < kCols; ++jj) v[jj] = __uint_as_float(rr[jj]);
          if (args.gn_stats) {
            const int cpg = args.gn_cpg;
            const int g_glob = (n0 + c0) / cpg, g_loc = g_glob - n0 / cpg;
            gn_accumulate_chunk<kCols>(cpg, v, row_ok, gn_uniform, lane, warp_acc, g_loc, gn_row_stats, g_glob);
          }
          if (row_ok) {
            if (args.out_bf16) {
              uint4* dst = reinterpret_cast<uint4*>(args.out_bf16 + grow * args.cout + n0 + c0);
#pragma unroll
              for (int jj = 0; jj < kCols / 8; ++jj)
                dst[jj] = make_uint4(kpm::pack2(v[8 * jj], v[8 * jj + 1]), kpm::pack2(v[8 * jj + 2], v[8 * jj + 3]),
                                     kpm::pack2(v[8 * jj + 4], v[8 * jj + 5]), kpm::pack2(v[8 * jj + 6], v[8 * jj + 7]));
            } else {
              float4* dst = reinterpret_cast<float4*>(args.out + grow * args.cout + n0 + c0);
#pragma unroll
              for (int jj = 0; jj < kCols / 4; ++jj)
                dst[jj] = make_float4(v[4 * jj], v[4 * jj + 1], v[4 * jj + 2], v[4 * jj + 3]);
            }
          }
        }
        tc::tcgen05_fence_before_sync();
        if (args.gn_stats) {
          asm volatile("bar.sync 1, 128;" ::: "memory");
          if (gn_uniform) {
            const int e = warp * 32 + lane;
            const int ngr2 = 2 * ((n0 + BN - 1) / args.gn_cpg - n0 / args.gn_cpg + 1);
            for (int i = e; i < ngr2; i += 128) {
              const float t = gn_acc[i] + gn_acc[128 + i] + gn_acc[256 + i] + gn_acc[384 + i];
              atomicAdd(args.gn_stats + ((int64_t)gn_seg * args.gn_groups + n0 / args.gn_cpg) * 2 + i, (double)t);
            }
          }
          asm volatile("bar.sync 1, 128;" ::: "memory");  // accumulators are re-zeroed by the next tile
        }
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(tmem_empty);
      }
    }
  } else if (warp == kFProdWarps) {
    // =========================================== MMA issuer ==================================================
    if (lane == 0) {
      constexpr uint32_t idesc = tc::umma_idesc_bf16(128, BN);
      uint32_t gc = 0, titer = 0, ws = 0, wphase = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++titer) {
        tc::mbar_wait_long(tmem_empty, (titer & 1) ^ 1);  // the epilogue has drained the previous tile's accumulators
        tc::tcgen05_fence_after_sync();
        for (int chunk = 0; chunk < nchunks; ++chunk, ++gc) {
          tc::mbar_wait_long(a_full, gc & 1);
          tc::tcgen05_fence_after_sync();
          for (int kb = 0; kb < kFKBlocks; ++kb) {
            tc::mbar_wait_long(&w_full[ws], wphase);
            tc::tcgen05_fence_after_sync();
            const uint64_t a_desc = tc::umma_desc_sw128(tc::smem_u32(a_tile + kb * kFKBlockBytes));
            const uint64_t b_desc = tc::umma_desc_sw128(tc::smem_u32(w_tile + ws * S::kWStage));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              tc::umma_bf16(tmem_base, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc,
                            (chunk | kb | k) != 0);
            tc::umma_commit(&w_empty[ws]);
            if (++ws == (uint32_t)wstages) { ws = 0; wphase ^= 1; }
          }
          tc::umma_commit(a_empty);
        }
        tc::umma_commit(tmem_full);
      }
    }
  } else {
    // =========================================== weight TMA ==================================================
    if (lane == 0) {
      uint32_t ws = 0, wphase = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        for (int chunk = 0; chunk < nchunks; ++chunk) {
          for (int kb = 0; kb < kFKBlocks; ++kb) {
            tc::mbar_wait_long(&w_empty[ws], wphase ^ 1);
            tc::mbar_arrive_expect_tx(&w_full[ws], BN * 128);
            tc::tma_load_2d(w_tile + ws * S::kWStage, &tma_w, &w_full[ws], (chunk * kFKBlocks + kb) * 64, n0);
            if (++ws == (uint32_t)wstages) { ws = 0; wphase ^= 1; }
          }
        }
      }
    }
  }
  tc::tcgen05_fence_before_sync();
  __syncthreads();
  if (warp == kFProdWarps) tc::tmem_dealloc<kTmemCols>(tmem_base);
}

int make_tmap_bf16_2d(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows);  // gemm.cu

template <int BN, int NH>
static int launch_fused(const CUtensorMap& tw, FusedArgs args, cudaStream_t st) {
  using S = FusedSmem<BN>;
  args.wstages = S::max_wstages(args.HR);
  if (args.wstages < 2) return SE3ET_ERR_UNSUPPORTED;
  const int smem = S::total(args.HR, args.wstages);
  SE3ET_ENSURE_SMEM((kpconv_fused_kernel<BN, NH>), smem);
  const int64_t ntiles = ceil_div(args.nq, kFPts);
  dim3 grid((unsigned)(ntiles < kNumSMs ? ntiles : kNumSMs), (unsigned)(args.cout / BN));
  kpconv_fused_kernel<BN, NH><<<grid, kFThreads, smem, st>>>(tw, args);
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}

}  // namespace se3et

using namespace se3et;

extern "C" int se3et_kpconv_fused(const float* q_pts, const float* s_pts, const int64_t* neighbors, int64_t nq,
                                  int64_t ns, int64_t h, const void* x_bf16, int64_t cin, const void* w_bf16,
                                  int64_t cout, const float* kernel_points_15x3, float kp_extent, void* out,
                                  int out_bf16, double* stats, const int64_t* seg_offsets, int64_t nseg,
                                  int64_t groups, se3et_stream_t stream) {
  if (nq < 0 || ns <= 0 || h <= 0 || cin <= 0 || cout <= 0 || !(kp_extent > 0.f)) return SE3ET_ERR_ARG;
  if (h > 2 * kFKS * 16 || cin % kChunk != 0 || cout % 16 != 0) return SE3ET_ERR_UNSUPPORTED;
  int bn = 0;
  for (int c : {128, 64, 32, 16})
    if (cout % c == 0) { bn = c; break; }
  if (!q_pts || !s_pts || !neighbors || !x_bf16 || !w_bf16 || !kernel_points_15x3 || !out) return SE3ET_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  FusedArgs a;
  a.q_pts = q_pts; a.s_pts = s_pts; a.idx = neighbors; a.x = static_cast<const __nv_bfloat16*>(x_bf16);
  a.kernel_points = kernel_points_15x3; a.nq = nq;
  a.out = out_bf16 ? nullptr : static_cast<float*>(out);
  a.out_bf16 = out_bf16 ? static_cast<__nv_bfloat16*>(out) : nullptr;
  a.ns = ns; a.H = (int)h;
  const bool halves = h > kFKS * 16;   // two neighbour halves per (chunk, anchor)
  a.HR = halves ? (((int)h + 1) / 2 + 7) / 8 * 8 : ((int)h + 7) / 8 * 8;
  if (a.HR < 16) a.HR = 16;  // the per-warp ring doubles as the W16 scratch (16 x 112 bytes)
  a.cin = (int)cin; a.cout = (int)cout; a.inv_extent = 1.f / kp_extent;
  a.gn_stats = nullptr; a.gn_seg_off = nullptr; a.gn_nseg = 0; a.gn_cpg = 1; a.gn_groups = 0;
  if (stats) {
    if (!seg_offsets || nseg <= 0 || groups <= 0 || cout % groups) return SE3ET_ERR_ARG;
    const int64_t cpg = cout / groups;
    const int chunk = bn >= 32 ? 32 : bn;
    const bool pow2 = (cpg & (cpg - 1)) == 0;
    if (!((pow2 && cpg <= chunk) || cpg % chunk == 0) || bn / cpg > 64) return SE3ET_ERR_UNSUPPORTED;
    SE3ET_CUDA_CHECK(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * nseg * groups, st));
    a.gn_stats = stats; a.gn_seg_off = seg_offsets; a.gn_nseg = (int)nseg; a.gn_cpg = (int)cpg; a.gn_groups = (int)groups;
  }
  if (nq == 0) return SE3ET_OK;
  CUtensorMap tw;
  int rc = make_tmap_bf16_2d(&tw, w_bf16, cout, 36 * cin, 36 * cin, bn);
  if (rc) return rc;
  if (halves) {
    switch (bn) {
      case 128: return launch_fused<128, 2>(tw, a, st);
      case 64: return launch_fused<64, 2>(tw, a, st);
      case 32: return launch_fused<32, 2>(tw, a, st);
      default: return launch_fused<16, 2>(tw, a, st);
    }
  }
  switch (bn) {
    case 128: return launch_fused<128, 1>(tw, a, st);
    case 64: return launch_fused<64, 1>(tw, a, st);
    case 32: return launch_fused<32, 1>(tw, a, st);
    default: return launch_fused<16, 1>(tw, a, st);
  }
}

// Diagnostics: attributes of the fused kernel instantiation for tile width bn (registers, static smem, max threads).
extern "C" int se3et_kpconv_fused_attrs(int bn, int* out5) {
  cudaFuncAttributes at;
  cudaError_t e;
  switch (bn) {
    case 128: e = cudaFuncGetAttributes(&at, kpconv_fused_kernel<128, 1>); break;
    case 64: e = cudaFuncGetAttributes(&at, kpconv_fused_kernel<64, 1>); break;
    case 32: e = cudaFuncGetAttributes(&at, kpconv_fused_kernel<32, 1>); break;
    default: e = cudaFuncGetAttributes(&at, kpconv_fused_kernel<16, 1>); break;
  }
  if (e != cudaSuccess) { set_last_error("cudaFuncGetAttributes", e); return SE3ET_ERR_CUDA; }
  out5[0] = at.numRegs; out5[1] = (int)at.sharedSizeBytes; out5[2] = at.maxThreadsPerBlock;
  out5[3] = (int)at.localSizeBytes; out5[4] = at.maxDynamicSharedSizeBytes;
  if (bn == 32) {  // occupancy probe for the common configuration (HR = 40)
    const int smem = FusedSmem<32>::total(40, FusedSmem<32>::max_wstages(40));
    cudaFuncSetAttribute(kpconv_fused_kernel<32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    int nb = -1;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kpconv_fused_kernel<32, 1>, kFThreads, smem);
    out5[3] = nb;
    out5[4] = smem;
    if (e != cudaSuccess) { set_last_error("occupancy", e); return SE3ET_ERR_CUDA; }
  }
  return SE3ET_OK;
}
