// Warp-level producer of the KPConvInterSO3 operand A' (see kpconv_tables.cuh) on the legacy tensor path:
// per query point one m16 x k(neighbours) weight matrix W16 (bf16) times the gathered neighbour features
// x[idx[n]][a][16-channel chunk] (bf16, staged in shared memory with cp.async), mma.sync m16n8k16, fp32 accumulate.
// Every accumulator element is one A' entry; a Store policy decides where it goes (global rows for the stand-alone
// gather kernel, the swizzled tcgen05 operand tile for the fused kernel).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"
#include "kpconv_tables.cuh"

namespace se3et {
namespace kpm {

constexpr int kChunk = 16;           // channels per gathered piece: 32 bytes = one DRAM/L2 sector
constexpr int kXRowBytes = 208;      // 6 anchors x 32 B + 16 B pad: ldmatrix rows 13 x 16 B apart -> conflict-free
constexpr int kPieces = kA * 2;      // 16-byte cp.async pieces per neighbour row and chunk

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// 16-byte async copy global -> shared; src_bytes = 0 writes zeros (shadow neighbours)
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void st_shared_b32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
// four / two 8x8 b16 matrices from mma accumulator-layout registers to shared-memory rows of 16 bytes; lanes 0-7,
// 8-15, 16-23, 24-31 give the row addresses of matrix 0, 1, 2, 3 (x2: lanes 0-7 and 8-15)
__device__ __forceinline__ void stmatrix_x4(uint32_t addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
  asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(r0), "r"(r1), "r"(r2),
               "r"(r3)
               : "memory");
}
__device__ __forceinline__ void stmatrix_x2(uint32_t addr, uint32_t r0, uint32_t r1) {
  asm volatile("stmatrix.sync.aligned.m8n8.x2.shared.b16 [%0], {%1, %2};" ::"r"(addr), "r"(r0), "r"(r1) : "memory");
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}

// Per-lane description of where its accumulator rows go: basis rows g = lane / 4 and g + 8, each copied to the
// (r, kc) pairs of basis_target(); `ridx` = column r of the anchor permutation table packed 3 bits per input anchor.
struct LaneTargets {
  uint32_t r[2][2], kc[2][2], ridx[2][2];  // [row half][target]
  bool centre;                             // this lane's upper row is the centre row 15: four more targets r = 2..5
};

__device__ __forceinline__ LaneTargets make_lane_targets(int lane, const int8_t (*target_tab)[6],
                                                         const uint32_t* ridx_cols) {
  LaneTargets t;
  const int g = lane >> 2;
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int v = target_tab[g + 8 * h][i];
      t.r[h][i] = (uint32_t)(v >> 3);
      t.kc[h][i] = (uint32_t)(v & 7);
      t.ridx[h][i] = ridx_cols[v >> 3];
    }
  t.centre = (g + 8) == 15;
  return t;
}

// Influence weights of one neighbour (blocks_epn.py:341-353, 'linear') folded into the 16 basis rows.
// d = s[j] - q (the neighbour in the query's frame); kp = the 15 kernel points (shared or constant memory).
__device__ __forceinline__ void basis_weights(float dx, float dy, float dz, const float* __restrict__ kp,
                                              float inv_extent, bool valid, float (&row)[16]) {
  float w[kKP];
#pragma unroll
  for (int k = 0; k < kKP; ++k) {
    const float ex = dx - kp[3 * k], ey = dy - kp[3 * k + 1], ez = dz - kp[3 * k + 2];
    w[k] = valid ? fmaxf(0.f, 1.f - sqrtf(ex * ex + ey * ey + ez * ez) * inv_extent) : 0.f;
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

}  // namespace kpm
}  // namespace se3et
