// Fused (flash-style) multi-head attention over variable-size clouds, bf16 tensor cores (mma.sync m16n8k16),
// fp32 softmax statistics and accumulation; scores and probabilities never leave the SM.
//
// One kernel serves both attention flavours of the SE3ET transformer:
//   * equivariant RPE self-attention (rpe_transformer.py:56-131): per anchor a and head h,
//       softmax_m( (q_a . k_a + bias[n, (a,h), m]) / sqrt(c) ) v_a,   bias = q . proj_p(embedding) precomputed
//   * invariant-q/k, equivariant-v cross attention (vanilla_transformer.py:58-85): q, k have no anchor axis
//       (anchor stride 0), the same softmax is applied to the 6 anchor slabs of v.
// A "problem" is one (query cloud, key cloud) pair; grid = (query tiles of 64, anchors * heads, problems).
// Each warp owns 16 query rows; K/V tiles of 64 keys are staged in shared memory.
#include <cuda_bf16.h>

#include "common.cuh"

namespace se3et {

struct AttnProblem {  // device table, one per blockIdx.z
  int64_t q_start, n_q, kv_start, n_kv, bias_off;
};

struct AttnTensors {
  const __nv_bfloat16 *q, *k, *v;
  int64_t q_pt, q_an, k_pt, k_an, v_pt, v_an;  // element strides per point / per anchor (0 = no anchor axis)
  const float* bias;                           // nullable; [bias_off + ((i*A + a)*H + h) * n_kv + m]
  __nv_bfloat16* out;                          // [(q_start + i)*A + a][h*D + c], row pitch ldo
  int64_t ldo;
  int A, H;
  float scale;
};

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, const void* smem_row) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row);
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}

constexpr int kAttnWarps = 4;
constexpr int kAttnBQ = 16 * kAttnWarps;  // 64 queries per CTA
constexpr int kAttnBK = 64;               // keys per tile

template <int D>
__global__ void __launch_bounds__(kAttnWarps * 32) flash_attention_kernel(AttnTensors t,
                                                                           const AttnProblem* __restrict__ problems) {
  constexpr int kLd = D + 8;  // padded smem row (elements): conflict-free fragment reads, 16-byte aligned rows
  __shared__ __align__(16) __nv_bfloat16 sk_buf[2][kAttnBK * kLd];  // double-buffered: tile i + 1 lands under tile i's math
  __shared__ __align__(16) __nv_bfloat16 sv_buf[2][kAttnBK * kLd];
  const AttnProblem pr = problems[blockIdx.z];
  const int q0 = blockIdx.x * kAttnBQ;
  if (q0 >= pr.n_q) return;
  const int a = blockIdx.y / t.H, h = blockIdx.y % t.H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tq = lane & 3;
  const int nq = (int)pr.n_q, nkv = (int)pr.n_kv;

  // ---- Q fragments (registers, loaded once)
  const int r_lo = q0 + warp * 16 + g, r_hi = r_lo + 8;
  const int rl = min(r_lo, nq - 1), rh = min(r_hi, nq - 1);
  const __nv_bfloat16* q_lo = t.q + (pr.q_start + rl) * t.q_pt + a * t.q_an + h * D;
  const __nv_bfloat16* q_hi = t.q + (pr.q_start + rh) * t.q_pt + a * t.q_an + h * D;
  uint32_t qa[D / 16][4];
#pragma unroll
  for (int ks = 0; ks < D / 16; ++ks) {
    qa[ks][0] = *reinterpret_cast<const uint32_t*>(q_lo + ks * 16 + 2 * tq);
    qa[ks][1] = *reinterpret_cast<const uint32_t*>(q_hi + ks * 16 + 2 * tq);
    qa[ks][2] = *reinterpret_cast<const uint32_t*>(q_lo + ks * 16 + 8 + 2 * tq);
    qa[ks][3] = *reinterpret_cast<const uint32_t*>(q_hi + ks * 16 + 8 + 2 * tq);
  }
  const float* bias_lo = nullptr;
  const float* bias_hi = nullptr;
  if (t.bias) {
    bias_lo = t.bias + pr.bias_off + (((int64_t)rl * t.A + a) * t.H + h) * nkv;
    bias_hi = t.bias + pr.bias_off + (((int64_t)rh * t.A + a) * t.H + h) * nkv;
  }

  float o[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;
  const float sc = t.scale * 1.4426950408889634f;  // scores are kept in log2 units

  // ---- K and V tiles: 64 rows x D bf16 as 16-byte cp.async chunks (keys past the end repeat the last one; masked below)
  auto stage = [&](int k0, int buf) {
    constexpr int kChunks = D / 8;
    for (int c = threadIdx.x; c < kAttnBK * kChunks; c += kAttnWarps * 32) {
      const int row = c / kChunks, ch = c - row * kChunks;
      const int key = min(k0 + row, nkv - 1);
      const uint32_t dk = (uint32_t)__cvta_generic_to_shared(sk_buf[buf] + row * kLd + ch * 8);
      const uint32_t dv = (uint32_t)__cvta_generic_to_shared(sv_buf[buf] + row * kLd + ch * 8);
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dk),
                   "l"(t.k + (pr.kv_start + key) * t.k_pt + a * t.k_an + h * D + ch * 8) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dv),
                   "l"(t.v + (pr.kv_start + key) * t.v_pt + a * t.v_an + h * D + ch * 8) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (nkv > 0) stage(0, 0);
  int buf = 0;
  for (int k0 = 0; k0 < nkv; k0 += kAttnBK, buf ^= 1) {
    __syncthreads();  // the other buffer's tile is fully consumed
    if (k0 + kAttnBK < nkv) {
      stage(k0 + kAttnBK, buf ^ 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const __nv_bfloat16* sk = sk_buf[buf];
    const __nv_bfloat16* sv = sv_buf[buf];

    // ---- S = Q K^T  (16 x 64 per warp)
    float s[kAttnBK / 8][4];
#pragma unroll
    for (int nt = 0; nt < kAttnBK / 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      const __nv_bfloat16* krow = sk + (nt * 8 + g) * kLd + 2 * tq;
#pragma unroll
      for (int ks = 0; ks < D / 16; ++ks) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(krow + ks * 16);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(krow + ks * 16 + 8);
        mma_bf16_16816(s[nt], qa[ks], b0, b1);
      }
    }
    // ---- bias, scale, key mask, running max
    float mx_lo = m_lo, mx_hi = m_hi;
#pragma unroll
    for (int nt = 0; nt < kAttnBK / 8; ++nt) {
      const int key = k0 + nt * 8 + 2 * tq;
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const bool ok = key + e < nkv;
        float b_lo = 0.f, b_hi = 0.f;
        if (t.bias && ok) { b_lo = bias_lo[key + e]; b_hi = bias_hi[key + e]; }
        s[nt][e] = ok ? (s[nt][e] + b_lo) * sc : -INFINITY;
        s[nt][2 + e] = ok ? (s[nt][2 + e] + b_hi) * sc : -INFINITY;
        mx_lo = fmaxf(mx_lo, s[nt][e]);
        mx_hi = fmaxf(mx_hi, s[nt][2 + e]);
      }
    }
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
    const float corr_lo = exp2f(m_lo - mx_lo), corr_hi = exp2f(m_hi - mx_hi);  // exp2(-inf) = 0 on the first tile
    m_lo = mx_lo;
    m_hi = mx_hi;
    l_lo *= corr_lo;
    l_hi *= corr_hi;
#pragma unroll
    for (int i = 0; i < D / 8; ++i) {
      o[i][0] *= corr_lo; o[i][1] *= corr_lo; o[i][2] *= corr_hi; o[i][3] *= corr_hi;
    }
    // ---- P = exp2(S - max), packed straight into A fragments of the P V product
    uint32_t pa[kAttnBK / 16][4];
#pragma unroll
    for (int nt = 0; nt < kAttnBK / 8; ++nt) {
      const float p0 = exp2f(s[nt][0] - m_lo), p1 = exp2f(s[nt][1] - m_lo);
      const float p2 = exp2f(s[nt][2] - m_hi), p3 = exp2f(s[nt][3] - m_hi);
      l_lo += p0 + p1;
      l_hi += p2 + p3;
      pa[nt >> 1][(nt & 1) * 2 + 0] = pack_bf16x2(p0, p1);
      pa[nt >> 1][(nt & 1) * 2 + 1] = pack_bf16x2(p2, p3);
    }
    // ---- O += P V
#pragma unroll
    for (int ks = 0; ks < kAttnBK / 16; ++ks) {
#pragma unroll
      for (int nt = 0; nt < D / 8; ++nt) {
        uint32_t b0, b1;
        ldmatrix_x2_trans(b0, b1, sv + (ks * 16 + (lane & 15)) * kLd + nt * 8);
        mma_bf16_16816(o[nt], pa[ks], b0, b1);
      }
    }
  }
  // ---- finalize
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);
  const float inv_lo = l_lo > 0.f ? 1.f / l_lo : 0.f, inv_hi = l_hi > 0.f ? 1.f / l_hi : 0.f;
  __nv_bfloat16* o_lo = t.out + ((pr.q_start + r_lo) * t.A + a) * t.ldo + h * D + 2 * tq;
  __nv_bfloat16* o_hi = t.out + ((pr.q_start + r_hi) * t.A + a) * t.ldo + h * D + 2 * tq;
#pragma unroll
  for (int nt = 0; nt < D / 8; ++nt) {
    if (r_lo < nq) *reinterpret_cast<uint32_t*>(o_lo + nt * 8) = pack_bf16x2(o[nt][0] * inv_lo, o[nt][1] * inv_lo);
    if (r_hi < nq) *reinterpret_cast<uint32_t*>(o_hi + nt * 8) = pack_bf16x2(o[nt][2] * inv_hi, o[nt][3] * inv_hi);
  }
}

}  // namespace se3et

using namespace se3et;

extern "C" int se3et_flash_attention(const void* q, int64_t q_pt, int64_t q_an, const void* k, int64_t k_pt,
                                     int64_t k_an, const void* v, int64_t v_pt, int64_t v_an, const float* bias,
                                     const int64_t* problems, int64_t num_problems, int64_t max_q, int64_t anchors,
                                     int64_t heads, int64_t head_dim, float scale, void* out_bf16, int64_t ldo,
                                     se3et_stream_t stream) {
  if (num_problems < 0 || max_q < 0 || anchors <= 0 || heads <= 0 || anchors * heads > 65535 || num_problems > 65535)
    return SE3ET_ERR_ARG;
  if (num_problems == 0 || max_q == 0) return SE3ET_OK;
  if (!q || !k || !v || !problems || !out_bf16) return SE3ET_ERR_ARG;
  // 16-byte K/V chunks and 4-byte Q/O accesses
  if ((k_pt | k_an | v_pt | v_an) % 8 || (q_pt | q_an | ldo) % 2) return SE3ET_ERR_ARG;
  AttnTensors t;
  t.q = static_cast<const __nv_bfloat16*>(q);
  t.k = static_cast<const __nv_bfloat16*>(k);
  t.v = static_cast<const __nv_bfloat16*>(v);
  t.q_pt = q_pt; t.q_an = q_an; t.k_pt = k_pt; t.k_an = k_an; t.v_pt = v_pt; t.v_an = v_an;
  t.bias = bias;
  t.out = static_cast<__nv_bfloat16*>(out_bf16);
  t.ldo = ldo;
  t.A = (int)anchors;
  t.H = (int)heads;
  t.scale = scale;
  dim3 grid((unsigned)ceil_div(max_q, kAttnBQ), (unsigned)(anchors * heads), (unsigned)num_problems);
  const auto* pr = reinterpret_cast<const AttnProblem*>(problems);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (head_dim) {
    case 16: flash_attention_kernel<16><<<grid, kAttnWarps * 32, 0, st>>>(t, pr); break;
    case 32: flash_attention_kernel<32><<<grid, kAttnWarps * 32, 0, st>>>(t, pr); break;
    case 64: flash_attention_kernel<64><<<grid, kAttnWarps * 32, 0, st>>>(t, pr); break;
    default: return SE3ET_ERR_UNSUPPORTED;
  }
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}
