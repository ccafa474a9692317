// Kernels of the SE3ET-E equivariant cross attention (MultiHeadAttentionEQ, modes a_soft / r_soft) and of the
// spherical-harmonics score term of the equivariant self attention:
//
//   anchor_pair_stats   G[p][a][e] = sum_{n,m} f( q_a[n] . k_e[m] / (H sqrt(c)) )   (vanilla_transformer.py:289-290,
//                       380-399, 425-431: head-mean of the local scores, made non-negative, pooled over all point
//                       pairs); mma.sync m16n8k16 over the full channel width, one CTA per (query tile, a, e, pair)
//   anchor_mix_weights  W[p][a][e]: a_soft = G normalised over e (:466-476); r_soft = rotation weights
//                       attn_r[r] = mean_a G[a][perm[r][a]] normalised over the 24 rotations (:560-575), folded to
//                       W[a][e] = sum_{r : perm[r][a] == e} attn_r[r] (what the brahnm gather + sum computes, :839-845)
//   anchor_mix          out[(n, a)] = sum_e W[pair(n)][a][e] * in_e[(n, a)]  (the weighted sum over key anchors of the
//                       per-(a, e) attention outputs, and eq2inv_soft, conditional_transformer.py:209-249)
//   sh_bias_add         bias[n, (a, h), m] += Y1(p_n - p_m) rotated by the anchor . u[(n, a), h]   (rpe_transformer.py:
//                       76-79 with geotransformer.py:57-67, l = 1 part; the l = 0 part is constant along m)
// Equivariant states are stored [point, anchor(6), channel] (bf16); a "problem" is one (query cloud, key cloud) pair.
#include <cuda_bf16.h>

#include "common.cuh"

namespace se3et {

struct EqProblem {  // same table as the attention kernel: {q_start, n_q, kv_start, n_kv, bias_off}
  int64_t q_start, n_q, kv_start, n_kv, bias_off;
};

__device__ __forceinline__ void eq_mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float positive_fn(float x, int mode) {
  switch (mode) {
    case 0: return x * x;                                   // 'sq'
    case 1: return x > 20.f ? x : log1pf(__expf(x));        // 'softplus' (beta 1, threshold 20 as torch)
    case 2: return 1.f / (1.f + __expf(-x));                // 'sigmoid'
    case 3: return fmaxf(x, 0.f);                           // 'relu'
    default: return fabsf(x);                               // 'abs'
  }
}

constexpr int kStatWarps = 4;
constexpr int kStatBQ = 16 * kStatWarps;
constexpr int kStatBK = 64;

template <int C>
__global__ void __launch_bounds__(kStatWarps * 32)
anchor_pair_stats_kernel(const __nv_bfloat16* __restrict__ q, int64_t q_pt, int64_t q_an,
                         const __nv_bfloat16* __restrict__ k, int64_t k_pt, int64_t k_an,
                         const EqProblem* __restrict__ problems, int A, float scale, int mode,
                         float* __restrict__ g /* [P][A][A] */) {
  constexpr int kLd = C + 8;
  extern __shared__ __align__(16) uint8_t stat_smem[];
  __nv_bfloat16* sk = reinterpret_cast<__nv_bfloat16*>(stat_smem);
  __shared__ float sh_red[kStatWarps];
  const EqProblem pr = problems[blockIdx.z];
  const int q0 = blockIdx.x * kStatBQ;
  if (q0 >= pr.n_q) return;
  const int a = blockIdx.y / A, e = blockIdx.y % A;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, tq = lane & 3;
  const int nq = (int)pr.n_q, nkv = (int)pr.n_kv;
  const int r_lo = q0 + warp * 16 + gq, r_hi = r_lo + 8;
  const int rl = min(r_lo, nq - 1), rh = min(r_hi, nq - 1);
  const __nv_bfloat16* q_lo = q + (pr.q_start + rl) * q_pt + a * q_an;
  const __nv_bfloat16* q_hi = q + (pr.q_start + rh) * q_pt + a * q_an;
  uint32_t qa[C / 16][4];
#pragma unroll
  for (int ks = 0; ks < C / 16; ++ks) {
    qa[ks][0] = *reinterpret_cast<const uint32_t*>(q_lo + ks * 16 + 2 * tq);
    qa[ks][1] = *reinterpret_cast<const uint32_t*>(q_hi + ks * 16 + 2 * tq);
    qa[ks][2] = *reinterpret_cast<const uint32_t*>(q_lo + ks * 16 + 8 + 2 * tq);
    qa[ks][3] = *reinterpret_cast<const uint32_t*>(q_hi + ks * 16 + 8 + 2 * tq);
  }
  float acc = 0.f;
  for (int k0 = 0; k0 < nkv; k0 += kStatBK) {
    __syncthreads();
    constexpr int kChunks = C / 8;
    for (int c = threadIdx.x; c < kStatBK * kChunks; c += kStatWarps * 32) {
      const int row = c / kChunks, ch = c - row * kChunks;
      const int key = min(k0 + row, nkv - 1);
      *reinterpret_cast<uint4*>(sk + row * kLd + ch * 8) =
          *reinterpret_cast<const uint4*>(k + (pr.kv_start + key) * k_pt + e * k_an + ch * 8);
    }
    __syncthreads();
#pragma unroll
    for (int nt = 0; nt < kStatBK / 8; ++nt) {
      float s[4] = {0.f, 0.f, 0.f, 0.f};
      const __nv_bfloat16* krow = sk + (nt * 8 + gq) * kLd + 2 * tq;
#pragma unroll
      for (int ks = 0; ks < C / 16; ++ks)
        eq_mma_16816(s, qa[ks], *reinterpret_cast<const uint32_t*>(krow + ks * 16),
                     *reinterpret_cast<const uint32_t*>(krow + ks * 16 + 8));
      const int key = k0 + nt * 8 + 2 * tq;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        if (key + i < nkv) {
          if (r_lo < nq) acc += positive_fn(s[i] * scale, mode);
          if (r_hi < nq) acc += positive_fn(s[2 + i] * scale, mode);
        }
      }
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) sh_red[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < kStatWarps; ++w) t += sh_red[w];
    atomicAdd(g + ((int64_t)blockIdx.z * A + a) * A + e, t);
  }
}

// one CTA of 32 threads per problem
__global__ void anchor_mix_weights_kernel(const float* __restrict__ g, const EqProblem* __restrict__ problems,
                                          const int32_t* __restrict__ perms, int num_rot, int A, int r_soft,
                                          float* __restrict__ w /* [P][A][A] */, float* __restrict__ attn_r /* [P][R] */) {
  __shared__ float sg[8][8];
  __shared__ float sr[64];
  const int p = blockIdx.x, t = threadIdx.x;
  const float inv = 1.f / ((float)problems[p].n_q * (float)problems[p].n_kv);
  for (int i = t; i < A * A; i += 32) sg[i / A][i % A] = g[(int64_t)p * A * A + i] * inv;
  __syncwarp();
  if (!r_soft) {
    if (t < A) {
      float s = 0.f;
      for (int e = 0; e < A; ++e) s += sg[t][e];
      for (int e = 0; e < A; ++e) w[((int64_t)p * A + t) * A + e] = sg[t][e] / s;
    }
    return;
  }
  for (int r = t; r < num_rot; r += 32) {
    float s = 0.f;
    for (int a = 0; a < A; ++a) s += sg[a][perms[r * A + a]];
    sr[r] = s / (float)A;
  }
  __syncwarp();
  float tot = 0.f;
  for (int r = 0; r < num_rot; ++r) tot += sr[r];
  for (int r = t; r < num_rot; r += 32) {
    const float v = sr[r] / tot;
    if (attn_r) attn_r[(int64_t)p * num_rot + r] = v;
  }
  for (int i = t; i < A * A; i += 32) {
    const int a = i / A, e = i % A;
    float s = 0.f;
    for (int r = 0; r < num_rot; ++r)
      if (perms[r * A + a] == e) s += sr[r];
    w[(int64_t)p * A * A + i] = s / tot;
  }
}

// out[(n, a)][c] = sum_e W[pair(n)][a][e] * in[e * stride_e + n * stride_n + a * stride_a + c]
__global__ void __launch_bounds__(256) anchor_mix_kernel(const __nv_bfloat16* __restrict__ in, int64_t stride_e,
                                                          int64_t stride_n, int64_t stride_a,
                                                          const float* __restrict__ w,
                                                          const int64_t* __restrict__ cloud_off, int nclouds, int A,
                                                          int C, int64_t n_points, __nv_bfloat16* __restrict__ out) {
  const int half = C / 2;
  const int64_t total = n_points * A * half;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % half) * 2;
    const int64_t row = i / half;  // (n, a)
    const int a = (int)(row % A);
    const int64_t n = row / A;
    const int p = segment_of(cloud_off, nclouds, n);
    const float* wr = w + ((int64_t)p * A + a) * A;
    float s0 = 0.f, s1 = 0.f;
    for (int e = 0; e < A; ++e) {
      const uint32_t v = *reinterpret_cast<const uint32_t*>(in + e * stride_e + n * stride_n + a * stride_a + c);
      const float we = wr[e];
      s0 = fmaf(we, __uint_as_float(v << 16), s0);
      s1 = fmaf(we, __uint_as_float(v & 0xffff0000u), s1);
    }
    __nv_bfloat162 o = __floats2bfloat162_rn(s0, s1);
    *reinterpret_cast<__nv_bfloat162*>(out + row * C + c) = o;
  }
}

// bias[bias_off + ((i * A + a) * H + h) * n + m] += c1 * (R_a^T unit(p_i - p_m)) . u[((start + i) * A + a) * ldu + h * 3 ..]
__global__ void __launch_bounds__(256) sh_bias_add_kernel(const float* __restrict__ pts,
                                                           const EqProblem* __restrict__ problems,
                                                           const float* __restrict__ u, int ldu,
                                                           const float* __restrict__ anchors /* [A][3][3] */, int A,
                                                           int H, float c1, float* __restrict__ bias) {
  const EqProblem pr = problems[blockIdx.y];
  const int n = (int)pr.n_q;
  const int64_t total = (int64_t)n * A * H * n;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(t % n);
    int64_t r = t / n;
    const int h = (int)(r % H);
    r /= H;
    const int a = (int)(r % A);
    const int i = (int)(r / A);
    const float* pi = pts + 3 * (pr.q_start + i);
    const float* pm = pts + 3 * (pr.q_start + m);
    float dx = pi[0] - pm[0], dy = pi[1] - pm[1], dz = pi[2] - pm[2];
    const float nrm = fmaxf(sqrtf(dx * dx + dy * dy + dz * dz), 1e-12f);  // F.normalize eps
    dx /= nrm; dy /= nrm; dz /= nrm;
    // D^1(anchors^T) = anchors^T:  y_c = sum_d anchors[a][d][c] * unit_d
    const float* R = anchors + 9 * a;
    const float y0 = R[0] * dx + R[3] * dy + R[6] * dz;
    const float y1 = R[1] * dx + R[4] * dy + R[7] * dz;
    const float y2 = R[2] * dx + R[5] * dy + R[8] * dz;
    const float* uu = u + ((pr.q_start + i) * A + a) * (int64_t)ldu + h * 3;
    bias[pr.bias_off + t] += c1 * (y0 * uu[0] + y1 * uu[1] + y2 * uu[2]);
  }
}

}  // namespace se3et

using namespace se3et;

extern "C" int se3et_anchor_pair_stats(const void* q_bf16, int64_t q_pt, int64_t q_an, const void* k_bf16,
                                       int64_t k_pt, int64_t k_an, const int64_t* problems, int64_t num_problems,
                                       int64_t max_q, int64_t anchors, int64_t channels, float scale, int positive,
                                       float* g, se3et_stream_t stream) {
  if (num_problems < 0 || anchors <= 0 || anchors > 8 || max_q < 0 || positive < 0 || positive > 4) return SE3ET_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!g) return SE3ET_ERR_ARG;
  SE3ET_CUDA_CHECK(cudaMemsetAsync(g, 0, sizeof(float) * num_problems * anchors * anchors, st));
  if (num_problems == 0 || max_q == 0) return SE3ET_OK;
  if (!q_bf16 || !k_bf16 || !problems) return SE3ET_ERR_ARG;
  dim3 grid((unsigned)ceil_div(max_q, kStatBQ), (unsigned)(anchors * anchors), (unsigned)num_problems);
  const auto* q = static_cast<const __nv_bfloat16*>(q_bf16);
  const auto* k = static_cast<const __nv_bfloat16*>(k_bf16);
  const auto* pr = reinterpret_cast<const EqProblem*>(problems);
#define SE3ET_STATS(C)                                                                                              \
  {                                                                                                                 \
    const int smem = kStatBK * ((C) + 8) * 2;                                                                       \
    SE3ET_ENSURE_SMEM(anchor_pair_stats_kernel<C>, smem);                                                           \
    anchor_pair_stats_kernel<C><<<grid, kStatWarps * 32, smem, st>>>(q, q_pt, q_an, k, k_pt, k_an, pr, (int)anchors, \
                                                                      scale, positive, g);                          \
  }
  switch (channels) {
    case 64: SE3ET_STATS(64) break;
    case 128: SE3ET_STATS(128) break;
    case 256: SE3ET_STATS(256) break;
    default: return SE3ET_ERR_UNSUPPORTED;
  }
#undef SE3ET_STATS
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}

extern "C" int se3et_anchor_mix_weights(const float* g, const int64_t* problems, int64_t num_problems,
                                        const int32_t* perms, int64_t num_rotations, int64_t anchors, int r_soft,
                                        float* w, float* attn_r, se3et_stream_t stream) {
  if (num_problems < 0 || anchors <= 0 || anchors > 8 || num_rotations < 0 || num_rotations > 64) return SE3ET_ERR_ARG;
  if (num_problems == 0) return SE3ET_OK;
  if (!g || !problems || !w || (r_soft && (!perms || num_rotations == 0))) return SE3ET_ERR_ARG;
  anchor_mix_weights_kernel<<<(unsigned)num_problems, 32, 0, static_cast<cudaStream_t>(stream)>>>(
      g, reinterpret_cast<const EqProblem*>(problems), perms, (int)num_rotations, (int)anchors, r_soft, w, attn_r);
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}

extern "C" int se3et_anchor_mix(const void* in_bf16, int64_t stride_e, int64_t stride_n, int64_t stride_a,
                                const float* w, const int64_t* cloud_offsets, int64_t nclouds, int64_t anchors,
                                int64_t channels, int64_t n_points, void* out_bf16, se3et_stream_t stream) {
  if (n_points < 0 || anchors <= 0 || channels <= 0 || channels % 2 || nclouds <= 0) return SE3ET_ERR_ARG;
  if (n_points == 0) return SE3ET_OK;
  if (!in_bf16 || !w || !cloud_offsets || !out_bf16) return SE3ET_ERR_ARG;
  int64_t blocks = ceil_div(n_points * anchors * channels / 2, 256);
  if (blocks > (int64_t)kNumSMs * 16) blocks = (int64_t)kNumSMs * 16;
  anchor_mix_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(in_bf16), stride_e, stride_n, stride_a, w, cloud_offsets, (int)nclouds,
      (int)anchors, (int)channels, n_points, static_cast<__nv_bfloat16*>(out_bf16));
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}

extern "C" int se3et_sh_bias_add(const float* points, const int64_t* problems, int64_t num_problems, int64_t max_n,
                                 const float* u, int64_t ldu, const float* anchors_Ax3x3, int64_t anchors,
                                 int64_t heads, float c1, float* bias, se3et_stream_t stream) {
  if (num_problems < 0 || max_n < 0 || anchors <= 0 || heads <= 0 || ldu < heads * 3) return SE3ET_ERR_ARG;
  if (num_problems == 0 || max_n == 0) return SE3ET_OK;
  if (!points || !problems || !u || !anchors_Ax3x3 || !bias) return SE3ET_ERR_ARG;
  int64_t bx = ceil_div(max_n * anchors * heads * max_n, 256);
  if (bx > (int64_t)kNumSMs * 8) bx = (int64_t)kNumSMs * 8;
  dim3 grid((unsigned)bx, (unsigned)num_problems);
  sh_bias_add_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      points, reinterpret_cast<const EqProblem*>(problems), u, (int)ldu, anchors_Ax3x3, (int)anchors, (int)heads, c1,
      bias);
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}
