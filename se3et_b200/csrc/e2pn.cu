// E2PN backbone kernels other than the tensor-core GEMM (sm_100a):
//   kpconv_gather   neighbour gather + kernel-point influence + rotate-by-permute  (blocks_epn.py:334-390,471-478,503-505)
//   groupnorm_stats / groupnorm_apply   GroupNormEPN over (C/G x A x N_pair)        (blocks_epn.py:684-701)
//   maxpool_nbr     strided shortcut max pooling                                    (e2pn/blocks.py:93-110)
//   anchor_max      InvOutBlockEPN / eq->inv pooling, max over the anchor axis      (blocks_epn.py:924)
//   upsample_concat nearest_upsample + torch.cat of the decoder                     (kpconv/functional.py:6-22)
// Features are stored [point, anchor(6), channel]; activations bf16, pre-norm GEMM outputs fp32.
#include <cuda_bf16.h>

#include "common.cuh"
#include "kpconv_tables.cuh"

namespace se3et {

__constant__ int c_ridx[6][6] = {{0, 3, 3, 3, 3, 5}, {1, 0, 4, 5, 2, 1}, {2, 2, 0, 4, 5, 4},
                                 {3, 5, 2, 0, 4, 3}, {4, 4, 5, 2, 0, 2}, {5, 1, 1, 1, 1, 0}};

constexpr int kMaxH = 96;    // neighbour columns supported by the gather / pooling kernels

int kpconv_gather_mma(const float* q_pts, const float* s_pts, const int64_t* neighbors, int64_t nq, int64_t ns, int h,
                      const __nv_bfloat16* x, int cin, const float* kp, float inv_extent, __nv_bfloat16* out, int kpad,
                      cudaStream_t st);  // kpconv.cu

__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&p);
}

// One CTA per query point.  Threads own column pairs (a, c..c+1) of the gathered feature rows.
//   wf[k][col]  = sum_n w[n][k] * x[idx[n]][col]                       (einsum 'pnac,pnk->pkac')
//   A'[(p,r)][(kc, a', c)] = sum_{k : kidx[k][r] == kc} wf[k][(a, c)],  a' = ridx[a][r]
// so that  out[(p,r), d] = sum_K A'[(p,r), K] * W[kc][a'][c][d]  is ONE GEMM with the untouched weights
// (the reference gathers the weights instead: 'kpac,karcd->prd' with W_eff = W[kidx_rot, ridx_rot]).
template <bool kPair>
__global__ void __launch_bounds__(256) kpconv_gather_kernel(const float* __restrict__ q_pts,
                                                             const float* __restrict__ s_pts,
                                                             const int64_t* __restrict__ idx, int H, int64_t ns,
                                                             const __nv_bfloat16* __restrict__ x, int cin,
                                                             const float* __restrict__ kernel_points, float inv_extent,
                                                             __nv_bfloat16* __restrict__ out, int kpad) {
  __shared__ float sh_w[kMaxH][kKP + 1];
  __shared__ int sh_idx[kMaxH];
  const int64_t p = blockIdx.x;
  const int tid = threadIdx.x;
  if (tid < H) {
    const int64_t j = idx[p * H + tid];
    sh_idx[tid] = (j >= 0 && j < ns) ? (int)j : -1;  // shadow neighbour: zero feature, far point
  }
  __syncthreads();
  for (int t = tid; t < H * kKP; t += blockDim.x) {
    const int n = t / kKP, k = t - n * kKP;
    const int j = sh_idx[n];
    float w = 0.f;
    if (j >= 0) {
      const float dx = (s_pts[3 * (int64_t)j + 0] - q_pts[3 * p + 0]) - kernel_points[3 * k + 0];
      const float dy = (s_pts[3 * (int64_t)j + 1] - q_pts[3 * p + 1]) - kernel_points[3 * k + 1];
      const float dz = (s_pts[3 * (int64_t)j + 2] - q_pts[3 * p + 2]) - kernel_points[3 * k + 2];
      w = fmaxf(0.f, 1.f - sqrtf(dx * dx + dy * dy + dz * dz) * inv_extent);  // 'linear' influence (blocks_epn.py:351)
    }
    sh_w[n][k] = w;
  }
  __syncthreads();

  const int width = kA * cin;                   // columns of one feature row
  const int units = kPair ? width / 2 : width;  // work items per point
  for (int u = tid; u < units; u += blockDim.x) {
    const int col = kPair ? 2 * u : u;
    float acc0[kKP], acc1[kKP];
#pragma unroll
    for (int k = 0; k < kKP; ++k) acc0[k] = acc1[k] = 0.f;
    for (int n = 0; n < H; ++n) {
      const int j = sh_idx[n];
      if (j < 0) continue;
      float x0, x1 = 0.f;
      if (kPair) {
        const uint32_t v = *reinterpret_cast<const uint32_t*>(x + (int64_t)j * width + col);
        x0 = bf_lo(v);
        x1 = bf_hi(v);
      } else {
        x0 = __bfloat162float(x[(int64_t)j * width + col]);
      }
#pragma unroll
      for (int k = 0; k < kKP; ++k) {
        const float w = sh_w[n][k];
        acc0[k] = fmaf(w, x0, acc0[k]);
        if (kPair) acc1[k] = fmaf(w, x1, acc1[k]);
      }
    }
    const int a = col / cin, c = col - a * cin;
#pragma unroll
    for (int r = 0; r < kA; ++r) {
      float s0[kKC], s1[kKC];
#pragma unroll
      for (int kc = 0; kc < kKC; ++kc) s0[kc] = s1[kc] = 0.f;
#pragma unroll
      for (int k = 0; k < kKP; ++k) {
        // tables are compile-time constants after unrolling: s0[] stays in registers
        const int kc = kidx_tab(k, r);
        s0[kc] += acc0[k];
        if (kPair) s1[kc] += acc1[k];
      }
      const int ap = c_ridx[a][r];
      __nv_bfloat16* row = out + (p * kA + r) * (int64_t)kpad;
#pragma unroll
      for (int kc = 0; kc < kKC; ++kc) {
        const int kk = (kc * kA + ap) * cin + c;
        if (kPair) *reinterpret_cast<uint32_t*>(row + kk) = pack_bf16(s0[kc], s1[kc]);
        else row[kk] = __float2bfloat16(s0[kc]);
      }
    }
  }
  // zero the K padding (only when 36*cin is not a multiple of 64)
  const int kreal = kKC * kA * cin;
  for (int t = tid; t < kA * (kpad - kreal); t += blockDim.x) {
    const int r = t / (kpad - kreal), j = t - r * (kpad - kreal);
    out[(p * kA + r) * (int64_t)kpad + kreal + j] = __float2bfloat16(0.f);
  }
}

// ---------------------------------------------------------------------------------------------
// GroupNorm over (channels of a group) x (rows of a segment); a segment = one point-cloud pair.
// rows_per_point = 6 for equivariant features (N, A, C), 1 for invariant ones (N, C).
// stats[seg][group] = {sum, sum of squares} in double.
// ---------------------------------------------------------------------------------------------
constexpr int kStatRows = 128;

__global__ void __launch_bounds__(256) groupnorm_stats_kernel(const float* __restrict__ y, int64_t rows, int C,
                                                               int cpg, const int64_t* __restrict__ seg_off, int nseg,
                                                               int rows_per_point, double* __restrict__ stats) {
  extern __shared__ double sh_part[];  // [G_block][2]; fp64 everywhere: the result does not depend on how rows
                                        // are partitioned into CTAs (batched pairs == single pairs to the last bit)
  const int c = blockIdx.y * blockDim.x + threadIdx.x;
  const int g_first = (blockIdx.y * blockDim.x) / cpg;
  const int g_count = (min(C, (int)((blockIdx.y + 1) * blockDim.x)) - 1) / cpg - g_first + 1;
  const int G = C / cpg;
  const int64_t r0 = (int64_t)blockIdx.x * kStatRows;
  const int64_t r1 = min(rows, r0 + kStatRows);
  int seg = segment_of(seg_off, nseg, r0 / rows_per_point);
  int64_t r = r0;
  while (r < r1) {
    const int64_t seg_end = min(r1, seg_off[seg + 1] * rows_per_point);
    double s = 0.0, ss = 0.0;
    if (c < C)
      for (int64_t i = r; i < seg_end; ++i) {
        const double v = (double)y[i * C + c];
        s += v;
        ss += v * v;
      }
    for (int i = threadIdx.x; i < 2 * g_count; i += blockDim.x) sh_part[i] = 0.0;
    __syncthreads();
    if (c < C) {
      atomicAdd(&sh_part[2 * (c / cpg - g_first)], s);
      atomicAdd(&sh_part[2 * (c / cpg - g_first) + 1], ss);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * g_count; i += blockDim.x)
      atomicAdd(&stats[((int64_t)seg * G + g_first) * 2 + i], sh_part[i]);
    __syncthreads();
    r = seg_end;
    ++seg;
    while (r < r1 && seg < nseg && seg_off[seg + 1] * rows_per_point <= r) ++seg;  // skip empty segments
  }
}

struct NormSide {
  const float* y;       // fp32 pre-norm values [rows, C] (nullable => side absent)
  const double* stats;  // [nseg, G, 2]
  const float* gamma;   // [C]
  const float* beta;    // [C]
};

// out = act( norm_a(ya) [+ norm_b(yb)] [+ resid] ),  act = LeakyReLU(slope) when slope != 1
__global__ void __launch_bounds__(256) groupnorm_apply_kernel(NormSide a, NormSide b,
                                                               const __nv_bfloat16* __restrict__ resid, int64_t rows,
                                                               int C, int cpg, const int64_t* __restrict__ seg_off,
                                                               int nseg, int rows_per_point, float eps, float slope,
                                                               float* __restrict__ out_f32,
                                                               __nv_bfloat16* __restrict__ out_bf16) {
  const int vec_per_row = C / 4;
  const int64_t total = rows * vec_per_row;
  const int G = C / cpg;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / vec_per_row;
    const int c0 = (int)(i - row * vec_per_row) * 4;
    const int seg = segment_of(seg_off, nseg, row / rows_per_point);
    const double cnt = (double)(seg_off[seg + 1] - seg_off[seg]) * rows_per_point * cpg;
    float v[4];
    {
      const float4 yv = *reinterpret_cast<const float4*>(a.y + row * C + c0);
      const float in[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int g = (c0 + j) / cpg;
        const double mean = a.stats[((int64_t)seg * G + g) * 2] / cnt;
        const double var = a.stats[((int64_t)seg * G + g) * 2 + 1] / cnt - mean * mean;
        const float rstd = rsqrtf((float)fmax(var, 0.0) + eps);
        v[j] = (in[j] - (float)mean) * rstd * a.gamma[c0 + j] + a.beta[c0 + j];
      }
    }
    if (b.y) {
      const float4 yv = *reinterpret_cast<const float4*>(b.y + row * C + c0);
      const float in[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int g = (c0 + j) / cpg;
        const double mean = b.stats[((int64_t)seg * G + g) * 2] / cnt;
        const double var = b.stats[((int64_t)seg * G + g) * 2 + 1] / cnt - mean * mean;
        const float rstd = rsqrtf((float)fmax(var, 0.0) + eps);
        v[j] += (in[j] - (float)mean) * rstd * b.gamma[c0 + j] + b.beta[c0 + j];
      }
    }
    if (resid) {
      const uint2 rv = *reinterpret_cast<const uint2*>(resid + row * C + c0);
      v[0] += bf_lo(rv.x); v[1] += bf_hi(rv.x); v[2] += bf_lo(rv.y); v[3] += bf_hi(rv.y);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = v[j] >= 0.f ? v[j] : v[j] * slope;
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + row * C + c0) = make_float4(v[0], v[1], v[2], v[3]);
    if (out_bf16) {
      uint2 o;
      o.x = pack_bf16(v[0], v[1]);
      o.y = pack_bf16(v[2], v[3]);
      *reinterpret_cast<uint2*>(out_bf16 + row * C + c0) = o;
    }
  }
}

// Streaming form of the same operation for power-of-two row widths (every SE3ET layer): a thread owns a fixed
// group of 4 channels and walks down the rows, so the per-(pair, group) mean / rstd and the affine parameters are
// loaded and folded once per pair instead of once per element; the row -> pair lookup advances linearly.
struct NormCols {
  float mean[4], s[4], beta[4];
};
__device__ __forceinline__ void load_norm_cols(const NormSide& n, int seg, int G, int cpg, int c0, double cnt, float eps,
                                               NormCols& o) {
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int g = (c0 + j) / cpg;
    const double mean = n.stats[((int64_t)seg * G + g) * 2] / cnt;
    const double var = n.stats[((int64_t)seg * G + g) * 2 + 1] / cnt - mean * mean;
    const float rstd = rsqrtf((float)fmax(var, 0.0) + eps);
    o.mean[j] = (float)mean;
    o.s[j] = rstd * n.gamma[c0 + j];
    o.beta[j] = n.beta[c0 + j];
  }
}

template <bool kHasB>
__global__ void __launch_bounds__(256) groupnorm_apply_rows_kernel(NormSide a, NormSide b,
                                                                    const __nv_bfloat16* __restrict__ resid,
                                                                    int64_t rows, int C, int cpg,
                                                                    const int64_t* __restrict__ seg_off, int nseg,
                                                                    int rows_per_point, float eps, float slope,
                                                                    float* __restrict__ out_f32,
                                                                    __nv_bfloat16* __restrict__ out_bf16) {
  const int V = C >> 2;  // float4 vectors per row; a power of two <= 256, so it divides the thread count
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int c0 = (int)(tid % V) * 4;
  const int64_t row_step = (int64_t)gridDim.x * blockDim.x / V;
  const int G = C / cpg;
  int seg = -1;
  int64_t seg_end = 0;  // first row past the current pair
  NormCols na, nb;
  for (int64_t row = tid / V; row < rows; row += row_step) {
    if (row >= seg_end) {
      if (seg < 0) seg = segment_of(seg_off, nseg, row / rows_per_point);
      while (seg + 1 < nseg && seg_off[seg + 1] * rows_per_point <= row) ++seg;
      seg_end = seg_off[seg + 1] * rows_per_point;
      if (seg == nseg - 1) seg_end = rows;
      const double cnt = (double)(seg_off[seg + 1] - seg_off[seg]) * rows_per_point * cpg;
      load_norm_cols(a, seg, G, cpg, c0, cnt, eps, na);
      if (kHasB) load_norm_cols(b, seg, G, cpg, c0, cnt, eps, nb);
    }
    const float4 ya = __ldcs(reinterpret_cast<const float4*>(a.y + row * C + c0));
    float v[4] = {(ya.x - na.mean[0]) * na.s[0] + na.beta[0], (ya.y - na.mean[1]) * na.s[1] + na.beta[1],
                  (ya.z - na.mean[2]) * na.s[2] + na.beta[2], (ya.w - na.mean[3]) * na.s[3] + na.beta[3]};
    if (kHasB) {
      const float4 yb = __ldcs(reinterpret_cast<const float4*>(b.y + row * C + c0));
      v[0] += (yb.x - nb.mean[0]) * nb.s[0] + nb.beta[0];
      v[1] += (yb.y - nb.mean[1]) * nb.s[1] + nb.beta[1];
      v[2] += (yb.z - nb.mean[2]) * nb.s[2] + nb.beta[2];
      v[3] += (yb.w - nb.mean[3]) * nb.s[3] + nb.beta[3];
    }
    if (resid) {
      const uint2 rv = *reinterpret_cast<const uint2*>(resid + row * C + c0);
      v[0] += bf_lo(rv.x); v[1] += bf_hi(rv.x); v[2] += bf_lo(rv.y); v[3] += bf_hi(rv.y);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = v[j] >= 0.f ? v[j] : v[j] * slope;
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + row * C + c0) = make_float4(v[0], v[1], v[2], v[3]);
    if (out_bf16) {
      uint2 o;
      o.x = pack_bf16(v[0], v[1]);
      o.y = pack_bf16(v[2], v[3]);
      *reinterpret_cast<uint2*>(out_bf16 + row * C + c0) = o;
    }
  }
}


// Two GroupNorm + LeakyReLU stages back to back (KPConvInterSO3Block's norm followed by the enclosing block's
// norm, blocks_epn.py:737-741 + 790-794 / 841-843) without materialising the intermediate:
//   f = LeakyReLU(GN_1(y)),  out = LeakyReLU(GN_2(f))
// kMode 0 (pass A): accumulates the statistics of f per (pair, group) into stats2 (fp64 atomics)
// kMode 1 (pass B): recomputes f and writes out (bf16)
// kMode 2         : statistics of y itself into stats2 (the first norm's statistics, for producers whose epilogue
//                   does not deliver them: cheaper here, as a streaming pass, than on the fused KPConv's critical path)
// kBf16: y is bf16 (the inference path's conv kernels write their pre-norm output in bf16: half the bytes of all three
// passes); a thread then owns eight columns, so that every load stays 16 bytes wide.
// Per-thread constants of one GroupNorm for N columns in the folded form  y * s + t  (t = beta - mean * s).
template <int N>
struct NormColsN {
  float s[N], t[N];
};
template <int N>
__device__ __forceinline__ void load_norm_cols_n(const NormSide& n, int seg, int G, int cpg, int c0, double cnt,
                                                 float eps, NormColsN<N>& o) {
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const int g = (c0 + j) / cpg;
    const double mean = n.stats[((int64_t)seg * G + g) * 2] / cnt;
    const double var = n.stats[((int64_t)seg * G + g) * 2 + 1] / cnt - mean * mean;
    const float rstd = rsqrtf((float)fmax(var, 0.0) + eps);
    o.s[j] = rstd * n.gamma[c0 + j];
    o.t[j] = (float)((double)n.beta[c0 + j] - mean * (double)o.s[j]);
  }
}

template <int kMode, bool kBf16>
__global__ void __launch_bounds__(256, (kMode == 1 || (kBf16 && kMode == 0)) ? 2 : 3) groupnorm_double_kernel(NormSide a, NormSide b2, double* __restrict__ stats2_acc,
                                                                int64_t rows, int C, int cpg,
                                                                const int64_t* __restrict__ seg_off, int nseg,
                                                                int rows_per_point, float eps, float slope,
                                                                __nv_bfloat16* __restrict__ out_bf16) {
  // every CTA owns a contiguous range of rows and walks the pairs it intersects one after the other, so the
  // statistics of a pair are reduced inside the CTA (shared-memory atomics) before one set of fp64 atomics
  __shared__ float sh_acc[2 * 256];  // [group][sum, sum sq], G <= 256
  constexpr int N = kBf16 ? 8 : 4;   // columns per thread
  const int V = C / N;
  const int G = C / cpg;
  const int c0 = (threadIdx.x % V) * N;
  const int rsub = threadIdx.x / V, rstep = blockDim.x / V;
  const int64_t per_cta = (rows + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = (int64_t)blockIdx.x * per_cta, r1 = min(rows, r0 + per_cta);
  if (r0 >= r1) return;
  int seg = segment_of(seg_off, nseg, r0 / rows_per_point);
  int64_t row0 = r0;
  while (row0 < r1) {
    while (seg + 1 < nseg && seg_off[seg + 1] * rows_per_point <= row0) ++seg;
    int64_t seg_end = seg == nseg - 1 ? rows : seg_off[seg + 1] * rows_per_point;
    const int64_t row1 = min(r1, seg_end);
    const double cnt = (double)(seg_off[seg + 1] - seg_off[seg]) * rows_per_point * cpg;
    constexpr bool kApply = kMode == 1;
    NormColsN<N> n1, n2;
    if (kMode != 2) load_norm_cols_n<N>(a, seg, G, cpg, c0, cnt, eps, n1);
    float s2b[N];
    if (kApply) {
      load_norm_cols_n<N>(b2, seg, G, cpg, c0, cnt, eps, n2);
#pragma unroll
      for (int j = 0; j < N; ++j) {
        s2b[j] = n2.s[j] * (0.5f * (1.f - slope));
        n2.s[j] *= 0.5f * (1.f + slope);
      }
    }
    float s[N], ss[N];
#pragma unroll
    for (int j = 0; j < N; ++j) s[j] = ss[j] = 0.f;
    // four rows per thread and iteration, software pipelined: the loads of the next four rows are issued before this
    // iteration's arithmetic (this kernel is pure streaming: the bytes in flight per SM decide its speed, and the bf16
    // form has twice the arithmetic per byte)
    constexpr int kU = 4;
    const int64_t step = (int64_t)kU * rstep;
    auto load_rows = [&](int64_t rowb, uint4 (&dst)[kU]) {
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int64_t row = rowb + (int64_t)u * rstep;
        if (row < row1) {
          if (kBf16)
            dst[u] = __ldcs(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(a.y) + row * C + c0));
          else
            dst[u] = __ldcs(reinterpret_cast<const uint4*>(a.y + row * C + c0));
        }
      }
    };
    uint4 yv[kU], yn[kU];
    load_rows(row0 + rsub, yv);
    for (int64_t rowb = row0 + rsub; rowb < row1; rowb += step) {
      load_rows(rowb + step, yn);
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        const int64_t row = rowb + (int64_t)u * rstep;
        if (row >= row1) continue;
        float v[N];
        if (kBf16) {
          v[0] = bf_lo(yv[u].x); v[1] = bf_hi(yv[u].x); v[2] = bf_lo(yv[u].y); v[3] = bf_hi(yv[u].y);
          v[N - 4] = bf_lo(yv[u].z); v[N - 3] = bf_hi(yv[u].z); v[N - 2] = bf_lo(yv[u].w); v[N - 1] = bf_hi(yv[u].w);
        } else {
          v[0] = __uint_as_float(yv[u].x); v[1] = __uint_as_float(yv[u].y);
          v[2] = __uint_as_float(yv[u].z); v[3] = __uint_as_float(yv[u].w);
        }
        // the instruction count per element decides the speed of the bf16 form (half the bytes per element):
        // LeakyReLU(w) = max(w, slope w) for 0 <= slope <= 1 (checked by the host), and in the apply pass the first
        // activation is folded into the second norm:  GN_2(LeakyReLU(w)) = s2 (a w + b |w|) + t2,  a, b = (1 +- slope) / 2
        if (kMode == 0) {
#pragma unroll
          for (int j = 0; j < N; ++j) {
            const float w = fmaf(v[j], n1.s[j], n1.t[j]);
            v[j] = fmaxf(w, w * slope);
          }
        }
        if (!kApply) {
#pragma unroll
          for (int j = 0; j < N; ++j) {
            s[j] += v[j];
            ss[j] = fmaf(v[j], v[j], ss[j]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < N; ++j) {
            const float w = fmaf(v[j], n1.s[j], n1.t[j]);
            const float z = fmaf(w, n2.s[j], fmaf(fabsf(w), s2b[j], n2.t[j]));
            v[j] = fmaxf(z, z * slope);
          }
          if (kBf16) {
            *reinterpret_cast<uint4*>(out_bf16 + row * C + c0) =
                make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[N - 4], v[N - 3]),
                           pack_bf16(v[N - 2], v[N - 1]));
          } else {
            uint2 o;
            o.x = pack_bf16(v[0], v[1]);
            o.y = pack_bf16(v[2], v[3]);
            *reinterpret_cast<uint2*>(out_bf16 + row * C + c0) = o;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u) yv[u] = yn[u];
    }
    if (kMode != 1) {
      for (int i = threadIdx.x; i < 2 * G; i += blockDim.x) sh_acc[i] = 0.f;
      __syncthreads();
      // lanes V apart own the same columns: reduce them in registers first (float atomics on shared memory are
      // compare-and-swap loops: 64 threads per address made this flush 40 % of the kernel's stall cycles)
      for (int o = V; o < 32; o <<= 1) {
#pragma unroll
        for (int j = 0; j < N; ++j) {
          s[j] += __shfl_xor_sync(0xffffffffu, s[j], o);
          ss[j] += __shfl_xor_sync(0xffffffffu, ss[j], o);
        }
      }
      if (V >= 32 || (threadIdx.x & 31) < V) {
        if (cpg >= N) {  // the thread's columns share one group
          float ts = 0.f, tss = 0.f;
#pragma unroll
          for (int j = 0; j < N; ++j) {
            ts += s[j];
            tss += ss[j];
          }
          atomicAdd(&sh_acc[2 * (c0 / cpg)], ts);
          atomicAdd(&sh_acc[2 * (c0 / cpg) + 1], tss);
        } else {
#pragma unroll
          for (int j = 0; j < N; ++j) {
            atomicAdd(&sh_acc[2 * ((c0 + j) / cpg)], s[j]);
            atomicAdd(&sh_acc[2 * ((c0 + j) / cpg) + 1], ss[j]);
          }
        }
      }
      __syncthreads();
      for (int i = threadIdx.x; i < 2 * G; i += blockDim.x)
        atomicAdd(stats2_acc + (int64_t)seg * G * 2 + i, (double)sh_acc[i]);
      __syncthreads();
    }
    row0 = row1;
  }
}

// out[q][col] = max_n xpad[idx[q][n]][col]; shadow neighbours contribute 0 (blocks.py:100-109)
// A CTA handles kPoolQ query points; a thread owns 8 channels (one 16-byte load per neighbour row, two neighbours in
// flight) of one query.
constexpr int kPoolQ = 4;

__device__ __forceinline__ uint32_t max_bf16x2(uint32_t a, uint32_t b) {
  const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a), *reinterpret_cast<const __nv_bfloat162*>(&b));
  return *reinterpret_cast<const uint32_t*>(&r);
}
__device__ __forceinline__ void max_bf16x8_packed(uint4& m, const uint4& v) {
  m.x = max_bf16x2(m.x, v.x); m.y = max_bf16x2(m.y, v.y); m.z = max_bf16x2(m.z, v.z); m.w = max_bf16x2(m.w, v.w);
}
__device__ __forceinline__ void max_bf16x8(float (&m)[8], const uint4& v) {
  m[0] = fmaxf(m[0], bf_lo(v.x)); m[1] = fmaxf(m[1], bf_hi(v.x));
  m[2] = fmaxf(m[2], bf_lo(v.y)); m[3] = fmaxf(m[3], bf_hi(v.y));
  m[4] = fmaxf(m[4], bf_lo(v.z)); m[5] = fmaxf(m[5], bf_hi(v.z));
  m[6] = fmaxf(m[6], bf_lo(v.w)); m[7] = fmaxf(m[7], bf_hi(v.w));
}

__global__ void __launch_bounds__(256) maxpool_nbr_kernel(const __nv_bfloat16* __restrict__ x, int64_t ns, int width,
                                                           const int64_t* __restrict__ idx, int H_full, int64_t nq,
                                                           const int64_t* __restrict__ seg_off,
                                                           const int32_t* __restrict__ seg_width, int nseg,
                                                           __nv_bfloat16* __restrict__ out) {
  __shared__ int sh_idx[kPoolQ][kMaxH];
  __shared__ int sh_h[kPoolQ];
  const int64_t q0 = (int64_t)blockIdx.x * kPoolQ;
  for (int t = threadIdx.x; t < kPoolQ * H_full; t += blockDim.x) {
    const int qi = t / H_full, n = t - qi * H_full;
    const int64_t q = q0 + qi;
    int j = -1;
    if (q < nq) {
      const int64_t jj = idx[q * H_full + n];
      j = (jj >= 0 && jj < ns) ? (int)jj : -1;
    }
    sh_idx[qi][n] = j;
  }
  if (threadIdx.x < kPoolQ) {
    const int64_t q = q0 + threadIdx.x;
    int H = H_full;
    if (seg_width && q < nq) H = min(H_full, max(1, seg_width[segment_of(seg_off, nseg, q)]));
    sh_h[threadIdx.x] = H;
  }
  __syncthreads();
  const int vecs = width / 8;  // 16-byte chunks per row
  for (int u = threadIdx.x; u < kPoolQ * vecs; u += blockDim.x) {
    const int qi = u / vecs, v = u - qi * vecs;
    const int64_t q = q0 + qi;
    if (q >= nq) continue;
    const int H = sh_h[qi];
    // running maximum as four packed bf16 pairs (HMNMX2.BF16: the maximum of bf16 values is exact, so this equals the
    // fp32 maximum rounded back; 4 instructions per 16-byte row piece instead of 16), four neighbour rows in flight
    uint4 m = make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u);  // -inf
    bool shadow = false;
    const __nv_bfloat16* xv = x + 8 * v;
    int n = 0;
    for (; n + 3 < H; n += 4) {
      int j[4];
      uint4 r[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        j[t] = sh_idx[qi][n + t];
        r[t] = m;
        if (j[t] >= 0) r[t] = *reinterpret_cast<const uint4*>(xv + (int64_t)j[t] * width);
        shadow |= j[t] < 0;
      }
#pragma unroll
      for (int t = 0; t < 4; ++t) max_bf16x8_packed(m, r[t]);
    }
    for (; n < H; ++n) {
      const int j0 = sh_idx[qi][n];
      if (j0 >= 0) max_bf16x8_packed(m, *reinterpret_cast<const uint4*>(xv + (int64_t)j0 * width));
      else shadow = true;
    }
    if (shadow) max_bf16x8_packed(m, make_uint4(0u, 0u, 0u, 0u));  // the zero shadow row takes part in the max
    const uint4 o = m;
    *reinterpret_cast<uint4*>(out + q * width + 8 * v) = o;
  }
}

// two channels per thread: row widths that are not a multiple of 8
__global__ void __launch_bounds__(256) maxpool_nbr_scalar_kernel(const __nv_bfloat16* __restrict__ x, int64_t ns, int width,
                                                           const int64_t* __restrict__ idx, int H_full,
                                                           const int64_t* __restrict__ seg_off,
                                                           const int32_t* __restrict__ seg_width, int nseg,
                                                           __nv_bfloat16* __restrict__ out) {
  __shared__ int sh_idx[kMaxH];
  const int64_t q = blockIdx.x;
  int H = H_full;
  if (seg_width) H = min(H_full, max(1, seg_width[segment_of(seg_off, nseg, q)]));
  if (threadIdx.x < H) {
    const int64_t j = idx[q * H_full + threadIdx.x];
    sh_idx[threadIdx.x] = (j >= 0 && j < ns) ? (int)j : -1;
  }
  __syncthreads();
  for (int u = threadIdx.x; u < width / 2; u += blockDim.x) {
    float m0 = -INFINITY, m1 = -INFINITY;
    for (int n = 0; n < H; ++n) {
      const int j = sh_idx[n];
      float x0 = 0.f, x1 = 0.f;
      if (j >= 0) {
        const uint32_t v = *reinterpret_cast<const uint32_t*>(x + (int64_t)j * width + 2 * u);
        x0 = bf_lo(v);
        x1 = bf_hi(v);
      }
      m0 = fmaxf(m0, x0);
      m1 = fmaxf(m1, x1);
    }
    *reinterpret_cast<uint32_t*>(out + q * width + 2 * u) = pack_bf16(m0, m1);
  }
}

// [N, A, C] -> [N, C], max over anchors
__global__ void __launch_bounds__(256) anchor_max_kernel(const __nv_bfloat16* __restrict__ x, int64_t n, int A, int C,
                                                          __nv_bfloat16* __restrict__ out, int64_t out_ld) {
  const int half = C / 2;
  const int64_t total = n * half;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / half;
    const int c = (int)(i - p * half) * 2;
    float m0 = -INFINITY, m1 = -INFINITY;
    for (int a = 0; a < A; ++a) {
      const uint32_t v = *reinterpret_cast<const uint32_t*>(x + (p * A + a) * C + c);
      m0 = fmaxf(m0, bf_lo(v));
      m1 = fmaxf(m1, bf_hi(v));
    }
    *reinterpret_cast<uint32_t*>(out + p * out_ld + c) = pack_bf16(m0, m1);
  }
}

// out[i] = [ xpad[up_idx[i][0]] (c1) | y[i] (c2) ]   (nearest_upsample + cat)
__global__ void __launch_bounds__(256) upsample_concat_kernel(const __nv_bfloat16* __restrict__ x, int64_t nx, int c1,
                                                               const int64_t* __restrict__ up_idx, int up_ld,
                                                               const __nv_bfloat16* __restrict__ y, int c2, int64_t n,
                                                               __nv_bfloat16* __restrict__ out) {
  const int w = (c1 + c2) / 2;
  const int64_t total = n * w;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t p = i / w;
    const int c = (int)(i - p * w) * 2;
    uint32_t v = 0;
    if (c < c1) {
      const int64_t j = up_idx[p * up_ld];
      if (j >= 0 && j < nx) v = *reinterpret_cast<const uint32_t*>(x + j * c1 + c);
    } else {
      v = *reinterpret_cast<const uint32_t*>(y + p * c2 + (c - c1));
    }
    *reinterpret_cast<uint32_t*>(out + p * (c1 + c2) + c) = v;
  }
}

static inline int elementwise_blocks(int64_t work) {
  int64_t b = ceil_div(work, 256);
  const int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace se3et

using namespace se3et;

extern "C" int se3et_kpconv_tables(int32_t* kidx_15x6, int32_t* ridx_6x6) {
  if (!kidx_15x6 || !ridx_6x6) return SE3ET_ERR_ARG;
  for (int k = 0; k < 15; ++k)
    for (int r = 0; r < 6; ++r) kidx_15x6[k * 6 + r] = kidx_tab(k, r);
  for (int a = 0; a < 6; ++a)
    for (int r = 0; r < 6; ++r) ridx_6x6[a * 6 + r] = ridx_tab(a, r);
  return SE3ET_OK;
}

extern "C" int se3et_kpconv_gather(const float* q_pts, const float* s_pts, const int64_t* neighbors, int64_t nq,
                                   int64_t ns, int64_t h, const void* x_bf16, int64_t cin,
                                   const float* kernel_points_15x3, float kp_extent, void* out_bf16, int64_t kpad,
                                   se3et_stream_t stream) {
  if (nq < 0 || ns < 0 || h <= 0 || h > kMaxH || cin <= 0 || !(kp_extent > 0.f)) return SE3ET_ERR_ARG;
  if (kpad < 36 * cin || kpad % 8 != 0) return SE3ET_ERR_ARG;
  if (nq == 0) return SE3ET_OK;
  if (!q_pts || !s_pts || !neighbors || !x_bf16 || !kernel_points_15x3 || !out_bf16) return SE3ET_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const auto* x = static_cast<const __nv_bfloat16*>(x_bf16);
  auto* out = static_cast<__nv_bfloat16*>(out_bf16);
  const int width = 6 * (int)cin;
  if (cin % 16 == 0 && kpad == 36 * cin && ns > 0)  // tensor-core producer (kpconv.cu)
    return kpconv_gather_mma(q_pts, s_pts, neighbors, nq, ns, (int)h, x, (int)cin, kernel_points_15x3, 1.f / kp_extent,
                             out, (int)kpad, st);
  if (cin % 2 == 0) {
    const int threads = width / 2 >= 256 ? 256 : (width / 2 <= 64 ? 64 : 128);
    kpconv_gather_kernel<true><<<(unsigned)nq, threads, 0, st>>>(q_pts, s_pts, neighbors, (int)h, ns, x, (int)cin,
                                                                kernel_points_15x3, 1.f / kp_extent, out, (int)kpad);
  } else {
    kpconv_gather_kernel<false><<<(unsigned)nq, 64, 0, st>>>(q_pts, s_pts, neighbors, (int)h, ns, x, (int)cin,
                                                            kernel_points_15x3, 1.f / kp_extent, out, (int)kpad);
  }
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}

extern "C" int se3et_groupnorm_stats(const float* y, int64_t rows, int64_t channels, int64_t groups,
                                     const int64_t* seg_offsets, int64_t nseg, int64_t rows_per_point, double* stats,
                                     se3et_stream_t stream) {
  if (rows < 0 || channels <= 0 || groups <= 0 || channels % groups || nseg <= 0 || rows_per_point <= 0 || !stats ||
      !seg_offsets)
    return SE3ET_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SE3ET_CUDA_CHECK(cudaMemsetAsync(stats, 0, sizeof(double) * 2 * nseg * groups, st));
  if (rows == 0) return SE3ET_OK;
  if (!y) return SE3ET_ERR_ARG;
  const int cpg = (int)(channels / groups);
  const int threads = channels >= 256 ? 256 : (int)((channels + 31) / 32 * 32);
  dim3 grid((unsigned)ceil_div(rows, kStatRows), (unsigned)ceil_div(channels, threads));
  const size_t smem = sizeof(double) * 2 * (threads / (cpg < threads ? cpg : threads) + 2);
  groupnorm_stats_kernel<<<grid, threads, smem, st>>>(y, rows, (int)channels, cpg, seg_offsets, (int)nseg,
                                                  (int)rows_per_point, stats);
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}

extern "C" int se3et_groupnorm_apply(const float* ya, const double* stats_a, const float* gamma_a, const float* beta_a,
                                     const float* yb, const double* stats_b, const float* gamma_b, const float* beta_b,
                                     const void* resid_bf16, int64_t rows, int64_t channels, int64_t groups,
                                     const int64_t* seg_offsets, int64_t nseg, int64_t rows_per_point, float eps,
                                     float leaky_slope, float* out_f32, void* out_bf16, se3et_stream_t stream) {
  if (rows < 0 || channels <= 0 || channels % 4 || groups <= 0 || channels % groups || nseg <= 0 ||
      rows_per_point <= 0 || !seg_offsets)
    return SE3ET_ERR_ARG;
  if (rows == 0) return SE3ET_OK;
  if (!ya || !stats_a || !gamma_a || !beta_a || (!out_f32 && !out_bf16)) return SE3ET_ERR_ARG;
  if (yb && (!stats_b || !gamma_b || !beta_b)) return SE3ET_ERR_ARG;
  NormSide a{ya, stats_a, gamma_a, beta_a};
  NormSide b{yb, stats_b, gamma_b, beta_b};
  const int64_t vecs = channels / 4;
  if (vecs <= 256 && (vecs & (vecs - 1)) == 0) {
    // fixed-column streaming kernel; grid sized to a multiple of the SM count, every thread gets >= 1 row
    int64_t blocks = ceil_div(rows * vecs, 256);
    const int64_t cap = (int64_t)kNumSMs * 8;
    if (blocks > cap) blocks = cap;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto* rb = static_cast<const __nv_bfloat16*>(resid_bf16);
    auto* ob = static_cast<__nv_bfloat16*>(out_bf16);
    if (yb)
      groupnorm_apply_rows_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(
          a, b, rb, rows, (int)channels, (int)(channels / groups), seg_offsets, (int)nseg, (int)rows_per_point, eps,
          leaky_slope, out_f32, ob);
    else
      groupnorm_apply_rows_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(
          a, b, rb, rows, (int)channels, (int)(channels / groups), seg_offsets, (int)nseg, (int)rows_per_point, eps,
          leaky_slope, out_f32, ob);
    SE3ET_LAUNCH_CHECK();
    return SE3ET_OK;
  }
  groupnorm_apply_kernel<<<elementwise_blocks(rows * channels / 4), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      a, b, static_cast<const __nv_bfloat16*>(resid_bf16), rows, (int)channels, (int)(channels / groups), seg_offsets,
      (int)nseg, (int)rows_per_point, eps, leaky_slope, out_f32, static_cast<__nv_bfloat16*>(out_bf16));
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}

extern "C" int se3et_maxpool_nbr(const void* x_bf16, int64_t ns, int64_t width, const int64_t* neighbors, int64_t nq,
                                 int64_t h, const int64_t* seg_offsets, const int32_t* seg_width, int64_t nseg,
                                 void* out_bf16, se3et_stream_t stream) {
  if (seg_width && (!seg_offsets || nseg <= 0)) return SE3ET_ERR_ARG;
  if (nq < 0 || ns < 0 || h <= 0 || h > kMaxH || width <= 0 || width % 2) return SE3ET_ERR_ARG;
  if (nq == 0) return SE3ET_OK;
  if (!x_bf16 || !neighbors || !out_bf16) return SE3ET_ERR_ARG;
  if (width % 8 != 0) {
    const int threads = width / 2 >= 256 ? 256 : (width / 2 <= 64 ? 64 : 128);
    maxpool_nbr_scalar_kernel<<<(unsigned)nq, threads, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(x_bf16), ns, (int)width, neighbors, (int)h, seg_offsets, seg_width, (int)nseg,
        static_cast<__nv_bfloat16*>(out_bf16));
    SE3ET_LAUNCH_CHECK();
    return SE3ET_OK;
  }
  const int64_t work = kPoolQ * (width / 8);
  const int threads = work >= 256 ? 256 : (work <= 64 ? 64 : 128);
  maxpool_nbr_kernel<<<(unsigned)ceil_div(nq, kPoolQ), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x_bf16), ns, (int)width, neighbors, (int)h, nq, seg_offsets, seg_width,
      (int)nseg, static_cast<__nv_bfloat16*>(out_bf16));
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}

extern "C" int se3et_anchor_max(const void* x_bf16, int64_t n, int64_t anchors, int64_t channels, void* out_bf16,
                                int64_t out_ld, se3et_stream_t stream) {
  if (n < 0 || anchors <= 0 || channels <= 0 || channels % 2 || out_ld < channels || out_ld % 2) return SE3ET_ERR_ARG;
  if (n == 0) return SE3ET_OK;
  if (!x_bf16 || !out_bf16) return SE3ET_ERR_ARG;
  anchor_max_kernel<<<elementwise_blocks(n * channels / 2), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x_bf16), n, (int)anchors, (int)channels, static_cast<__nv_bfloat16*>(out_bf16),
      out_ld);
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}

extern "C" int se3et_upsample_concat(const void* x_bf16, int64_t nx, int64_t c1, const int64_t* up_idx, int64_t up_ld,
                                     const void* y_bf16, int64_t c2, int64_t n, void* out_bf16,
                                     se3et_stream_t stream) {
  if (n < 0 || nx < 0 || c1 <= 0 || c2 < 0 || c1 % 2 || c2 % 2 || up_ld <= 0) return SE3ET_ERR_ARG;
  if (n == 0) return SE3ET_OK;
  if (!x_bf16 || !up_idx || !out_bf16 || (c2 > 0 && !y_bf16)) return SE3ET_ERR_ARG;
  upsample_concat_kernel<<<elementwise_blocks(n * (c1 + c2) / 2), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const __nv_bfloat16*>(x_bf16), nx, (int)c1, up_idx, (int)up_ld,
      static_cast<const __nv_bfloat16*>(y_bf16), (int)c2, n, static_cast<__nv_bfloat16*>(out_bf16));
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}

extern "C" int se3et_groupnorm_double(const void* y_in, int y_bf16, const double* stats1, const float* gamma1, const float* beta1,
                                      double* stats2, const float* gamma2, const float* beta2, int64_t rows,
                                      int64_t channels, int64_t groups, const int64_t* seg_offsets, int64_t nseg,
                                      int64_t rows_per_point, float eps, float leaky_slope, int apply, void* out_bf16,
                                      se3et_stream_t stream) {
  if (rows < 0 || channels <= 0 || channels % 4 || groups <= 0 || channels % groups || nseg <= 0 ||
      rows_per_point <= 0 || !seg_offsets || !stats2)
    return SE3ET_ERR_ARG;
  const int64_t vecs = channels / 4;
  if (vecs > 256 || (vecs & (vecs - 1)) != 0 || groups > 256) return SE3ET_ERR_UNSUPPORTED;
  const float* y = static_cast<const float*>(y_in);   // reinterpreted by the kernel when y_bf16
  if (y_bf16 && (channels % 8 || vecs < 2)) return SE3ET_ERR_UNSUPPORTED;
  if (!(leaky_slope >= 0.f && leaky_slope <= 1.f)) return SE3ET_ERR_UNSUPPORTED;   // max(w, slope * w) form
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (apply != 1) SE3ET_CUDA_CHECK(cudaMemsetAsync(stats2, 0, sizeof(double) * 2 * nseg * groups, st));
  if (rows == 0) return SE3ET_OK;
  if (apply < 0 || apply > 2) return SE3ET_ERR_ARG;
  if (!y || (apply != 2 && (!stats1 || !gamma1 || !beta1)) || (apply == 1 && (!gamma2 || !beta2 || !out_bf16)))
    return SE3ET_ERR_ARG;
  int64_t blocks = ceil_div(rows * (y_bf16 ? vecs / 2 : vecs), 256);
  const int64_t cap = (int64_t)kNumSMs * 8;
  if (blocks > cap) blocks = cap;
  NormSide a{y, stats1, gamma1, beta1};
  NormSide b{nullptr, stats2, gamma2, beta2};
  const int cpg = (int)(channels / groups);
#define SE3ET_GND(MODE, OUT)                                                                                          \
  do {                                                                                                                \
    if (y_bf16)                                                                                                       \
      groupnorm_double_kernel<MODE, true><<<(unsigned)blocks, 256, 0, st>>>(                                          \
          a, b, stats2, rows, (int)channels, cpg, seg_offsets, (int)nseg, (int)rows_per_point, eps, leaky_slope, OUT); \
    else                                                                                                              \
      groupnorm_double_kernel<MODE, false><<<(unsigned)blocks, 256, 0, st>>>(                                         \
          a, b, stats2, rows, (int)channels, cpg, seg_offsets, (int)nseg, (int)rows_per_point, eps, leaky_slope, OUT); \
  } while (0)
  if (apply == 1)
    SE3ET_GND(1, static_cast<__nv_bfloat16*>(out_bf16));
  else if (apply == 0)
    SE3ET_GND(0, nullptr);
  else
    SE3ET_GND(2, nullptr);
#undef SE3ET_GND
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}
