// SuperPointMatching (geotransformer/superpoint_matching.py:13-55) for a batch of pairs:
//   E = exp(-clamp(2 - 2 <r, s>, 0))  on unit features, masked rows / columns removed,
//   score = (E / rowsum) * (E / colsum)            (dual normalisation)
//   global top-k over the flattened matrix with the canonical order (score desc, flat index asc).
// Deterministic: row / column sums are fixed-order reductions (no atomics), selection is an exact radix select.
#include "common.cuh"

namespace se3et {

struct MatchProblem {  // device table, one per pair
  int64_t ref_start, n_ref, src_start, n_src, e_off;  // e_off: offset of this pair's n_ref x n_src matrix
};

// one CTA per (ref row, pair): E row + its sum
__global__ void __launch_bounds__(256) spm_exp_rows_kernel(const float* __restrict__ ref, const float* __restrict__ src,
                                                            int C, const uint8_t* __restrict__ ref_mask,
                                                            const uint8_t* __restrict__ src_mask,
                                                            const MatchProblem* __restrict__ problems,
                                                            float* __restrict__ e, float* __restrict__ row_sum) {
  extern __shared__ float sh_row[];  // ref feature row [C], then per-warp partial sums
  const MatchProblem pr = problems[blockIdx.y];
  const int n = blockIdx.x;
  if (n >= pr.n_ref) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const float* r = ref + (pr.ref_start + n) * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) sh_row[c] = r[c];
  __syncthreads();
  const bool row_on = !ref_mask || ref_mask[pr.ref_start + n];
  float* erow = e + pr.e_off + (int64_t)n * pr.n_src;
  float part = 0.f;  // this warp's columns, ascending order
  for (int m = warp; m < pr.n_src; m += nwarps) {
    const float* s = src + (pr.src_start + m) * C;
    float dot = 0.f;
    for (int c = lane; c < C; c += 32) dot = fmaf(sh_row[c], s[c], dot);
    dot = warp_sum(dot);
    const bool on = row_on && (!src_mask || src_mask[pr.src_start + m]);
    const float v = on ? expf(-fmaxf(2.f - 2.f * dot, 0.f)) : 0.f;
    if (lane == 0) erow[m] = v;
    part += v;
  }
  float* sh_part = sh_row + C;
  if (lane == 0) sh_part[warp] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < nwarps; ++w) t += sh_part[w];
    row_sum[pr.ref_start + n] = t;
  }
}

// Register-tiled form of the same pass (round 2): a CTA owns 32 ref rows and walks the src rows 128 at a time; the ref
// block stays in shared memory (k-major), a thread accumulates a 4 x 4 block over k in ascending order (fp32 FMA: the
// dot products of unit features need fp32 for the 1e-4 score tolerance, so no bf16 tensor-core product), the row sums
// are a fixed-order reduction (warp tree per column tile, tiles in ascending order): deterministic, 8x fewer
// instructions than one warp-reduced dot product per matrix entry.
constexpr int kMatRows = 32, kMatCols = 128, kMatK = 16, kMatMaxC = 256;

__global__ void __launch_bounds__(256) spm_exp_tiles_kernel(const float* __restrict__ ref, const float* __restrict__ src,
                                                             int C, const uint8_t* __restrict__ ref_mask,
                                                             const uint8_t* __restrict__ src_mask,
                                                             const MatchProblem* __restrict__ problems,
                                                             float* __restrict__ e, float* __restrict__ row_sum) {
  __shared__ __align__(16) float As[kMatMaxC][kMatRows];  // [k][ref row]
  __shared__ __align__(16) float Bs[kMatK][kMatCols];     // [k][src row]
  const MatchProblem pr = problems[blockIdx.y];
  const int n0 = blockIdx.x * kMatRows;
  if (n0 >= pr.n_ref) return;
  const int tid = threadIdx.x, lane = tid & 31, ty = tid >> 5;  // rows 4 ty .. 4 ty + 3, columns 4 lane .. 4 lane + 3
  const int n_src = (int)pr.n_src, n_ref = (int)pr.n_ref;
  // the ref block, transposed: thread -> (row = i % 32, four consecutive k)
  for (int i = tid; i < kMatRows * (C / 4); i += 256) {
    const int row = i & 31, k4 = (i >> 5) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n0 + row < n_ref) v = *reinterpret_cast<const float4*>(ref + (pr.ref_start + n0 + row) * C + k4);
    As[k4][row] = v.x; As[k4 + 1][row] = v.y; As[k4 + 2][row] = v.z; As[k4 + 3][row] = v.w;
  }
  bool row_on[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + 4 * ty + i;
    row_on[i] = n < n_ref && (!ref_mask || ref_mask[pr.ref_start + n]);
  }
  float rsum[4] = {0.f, 0.f, 0.f, 0.f};
  for (int m0 = 0; m0 < n_src; m0 += kMatCols) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < C; k0 += kMatK) {
      __syncthreads();  // Bs (and, first time, As) ready to be overwritten / read
      for (int i = tid; i < kMatCols * (kMatK / 4); i += 256) {
        const int col = i & (kMatCols - 1), k4 = (i / kMatCols) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m0 + col < n_src) v = *reinterpret_cast<const float4*>(src + (pr.src_start + m0 + col) * C + k0 + k4);
        Bs[k4][col] = v.x; Bs[k4 + 1][col] = v.y; Bs[k4 + 2][col] = v.z; Bs[k4 + 3][col] = v.w;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < kMatK; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(&As[k0 + k][4 * ty]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[k][4 * lane]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
    }
    bool col_on[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + 4 * lane + j;
      col_on[j] = m < n_src && (!src_mask || src_mask[pr.src_start + m]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = n0 + 4 * ty + i;
      float part = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int m = m0 + 4 * lane + j;
        const float v = (row_on[i] && col_on[j]) ? expf(-fmaxf(2.f - 2.f * acc[i][j], 0.f)) : 0.f;
        if (n < n_ref && m < n_src) e[pr.e_off + (int64_t)n * n_src + m] = v;
        part += v;
      }
      rsum[i] += warp_sum(part);  // the same value in every lane; column tiles in ascending order
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (n0 + 4 * ty + i < n_ref) row_sum[pr.ref_start + n0 + 4 * ty + i] = rsum[i];
  }
}

// one thread per column (fixed-order sum over rows)
__global__ void __launch_bounds__(128) spm_col_sums_kernel(const float* __restrict__ e,
                                                            const MatchProblem* __restrict__ problems,
                                                            float* __restrict__ col_sum) {
  const MatchProblem pr = problems[blockIdx.y];
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= pr.n_src) return;
  const float* col = e + pr.e_off + m;
  float t = 0.f;
  for (int n = 0; n < pr.n_ref; ++n) t += col[(int64_t)n * pr.n_src];
  col_sum[pr.src_start + m] = t;
}

constexpr int kSelThreads = 1024;
constexpr int kSelMaxK = 1024;

__device__ __forceinline__ float spm_score(const float* __restrict__ e, const float* __restrict__ rs,
                                           const float* __restrict__ cs, int64_t i, int n_src, bool dual) {
  const float v = e[i];
  if (!dual) return v;
  const int n = (int)(i / n_src), m = (int)(i - (int64_t)n * n_src);
  const float r = rs[n], c = cs[m];
  if (!(r > 0.f) || !(c > 0.f)) return 0.f;  // masked row / column
  return (v / r) * (v / c);
}

// block-wide exclusive prefix sum of a 0/1 flag, returns total through `total`
__device__ __forceinline__ int block_prefix(int flag, int* sh_warp, int& total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned ballot = __ballot_sync(0xffffffffu, flag);
  const int in_warp = __popc(ballot & ((1u << lane) - 1u));
  if (lane == 0) sh_warp[warp] = __popc(ballot);
  __syncthreads();
  int off = 0, tot = 0;
  for (int w = 0; w < kSelThreads / 32; ++w) {
    const int c = sh_warp[w];
    if (w < warp) off += c;
    tot += c;
  }
  __syncthreads();
  total = tot;
  return off + in_warp;
}

constexpr int kSelSlices = 8;  // CTAs per pair in the selection (a pair alone would leave most of the 148 SMs idle)

// one CTA per (slice, pair): the slice's top-k by the canonical order (exact k-th largest by 3-pass radix select on the
// float bits, ordered collection, bitonic sort) as 64-bit keys (~score bits, flat index); the global top-k is a subset
// of the union of the slices' top-k lists, spm_merge_kernel picks it
__global__ void __launch_bounds__(kSelThreads) spm_topk_kernel(const float* __restrict__ e,
                                                                const float* __restrict__ row_sum,
                                                                const float* __restrict__ col_sum,
                                                                const uint8_t* __restrict__ ref_mask,
                                                                const uint8_t* __restrict__ src_mask,
                                                                const MatchProblem* __restrict__ problems, int k_req,
                                                                int dual, unsigned long long* __restrict__ slice_keys,
                                                                int32_t* __restrict__ slice_counts) {
  __shared__ unsigned int hist[2048];
  __shared__ int sh_warp[kSelThreads / 32];
  __shared__ unsigned long long cand[kSelMaxK];
  __shared__ unsigned int sh_prefix, sh_remaining;
  __shared__ int sh_valid;
  const MatchProblem pr = problems[blockIdx.y];
  const int n_src = (int)pr.n_src;
  const int64_t all = pr.n_ref * pr.n_src;
  const int64_t i_begin = all * blockIdx.x / gridDim.x, i_end = all * (blockIdx.x + 1) / gridDim.x;
  const float* ep = e + pr.e_off;
  const float* rs = row_sum + pr.ref_start;
  const float* cs = col_sum + pr.src_start;
  unsigned long long* out_k = slice_keys + ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * k_req;

  // number of unmasked entries of the slice
  if (threadIdx.x == 0) sh_valid = 0;
  __syncthreads();
  {
    int v = 0;
    for (int64_t i = i_begin + threadIdx.x; i < i_end; i += kSelThreads) {
      const int n = (int)(i / n_src), m = (int)(i - (int64_t)n * n_src);
      v += ((ref_mask && !ref_mask[pr.ref_start + n]) || (src_mask && !src_mask[pr.src_start + m])) ? 0 : 1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(&sh_valid, v);
    __syncthreads();
    const int nvalid = sh_valid;
    __syncthreads();
    if (threadIdx.x == 0) sh_valid = nvalid < k_req ? nvalid : k_req;
    __syncthreads();
  }
  const int k = sh_valid;  // this slice contributes at most min(num_correspondences, its unmasked entries) candidates
  if (threadIdx.x == 0) slice_counts[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = k;
  if (k == 0) return;

  // ---- radix select: bits of a non-negative float order like unsigned integers
  unsigned int prefix = 0, prefix_mask = 0;
  unsigned int remaining = (unsigned)k;  // rank (1-based, from the top) inside the current bucket
  const int shifts[3] = {21, 10, 0};
  const int widths[3] = {11, 11, 10};
  for (int pass = 0; pass < 3; ++pass) {
    const int shift = shifts[pass], nb = 1 << widths[pass];
    for (int i = threadIdx.x; i < nb; i += kSelThreads) hist[i] = 0;
    __syncthreads();
    for (int64_t i = i_begin + threadIdx.x; i < i_end; i += kSelThreads) {
      const int n = (int)(i / n_src), m = (int)(i - (int64_t)n * n_src);
      if ((ref_mask && !ref_mask[pr.ref_start + n]) || (src_mask && !src_mask[pr.src_start + m])) continue;
      const unsigned int bits = __float_as_uint(spm_score(ep, rs, cs, i, n_src, dual));
      if ((bits & prefix_mask) == prefix) atomicAdd(&hist[(bits >> shift) & (nb - 1)], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int acc = 0;
      int b = nb - 1;
      for (; b > 0; --b) {
        if (acc + hist[b] >= remaining) break;
        acc += hist[b];
      }
      sh_prefix = prefix | ((unsigned)b << shift);
      sh_remaining = remaining - acc;
    }
    __syncthreads();
    prefix = sh_prefix;
    remaining = sh_remaining;
    prefix_mask |= (unsigned)(nb - 1) << shift;
    __syncthreads();
  }
  const unsigned int thr_bits = prefix;  // bits of the k-th largest score; `remaining` of the ties are taken

  // ---- ordered collection (ascending flat index): everything above the threshold, then the first ties
  int n_above = 0, n_tie = 0;
  for (int64_t base = i_begin; base < i_end; base += kSelThreads) {
    const int64_t i = base + threadIdx.x;
    unsigned int bits = 0;
    bool valid = false;
    if (i < i_end) {
      const int n = (int)(i / n_src), m = (int)(i - (int64_t)n * n_src);
      valid = !((ref_mask && !ref_mask[pr.ref_start + n]) || (src_mask && !src_mask[pr.src_start + m]));
      if (valid) bits = __float_as_uint(spm_score(ep, rs, cs, i, n_src, dual));
    }
    int tot;
    const int above = valid && bits > thr_bits;
    int pos = block_prefix(above, sh_warp, tot);
    if (above) cand[n_above + pos] = ((unsigned long long)(~bits) << 32) | (unsigned int)i;
    n_above += tot;
    const int tie = valid && bits == thr_bits;
    pos = block_prefix(tie, sh_warp, tot);
    // ties go after all "above" entries: slots [k - remaining, k)
    if (tie && n_tie + pos < (int)remaining)
      cand[k - (int)remaining + n_tie + pos] = ((unsigned long long)(~bits) << 32) | (unsigned int)i;
    n_tie += tot;
  }
  __syncthreads();
  // ---- sort the k candidates by (score desc, flat index asc): key = (~bits, index) ascending
  int n2 = 1;
  while (n2 < k) n2 <<= 1;
  for (int i = k + threadIdx.x; i < n2; i += kSelThreads) cand[i] = ~0ull;
  __syncthreads();
  for (int kk = 2; kk <= n2; kk <<= 1) {
    for (int j = kk >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n2; i += kSelThreads) {
        const int p = i ^ j;
        if (p > i) {
          const unsigned long long x = cand[i], y = cand[p];
          const bool up = (i & kk) == 0;
          if ((x > y) == up) { cand[i] = y; cand[p] = x; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < k; i += kSelThreads) out_k[i] = cand[i];
}

// one CTA per pair: the sorted candidate lists of its slices -> the pair's top-k in the canonical order
__global__ void __launch_bounds__(kSelThreads) spm_merge_kernel(const unsigned long long* __restrict__ slice_keys,
                                                                 const int32_t* __restrict__ slice_counts, int nslices,
                                                                 const MatchProblem* __restrict__ problems, int k_req,
                                                                 int64_t* __restrict__ ref_idx,
                                                                 int64_t* __restrict__ src_idx, float* __restrict__ scores,
                                                                 int32_t* __restrict__ counts) {
  extern __shared__ unsigned long long keys[];  // [nslices * k_req rounded up to a power of two]
  const MatchProblem pr = problems[blockIdx.x];
  const int n_src = (int)pr.n_src;
  int n = 0;
  for (int sl = 0; sl < nslices; ++sl) {
    const int c = slice_counts[(int64_t)blockIdx.x * nslices + sl];
    const unsigned long long* src = slice_keys + ((int64_t)blockIdx.x * nslices + sl) * k_req;
    for (int i = threadIdx.x; i < c; i += kSelThreads) keys[n + i] = src[i];
    n += c;
  }
  int n2 = 1;
  while (n2 < n) n2 <<= 1;
  for (int i = n + threadIdx.x; i < n2; i += kSelThreads) keys[i] = ~0ull;
  __syncthreads();
  for (int kk = 2; kk <= n2; kk <<= 1) {
    for (int j = kk >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n2; i += kSelThreads) {
        const int p = i ^ j;
        if (p > i) {
          const unsigned long long x = keys[i], y = keys[p];
          const bool up = (i & kk) == 0;
          if ((x > y) == up) { keys[i] = y; keys[p] = x; }
        }
      }
      __syncthreads();
    }
  }
  const int k = n < k_req ? n : k_req;  // min(num_correspondences, #unmasked entries) (superpoint_matching.py:43)
  if (threadIdx.x == 0) counts[blockIdx.x] = k;
  int64_t* out_r = ref_idx + (int64_t)blockIdx.x * k_req;
  int64_t* out_s = src_idx + (int64_t)blockIdx.x * k_req;
  float* out_v = scores + (int64_t)blockIdx.x * k_req;
  for (int i = threadIdx.x; i < k_req; i += kSelThreads) {
    if (i < k) {
      const unsigned long long c = keys[i];
      const unsigned int flat = (unsigned int)(c & 0xffffffffull);
      out_r[i] = flat / n_src;
      out_s[i] = flat % n_src;
      out_v[i] = __uint_as_float(~(unsigned int)(c >> 32));
    } else {
      out_r[i] = -1;
      out_s[i] = -1;
      out_v[i] = 0.f;
    }
  }
}

}  // namespace se3et

using namespace se3et;

extern "C" int se3et_superpoint_matching_workspace_floats(int64_t num_pairs, int64_t num_correspondences, int64_t e_total,
                                                          int64_t* floats) {
  if (!floats || num_pairs < 0 || num_correspondences <= 0 || e_total < 0) return SE3ET_ERR_ARG;
  // score matrices, then per (pair, slice): k 64-bit candidate keys; then the slice counts
  *floats = (e_total + 3) / 4 * 4 + num_pairs * kSelSlices * (2 * num_correspondences + 1) + 4;
  return SE3ET_OK;
}

extern "C" int se3et_superpoint_matching(const float* ref_feats, const float* src_feats, int64_t channels,
                                         const uint8_t* ref_masks, const uint8_t* src_masks, const int64_t* problems,
                                         int64_t num_pairs, int64_t max_ref, int64_t max_src,
                                         int64_t num_correspondences, int dual_normalization, float* e_workspace,
                                         int64_t e_total, float* row_sums, float* col_sums, int64_t* ref_idx, int64_t* src_idx,
                                         float* scores, int32_t* counts, se3et_stream_t stream) {
  if (num_pairs < 0 || channels <= 0 || max_ref < 0 || max_src < 0 || num_correspondences <= 0 ||
      num_correspondences > kSelMaxK || num_pairs > 65535)
    return SE3ET_ERR_ARG;
  if (num_pairs == 0) return SE3ET_OK;
  if (!ref_feats || !src_feats || !problems || !e_workspace || !row_sums || !col_sums || !ref_idx || !src_idx ||
      !scores || !counts)
    return SE3ET_ERR_ARG;
  if (max_ref * max_src >= ((int64_t)1 << 32)) return SE3ET_ERR_UNSUPPORTED;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const auto* pr = reinterpret_cast<const MatchProblem*>(problems);
  if (max_ref > 0 && max_src > 0) {
    const bool tiles = channels % 16 == 0 && channels <= kMatMaxC && !(reinterpret_cast<uintptr_t>(ref_feats) & 15) &&
                       !(reinterpret_cast<uintptr_t>(src_feats) & 15);
    if (tiles) {
      dim3 g1((unsigned)ceil_div(max_ref, kMatRows), (unsigned)num_pairs);
      spm_exp_tiles_kernel<<<g1, 256, 0, st>>>(ref_feats, src_feats, (int)channels, ref_masks, src_masks, pr,
                                               e_workspace, row_sums);
    } else {
      dim3 g1((unsigned)max_ref, (unsigned)num_pairs);
      spm_exp_rows_kernel<<<g1, 256, sizeof(float) * (channels + 8), st>>>(ref_feats, src_feats, (int)channels,
                                                                           ref_masks, src_masks, pr, e_workspace,
                                                                           row_sums);
    }
    SE3ET_LAUNCH_CHECK();
    dim3 g2((unsigned)ceil_div(max_src, 128), (unsigned)num_pairs);
    spm_col_sums_kernel<<<g2, 128, 0, st>>>(e_workspace, pr, col_sums);
    SE3ET_LAUNCH_CHECK();
  }
  if ((reinterpret_cast<uintptr_t>(e_workspace) & 15) || e_total < 0) return SE3ET_ERR_ARG;
  unsigned long long* slice_keys = reinterpret_cast<unsigned long long*>(e_workspace + (e_total + 3) / 4 * 4);
  int32_t* slice_counts = reinterpret_cast<int32_t*>(slice_keys + num_pairs * kSelSlices * num_correspondences);
  spm_topk_kernel<<<dim3(kSelSlices, (unsigned)num_pairs), kSelThreads, 0, st>>>(
      e_workspace, row_sums, col_sums, ref_masks, src_masks, pr, (int)num_correspondences, dual_normalization,
      slice_keys, slice_counts);
  SE3ET_LAUNCH_CHECK();
  int n2 = 1;
  while (n2 < kSelSlices * num_correspondences) n2 <<= 1;
  const size_t msmem = sizeof(unsigned long long) * (size_t)n2;
  if (msmem > 48 * 1024) SE3ET_ENSURE_SMEM(spm_merge_kernel, msmem);
  spm_merge_kernel<<<(unsigned)num_pairs, kSelThreads, msmem, st>>>(slice_keys, slice_counts, kSelSlices, pr,
                                                                   (int)num_correspondences, ref_idx, src_idx, scores,
                                                                   counts);
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}
