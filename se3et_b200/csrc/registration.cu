// Fine-stage registration on the GPU: LocalGlobalRegistration.forward
// (geotransformer/modules/geotransformer/local_global_registration.py:49-235, call site
// experiments/se3eti.3dmatch/model.py:208-224) with weighted_procrustes
// (geotransformer/modules/registration/procrustes.py:6-73), for all patch correspondences of all pairs of a launch
// sequence at once.
//
//   lgr_corr_kernel    one CTA per patch correspondence: exp(log score), mutual top-k (score desc, index asc),
//                      confidence threshold and masks -> the patch's correspondences in row-major order (what
//                      torch.nonzero gives), in fixed slots [patch][k * K]
//   lgr_local_kernel   one CTA per patch: weighted Procrustes on its correspondences (if >= correspondence_threshold),
//                      then the number of correspondences OF THE WHOLE PAIR the transform brings within the acceptance
//                      radius
//   lgr_global_kernel  one CTA per pair: first best patch transform -> inlier-weighted Procrustes over all
//                      correspondences, num_refinement_steps times; writes the compacted correspondences
// The rotation is the proper rotation maximising trace(R H) (what V diag(1, 1, det) U^T of the reference's SVD is),
// found as the dominant eigenvector of Horn's 4 x 4 quaternion matrix by cyclic Jacobi sweeps in fp64; all moment sums
// are fp64 (the reference runs fp32 sums and a CPU LAPACK SVD: agreement ~1e-5 on the transform).
#include <cuda_runtime.h>
#include <math.h>

#include "common.cuh"

namespace se3et {
namespace lgr {

constexpr int kMaxTopk = 4;
constexpr int kThreads = 128;

struct Params {
  const float* log_scores;  // [B, ld, ld] (top-left K x K used)
  const float* ref_pts;     // [B, K, 3]
  const float* src_pts;     // [B, K, 3]
  const uint8_t* ref_mask;  // [B, K]
  const uint8_t* src_mask;  // [B, K]
  const int64_t* patch_off; // [P + 1] patches of pair i: [patch_off[i], patch_off[i + 1])
  int B, K, ld, P, topk, cap;
  float conf_thr, radius;
  int corr_thr, steps;
  // slotted correspondences
  int32_t* slot_ref;   // [B, cap]
  int32_t* slot_src;   // [B, cap]
  float* slot_score;   // [B, cap]
  int32_t* count;      // [B]
  // local stage
  double* local_T;     // [B, 12] row-major R | t
  int32_t* inliers;    // [B]  (-1: patch below the correspondence threshold)
  // outputs
  const int64_t* corr_off;  // [B + 1] exclusive scan of count (host: torch.cumsum)
  float* out_ref;      // [sum count, 3]
  float* out_src;      // [sum count, 3]
  float* out_score;    // [sum count]
  float* out_T;        // [P, 16]
};

// ---- weighted Procrustes from moment sums ---------------------------------------------------------------------
// m[0] = sum w, m[1..3] = sum w s, m[4..6] = sum w r, m[7..15] = sum w s r^T (row-major, s index first)
__device__ void solve_procrustes(const double* m, double eps, double* T12) {
  const double denom = m[0] + eps;
  double cs[3], cr[3], H[3][3];
  for (int i = 0; i < 3; ++i) { cs[i] = m[1 + i] / denom; cr[i] = m[4 + i] / denom; }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      H[i][j] = (m[7 + 3 * i + j] - cs[i] * m[4 + j] - m[1 + i] * cr[j] + m[0] * cs[i] * cr[j]) / denom;
  // Horn's quaternion matrix for the rotation s -> r maximising sum w r . (R s) = trace(R H)
  double N[4][4];
  N[0][0] = H[0][0] + H[1][1] + H[2][2];
  N[1][1] = H[0][0] - H[1][1] - H[2][2];
  N[2][2] = -H[0][0] + H[1][1] - H[2][2];
  N[3][3] = -H[0][0] - H[1][1] + H[2][2];
  N[0][1] = N[1][0] = H[1][2] - H[2][1];
  N[0][2] = N[2][0] = H[2][0] - H[0][2];
  N[0][3] = N[3][0] = H[0][1] - H[1][0];
  N[1][2] = N[2][1] = H[0][1] + H[1][0];
  N[1][3] = N[3][1] = H[2][0] + H[0][2];
  N[2][3] = N[3][2] = H[1][2] + H[2][1];
  double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
  for (int sweep = 0; sweep < 24; ++sweep) {
    double off = 0.0;
    for (int p = 0; p < 4; ++p)
      for (int q = p + 1; q < 4; ++q) off += N[p][q] * N[p][q];
    if (off < 1e-300) break;
    for (int p = 0; p < 4; ++p)
      for (int q = p + 1; q < 4; ++q) {
        if (fabs(N[p][q]) < 1e-300) continue;
        const double theta = (N[q][q] - N[p][p]) / (2.0 * N[p][q]);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 4; ++k) {  // columns p, q
          const double akp = N[k][p], akq = N[k][q];
          N[k][p] = c * akp - s * akq;
          N[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 4; ++k) {  // rows p, q
          const double apk = N[p][k], aqk = N[q][k];
          N[p][k] = c * apk - s * aqk;
          N[q][k] = s * apk + c * aqk;
        }
        for (int k = 0; k < 4; ++k) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int best = 0;
  for (int i = 1; i < 4; ++i)
    if (N[i][i] > N[best][best]) best = i;
  double qw = V[0][best], qx = V[1][best], qy = V[2][best], qz = V[3][best];
  const double qn = sqrt(qw * qw + qx * qx + qy * qy + qz * qz);
  qw /= qn; qx /= qn; qy /= qn; qz /= qn;
  double R[3][3] = {{1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qz * qw), 2 * (qx * qz + qy * qw)},
                    {2 * (qx * qy + qz * qw), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qx * qw)},
                    {2 * (qx * qz - qy * qw), 2 * (qy * qz + qx * qw), 1 - 2 * (qx * qx + qy * qy)}};
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) T12[4 * i + j] = R[i][j];
    T12[4 * i + 3] = cr[i] - (R[i][0] * cs[0] + R[i][1] * cs[1] + R[i][2] * cs[2]);
  }
}

__device__ __forceinline__ void add_moments(double* m, double w, const float* s, const float* r) {
  m[0] += w;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    m[1 + i] += w * s[i];
    m[4 + i] += w * r[i];
#pragma unroll
    for (int j = 0; j < 3; ++j) m[7 + 3 * i + j] += w * s[i] * r[j];
  }
}

// block-wide sum of 16 doubles per thread -> sh_out[16] (all threads may read it after the call)
__device__ void block_sum16(double* m, double* sh_red, double* sh_out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    double v = m[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) sh_red[warp * 16 + i] = v;
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    double v = 0.0;
    for (int w = 0; w < nw; ++w) v += sh_red[w * 16 + threadIdx.x];
    sh_out[threadIdx.x] = v;
  }
  __syncthreads();
}

__device__ __forceinline__ bool within(const double* T, const float* s, const float* r, float radius) {
  // fp32 like the reference's apply_transform + norm on fp32 tensors
  const float x = (float)T[0] * s[0] + (float)T[1] * s[1] + (float)T[2] * s[2] + (float)T[3] - r[0];
  const float y = (float)T[4] * s[0] + (float)T[5] * s[1] + (float)T[6] * s[2] + (float)T[7] - r[1];
  const float z = (float)T[8] * s[0] + (float)T[9] * s[1] + (float)T[10] * s[2] + (float)T[11] - r[2];
  return sqrtf(x * x + y * y + z * z) < radius;
}

// ---- kernel 1: correspondences of one patch pair ----------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) lgr_corr_kernel(Params p) {
  extern __shared__ float sh[];                 // [K][K + 1] exp scores
  const int K = p.K, pitch = K + 1;
  int* rowtop = reinterpret_cast<int*>(sh + K * pitch);   // [K][kMaxTopk]
  int* coltop = rowtop + K * kMaxTopk;                    // [K][kMaxTopk]
  int* scan = coltop + K * kMaxTopk;                      // [K + 1]
  const int b = blockIdx.x;
  const float* src = p.log_scores + (int64_t)b * p.ld * p.ld;
  for (int i = threadIdx.x; i < K * K; i += blockDim.x) {
    const int r = i / K, c = i - r * K;
    sh[r * pitch + c] = expf(src[(int64_t)r * p.ld + c]);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 2 * K; t += blockDim.x) {
    const bool is_row = t < K;
    const int line = is_row ? t : t - K;
    float bv[kMaxTopk];
    int bi[kMaxTopk];
#pragma unroll
    for (int j = 0; j < kMaxTopk; ++j) { bv[j] = -INFINITY; bi[j] = -1; }
    for (int e = 0; e < K; ++e) {
      const float v = is_row ? sh[line * pitch + e] : sh[e * pitch + line];
      // insert keeping (value desc, index asc): strict > leaves earlier indices ahead on ties
      if (v > bv[p.topk - 1] || bi[p.topk - 1] < 0) {
        int pos = p.topk - 1;
        while (pos > 0 && (bi[pos - 1] < 0 || v > bv[pos - 1])) { bv[pos] = bv[pos - 1]; bi[pos] = bi[pos - 1]; --pos; }
        bv[pos] = v;
        bi[pos] = e;
      }
    }
    int* dst = (is_row ? rowtop : coltop) + line * kMaxTopk;
    for (int j = 0; j < kMaxTopk; ++j) dst[j] = j < p.topk ? bi[j] : -1;
  }
  __syncthreads();
  // kept columns of row r, ascending
  int kept[kMaxTopk], nk = 0;
  const int r = threadIdx.x;
  if (r < K && p.ref_mask[(int64_t)b * K + r]) {
    for (int j = 0; j < p.topk; ++j) {
      const int c = rowtop[r * kMaxTopk + j];
      if (c < 0) continue;
      if (!(sh[r * pitch + c] > p.conf_thr) || !p.src_mask[(int64_t)b * K + c]) continue;
      bool mutual = false;
      for (int q = 0; q < p.topk; ++q) mutual |= coltop[c * kMaxTopk + q] == r;
      if (!mutual) continue;
      int pos = nk++;
      while (pos > 0 && kept[pos - 1] > c) { kept[pos] = kept[pos - 1]; --pos; }
      kept[pos] = c;
    }
  }
  if (r < K) scan[r + 1] = nk;
  if (r == 0) scan[0] = 0;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < K; ++i) scan[i + 1] += scan[i];
    p.count[b] = scan[K];
  }
  __syncthreads();
  if (r < K) {
    const int base = scan[r];
    for (int j = 0; j < nk; ++j) {
      p.slot_ref[(int64_t)b * p.cap + base + j] = r;
      p.slot_src[(int64_t)b * p.cap + base + j] = kept[j];
      p.slot_score[(int64_t)b * p.cap + base + j] = sh[r * pitch + kept[j]];
    }
  }
}

__device__ __forceinline__ int pair_of_patch(const int64_t* off, int P, int b) {
  int lo = 0, hi = P - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (off[mid] <= b) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// ---- kernel 2: local transforms and their support ---------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) lgr_local_kernel(Params p) {
  __shared__ double sh_red[(kThreads / 32) * 16], sh_m[16], sh_T[12];
  __shared__ int sh_cnt[kThreads / 32];
  const int b = blockIdx.x;
  const int n = p.count[b];
  if (n < p.corr_thr) {  // uniform per block
    if (threadIdx.x == 0) p.inliers[b] = -1;
    return;
  }
  double m[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) m[i] = 0.0;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const int64_t s = (int64_t)b * p.cap + j;
    add_moments(m, (double)fmaxf(p.slot_score[s], 0.f), p.src_pts + ((int64_t)b * p.K + p.slot_src[s]) * 3,
                p.ref_pts + ((int64_t)b * p.K + p.slot_ref[s]) * 3);
  }
  block_sum16(m, sh_red, sh_m);
  if (threadIdx.x == 0) {
    solve_procrustes(sh_m, 1e-5, sh_T);
    for (int i = 0; i < 12; ++i) p.local_T[(int64_t)b * 12 + i] = sh_T[i];
  }
  __syncthreads();
  const int pair = pair_of_patch(p.patch_off, p.P, b);
  const int b0 = (int)p.patch_off[pair], b1 = (int)p.patch_off[pair + 1];
  int cnt = 0;
  for (int bb = b0 + threadIdx.x; bb < b1; bb += blockDim.x) {
    const int nn = p.count[bb];
    for (int j = 0; j < nn; ++j) {
      const int64_t s = (int64_t)bb * p.cap + j;
      cnt += within(sh_T, p.src_pts + ((int64_t)bb * p.K + p.slot_src[s]) * 3,
                    p.ref_pts + ((int64_t)bb * p.K + p.slot_ref[s]) * 3, p.radius);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if ((threadIdx.x & 31) == 0) sh_cnt[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < kThreads / 32; ++w) t += sh_cnt[w];
    p.inliers[b] = t;
  }
}

// ---- kernel 3: global refinement, one CTA per pair ----------------------------------------------------------------
__global__ void __launch_bounds__(256) lgr_global_kernel(Params p) {
  __shared__ double sh_red[8 * 16], sh_m[16], sh_T[12];
  __shared__ int sh_best;
  const int pair = blockIdx.x;
  const int b0 = (int)p.patch_off[pair], b1 = (int)p.patch_off[pair + 1];
  // compacted correspondences (row-major over patches, as torch.nonzero orders them)
  for (int bb = b0 + threadIdx.x; bb < b1; bb += blockDim.x) {
    const int nn = p.count[bb];
    const int64_t o = p.corr_off[bb];
    for (int j = 0; j < nn; ++j) {
      const int64_t s = (int64_t)bb * p.cap + j;
      const float* rp = p.ref_pts + ((int64_t)bb * p.K + p.slot_ref[s]) * 3;
      const float* sp = p.src_pts + ((int64_t)bb * p.K + p.slot_src[s]) * 3;
      for (int c = 0; c < 3; ++c) { p.out_ref[(o + j) * 3 + c] = rp[c]; p.out_src[(o + j) * 3 + c] = sp[c]; }
      p.out_score[o + j] = p.slot_score[s];
    }
  }
  if (threadIdx.x == 0) {
    int best = -1, best_cnt = -1;
    for (int bb = b0; bb < b1; ++bb)
      if (p.inliers[bb] > best_cnt) { best_cnt = p.inliers[bb]; best = bb; }
    sh_best = best;
    if (best >= 0)
      for (int i = 0; i < 12; ++i) sh_T[i] = p.local_T[(int64_t)best * 12 + i];
  }
  __syncthreads();
  // mode 0: all scores (degenerate start); mode 1: scores of the correspondences within the radius of sh_T
  auto moments = [&](int mode) {
    double m[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) m[i] = 0.0;
    for (int bb = b0 + threadIdx.x; bb < b1; bb += blockDim.x) {
      const int nn = p.count[bb];
      for (int j = 0; j < nn; ++j) {
        const int64_t s = (int64_t)bb * p.cap + j;
        const float* rp = p.ref_pts + ((int64_t)bb * p.K + p.slot_ref[s]) * 3;
        const float* sp = p.src_pts + ((int64_t)bb * p.K + p.slot_src[s]) * 3;
        if (mode == 0 || within(sh_T, sp, rp, p.radius)) add_moments(m, (double)fmaxf(p.slot_score[s], 0.f), sp, rp);
      }
    }
    block_sum16(m, sh_red, sh_m);
    if (threadIdx.x == 0) solve_procrustes(sh_m, 1e-5, sh_T);
    __syncthreads();
  };
  if (sh_best < 0) moments(0);
  for (int it = 0; it < p.steps; ++it) moments(1);
  if (threadIdx.x < 16) {
    const int i = threadIdx.x;
    p.out_T[(int64_t)pair * 16 + i] = i < 12 ? (float)sh_T[i] : (i == 15 ? 1.f : 0.f);
  }
}

}  // namespace lgr
}  // namespace se3et

using namespace se3et;

extern "C" int se3et_lgr_workspace_bytes(int64_t num_patches, int64_t k_points, int64_t topk, size_t* bytes) {
  if (!bytes || num_patches < 0 || k_points <= 0 || topk <= 0 || topk > lgr::kMaxTopk) return SE3ET_ERR_ARG;
  const size_t cap = (size_t)(topk * k_points);
  size_t b = 0;
  b += align_up(sizeof(int32_t) * num_patches * cap, 256) * 2;  // slot_ref, slot_src
  b += align_up(sizeof(float) * num_patches * cap, 256);        // slot_score
  b += align_up(sizeof(double) * num_patches * 12, 256);        // local_T
  b += align_up(sizeof(int32_t) * num_patches, 256);            // inliers
  *bytes = b + 1024;
  return SE3ET_OK;
}

// Stage 1: the correspondences of every patch pair (slots in the workspace) and their number per patch.
extern "C" int se3et_lgr_correspondences(const float* log_scores, int64_t ld, const uint8_t* ref_masks,
                                         const uint8_t* src_masks, int64_t num_patches, int64_t k_points, int64_t topk,
                                         float confidence_threshold, void* workspace, size_t workspace_bytes,
                                         int32_t* counts, se3et_stream_t stream) {
  if (!log_scores || !ref_masks || !src_masks || !workspace || !counts || num_patches < 0 || k_points <= 0 ||
      ld < k_points || topk <= 0 || topk > lgr::kMaxTopk)
    return SE3ET_ERR_ARG;
  if (k_points > lgr::kThreads) return SE3ET_ERR_UNSUPPORTED;
  size_t need = 0;
  se3et_lgr_workspace_bytes(num_patches, k_points, topk, &need);
  if (workspace_bytes < need) return SE3ET_ERR_WORKSPACE;
  if (num_patches == 0) return SE3ET_OK;
  Carver cv(workspace, workspace_bytes);
  lgr::Params p{};
  p.log_scores = log_scores; p.ref_mask = ref_masks; p.src_mask = src_masks;
  p.B = (int)num_patches; p.K = (int)k_points; p.ld = (int)ld; p.topk = (int)topk; p.cap = (int)(topk * k_points);
  p.conf_thr = confidence_threshold;
  p.slot_ref = cv.take<int32_t>((size_t)p.B * p.cap);
  p.slot_src = cv.take<int32_t>((size_t)p.B * p.cap);
  p.slot_score = cv.take<float>((size_t)p.B * p.cap);
  p.count = counts;
  const size_t smem = sizeof(float) * p.K * (p.K + 1) + sizeof(int) * (2 * p.K * lgr::kMaxTopk + p.K + 1);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (smem > 48 * 1024) SE3ET_ENSURE_SMEM(lgr::lgr_corr_kernel, smem);
  lgr::lgr_corr_kernel<<<p.B, lgr::kThreads, smem, st>>>(p);
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}

// Stage 2: local-to-global registration on the slots of stage 1 (same workspace); corr_offsets = exclusive scan of
// counts (num_patches + 1 entries, on the device).
extern "C" int se3et_lgr_register(const float* ref_knn_points, const float* src_knn_points, const int64_t* patch_offsets,
                                  int64_t num_pairs, int64_t num_patches, int64_t k_points, int64_t topk,
                                  float acceptance_radius, int64_t correspondence_threshold,
                                  int64_t num_refinement_steps, void* workspace, size_t workspace_bytes,
                                  const int32_t* counts, const int64_t* corr_offsets, float* out_ref_points,
                                  float* out_src_points, float* out_scores, float* out_transforms,
                                  se3et_stream_t stream) {
  if (!ref_knn_points || !src_knn_points || !patch_offsets || !workspace || !counts || !corr_offsets ||
      !out_ref_points || !out_src_points || !out_scores || !out_transforms || num_pairs <= 0 || num_patches < 0 ||
      k_points <= 0 || topk <= 0 || topk > lgr::kMaxTopk || num_refinement_steps < 1)
    return SE3ET_ERR_ARG;
  size_t need = 0;
  se3et_lgr_workspace_bytes(num_patches, k_points, topk, &need);
  if (workspace_bytes < need) return SE3ET_ERR_WORKSPACE;
  Carver cv(workspace, workspace_bytes);
  lgr::Params p{};
  p.ref_pts = ref_knn_points; p.src_pts = src_knn_points; p.patch_off = patch_offsets;
  p.B = (int)num_patches; p.K = (int)k_points; p.P = (int)num_pairs; p.topk = (int)topk; p.cap = (int)(topk * k_points);
  p.radius = acceptance_radius; p.corr_thr = (int)correspondence_threshold; p.steps = (int)num_refinement_steps;
  p.slot_ref = cv.take<int32_t>((size_t)p.B * p.cap);
  p.slot_src = cv.take<int32_t>((size_t)p.B * p.cap);
  p.slot_score = cv.take<float>((size_t)p.B * p.cap);
  p.local_T = cv.take<double>((size_t)p.B * 12);
  p.inliers = cv.take<int32_t>((size_t)p.B);
  p.count = const_cast<int32_t*>(counts);
  p.corr_off = corr_offsets;
  p.out_ref = out_ref_points; p.out_src = out_src_points; p.out_score = out_scores; p.out_T = out_transforms;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (p.B > 0) {
    lgr::lgr_local_kernel<<<p.B, lgr::kThreads, 0, st>>>(p);
    SE3ET_LAUNCH_CHECK();
  }
  lgr::lgr_global_kernel<<<p.P, 256, 0, st>>>(p);
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}

// weighted_procrustes (procrustes.py:6-73) for `batch` independent point sets of n points each: one CTA per set.
namespace se3et {
namespace lgr {
__global__ void __launch_bounds__(256) procrustes_kernel(const float* src, const float* ref, const float* w, int n,
                                                         float weight_thresh, float eps, float* out_T) {
  __shared__ double sh_red[8 * 16], sh_m[16], sh_T[12];
  const int b = blockIdx.x;
  double m[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) m[i] = 0.0;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    float wj = w ? w[(int64_t)b * n + j] : 1.f;
    if (wj < weight_thresh) wj = 0.f;
    add_moments(m, (double)wj, src + ((int64_t)b * n + j) * 3, ref + ((int64_t)b * n + j) * 3);
  }
  block_sum16(m, sh_red, sh_m);
  if (threadIdx.x == 0) solve_procrustes(sh_m, (double)eps, sh_T);
  __syncthreads();
  if (threadIdx.x < 16) {
    const int i = threadIdx.x;
    out_T[(int64_t)b * 16 + i] = i < 12 ? (float)sh_T[i] : (i == 15 ? 1.f : 0.f);
  }
}
}  // namespace lgr
}  // namespace se3et

extern "C" int se3et_weighted_procrustes(const float* src_points, const float* ref_points, const float* weights,
                                         int64_t batch, int64_t n, float weight_thresh, float eps,
                                         float* out_transforms, se3et_stream_t stream) {
  if (!src_points || !ref_points || !out_transforms || batch <= 0 || n < 0) return SE3ET_ERR_ARG;
  lgr::procrustes_kernel<<<(unsigned)batch, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      src_points, ref_points, weights, (int)n, weight_thresh, eps, out_transforms);
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}
