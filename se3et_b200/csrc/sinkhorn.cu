// LearnableLogOptimalTransport.forward (geotransformer/modules/sinkhorn/learnable_sinkhorn.py:13-66): log-domain Sinkhorn
// with a dustbin row / column (SuperGlue style) on the (P, K, K) patch score matrices of the fine matching stage
// (experiments/se3eti.3dmatch/model.py:202-205).  The reference runs 2 x num_iterations torch.logsumexp launches over the
// (P, K+1, K+1) tensor; here one CTA owns one matrix, keeps it on chip for all iterations and writes the result once
// (generic kernel: shared memory; K = 64: registers, see log_sinkhorn_reg_kernel).  Row pass: one warp per row, lanes over columns; column pass: one warp per column, lanes over rows (odd pitch:
// conflict-free both ways).  fp32 throughout, logsumexp as max + log(sum(exp(x - max))) like torch.
#include "common.cuh"

namespace se3et {

constexpr int kOtThreads = 256;
constexpr float kOtInf = 1e12f;  // learnable_sinkhorn.py:6 (inf = 1e12): masked scores and marginals

__global__ void __launch_bounds__(kOtThreads) log_sinkhorn_kernel(const float* __restrict__ scores,
                                                                  const uint8_t* __restrict__ row_masks,
                                                                  const uint8_t* __restrict__ col_masks,
                                                                  const float* __restrict__ alpha_ptr, int M, int N,
                                                                  int iters, float* __restrict__ out) {
  extern __shared__ float sh[];
  const int R = M + 1, C = N + 1;
  const int pitch = C | 1;
  float* S = sh;                 // [R][pitch]
  float* u = S + R * pitch;      // [R]
  float* v = u + R;              // [C]
  float* log_mu = v + C;         // [R]
  float* log_nu = log_mu + R;    // [C]
  __shared__ int sh_cnt[2];
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const float alpha = *alpha_ptr;
  const uint8_t* rm = row_masks ? row_masks + (int64_t)b * M : nullptr;
  const uint8_t* cm = col_masks ? col_masks + (int64_t)b * N : nullptr;
  if (threadIdx.x < 2) sh_cnt[threadIdx.x] = 0;
  __syncthreads();
  // valid rows / columns (:46-47)
  int cr = 0, cc = 0;
  for (int i = threadIdx.x; i < M; i += blockDim.x) cr += (!rm || rm[i]) ? 1 : 0;
  for (int j = threadIdx.x; j < N; j += blockDim.x) cc += (!cm || cm[j]) ? 1 : 0;
  cr = (int)warp_sum((float)cr);
  cc = (int)warp_sum((float)cc);
  if (lane == 0) {
    atomicAdd(&sh_cnt[0], cr);
    atomicAdd(&sh_cnt[1], cc);
  }
  // padded scores (:40-43): dustbin column / row = alpha, masked rows / columns = -inf
  for (int t = threadIdx.x; t < R * C; t += blockDim.x) {
    const int i = t / C, j = t - i * C;
    float s = (i < M && j < N) ? scores[((int64_t)b * M + i) * N + j] : alpha;
    const bool masked = (i < M && rm && !rm[i]) || (j < N && cm && !cm[j]);
    S[i * pitch + j] = masked ? -kOtInf : s;
  }
  __syncthreads();
  const float nvr = (float)sh_cnt[0], nvc = (float)sh_cnt[1];
  const float norm = -logf(nvr + nvc);  // (:48)
  for (int i = threadIdx.x; i < R; i += blockDim.x) {
    float m = i < M ? norm : logf(nvc) + norm;  // (:50-52)
    if (i < M && rm && !rm[i]) m = -kOtInf;
    log_mu[i] = m;
    u[i] = 0.f;
  }
  for (int j = threadIdx.x; j < C; j += blockDim.x) {
    float m = j < N ? norm : logf(nvr) + norm;  // (:55-57)
    if (j < N && cm && !cm[j]) m = -kOtInf;
    log_nu[j] = m;
    v[j] = 0.f;
  }
  __syncthreads();
  for (int it = 0; it < iters; ++it) {  // (:13-18)
    for (int i = warp; i < R; i += nwarps) {
      const float* row = S + i * pitch;
      float mx = -INFINITY;
      for (int j = lane; j < C; j += 32) mx = fmaxf(mx, row[j] + v[j]);
      mx = warp_max(mx);
      float s = 0.f;
      for (int j = lane; j < C; j += 32) s += expf(row[j] + v[j] - mx);
      s = warp_sum(s);
      if (lane == 0) u[i] = log_mu[i] - (mx + logf(s));
    }
    __syncthreads();
    for (int j = warp; j < C; j += nwarps) {
      float mx = -INFINITY;
      for (int i = lane; i < R; i += 32) mx = fmaxf(mx, S[i * pitch + j] + u[i]);
      mx = warp_max(mx);
      float s = 0.f;
      for (int i = lane; i < R; i += 32) s += expf(S[i * pitch + j] + u[i] - mx);
      s = warp_sum(s);
      if (lane == 0) v[j] = log_nu[j] - (mx + logf(s));
    }
    __syncthreads();
  }
  float* o = out + (int64_t)b * R * C;
  for (int t = threadIdx.x; t < R * C; t += blockDim.x) {
    const int i = t / C, j = t - i * C;
    o[t] = S[i * pitch + j] + u[i] + v[j] - norm;  // (:19, :60)
  }
}

// K x K patches (K = 64, the 3DMatch configuration): thread-per-row / thread-per-column with the matrix held in registers.
// Thread t owns row t AND column t of the padded (K+1) x (K+1) matrix (two register copies); u and v live in shared
// memory and are read as broadcasts, so an iteration has no shuffles and no shared-memory traffic for S at all.
template <int K>
__global__ void __launch_bounds__(96) log_sinkhorn_reg_kernel(const float* __restrict__ scores,
                                                              const uint8_t* __restrict__ row_masks,
                                                              const uint8_t* __restrict__ col_masks,
                                                              const float* __restrict__ alpha_ptr, int iters,
                                                              float* __restrict__ out) {
  constexpr int R = K + 1;
  constexpr int kPitch = R | 1;
  constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
  __shared__ float S[R * kPitch];
  __shared__ float u[R], v[R];
  __shared__ int sh_cnt[2];
  const int b = blockIdx.x, t = threadIdx.x;
  const float alpha = *alpha_ptr;
  const uint8_t* rm = row_masks ? row_masks + (int64_t)b * K : nullptr;
  const uint8_t* cm = col_masks ? col_masks + (int64_t)b * K : nullptr;
  if (t < 2) sh_cnt[t] = 0;
  __syncthreads();
  if (t < K) {
    if (!rm || rm[t]) atomicAdd(&sh_cnt[0], 1);
    if (!cm || cm[t]) atomicAdd(&sh_cnt[1], 1);
  }
  for (int e = t; e < R * R; e += 96) {
    const int i = e / R, j = e - i * R;
    const float sc = (i < K && j < K) ? scores[((int64_t)b * K + i) * K + j] : alpha;
    const bool masked = (i < K && rm && !rm[i]) || (j < K && cm && !cm[j]);
    S[i * kPitch + j] = masked ? -kOtInf : sc;
  }
  if (t < R) u[t] = v[t] = 0.f;
  __syncthreads();
  const float nvr = (float)sh_cnt[0], nvc = (float)sh_cnt[1];
  const float norm = -logf(nvr + nvc);
  const bool active = t < R;
  const int me = active ? t : 0;
  float log_mu = me < K ? norm : logf(nvc) + norm;
  if (me < K && rm && !rm[me]) log_mu = -kOtInf;
  float log_nu = me < K ? norm : logf(nvr) + norm;
  if (me < K && cm && !cm[me]) log_nu = -kOtInf;
  float srow[R], scol[R];
#pragma unroll
  for (int j = 0; j < R; ++j) {
    srow[j] = S[me * kPitch + j];
    scol[j] = S[j * kPitch + me];
  }
  for (int it = 0; it < iters; ++it) {
    if (active) {  // u[i] = log_mu[i] - logsumexp_j(S[i][j] + v[j])
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < R; ++j) mx = fmaxf(mx, srow[j] + v[j]);
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < R; ++j) sum += exp2f((srow[j] + v[j] - mx) * kLog2e);
      u[me] = log_mu - (mx + log2f(sum) * kLn2);
    }
    __syncthreads();
    if (active) {  // v[j] = log_nu[j] - logsumexp_i(S[i][j] + u[i])
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < R; ++i) mx = fmaxf(mx, scol[i] + u[i]);
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < R; ++i) sum += exp2f((scol[i] + u[i] - mx) * kLog2e);
      v[me] = log_nu - (mx + log2f(sum) * kLn2);
    }
    __syncthreads();
  }
  if (active) {
    const float ui = u[me];
#pragma unroll
    for (int j = 0; j < R; ++j) S[me * kPitch + j] = srow[j] + ui + v[j] - norm;
  }
  __syncthreads();
  float* o = out + (int64_t)b * R * R;
  for (int e = t; e < R * R; e += 96) {
    const int i = e / R, j = e - i * R;
    o[e] = S[i * kPitch + j];
  }
}

}  // namespace se3et

using namespace se3et;

extern "C" int se3et_log_optimal_transport(const float* scores, const uint8_t* row_masks, const uint8_t* col_masks,
                                           const float* alpha, int64_t batch, int64_t num_row, int64_t num_col,
                                           int64_t num_iterations, float* out, se3et_stream_t stream) {
  if (batch < 0 || num_row <= 0 || num_col <= 0 || num_iterations < 0 || num_row > 1024 || num_col > 1024)
    return SE3ET_ERR_ARG;
  if (batch == 0) return SE3ET_OK;
  if (!scores || !alpha || !out) return SE3ET_ERR_ARG;
  if (num_row == 64 && num_col == 64) {
    log_sinkhorn_reg_kernel<64><<<(unsigned)batch, 96, 0, static_cast<cudaStream_t>(stream)>>>(
        scores, row_masks, col_masks, alpha, (int)num_iterations, out);
    SE3ET_LAUNCH_CHECK();
    return SE3ET_OK;
  }
  const int R = (int)num_row + 1, C = (int)num_col + 1;
  const size_t smem = sizeof(float) * ((size_t)R * (C | 1) + 2 * (size_t)R + 2 * (size_t)C);
  if (smem > 227 * 1024) return SE3ET_ERR_UNSUPPORTED;
  if (smem > 48 * 1024) SE3ET_ENSURE_SMEM(log_sinkhorn_kernel, smem);
  log_sinkhorn_kernel<<<(unsigned)batch, kOtThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      scores, row_masks, col_masks, alpha, (int)num_row, (int)num_col, (int)num_iterations, out);
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}
