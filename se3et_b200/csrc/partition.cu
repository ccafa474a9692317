// point_to_node_partition (geotransformer/modules/ops/pointcloud_partition.py:60-107) for stacked clouds:
// every fine point goes to its nearest superpoint ("node") of the same cloud, every node lists its point_limit nearest
// assigned points.  The reference builds the dense (M, N) distance matrix with pairwise_distance
// (modules/ops/pairwise_distance.py:4-31: x2 - 2 x.y + y2, clamped at 0), takes min over nodes, masks and topk's it.
//
// Here: assign_kernel - one thread per point scans the nodes of its cloud (a few hundred, broadcast loads) with the
// reference's expanded distance formula in a fixed fp32 operation order (no FMA contraction; the numpy oracle
// reproduces it bit for bit); ties go to the lowest node index.  knn_kernel - one warp per node compacts its points from
// the assignment array, ranks them by (distance, point index) and writes the first point_limit; the remaining slots
// are padded with the cloud's point count and masked out, exactly like the reference's masked_fill.
#include "common.cuh"

namespace se3et {

constexpr int kPartWarps = 8;
constexpr int kPartCap = 128;  // assigned points a warp ranks from shared memory (more: exact slow path)

__device__ __forceinline__ float part_sqnorm(float a, float b, float c) {
  return __fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c, c));
}
// pairwise_distance.py:27-30 with x = node, y = point: (x2 - 2 * xy) + y2, clamp(min = 0)
__device__ __forceinline__ float part_dist(float x0, float x1, float x2c, float xn, float y0, float y1, float y2c,
                                           float yn) {
  const float xy = __fadd_rn(__fadd_rn(__fmul_rn(x0, y0), __fmul_rn(x1, y1)), __fmul_rn(x2c, y2c));
  const float d = __fadd_rn(__fsub_rn(xn, __fmul_rn(2.f, xy)), yn);
  return fmaxf(d, 0.f);
}

__global__ void part_offsets_kernel(const int64_t* __restrict__ point_len, const int64_t* __restrict__ node_len,
                                    int batch, int64_t* __restrict__ point_off, int64_t* __restrict__ node_off) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int64_t p = 0, n = 0;
    for (int b = 0; b < batch; ++b) {
      point_off[b] = p;
      node_off[b] = n;
      p += point_len[b];
      n += node_len[b];
    }
    point_off[batch] = p;
    node_off[batch] = n;
  }
}

__global__ void __launch_bounds__(256) part_assign_kernel(const float* __restrict__ points, int64_t n_points,
                                                          const float* __restrict__ nodes,
                                                          const int64_t* __restrict__ point_off,
                                                          const int64_t* __restrict__ node_off, int batch,
                                                          int64_t* __restrict__ point_to_node,
                                                          int32_t* __restrict__ p2n32, float* __restrict__ dist) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_points) return;
  const int b = segment_of(point_off, batch, i);
  const int64_t lo = node_off[b], hi = node_off[b + 1];
  const float y0 = points[3 * i], y1 = points[3 * i + 1], y2 = points[3 * i + 2];
  const float yn = part_sqnorm(y0, y1, y2);
  float best = __int_as_float(0x7f800000);
  int64_t best_m = lo;
  for (int64_t m = lo; m < hi; ++m) {
    const float x0 = __ldg(nodes + 3 * m), x1 = __ldg(nodes + 3 * m + 1), x2 = __ldg(nodes + 3 * m + 2);
    const float d = part_dist(x0, x1, x2, part_sqnorm(x0, x1, x2), y0, y1, y2, yn);
    if (d < best) {  // strict: the lowest node index wins ties (torch.min returns the first minimum)
      best = d;
      best_m = m;
    }
  }
  point_to_node[i] = best_m - lo;
  p2n32[i] = (int32_t)(best_m - lo);
  dist[i] = best;
}

__global__ void __launch_bounds__(kPartWarps * 32) part_knn_kernel(
    const int32_t* __restrict__ p2n32, const float* __restrict__ dist, const int64_t* __restrict__ point_off,
    const int64_t* __restrict__ node_off, int batch, int64_t n_nodes, int K, uint8_t* __restrict__ node_masks,
    int64_t* __restrict__ node_sizes, int64_t* __restrict__ knn_idx, uint8_t* __restrict__ knn_mask) {
  __shared__ float sh_d[kPartWarps][kPartCap];
  __shared__ int32_t sh_i[kPartWarps][kPartCap];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m = (int64_t)blockIdx.x * kPartWarps + warp;
  if (m >= n_nodes) return;
  const int b = segment_of(node_off, batch, m);
  const int32_t ml = (int32_t)(m - node_off[b]);
  const int64_t p_lo = point_off[b], p_hi = point_off[b + 1];
  int c = 0;
  for (int64_t base = p_lo; base < p_hi; base += 32) {
    const int64_t i = base + lane;
    const bool hit = i < p_hi && p2n32[i] == ml;
    const unsigned ballot = __ballot_sync(0xffffffffu, hit);
    const int pos = c + __popc(ballot & ((1u << lane) - 1u));
    if (hit && pos < kPartCap) {
      sh_d[warp][pos] = dist[i];
      sh_i[warp][pos] = (int32_t)(i - p_lo);
    }
    c += __popc(ballot);
  }
  __syncwarp();
  int64_t* row = knn_idx + m * K;
  uint8_t* mrow = knn_mask + m * K;
  if (c <= kPartCap) {
    // the list is in ascending point order: rank = #{smaller distance} + #{equal distance, earlier point}
    for (int e = lane; e < c; e += 32) {
      const float d = sh_d[warp][e];
      int rank = 0;
      for (int f = 0; f < c; ++f) {
        const float df = sh_d[warp][f];
        rank += (df < d || (df == d && f < e)) ? 1 : 0;
      }
      if (rank < K) {
        row[rank] = sh_i[warp][e];
        mrow[rank] = 1;
      }
    }
  } else {
    // more assigned points than the shared list holds: exact ranks straight from global memory
    for (int64_t base = p_lo; base < p_hi; base += 32) {
      const int64_t i = base + lane;
      if (i < p_hi && p2n32[i] == ml) {
        const float d = dist[i];
        int rank = 0;
        for (int64_t j = p_lo; j < p_hi; ++j) {
          if (p2n32[j] != ml) continue;
          const float dj = dist[j];
          rank += (dj < d || (dj == d && j < i)) ? 1 : 0;
        }
        if (rank < K) {
          row[rank] = i - p_lo;
          mrow[rank] = 1;
        }
      }
    }
  }
  const int filled = c < K ? c : K;
  for (int k = filled + lane; k < K; k += 32) {
    row[k] = p_hi - p_lo;  // masked_fill_(~masks, points.shape[0])
    mrow[k] = 0;
  }
  if (lane == 0) {
    node_masks[m] = c > 0 ? 1 : 0;
    if (node_sizes) node_sizes[m] = c;
  }
}

}  // namespace se3et

using namespace se3et;

extern "C" int se3et_point_to_node_partition_workspace_bytes(int64_t n_points, int64_t batch, size_t* bytes) {
  if (!bytes || n_points < 0 || batch <= 0) return SE3ET_ERR_ARG;
  *bytes = align_up(sizeof(int64_t) * 2 * (size_t)(batch + 1), 256) + align_up(sizeof(int32_t) * (size_t)n_points, 256) +
           align_up(sizeof(float) * (size_t)n_points, 256) + 256;
  return SE3ET_OK;
}

extern "C" int se3et_point_to_node_partition(const float* points, const int64_t* point_lengths, int64_t n_points,
                                             const float* nodes, const int64_t* node_lengths, int64_t n_nodes,
                                             int64_t batch, int64_t point_limit, int64_t* point_to_node,
                                             uint8_t* node_masks, int64_t* node_sizes, int64_t* node_knn_indices,
                                             uint8_t* node_knn_masks, void* workspace, size_t workspace_bytes,
                                             se3et_stream_t stream) {
  if (n_points < 0 || n_nodes < 0 || batch <= 0 || batch > INT32_MAX || point_limit <= 0 || point_limit > INT32_MAX)
    return SE3ET_ERR_ARG;
  if (!point_lengths || !node_lengths || !workspace) return SE3ET_ERR_ARG;
  if ((n_points > 0 && (!points || !point_to_node)) ||
      (n_nodes > 0 && (!nodes || !node_masks || !node_knn_indices || !node_knn_masks)))
    return SE3ET_ERR_ARG;
  size_t need = 0;
  se3et_point_to_node_partition_workspace_bytes(n_points, batch, &need);
  if (workspace_bytes < need || (reinterpret_cast<uintptr_t>(workspace) & 255)) return SE3ET_ERR_WORKSPACE;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  int64_t* point_off = reinterpret_cast<int64_t*>(ws);
  int64_t* node_off = point_off + (batch + 1);
  ws += align_up(sizeof(int64_t) * 2 * (size_t)(batch + 1), 256);
  int32_t* p2n32 = reinterpret_cast<int32_t*>(ws);
  ws += align_up(sizeof(int32_t) * (size_t)n_points, 256);
  float* dist = reinterpret_cast<float*>(ws);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  part_offsets_kernel<<<1, 32, 0, st>>>(point_lengths, node_lengths, (int)batch, point_off, node_off);
  SE3ET_LAUNCH_CHECK();
  if (n_points > 0) {
    part_assign_kernel<<<(unsigned)ceil_div(n_points, 256), 256, 0, st>>>(points, n_points, nodes, point_off, node_off,
                                                                         (int)batch, point_to_node, p2n32, dist);
    SE3ET_LAUNCH_CHECK();
  }
  if (n_nodes > 0) {
    part_knn_kernel<<<(unsigned)ceil_div(n_nodes, kPartWarps), kPartWarps * 32, 0, st>>>(
        p2n32, dist, point_off, node_off, (int)batch, n_nodes, (int)point_limit, node_masks, node_sizes,
        node_knn_indices, node_knn_masks);
    SE3ET_LAUNCH_CHECK();
  }
  return SE3ET_OK;
}
