// point_to_node_partition (geotransformer/modules/ops/pointcloud_partition.py:60-107) for stacked clouds:
// every fine point goes to its nearest superpoint ("node") of the same cloud, every node lists its point_limit nearest
// assigned points.  The reference builds the dense (M, N) distance matrix with pairwise_distance
// (modules/ops/pairwise_distance.py:4-31: x2 - 2 x.y + y2, clamped at 0), takes min over nodes, masks and topk's it.
//
// Here: assign_kernel - one thread per point scans the nodes of its cloud (a few hundred, broadcast loads) with the
// reference's expanded distance formula in a fixed fp32 operation order (no FMA contraction; the numpy oracle
// reproduces it bit for bit); ties go to the lowest node index; the node's size is counted on the way.  A scan and a
// scatter turn the assignment into per-node point lists (counting sort, O(N): a KITTI-shaped cloud has 20k fine points
// and 800 nodes, so the first version -- every node's warp scanning its whole cloud -- took 2.6 ms per pair).
// knn_kernel - one warp per node ranks its list by (distance, point index) and writes the first point_limit; the
// remaining slots are padded with the cloud's point count and masked out, exactly like the reference's masked_fill.
#include "common.cuh"

namespace se3et {

constexpr int kPartWarps = 8;
constexpr int kPartCap = 256;  // assigned points a warp ranks from shared memory (more: exact slow path)

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
                                                          int32_t* __restrict__ p2n32, float* __restrict__ dist,
                                                          int32_t* __restrict__ node_count) {
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
  p2n32[i] = (int32_t)best_m;   // global node row
  dist[i] = best;
  if (hi > lo) atomicAdd(&node_count[best_m], 1);
}

// exclusive scan of the node sizes (one block; a launch sequence has a few ten thousand nodes)
__global__ void __launch_bounds__(1024) part_scan_kernel(const int32_t* __restrict__ node_count, int64_t n_nodes,
                                                         int32_t* __restrict__ node_start) {
  __shared__ int32_t sh[1024];
  __shared__ int32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t base = 0; base < n_nodes; base += 1024) {
    const int64_t i = base + threadIdx.x;
    const int32_t v = i < n_nodes ? node_count[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int32_t t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < n_nodes) node_start[i] = carry + sh[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += sh[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) node_start[n_nodes] = carry;
}

__global__ void __launch_bounds__(256) part_scatter_kernel(const int32_t* __restrict__ p2n32, int64_t n_points,
                                                           const int64_t* __restrict__ point_off,
                                                           const int64_t* __restrict__ node_off, int batch,
                                                           const int32_t* __restrict__ node_start,
                                                           int32_t* __restrict__ cursor, int32_t* __restrict__ list) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_points) return;
  const int b = segment_of(point_off, batch, i);
  if (node_off[b + 1] <= node_off[b]) return;  // a cloud without nodes has no lists
  const int32_t m = p2n32[i];
  list[node_start[m] + atomicAdd(&cursor[m], 1)] = (int32_t)(i - point_off[b]);  // cloud-local point id
}

__global__ void __launch_bounds__(kPartWarps * 32) part_knn_kernel(
    const int32_t* __restrict__ list, const int32_t* __restrict__ node_start, const float* __restrict__ dist,
    const int64_t* __restrict__ point_off, const int64_t* __restrict__ node_off, int batch, int64_t n_nodes, int K,
    uint8_t* __restrict__ node_masks, int64_t* __restrict__ node_sizes, int64_t* __restrict__ knn_idx,
    uint8_t* __restrict__ knn_mask) {
  __shared__ float sh_d[kPartWarps][kPartCap];
  __shared__ int32_t sh_i[kPartWarps][kPartCap];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m = (int64_t)blockIdx.x * kPartWarps + warp;
  if (m >= n_nodes) return;
  const int b = segment_of(node_off, batch, m);
  const int64_t p_lo = point_off[b], p_hi = point_off[b + 1];
  const int32_t l0 = node_start[m];
  const int c = node_start[m + 1] - l0;
  int64_t* row = knn_idx + m * K;
  uint8_t* mrow = knn_mask + m * K;
  // the list is in scatter order: rank = #{smaller distance} + #{equal distance, lower point index}
  if (c <= kPartCap) {
    for (int e = lane; e < c; e += 32) {
      const int32_t i = list[l0 + e];
      sh_i[warp][e] = i;
      sh_d[warp][e] = dist[p_lo + i];
    }
    __syncwarp();
    for (int e = lane; e < c; e += 32) {
      const float d = sh_d[warp][e];
      const int32_t i = sh_i[warp][e];
      int rank = 0;
      for (int f = 0; f < c; ++f) {
        const float df = sh_d[warp][f];
        rank += (df < d || (df == d && sh_i[warp][f] < i)) ? 1 : 0;
      }
      if (rank < K) {
        row[rank] = i;
        mrow[rank] = 1;
      }
    }
  } else {
    // more assigned points than the shared list holds: exact ranks straight from the global list
    for (int e = lane; e < c; e += 32) {
      const int32_t i = list[l0 + e];
      const float d = dist[p_lo + i];
      int rank = 0;
      for (int f = 0; f < c; ++f) {
        const int32_t j = list[l0 + f];
        const float dj = dist[p_lo + j];
        rank += (dj < d || (dj == d && j < i)) ? 1 : 0;
      }
      if (rank < K) {
        row[rank] = i;
        mrow[rank] = 1;
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
  // offsets, point -> node, distance, point lists; the node-sized arrays (count, cursor, start) are bounded by the
  // number of points (a node without points is legal, more nodes than points per launch sequence is not expected)
  *bytes = align_up(sizeof(int64_t) * 2 * (size_t)(batch + 1), 256) + 2 * align_up(sizeof(int32_t) * (size_t)n_points, 256) +
           align_up(sizeof(float) * (size_t)n_points, 256) + 3 * align_up(sizeof(int32_t) * (size_t)(n_points + batch + 2), 256) +
           256;
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
  int32_t* list = reinterpret_cast<int32_t*>(ws);
  ws += align_up(sizeof(int32_t) * (size_t)n_points, 256);
  float* dist = reinterpret_cast<float*>(ws);
  ws += align_up(sizeof(float) * (size_t)n_points, 256);
  if (n_nodes > n_points + batch) return SE3ET_ERR_WORKSPACE;
  const size_t node_arr = align_up(sizeof(int32_t) * (size_t)(n_points + batch + 2), 256);
  int32_t* node_count = reinterpret_cast<int32_t*>(ws);
  int32_t* cursor = reinterpret_cast<int32_t*>(ws + node_arr);
  int32_t* node_start = reinterpret_cast<int32_t*>(ws + 2 * node_arr);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  part_offsets_kernel<<<1, 32, 0, st>>>(point_lengths, node_lengths, (int)batch, point_off, node_off);
  SE3ET_LAUNCH_CHECK();
  SE3ET_CUDA_CHECK(cudaMemsetAsync(node_count, 0, 2 * node_arr, st));  // counts and cursors
  if (n_points > 0) {
    part_assign_kernel<<<(unsigned)ceil_div(n_points, 256), 256, 0, st>>>(points, n_points, nodes, point_off, node_off,
                                                                         (int)batch, point_to_node, p2n32, dist,
                                                                         node_count);
    SE3ET_LAUNCH_CHECK();
  }
  if (n_nodes > 0) {
    part_scan_kernel<<<1, 1024, 0, st>>>(node_count, n_nodes, node_start);
    SE3ET_LAUNCH_CHECK();
    if (n_points > 0) {
      part_scatter_kernel<<<(unsigned)ceil_div(n_points, 256), 256, 0, st>>>(p2n32, n_points, point_off, node_off,
                                                                            (int)batch, node_start, cursor, list);
      SE3ET_LAUNCH_CHECK();
    }
    part_knn_kernel<<<(unsigned)ceil_div(n_nodes, kPartWarps), kPartWarps * 32, 0, st>>>(
        list, node_start, dist, point_off, node_off, (int)batch, n_nodes, (int)point_limit, node_masks, node_sizes,
        node_knn_indices, node_knn_masks);
    SE3ET_LAUNCH_CHECK();
  }
  return SE3ET_OK;
}
