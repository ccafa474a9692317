// TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// extern "C" entry points around the *unmodified* reference CPU implementation,
// compiled from the sources where they lie under /root/reference:
//   geotransformer/extensions/cpu/grid_subsampling/grid_subsampling_cpu.cpp
//   geotransformer/extensions/cpu/radius_neighbors/radius_neighbors_cpu.cpp
//   geotransformer/extensions/extra/cloud/cloud.cpp
// This file contains no reference code: it only marshals flat buffers into the
// std::vector arguments those functions take, the same way the reference's own
// torch wrappers do (grid_subsampling.cpp:26-56, radius_neighbors.cpp:25-51),
// so that the reference can be called without torch through ctypes.
//
// Output order is the reference's own (unordered_map iteration order for the
// subsample, std::sort on d2 only for the neighbours); canonicalisation is done
// by the caller (oracle/points.py), never here.
#include <cstdint>
#include <cstring>
#include <vector>

#include "cpu/grid_subsampling/grid_subsampling_cpu.h"
#include "cpu/radius_neighbors/radius_neighbors_cpu.h"

extern "C" {

// Returns total number of subsampled points M (or -1 if cap is too small).
// s_points/s_normals must hold cap*3 floats; s_lengths holds B longs.
long ref_grid_subsampling(const float* points, const long* lengths, const float* normals,
                          long n_total, long batch, float voxel_size,
                          float* s_points, long* s_lengths, float* s_normals, long cap) {
  std::vector<PointXYZ> vp(reinterpret_cast<const PointXYZ*>(points),
                           reinterpret_cast<const PointXYZ*>(points) + n_total);
  std::vector<PointXYZ> vn(reinterpret_cast<const PointXYZ*>(normals),
                           reinterpret_cast<const PointXYZ*>(normals) + n_total);
  std::vector<long> vl(lengths, lengths + batch);
  std::vector<PointXYZ> sp, sn;
  std::vector<long> sl;
  grid_subsampling_cpu(vp, sp, vl, sl, vn, sn, voxel_size);
  long m = static_cast<long>(sp.size());
  if (m > cap) return -1;
  std::memcpy(s_points, sp.data(), sizeof(float) * 3 * m);
  std::memcpy(s_normals, sn.data(), sizeof(float) * 3 * m);
  std::memcpy(s_lengths, sl.data(), sizeof(long) * batch);
  return m;
}

// Two-phase: call with out == nullptr to get the width (max neighbour count),
// the result is cached in a thread-local vector and copied by the second call.
static thread_local std::vector<long> g_last_neighbors;

long ref_radius_neighbors(const float* q_points, const float* s_points, const long* q_lengths,
                          const long* s_lengths, long nq, long ns, long batch, float radius,
                          long* out, long out_width) {
  if (out == nullptr) {
    std::vector<PointXYZ> vq(reinterpret_cast<const PointXYZ*>(q_points),
                             reinterpret_cast<const PointXYZ*>(q_points) + nq);
    std::vector<PointXYZ> vs(reinterpret_cast<const PointXYZ*>(s_points),
                             reinterpret_cast<const PointXYZ*>(s_points) + ns);
    std::vector<long> ql(q_lengths, q_lengths + batch);
    std::vector<long> sl(s_lengths, s_lengths + batch);
    g_last_neighbors.clear();
    radius_neighbors_cpu(vq, vs, ql, sl, g_last_neighbors, radius);
    return nq > 0 ? static_cast<long>(g_last_neighbors.size()) / nq : 0;
  }
  long width = nq > 0 ? static_cast<long>(g_last_neighbors.size()) / nq : 0;
  if (width != out_width) return -1;
  std::memcpy(out, g_last_neighbors.data(), sizeof(long) * g_last_neighbors.size());
  return width;
}

}  // extern "C"
