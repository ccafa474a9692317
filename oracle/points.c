/* TEST INFRASTRUCTURE ONLY -- never linked, imported or executed by the product path.
 *
 * Plain-C restatement of the reference's CPU point ops, written from the
 * algorithm (not from the reference's code structure):
 *
 *   oracle_grid_subsample   <-  geotransformer/extensions/cpu/grid_subsampling/
 *                               grid_subsampling_cpu.cpp:3-71 (voxel keys, origin, NX/NY)
 *                               grid_subsampling_cpu.h:49-73  (closest-to-barycentre choice)
 *                               extra/cloud/cloud.cpp:4-40    (min/max corner)
 *   oracle_radius_neighbors <-  geotransformer/extensions/cpu/radius_neighbors/
 *                               radius_neighbors_cpu.cpp:3-91 (per-cloud search, padding)
 *                               extra/nanoflann/nanoflann.hpp:249-253 (d2 <  r2, strict)
 *                               extra/nanoflann/nanoflann.hpp:432-440 (d2 = ((0+dx2)+dy2)+dz2)
 *
 * Where the reference's result order is implementation-defined we use the
 * canonical rules of SURVEY.md 8(c):
 *   - subsampled points: ascending voxel key inside each cloud
 *     (reference: libstdc++ unordered_map iteration order);
 *   - neighbours: (d2 ascending, support index ascending)
 *     (reference: std::sort on d2 only => ties unspecified).
 *
 * Parity pinned: tests/test_oracle_points.py checks this file against
 * oracle/_ref/libse3et_ref.so (the unmodified reference compiled here) on the
 * reference's data/demo pair and on seeded synthetic clouds, and against the
 * committed fixtures in tests/golden/.
 *
 * All arithmetic is IEEE fp32 without contraction (build with -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
  uint64_t key;
  int64_t idx;
} key_idx_t;

static int cmp_key_idx(const void* a, const void* b) {
  const key_idx_t* x = (const key_idx_t*)a;
  const key_idx_t* y = (const key_idx_t*)b;
  if (x->key != y->key) return x->key < y->key ? -1 : 1;
  if (x->idx != y->idx) return x->idx < y->idx ? -1 : 1;
  return 0;
}

/* One cloud.  Returns number of voxels written. */
static int64_t subsample_one(const float* p, const float* nrm, int64_t n, float voxel,
                             float* sp, float* sn) {
  if (n == 0) return 0;
  float mn[3] = {p[0], p[1], p[2]}, mx[3] = {p[0], p[1], p[2]};
  for (int64_t i = 0; i < n; ++i)
    for (int d = 0; d < 3; ++d) {
      float v = p[3 * i + d];
      if (v < mn[d]) mn[d] = v;
      if (v > mx[d]) mx[d] = v;
    }
  /* origin = floor(min * (float)(1.0 / voxel)) * voxel  -- the scale is computed in
   * double and narrowed to float at the PointXYZ*float operator (cpu.cpp:13). */
  const float inv = (float)(1.0 / (double)voxel);
  float org[3];
  for (int d = 0; d < 3; ++d) org[d] = floorf(mn[d] * inv) * voxel;
  const uint64_t nx = (uint64_t)(floorf((mx[0] - org[0]) / voxel) + 1.0f);
  const uint64_t ny = (uint64_t)(floorf((mx[1] - org[1]) / voxel) + 1.0f);

  key_idx_t* ki = (key_idx_t*)malloc(sizeof(key_idx_t) * (size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    /* true fp32 division, then floor, then float->size_t (cpu.cpp:39-42) */
    uint64_t ix = (uint64_t)(int64_t)floorf((p[3 * i + 0] - org[0]) / voxel);
    uint64_t iy = (uint64_t)(int64_t)floorf((p[3 * i + 1] - org[1]) / voxel);
    uint64_t iz = (uint64_t)(int64_t)floorf((p[3 * i + 2] - org[2]) / voxel);
    ki[i].key = ix + nx * iy + nx * ny * iz;
    ki[i].idx = i;
  }
  qsort(ki, (size_t)n, sizeof(key_idx_t), cmp_key_idx);

  int64_t m = 0;
  int64_t a = 0;
  while (a < n) {
    int64_t b = a;
    while (b < n && ki[b].key == ki[a].key) ++b;
    /* running fp32 sum in input-index order (h:41-47); members are already index-sorted */
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (int64_t j = a; j < b; ++j) {
      const float* q = p + 3 * ki[j].idx;
      sx += q[0];
      sy += q[1];
      sz += q[2];
    }
    const float ic = (float)(1.0 / (double)(int)(b - a)); /* h:55 */
    const float ax = sx * ic, ay = sy * ic, az = sz * ic;
    int64_t best = ki[a].idx;
    float bestd = -1.f;
    for (int64_t j = a; j < b; ++j) {
      const float* q = p + 3 * ki[j].idx;
      float dx = q[0] - ax, dy = q[1] - ay, dz = q[2] - az;
      float d = sqrtf(dx * dx + dy * dy + dz * dz); /* h:58-63: float sqrt, strict < */
      if (j == a || d < bestd) {
        bestd = d;
        best = ki[j].idx;
      }
    }
    memcpy(sp + 3 * m, p + 3 * best, 3 * sizeof(float));
    memcpy(sn + 3 * m, nrm + 3 * best, 3 * sizeof(float));
    ++m;
    a = b;
  }
  free(ki);
  return m;
}

/* s_points / s_normals need room for n_total points.  Returns total M. */
int64_t oracle_grid_subsample(const float* points, const int64_t* lengths, const float* normals,
                              int64_t batch, float voxel, float* s_points, int64_t* s_lengths,
                              float* s_normals) {
  int64_t start = 0, m_total = 0;
  for (int64_t b = 0; b < batch; ++b) {
    int64_t m = subsample_one(points + 3 * start, normals + 3 * start, lengths[b], voxel,
                              s_points + 3 * m_total, s_normals + 3 * m_total);
    s_lengths[b] = m;
    m_total += m;
    start += lengths[b];
  }
  return m_total;
}

typedef struct {
  float d2;
  int64_t idx;
} dist_idx_t;

static int cmp_dist_idx(const void* a, const void* b) {
  const dist_idx_t* x = (const dist_idx_t*)a;
  const dist_idx_t* y = (const dist_idx_t*)b;
  if (x->d2 != y->d2) return x->d2 < y->d2 ? -1 : 1;
  if (x->idx != y->idx) return x->idx < y->idx ? -1 : 1;
  return 0;
}

/* Brute force per cloud.  counts[i] = number of support points with d2 < r2.
 * If out != NULL, row i receives its first `width` neighbours in canonical
 * order (global support indices), padded with ns_total.
 * Returns the maximum count. */
int64_t oracle_radius_neighbors(const float* q, const float* s, const int64_t* q_lengths,
                                const int64_t* s_lengths, int64_t batch, float radius,
                                int64_t* counts, int64_t* out, int64_t width) {
  int64_t nq_total = 0, ns_total = 0;
  for (int64_t b = 0; b < batch; ++b) {
    nq_total += q_lengths[b];
    ns_total += s_lengths[b];
  }
  const float r2 = radius * radius; /* fp32 (cpu.cpp:12) */
  int64_t qs = 0, ss = 0, max_count = 0;
  int64_t cap = 1024;
  dist_idx_t* buf = (dist_idx_t*)malloc(sizeof(dist_idx_t) * (size_t)cap);
  for (int64_t b = 0; b < batch; ++b) {
    const int64_t nq = q_lengths[b], ns = s_lengths[b];
    for (int64_t i = 0; i < nq; ++i) {
      const float* qp = q + 3 * (qs + i);
      int64_t c = 0;
      for (int64_t j = 0; j < ns; ++j) {
        const float* sp = s + 3 * (ss + j);
        float d2 = 0.f;
        float dx = qp[0] - sp[0];
        d2 += dx * dx;
        float dy = qp[1] - sp[1];
        d2 += dy * dy;
        float dz = qp[2] - sp[2];
        d2 += dz * dz;
        if (d2 < r2) {
          if (c == cap) {
            cap *= 2;
            buf = (dist_idx_t*)realloc(buf, sizeof(dist_idx_t) * (size_t)cap);
          }
          buf[c].d2 = d2;
          buf[c].idx = ss + j;
          ++c;
        }
      }
      if (counts) counts[qs + i] = c;
      if (c > max_count) max_count = c;
      if (out) {
        qsort(buf, (size_t)c, sizeof(dist_idx_t), cmp_dist_idx);
        int64_t* row = out + (qs + i) * width;
        for (int64_t k = 0; k < width; ++k) row[k] = k < c ? buf[k].idx : ns_total;
      }
    }
    qs += nq;
    ss += ns;
  }
  free(buf);
  return max_count;
}
