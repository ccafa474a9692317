// grid_subsample + radius_neighbors for sm_100a.
//
// Both are HBM/L2-bound integer + fp32-compare work: no tensor cores.
// Bit-exactness with the reference's x86 build requires IEEE fp32 without FMA
// contraction, so every value that decides an index is computed with
// __fsub_rn/__fmul_rn/__fadd_rn/__fdiv_rn/__fsqrt_rn (never contracted by nvcc).
//
// grid_subsample  (reference: grid_subsampling_cpu.cpp:3-71, grid_subsampling_cpu.h:49-73)
//   bounds -> per-cloud grid -> occupancy bitmap over (cloud, reference voxel key)
//   -> popcount prefix scan (rank of a voxel == its canonical output row, ascending key)
//   -> per-voxel member lists -> in-index-order fp32 barycentre -> closest member.
// radius_neighbors (reference: radius_neighbors_cpu.cpp:3-91, nanoflann.hpp:249-253,432-440)
//   hashed uniform grid (cell = radius) over the support set, counting sort into
//   float4 {x,y,z,index} cell runs, one warp per query scanning its 27 cells with
//   coalesced 16-byte loads, ballot compaction of hits into shared memory,
//   in-warp bitonic sort of (d2 bits << 32 | index) keys, int64 row store.
#include "common.cuh"

namespace se3et {

// =============================================================================================
// generic exclusive scan of uint32 items with a device-side length
// =============================================================================================
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;
constexpr int kScanTile = kScanThreads * kScanItems;  // 4096

struct PopcLoad {
  const uint32_t* w;
  __device__ __forceinline__ uint32_t operator()(int64_t i) const { return __popc(w[i]); }
};
struct U32Load {
  const uint32_t* w;
  __device__ __forceinline__ uint32_t operator()(int64_t i) const { return w[i]; }
};

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
  // 256 threads; returns exclusive prefix of v over the block
  __shared__ uint32_t warp_sums[kScanThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  uint32_t woff = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kScanThreads / 32; ++w) {
    uint32_t s = warp_sums[w];
    if (w < warp) woff += s;
    tot += s;
  }
  __syncthreads();
  if (total) *total = tot;
  return woff + incl - v;
}

template <class Load>
__global__ void __launch_bounds__(kScanThreads) scan_partials_kernel(Load ld, const int64_t* n_dev, int64_t cap,
                                                                      uint32_t* block_sums) {
  const int64_t n = n_dev ? min(*n_dev, cap) : cap;
  const int64_t start = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k)
    if (start + k < n) s += ld(start + k);
  uint32_t tot;
  block_exclusive_scan(s, &tot);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(1024) scan_block_sums_kernel(uint32_t* block_sums, int nblocks, int64_t* total_out) {
  __shared__ uint32_t sh[1024];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < nblocks; base += 1024) {
    int i = base + threadIdx.x;
    uint32_t v = i < nblocks ? block_sums[i] : 0;
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      uint32_t t = threadIdx.x >= o ? sh[threadIdx.x - o] : 0;
      __syncthreads();
      sh[threadIdx.x] += t;
      __syncthreads();
    }
    uint32_t incl = sh[threadIdx.x];
    if (i < nblocks) block_sums[i] = carry + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += incl;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

template <class Load>
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(Load ld, const int64_t* n_dev, int64_t cap,
                                                                   const uint32_t* block_sums, uint32_t* out) {
  const int64_t n = n_dev ? min(*n_dev, cap) : cap;
  const int64_t start = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    v[k] = (start + k < n) ? ld(start + k) : 0;
    s += v[k];
  }
  uint32_t run = block_exclusive_scan(s, nullptr) + block_sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (start + k < n) out[start + k] = run;
    run += v[k];
  }
}

static inline int scan_num_blocks(int64_t cap) { return (int)ceil_div(cap > 0 ? cap : 1, kScanTile); }

// out[i] = sum_{j<i} ld(j) for i < n; *total_out = sum of all.  block_sums: scan_num_blocks(cap) words.
template <class Load>
static int exclusive_scan(Load ld, const int64_t* n_dev, int64_t cap, uint32_t* block_sums, uint32_t* out,
                          int64_t* total_out, cudaStream_t st) {
  const int nb = scan_num_blocks(cap);
  scan_partials_kernel<<<nb, kScanThreads, 0, st>>>(ld, n_dev, cap, block_sums);
  SE3ET_LAUNCH_CHECK();
  scan_block_sums_kernel<<<1, 1024, 0, st>>>(block_sums, nb, total_out);
  SE3ET_LAUNCH_CHECK();
  scan_apply_kernel<<<nb, kScanThreads, 0, st>>>(ld, n_dev, cap, block_sums, out);
  SE3ET_LAUNCH_CHECK();
  return SE3ET_OK;
}

// =============================================================================================
// shared: stack-mode offsets and per-cloud bounding boxes
// =============================================================================================
struct Bounds {
  uint32_t mn[3], mx[3];  // order-preserving encodings
};

__global__ void init_offsets_bounds_kernel(const int64_t* __restrict__ len_a, int64_t* off_a,
                                           const int64_t* __restrict__ len_b, int64_t* off_b, int batch,
                                           Bounds* bounds, int32_t* status) {
  if (threadIdx.x == 0) {
    int64_t acc = 0;
    for (int b = 0; b < batch; ++b) { off_a[b] = acc; acc += len_a[b]; }
    off_a[batch] = acc;
    if (len_b) {
      acc = 0;
      for (int b = 0; b < batch; ++b) { off_b[b] = acc; acc += len_b[b]; }
      off_b[batch] = acc;
    }
  }
  for (int b = threadIdx.x; b < batch; b += blockDim.x)
    for (int d = 0; d < 3; ++d) { bounds[b].mn[d] = 0xffffffffu; bounds[b].mx[d] = 0u; }
  if (status)
    for (int i = threadIdx.x; i < SE3ET_STATUS_WORDS; i += blockDim.x) status[i] = 0;
}

__global__ void __launch_bounds__(256) bounds_kernel(const float* __restrict__ pts, int64_t n,
                                                      const int64_t* __restrict__ off, int batch, Bounds* bounds) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < n;
  int b = -1;
  uint32_t lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
  if (valid) {
    b = segment_of(off, batch, i);
    for (int d = 0; d < 3; ++d) lo[d] = hi[d] = float_to_ordered(pts[3 * i + d]);
  }
  const int b0 = __shfl_sync(0xffffffffu, b, 0);
  const bool uniform = __all_sync(0xffffffffu, b == b0) && b0 >= 0;
  if (uniform) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        lo[d] = min(lo[d], __shfl_xor_sync(0xffffffffu, lo[d], o));
        hi[d] = max(hi[d], __shfl_xor_sync(0xffffffffu, hi[d], o));
      }
    }
    if ((threadIdx.x & 31) == 0)
      for (int d = 0; d < 3; ++d) { atomicMin(&bounds[b0].mn[d], lo[d]); atomicMax(&bounds[b0].mx[d], hi[d]); }
  } else if (valid) {
    for (int d = 0; d < 3; ++d) { atomicMin(&bounds[b].mn[d], lo[d]); atomicMax(&bounds[b].mx[d], hi[d]); }
  }
}

// =============================================================================================
// grid_subsample
// =============================================================================================
struct CloudGrid {
  float org[3];
  uint64_t nx, nxny;     // reference key = ix + nx*iy + nxny*iz  (mod 2^64)
  int64_t ncell_pos;     // C = nx*ny*nz: keys of in-range voxels
  int64_t wrap_off;      // keys in [-wrap_off, -1] (a component == -1) map to C + key + wrap_off
  int64_t base;          // first global cell of this cloud
};

__global__ void grid_setup_kernel(const Bounds* __restrict__ bounds, const int64_t* __restrict__ off, int batch,
                                  float voxel, float inv_voxel, int64_t max_cells, CloudGrid* grids,
                                  int64_t* total_words, int32_t* status) {
  for (int b = threadIdx.x; b < batch; b += blockDim.x) {
    CloudGrid g;
    if (off[b + 1] == off[b]) {
      g.org[0] = g.org[1] = g.org[2] = 0.f; g.nx = g.nxny = 0; g.ncell_pos = 0; g.wrap_off = 0;
    } else {
      uint64_t dim[3];
      for (int d = 0; d < 3; ++d) {
        const float mn = ordered_to_float(bounds[b].mn[d]), mx = ordered_to_float(bounds[b].mx[d]);
        g.org[d] = __fmul_rn(floorf(__fmul_rn(mn, inv_voxel)), voxel);                        // cpu.cpp:13
        dim[d] = (uint64_t)__fadd_rn(floorf(__fdiv_rn(__fsub_rn(mx, g.org[d]), voxel)), 1.0f);  // cpu.cpp:15-22
      }
      g.nx = dim[0]; g.nxny = dim[0] * dim[1];
      g.ncell_pos = (int64_t)(dim[0] * dim[1] * dim[2]);
      g.wrap_off = (int64_t)(dim[0] * dim[1] + dim[0] + 1);
    }
    g.base = 0;
    grids[b] = g;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int64_t acc = 0;
    double need = 0.0;
    for (int b = 0; b < batch; ++b) {
      grids[b].base = acc;
      const int64_t c = grids[b].ncell_pos + grids[b].wrap_off;
      need += (double)(uint64_t)c;
      if (need <= (double)max_cells) acc += c;
    }
    if (need > (double)max_cells) {
      atomicOr(&status[SE3ET_STATUS_ERROR], SE3ET_DEV_GRID_TOO_LARGE);
      const double k = need / 1024.0 + 1.0;
      status[SE3ET_STATUS_REQ_KCELLS] = k > 2147483647.0 ? 2147483647 : (int32_t)k;
      acc = 0;
    }
    total_words[0] = (acc + 31) >> 5;
    total_words[1] = acc;  // total cells
  }
}

__global__ void __launch_bounds__(256) zero_words_kernel(uint32_t* w, const int64_t* n_dev) {
  const int64_t n = *n_dev;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) w[i] = 0;
}

__global__ void __launch_bounds__(256) voxel_mark_kernel(const float* __restrict__ pts, int64_t n,
                                                          const int64_t* __restrict__ off, int batch,
                                                          const CloudGrid* __restrict__ grids, float voxel,
                                                          const int64_t* __restrict__ total_words, int64_t* cell_of,
                                                          uint32_t* bitmap, int32_t* status) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (total_words[0] == 0) { cell_of[i] = -1; return; }
  const int b = segment_of(off, batch, i);
  const CloudGrid g = grids[b];
  // true fp32 division then floor then float->size_t (cpu.cpp:39-42); -1 wraps like x86 cvttss2si
  const int64_t ix = (int64_t)floorf(__fdiv_rn(__fsub_rn(pts[3 * i + 0], g.org[0]), voxel));
  const int64_t iy = (int64_t)floorf(__fdiv_rn(__fsub_rn(pts[3 * i + 1], g.org[1]), voxel));
  const int64_t iz = (int64_t)floorf(__fdiv_rn(__fsub_rn(pts[3 * i + 2], g.org[2]), voxel));
  const uint64_t key = (uint64_t)ix + g.nx * (uint64_t)iy + g.nxny * (uint64_t)iz;
  const int64_t sk = (int64_t)key;
  int64_t cell;
  if (sk >= 0) {
    cell = sk;
    if (cell >= g.ncell_pos) { atomicOr(&status[SE3ET_STATUS_ERROR], SE3ET_DEV_INDEX_RANGE); cell = 0; }
  } else {
    const int64_t t = sk + g.wrap_off;  // unsigned order: wrapped keys sort after all in-range keys
    if (t < 0) { atomicOr(&status[SE3ET_STATUS_ERROR], SE3ET_DEV_INDEX_RANGE); cell = 0; }
    else cell = g.ncell_pos + t;
  }
  const int64_t gc = g.base + cell;
  cell_of[i] = gc;
  atomicOr(&bitmap[gc >> 5], 1u << (gc & 31));
}

__device__ __forceinline__ uint32_t rank_of_cell(int64_t gc, const uint32_t* __restrict__ bitmap,
                                                 const uint32_t* __restrict__ word_rank, int64_t nwords,
                                                 uint32_t m_total) {
  const int64_t w = gc >> 5;
  if (w >= nwords) return m_total;
  return word_rank[w] + __popc(bitmap[w] & ((1u << (gc & 31)) - 1u));
}

__global__ void __launch_bounds__(256) voxel_rank_kernel(int64_t n, const int64_t* __restrict__ cell_of,
                                                          const uint32_t* __restrict__ bitmap,
                                                          const uint32_t* __restrict__ word_rank,
                                                          const int64_t* __restrict__ total_words,
                                                          const int64_t* __restrict__ m_total, int32_t* vox_of,
                                                          uint32_t* vox_count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t gc = cell_of[i];
  if (gc < 0) { vox_of[i] = -1; return; }
  const uint32_t v = rank_of_cell(gc, bitmap, word_rank, total_words[0], (uint32_t)*m_total);
  vox_of[i] = (int32_t)v;
  atomicAdd(&vox_count[v], 1u);
}

__global__ void subsample_lengths_kernel(const CloudGrid* __restrict__ grids, int batch,
                                         const uint32_t* __restrict__ bitmap, const uint32_t* __restrict__ word_rank,
                                         const int64_t* __restrict__ total_words, const int64_t* __restrict__ m_total,
                                         int64_t* s_lengths, int32_t* status) {
  const uint32_t m = (uint32_t)*m_total;
  for (int b = threadIdx.x; b < batch; b += blockDim.x) {
    const int64_t lo = grids[b].base;
    const int64_t hi = (b + 1 < batch) ? grids[b + 1].base : total_words[1];
    const int64_t len = (int64_t)rank_of_cell(hi, bitmap, word_rank, total_words[0], m) -
                        (int64_t)rank_of_cell(lo, bitmap, word_rank, total_words[0], m);
    s_lengths[b] = len;
    atomicMax(&status[SE3ET_STATUS_MAX_LENGTH], (int32_t)len);
  }
  if (threadIdx.x == 0) status[SE3ET_STATUS_M_TOTAL] = (int32_t)m;
}

__global__ void __launch_bounds__(256) voxel_fill_kernel(int64_t n, const int32_t* __restrict__ vox_of,
                                                          const uint32_t* __restrict__ vox_start,
                                                          uint32_t* vox_cursor, int32_t* members) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t v = vox_of[i];
  if (v < 0) return;
  const uint32_t slot = atomicAdd(&vox_cursor[v], 1u);
  members[vox_start[v] + slot] = (int32_t)i;
}

__global__ void __launch_bounds__(128) voxel_choose_kernel(const float* __restrict__ pts,
                                                            const float* __restrict__ nrm,
                                                            const int64_t* __restrict__ m_total,
                                                            const uint32_t* __restrict__ vox_start,
                                                            const uint32_t* __restrict__ vox_count, int32_t* members,
                                                            float* s_points, float* s_normals) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= *m_total) return;
  int32_t* mem = members + vox_start[v];
  const int cnt = (int)vox_count[v];
  // restore input order (the atomic fill order is arbitrary): insertion sort, lists are short
  for (int a = 1; a < cnt; ++a) {
    const int32_t x = mem[a];
    int c = a - 1;
    while (c >= 0 && mem[c] > x) { mem[c + 1] = mem[c]; --c; }
    mem[c + 1] = x;
  }
  float sx = 0.f, sy = 0.f, sz = 0.f;  // running fp32 sums in input order (h:41-47)
  for (int a = 0; a < cnt; ++a) {
    const float* p = pts + 3 * (int64_t)mem[a];
    sx = __fadd_rn(sx, p[0]); sy = __fadd_rn(sy, p[1]); sz = __fadd_rn(sz, p[2]);
  }
  const float ic = (float)(1.0 / (double)cnt);  // h:55: double reciprocal narrowed to float
  const float ax = __fmul_rn(sx, ic), ay = __fmul_rn(sy, ic), az = __fmul_rn(sz, ic);
  int32_t best = mem[0];
  float bestd = 0.f;
  for (int a = 0; a < cnt; ++a) {
    const float* p = pts + 3 * (int64_t)mem[a];
    const float dx = __fsub_rn(p[0], ax), dy = __fsub_rn(p[1], ay), dz = __fsub_rn(p[2], az);
    const float d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    if (a == 0 || d < bestd) { bestd = d; best = mem[a]; }  // strict <: lowest index wins ties (h:64)
  }
  for (int d = 0; d < 3; ++d) {
    s_points[3 * v + d] = pts[3 * (int64_t)best + d];
    s_normals[3 * v + d] = nrm[3 * (int64_t)best + d];
  }
}

struct SubsampleWs {
  int64_t* off;          // batch+1
  Bounds* bounds;        // batch
  CloudGrid* grids;      // batch
  int64_t* total_words;  // [0]=words [1]=cells
  int64_t* m_total;      // 1
  uint32_t* bitmap;      // max_cells/32
  uint32_t* word_rank;   // max_cells/32
  uint32_t* block_sums;  // scan scratch
  int64_t* cell_of;      // n
  int32_t* vox_of;       // n
  uint32_t* vox_count;   // n
  uint32_t* vox_start;   // n
  uint32_t* vox_cursor;  // n
  int32_t* members;      // n
};

static bool carve_subsample(Carver& c, int64_t n, int64_t batch, int64_t max_cells, SubsampleWs& w) {
  const int64_t words = ceil_div(max_cells, 32);
  const int64_t nn = n > 0 ? n : 1;
  w.off = c.take<int64_t>(batch + 1);
  w.bounds = c.take<Bounds>(batch);
  w.grids = c.take<CloudGrid>(batch);
  w.total_words = c.take<int64_t>(2);
  w.m_total = c.take<int64_t>(1);
  w.bitmap = c.take<uint32_t>(words);
  w.word_rank = c.take<uint32_t>(words);
  w.block_sums = c.take<uint32_t>(scan_num_blocks(words > nn ? words : nn));
  w.cell_of = c.take<int64_t>(nn);
  w.vox_of = c.take<int32_t>(nn);
  w.vox_count = c.take<uint32_t>(nn);
  w.vox_start = c.take<uint32_t>(nn);
  w.vox_cursor = c.take<uint32_t>(nn);
  w.members = c.take<int32_t>(nn);
  return c.fits();
}

// =============================================================================================
// radius_neighbors
// =============================================================================================
__device__ __forceinline__ uint32_t cell_hash(int b, int cx, int cy, int cz) {
  uint32_t h = (uint32_t)cx * 73856093u ^ (uint32_t)cy * 19349663u ^ (uint32_t)cz * 83492791u ^ (uint32_t)b * 2654435761u;
  h ^= h >> 15; h *= 0x2c1b3c6du; h ^= h >> 12;
  return h;
}

__device__ __forceinline__ void cell_of_point(float x, float y, float z, const float* mn, float inv_cell, int& cx,
                                              int& cy, int& cz) {
  cx = (int)floorf((x - mn[0]) * inv_cell);
  cy = (int)floorf((y - mn[1]) * inv_cell);
  cz = (int)floorf((z - mn[2]) * inv_cell);
}

__global__ void cloud_min_kernel(const Bounds* __restrict__ bounds, int batch, float* mins) {
  for (int b = threadIdx.x; b < batch; b += blockDim.x)
    for (int d = 0; d < 3; ++d) {
      const uint32_t u = bounds[b].mn[d];
      mins[3 * b + d] = (u == 0xffffffffu) ? 0.f : ordered_to_float(u);  // empty cloud
    }
}

__global__ void __launch_bounds__(256) hash_count_kernel(const float* __restrict__ s, int64_t ns,
                                                          const int64_t* __restrict__ s_off, int batch,
                                                          const float* __restrict__ mins, float inv_cell,
                                                          uint32_t mask, uint32_t* bucket_of, uint32_t* hist) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ns) return;
  const int b = segment_of(s_off, batch, i);
  int cx, cy, cz;
  cell_of_point(s[3 * i], s[3 * i + 1], s[3 * i + 2], mins + 3 * b, inv_cell, cx, cy, cz);
  const uint32_t h = cell_hash(b, cx, cy, cz) & mask;
  bucket_of[i] = h;
  atomicAdd(&hist[h], 1u);
}

__global__ void __launch_bounds__(256) hash_scatter_kernel(const float* __restrict__ s, int64_t ns,
                                                            const uint32_t* __restrict__ bucket_of,
                                                            const uint32_t* __restrict__ start, uint32_t* cursor,
                                                            float4* sorted) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ns) return;
  const uint32_t h = bucket_of[i];
  const uint32_t pos = start[h] + atomicAdd(&cursor[h], 1u);
  sorted[pos] = make_float4(s[3 * i], s[3 * i + 1], s[3 * i + 2], __int_as_float((int)i));
}

constexpr int kQueryWarps = 8;
constexpr int kHitCap = 256;  // hits per query kept in shared memory; beyond that a slow exact path runs

struct QueryCtx {
  float qx, qy, qz, r2, inv_cell;
  const float* mn;
  int cx, cy, cz;
  int s_lo, s_hi;
  const float4* sorted;
};

// evaluates candidate t of the concatenated 27 cell runs; returns true and the sort key on a hit
__device__ __forceinline__ bool eval_candidate(const QueryCtx& c, const int* sh_excl, const int* sh_start, int t,
                                               unsigned long long& key) {
  int lo = 0, hi = 26;  // largest cell index with excl <= t
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (sh_excl[mid] <= t) lo = mid; else hi = mid - 1;
  }
  const float4 sp = c.sorted[sh_start[lo] + (t - sh_excl[lo])];
  const int j = __float_as_int(sp.w);
  if (j < c.s_lo || j >= c.s_hi) return false;  // hash collision with another cloud
  int scx, scy, scz;
  cell_of_point(sp.x, sp.y, sp.z, c.mn, c.inv_cell, scx, scy, scz);
  // accept only from the probed cell: rejects bucket collisions and double visits
  if (scx != c.cx + (lo % 3) - 1 || scy != c.cy + ((lo / 3) % 3) - 1 || scz != c.cz + (lo / 9) - 1) return false;
  const float dx = __fsub_rn(c.qx, sp.x), dy = __fsub_rn(c.qy, sp.y), dz = __fsub_rn(c.qz, sp.z);
  // nanoflann.hpp:432-440: result = ((0 + dx*dx) + dy*dy) + dz*dz
  const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
  if (!(d2 < c.r2)) return false;  // nanoflann.hpp:249-253, strict
  key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned int)j;
  return true;
}

struct QueryArgs {
  const float* q;
  int64_t nq;
  const int64_t* q_off;
  const int64_t* s_off;
  int batch;
  const float* mins;
  const float4* sorted;
  const uint32_t* start;
  uint32_t mask;
  float inv_cell, r2;
  int32_t* counts;
  int64_t* out;
  int64_t width, ns_total;
  int32_t* status;
  int32_t* cloud_max;
};

// one query per warp: 27-cell scan of the hashed grid, exact for any neighbour count
__device__ __forceinline__ void query_one(const QueryArgs& A, int64_t qi, int lane, unsigned long long* keys,
                                          int* sh_excl_w, int* sh_start_w) {
  const float* __restrict__ q = A.q;
  const int64_t* __restrict__ q_off = A.q_off;
  const int64_t* __restrict__ s_off = A.s_off;
  const int batch = A.batch;
  const float* __restrict__ mins = A.mins;
  const float4* __restrict__ sorted = A.sorted;
  const uint32_t* __restrict__ start = A.start;
  const uint32_t mask = A.mask;
  const float inv_cell = A.inv_cell, r2 = A.r2;
  int32_t* counts = A.counts;
  int64_t* out = A.out;
  const int64_t width = A.width, ns_total = A.ns_total;
  int32_t* status = A.status;
  int32_t* cloud_max = A.cloud_max;
  const int b = segment_of(q_off, batch, qi);
  QueryCtx c;
  c.qx = q[3 * qi]; c.qy = q[3 * qi + 1]; c.qz = q[3 * qi + 2];
  c.r2 = r2; c.inv_cell = inv_cell; c.mn = mins + 3 * b; c.sorted = sorted;
  c.s_lo = (int)s_off[b]; c.s_hi = (int)s_off[b + 1];
  cell_of_point(c.qx, c.qy, c.qz, c.mn, inv_cell, c.cx, c.cy, c.cz);

  int st = 0, len = 0;
  if (lane < 27 && c.s_hi > c.s_lo) {
    const uint32_t h = cell_hash(b, c.cx + (lane % 3) - 1, c.cy + ((lane / 3) % 3) - 1, c.cz + (lane / 9) - 1) & mask;
    st = (int)start[h];
    len = (int)start[h + 1] - st;
  }
  int incl = len;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  const int total = __shfl_sync(0xffffffffu, incl, 31);
  sh_excl_w[lane] = incl - len;
  sh_start_w[lane] = st;
  __syncwarp();

  int nhit = 0;
  for (int base = 0; base < total; base += 32) {
    const int t = base + lane;
    unsigned long long key = 0;
    const bool hit = t < total && eval_candidate(c, sh_excl_w, sh_start_w, t, key);
    const unsigned ballot = __ballot_sync(0xffffffffu, hit);
    const int pos = nhit + __popc(ballot & ((1u << lane) - 1u));
    if (hit && pos < kHitCap) keys[pos] = key;
    nhit += __popc(ballot);
  }
  if (lane == 0) {
    if (counts) counts[qi] = nhit;
    atomicMax(&status[SE3ET_STATUS_MAX_COUNT], nhit);
    if (cloud_max) atomicMax(&cloud_max[b], nhit);
  }
  if (!out || width <= 0) return;
  int64_t* row = out + qi * width;
  if (nhit <= 32) {
    // at most one key per lane: rank it against all others (keys are unique, the rank is the sorted position)
    __syncwarp();
    const unsigned long long mine = lane < nhit ? keys[lane] : ~0ull;
    int rank = 0;
#pragma unroll 4
    for (int f = 0; f < nhit; ++f) rank += keys[f] < mine ? 1 : 0;
    if (lane < nhit && rank < width) row[rank] = (int64_t)(mine & 0xffffffffull);
    for (int k = nhit + lane; k < width; k += 32) row[k] = ns_total;
  } else if (nhit <= 64) {
    // the common case (neighbour limits are <= 40): rank every key against all others -- keys are unique, so the rank
    // is its sorted position -- instead of a shared-memory bitonic network (21 synchronised passes for 64 keys)
    __syncwarp();
    unsigned long long mine[2];
    int rank[2] = {0, 0};
#pragma unroll
    for (int u = 0; u < 2; ++u) mine[u] = lane + 32 * u < nhit ? keys[lane + 32 * u] : ~0ull;
#pragma unroll 4
    for (int f = 0; f < nhit; ++f) {
      const unsigned long long kf = keys[f];
      rank[0] += kf < mine[0] ? 1 : 0;
      rank[1] += kf < mine[1] ? 1 : 0;
    }
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (lane + 32 * u < nhit && rank[u] < width) row[rank[u]] = (int64_t)(mine[u] & 0xffffffffull);
    for (int k = nhit + lane; k < width; k += 32) row[k] = ns_total;
  } else if (nhit <= kHitCap) {
    int n2 = 32;
    while (n2 < nhit) n2 <<= 1;
    for (int i = nhit + lane; i < n2; i += 32) keys[i] = ~0ull;
    __syncwarp();
    for (int k = 2; k <= n2; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = lane; i < n2; i += 32) {
          const int p = i ^ j;
          if (p > i) {
            const unsigned long long a = keys[i], bb = keys[p];
            const bool up = (i & k) == 0;
            if ((a > bb) == up) { keys[i] = bb; keys[p] = a; }
          }
        }
        __syncwarp();
      }
    }
    for (int k = lane; k < width; k += 32) row[k] = k < nhit ? (int64_t)(keys[k] & 0xffffffffull) : ns_total;
  } else {
    // exact but slow: emit the next-smallest key per pass (only for > kHitCap neighbours)
    unsigned long long last = 0;
    const int nout = nhit < width ? nhit : (int)width;
    for (int k = 0; k < nout; ++k) {
      unsigned long long best = ~0ull;
      for (int t = lane; t < total; t += 32) {
        unsigned long long key;
        if (eval_candidate(c, sh_excl_w, sh_start_w, t, key) && (k == 0 || key > last) && key < best) best = key;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other < best ? other : best;
      }
      last = best;
      if (lane == 0) row[k] = (int64_t)(best & 0xffffffffull);
    }
    for (int k = nout + lane; k < width; k += 32) row[k] = ns_total;
  }
}

__global__ void __launch_bounds__(kQueryWarps * 32) radius_query_kernel(QueryArgs A) {
  __shared__ unsigned long long sh_keys[kQueryWarps][kHitCap];
  __shared__ int sh_excl[kQueryWarps][32];
  __shared__ int sh_start[kQueryWarps][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t qi = (int64_t)blockIdx.x * kQueryWarps + warp;
  if (qi >= A.nq) return;  // whole warp exits; only __syncwarp below
  query_one(A, qi, lane, sh_keys[warp], sh_excl[warp], sh_start[warp]);
}

// the same search for a device-side list of queries (the by-cell kernel's overflow cases)
__global__ void __launch_bounds__(kQueryWarps * 32) radius_query_list_kernel(QueryArgs A, const int32_t* __restrict__ list,
                                                                            const int32_t* __restrict__ count) {
  __shared__ unsigned long long sh_keys[kQueryWarps][kHitCap];
  __shared__ int sh_excl[kQueryWarps][32];
  __shared__ int sh_start[kQueryWarps][32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = *count;
  for (int i = blockIdx.x * kQueryWarps + warp; i < n; i += gridDim.x * kQueryWarps) {
    query_one(A, list[i], lane, sh_keys[warp], sh_excl[warp], sh_start[warp]);
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Query by cell (round 2).  All queries of one grid cell share their 27-cell candidate set, so the queries are bucketed
// with the support set's hash and a warp works through the non-empty query buckets: per cell the candidate runs are
// scanned, validated (cloud, cell coordinates) and compacted into shared memory ONCE; per query only the distance test
// over the staged candidates and the ranking remain (~2x fewer instructions per query than query_one).  Cells with more
// than kCandCap candidates and queries with more than kCellHitCap hits go to a device-side list that
// radius_query_list_kernel finishes with the exact general path.  Results are the same set ordered by the same key.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kCellWarps = 8;
constexpr int kCandCap = 320;
constexpr int kCellHitCap = 128;
constexpr int kCellWarpBytes = kCandCap * 16 + kCellHitCap * 8 + 2 * 32 * 4;

__device__ __forceinline__ void emit_row_small(const unsigned long long* keys, int nhit, int64_t* row, int64_t width,
                                               int64_t ns_total, int lane) {
  // nhit <= kCellHitCap = 128: rank every key against all others (keys are unique: the rank is the sorted position)
  unsigned long long mine[4];
  int rank[4] = {0, 0, 0, 0};
  const int per = (nhit + 31) >> 5;  // keys per lane, 1..4 (warp-uniform)
#pragma unroll
  for (int u = 0; u < 4; ++u) mine[u] = (u < per && lane + 32 * u < nhit) ? keys[lane + 32 * u] : ~0ull;
  if (per == 1) {
#pragma unroll 4
    for (int f = 0; f < nhit; ++f) rank[0] += keys[f] < mine[0] ? 1 : 0;
  } else if (per == 2) {
#pragma unroll 4
    for (int f = 0; f < nhit; ++f) {
      const unsigned long long kf = keys[f];
      rank[0] += kf < mine[0] ? 1 : 0;
      rank[1] += kf < mine[1] ? 1 : 0;
    }
  } else {
#pragma unroll 2
    for (int f = 0; f < nhit; ++f) {
      const unsigned long long kf = keys[f];
#pragma unroll
      for (int u = 0; u < 4; ++u) rank[u] += kf < mine[u] ? 1 : 0;
    }
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
    if (u < per && lane + 32 * u < nhit && rank[u] < width) row[rank[u]] = (int64_t)(mine[u] & 0xffffffffull);
  for (int k = nhit + lane; k < width; k += 32) row[k] = ns_total;
}

__global__ void __launch_bounds__(kCellWarps * 32) radius_cell_kernel(QueryArgs A, const float4* __restrict__ q_sorted,
                                                                      const uint32_t* __restrict__ q_start,
                                                                      uint32_t table, int32_t* __restrict__ fb_list,
                                                                      int32_t* __restrict__ fb_count) {
  extern __shared__ __align__(16) uint8_t cell_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* base = cell_smem + warp * kCellWarpBytes;
  float4* cand = reinterpret_cast<float4*>(base);
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(base + kCandCap * 16);
  int* sh_excl = reinterpret_cast<int*>(base + kCandCap * 16 + kCellHitCap * 8);
  int* sh_start = sh_excl + 32;
  const unsigned full = 0xffffffffu;
  const unsigned lt_mask = (1u << lane) - 1u;
  const float inv_cell = A.inv_cell, r2 = A.r2;
  int warp_max = 0;
  const uint32_t nchunks = (table + 31) / 32;
  for (;;) {
    // 32 consecutive buckets per trip (most are empty), handed out dynamically: cells differ a lot in work
    uint32_t chunk = 0;
    if (lane == 0) chunk = (uint32_t)atomicAdd(fb_count + 1, 1);
    chunk = __shfl_sync(full, chunk, 0);
    if (chunk >= nchunks) break;
    const uint32_t hb = chunk * 32 + lane;
    const uint32_t my_s = hb <= table ? q_start[hb] : 0u;
    uint32_t my_e = __shfl_down_sync(full, my_s, 1);
    if (lane == 31) my_e = hb + 1 <= table ? q_start[hb + 1] : my_s;
    if (hb >= table) my_e = my_s;
    unsigned nonempty = __ballot_sync(full, my_e > my_s);
    while (nonempty) {
      const int bl = __ffs(nonempty) - 1;
      nonempty &= nonempty - 1;
      const uint32_t qs = __shfl_sync(full, my_s, bl), qe = __shfl_sync(full, my_e, bl);
      for (uint32_t qb = qs; qb < qe; qb += 32) {
        const bool have = qb + lane < qe;
        const float4 qv = have ? q_sorted[qb + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
        const int qi = have ? __float_as_int(qv.w) : -1;
        int b = -1, cx = 0, cy = 0, cz = 0;
        if (have) {
          b = segment_of(A.q_off, A.batch, qi);
          cell_of_point(qv.x, qv.y, qv.z, A.mins + 3 * b, inv_cell, cx, cy, cz);
        }
        unsigned remaining = __ballot_sync(full, have);
        while (remaining) {
          const int leader = __ffs(remaining) - 1;
          const int lb = __shfl_sync(full, b, leader), lcx = __shfl_sync(full, cx, leader);
          const int lcy = __shfl_sync(full, cy, leader), lcz = __shfl_sync(full, cz, leader);
          const unsigned group = __ballot_sync(full, have && b == lb && cx == lcx && cy == lcy && cz == lcz);
          remaining &= ~group;
          // ---- candidates of the cell's 27 neighbours, validated and compacted once --------------------------------
          const int s_lo = (int)A.s_off[lb], s_hi = (int)A.s_off[lb + 1];
          const float* mn = A.mins + 3 * lb;
          int st = 0, len = 0;
          if (lane < 27 && s_hi > s_lo) {
            const uint32_t h = cell_hash(lb, lcx + (lane % 3) - 1, lcy + ((lane / 3) % 3) - 1, lcz + (lane / 9) - 1) & A.mask;
            st = (int)A.start[h];
            len = (int)A.start[h + 1] - st;
          }
          int incl = len;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(full, incl, o);
            if (lane >= o) incl += t;
          }
          const int total = __shfl_sync(full, incl, 31);
          __syncwarp();
          sh_excl[lane] = incl - len;
          sh_start[lane] = st;
          __syncwarp();
          int ncand = 0;
          for (int tb = 0; tb < total; tb += 32) {
            const int t = tb + lane;
            bool ok = false;
            float4 sp = make_float4(0.f, 0.f, 0.f, 0.f);
            if (t < total) {
              int lo = 0, hi = 26;  // largest cell index with excl <= t
              while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (sh_excl[mid] <= t) lo = mid; else hi = mid - 1;
              }
              sp = A.sorted[sh_start[lo] + (t - sh_excl[lo])];
              const int j = __float_as_int(sp.w);
              if (j >= s_lo && j < s_hi) {  // else: hash collision with another cloud
                int scx, scy, scz;
                cell_of_point(sp.x, sp.y, sp.z, mn, inv_cell, scx, scy, scz);
                // accept only from the probed cell: rejects bucket collisions and double visits
                ok = scx == lcx + (lo % 3) - 1 && scy == lcy + ((lo / 3) % 3) - 1 && scz == lcz + (lo / 9) - 1;
              }
            }
            const unsigned bal = __ballot_sync(full, ok);
            const int pos = ncand + __popc(bal & lt_mask);
            if (ok && pos < kCandCap) cand[pos] = sp;
            ncand += __popc(bal);
          }
          __syncwarp();
          if (ncand > kCandCap) {  // dense cell: the general kernel finishes these queries
            if ((group >> lane) & 1u) fb_list[atomicAdd(fb_count, 1)] = qi;
            continue;
          }
          // ---- the queries of the cell, one after the other ---------------------------------------------------------
          int gmax = 0;
          unsigned g = group;
          while (g) {
            const int ql = __ffs(g) - 1;
            g &= g - 1;
            const float qx = __shfl_sync(full, qv.x, ql), qy = __shfl_sync(full, qv.y, ql);
            const float qz = __shfl_sync(full, qv.z, ql);
            const int qq = __shfl_sync(full, qi, ql);
            int nhit = 0;
            for (int tb = 0; tb < ncand; tb += 32) {
              const int t = tb + lane;
              bool hit = false;
              unsigned long long key = 0;
              if (t < ncand) {
                const float4 sp = cand[t];
                const float dx = __fsub_rn(qx, sp.x), dy = __fsub_rn(qy, sp.y), dz = __fsub_rn(qz, sp.z);
                // nanoflann.hpp:432-440: result = ((0 + dx*dx) + dy*dy) + dz*dz
                const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
                hit = d2 < r2;  // nanoflann.hpp:249-253, strict
                key = ((unsigned long long)__float_as_uint(d2) << 32) | (unsigned int)__float_as_int(sp.w);
              }
              const unsigned bal = __ballot_sync(full, hit);
              const int pos = nhit + __popc(bal & lt_mask);
              if (hit && pos < kCellHitCap) keys[pos] = key;
              nhit += __popc(bal);
            }
            if (nhit > kCellHitCap) {  // general kernel (it also reports the count)
              if (lane == 0) fb_list[atomicAdd(fb_count, 1)] = qq;
              continue;
            }
            gmax = max(gmax, nhit);
            if (lane == 0 && A.counts) A.counts[qq] = nhit;
            if (A.out && A.width > 0) {
              __syncwarp();
              emit_row_small(keys, nhit, A.out + (int64_t)qq * A.width, A.width, A.ns_total, lane);
            }
            __syncwarp();  // keys are reused by the next query
          }
          if (lane == 0 && A.cloud_max && gmax > 0) atomicMax(&A.cloud_max[lb], gmax);
          warp_max = max(warp_max, gmax);
        }
      }
    }
  }
  if (lane == 0 && warp_max > 0) atomicMax(&A.status[SE3ET_STATUS_MAX_COUNT], warp_max);
}

struct RadiusWs {
  int64_t *q_off, *s_off;
  Bounds* bounds;
  float* mins;
  uint32_t *bucket_of, *hist, *start, *cursor, *block_sums;
  float4* sorted;
  int64_t table;
  // query side of the by-cell search
  uint32_t *q_bucket_of, *q_hist, *q_start, *q_cursor;
  float4* q_sorted;
  int32_t *fb_list, *fb_count;
};

static int64_t hash_table_size(int64_t ns) {
  int64_t t = 1024;
  while (t < 2 * ns) t <<= 1;
  return t;
}

static bool carve_radius(Carver& c, int64_t nq, int64_t ns, int64_t batch, RadiusWs& w) {
  w.table = hash_table_size(ns);
  const int64_t nn = ns > 0 ? ns : 1;
  w.q_off = c.take<int64_t>(batch + 1);
  w.s_off = c.take<int64_t>(batch + 1);
  w.bounds = c.take<Bounds>(batch);
  w.mins = c.take<float>(3 * batch);
  w.bucket_of = c.take<uint32_t>(nn);
  w.hist = c.take<uint32_t>(w.table + 1);
  w.start = c.take<uint32_t>(w.table + 1);
  w.cursor = c.take<uint32_t>(w.table + 1);
  w.block_sums = c.take<uint32_t>(scan_num_blocks(w.table + 1));
  w.sorted = c.take<float4>(nn);
  const int64_t nnq = nq > 0 ? nq : 1;
  w.q_bucket_of = c.take<uint32_t>(nnq);
  w.q_hist = c.take<uint32_t>(w.table + 1);
  w.q_start = c.take<uint32_t>(w.table + 1);
  w.q_cursor = c.take<uint32_t>(w.table + 1);
  w.q_sorted = c.take<float4>(nnq);
  w.fb_list = c.take<int32_t>(nnq);
  w.fb_count = c.take<int32_t>(4);
  return c.fits();
}

}  // namespace se3et

using namespace se3et;

extern "C" int se3et_grid_subsample_workspace_bytes(int64_t n_total, int64_t batch, int64_t max_cells,
                                                    size_t* bytes) {
  if (!bytes || n_total < 0 || batch <= 0 || max_cells <= 0) return SE3ET_ERR_ARG;
  Carver c(nullptr, 0);
  SubsampleWs w;
  carve_subsample(c, n_total, batch, max_cells, w);
  *bytes = c.off + 256;
  return SE3ET_OK;
}

extern "C" int se3et_grid_subsample(const float* points, const int64_t* lengths, const float* normals,
                                    int64_t n_total, int64_t batch, float voxel_size, float* s_points,
                                    int64_t* s_lengths, float* s_normals, int32_t* status, void* workspace,
                                    size_t workspace_bytes, int64_t max_cells, se3et_stream_t stream) {
  if (!lengths || !s_lengths || !status || !workspace || n_total < 0 || batch <= 0 || max_cells <= 0 ||
      !(voxel_size > 0.f) || n_total >= (int64_t)1 << 31)
    return SE3ET_ERR_ARG;
  if (n_total > 0 && (!points || !normals || !s_points || !s_normals)) return SE3ET_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Carver c(workspace, workspace_bytes);
  SubsampleWs w;
  if (!carve_subsample(c, n_total, batch, max_cells, w)) return SE3ET_ERR_WORKSPACE;
  const int64_t words = ceil_div(max_cells, 32);
  const float inv_voxel = (float)(1.0 / (double)voxel_size);  // cpu.cpp:13 `(1. / voxel_size)` narrowed at operator*
  const int nblk = (int)ceil_div(n_total > 0 ? n_total : 1, 256);

  init_offsets_bounds_kernel<<<1, 256, 0, st>>>(lengths, w.off, nullptr, nullptr, (int)batch, w.bounds, status);
  SE3ET_LAUNCH_CHECK();
  if (n_total > 0) {
    bounds_kernel<<<nblk, 256, 0, st>>>(points, n_total, w.off, (int)batch, w.bounds);
    SE3ET_LAUNCH_CHECK();
  }
  grid_setup_kernel<<<1, 256, 0, st>>>(w.bounds, w.off, (int)batch, voxel_size, inv_voxel, max_cells, w.grids,
                                       w.total_words, status);
  SE3ET_LAUNCH_CHECK();
  zero_words_kernel<<<kNumSMs * 4, 256, 0, st>>>(w.bitmap, w.total_words);
  SE3ET_LAUNCH_CHECK();
  SE3ET_CUDA_CHECK(cudaMemsetAsync(w.vox_count, 0, sizeof(uint32_t) * (n_total > 0 ? n_total : 1), st));
  SE3ET_CUDA_CHECK(cudaMemsetAsync(w.vox_cursor, 0, sizeof(uint32_t) * (n_total > 0 ? n_total : 1), st));
  if (n_total > 0) {
    voxel_mark_kernel<<<nblk, 256, 0, st>>>(points, n_total, w.off, (int)batch, w.grids, voxel_size, w.total_words,
                                            w.cell_of, w.bitmap, status);
    SE3ET_LAUNCH_CHECK();
  }
  int rc = exclusive_scan(PopcLoad{w.bitmap}, w.total_words, words, w.block_sums, w.word_rank, w.m_total, st);
  if (rc) return rc;
  if (n_total > 0) {
    voxel_rank_kernel<<<nblk, 256, 0, st>>>(n_total, w.cell_of, w.bitmap, w.word_rank, w.total_words, w.m_total,
                                            w.vox_of, w.vox_count);
    SE3ET_LAUNCH_CHECK();
  }
  subsample_lengths_kernel<<<1, 256, 0, st>>>(w.grids, (int)batch, w.bitmap, w.word_rank, w.total_words, w.m_total,
                                              s_lengths, status);
  SE3ET_LAUNCH_CHECK();
  if (n_total > 0) {
    rc = exclusive_scan(U32Load{w.vox_count}, w.m_total, n_total, w.block_sums, w.vox_start, nullptr, st);
    if (rc) return rc;
    voxel_fill_kernel<<<nblk, 256, 0, st>>>(n_total, w.vox_of, w.vox_start, w.vox_cursor, w.members);
    SE3ET_LAUNCH_CHECK();
    voxel_choose_kernel<<<(int)ceil_div(n_total, 128), 128, 0, st>>>(points, normals, w.m_total, w.vox_start,
                                                                    w.vox_count, w.members, s_points, s_normals);
    SE3ET_LAUNCH_CHECK();
  }
  return SE3ET_OK;
}

static int radius_mode_default() {
  const char* e = getenv("SE3ET_RADIUS_MODE");  // A/B measurements: 0 per query, 1 by cell, 2 automatic
  return (e && e[0] >= '0' && e[0] <= '2' && !e[1]) ? e[0] - '0' : 2;
}
static std::atomic<int> g_radius_mode{radius_mode_default()};

extern "C" int se3et_radius_set_mode(int mode) {
  if (mode < 0 || mode > 2) return SE3ET_ERR_ARG;
  g_radius_mode.store(mode, std::memory_order_relaxed);
  return SE3ET_OK;
}

extern "C" int se3et_radius_neighbors_workspace_bytes(int64_t nq_total, int64_t ns_total, int64_t batch,
                                                      size_t* bytes) {
  if (!bytes || nq_total < 0 || ns_total < 0 || batch <= 0) return SE3ET_ERR_ARG;
  Carver c(nullptr, 0);
  RadiusWs w;
  carve_radius(c, nq_total, ns_total, batch, w);
  *bytes = c.off + 256;
  return SE3ET_OK;
}

extern "C" int se3et_radius_neighbors(const float* q_points, const float* s_points, const int64_t* q_lengths,
                                      const int64_t* s_lengths, int64_t nq_total, int64_t ns_total, int64_t batch,
                                      float radius, int32_t* counts, int64_t* out, int64_t width, int32_t* status,
                                      int32_t* cloud_max, void* workspace, size_t workspace_bytes,
                                      se3et_stream_t stream) {
  if (!q_lengths || !s_lengths || !status || !workspace || nq_total < 0 || ns_total < 0 || batch <= 0 ||
      width < 0 || !(radius > 0.f) || ns_total >= (int64_t)1 << 31 || nq_total >= (int64_t)1 << 31)
    return SE3ET_ERR_ARG;
  if ((nq_total > 0 && !q_points) || (ns_total > 0 && !s_points)) return SE3ET_ERR_ARG;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Carver c(workspace, workspace_bytes);
  RadiusWs w;
  if (!carve_radius(c, nq_total, ns_total, batch, w)) return SE3ET_ERR_WORKSPACE;
  const uint32_t mask = (uint32_t)(w.table - 1);
  const float r2 = radius * radius;                  // fp32 product (radius_neighbors_cpu.cpp:12)
  const float inv_cell = 1.0f / (radius * 1.001f);   // cell slightly larger than the radius: 27 cells always suffice

  init_offsets_bounds_kernel<<<1, 256, 0, st>>>(q_lengths, w.q_off, s_lengths, w.s_off, (int)batch, w.bounds, status);
  SE3ET_LAUNCH_CHECK();
  if (cloud_max) SE3ET_CUDA_CHECK(cudaMemsetAsync(cloud_max, 0, sizeof(int32_t) * batch, st));
  SE3ET_CUDA_CHECK(cudaMemsetAsync(w.hist, 0, sizeof(uint32_t) * (w.table + 1), st));
  SE3ET_CUDA_CHECK(cudaMemsetAsync(w.cursor, 0, sizeof(uint32_t) * (w.table + 1), st));
  if (ns_total > 0) {
    const int sblk = (int)ceil_div(ns_total, 256);
    bounds_kernel<<<sblk, 256, 0, st>>>(s_points, ns_total, w.s_off, (int)batch, w.bounds);
    SE3ET_LAUNCH_CHECK();
    cloud_min_kernel<<<1, 256, 0, st>>>(w.bounds, (int)batch, w.mins);
    SE3ET_LAUNCH_CHECK();
    hash_count_kernel<<<sblk, 256, 0, st>>>(s_points, ns_total, w.s_off, (int)batch, w.mins, inv_cell, mask,
                                            w.bucket_of, w.hist);
    SE3ET_LAUNCH_CHECK();
  } else {
    cloud_min_kernel<<<1, 256, 0, st>>>(w.bounds, (int)batch, w.mins);
    SE3ET_LAUNCH_CHECK();
  }
  int rc = exclusive_scan(U32Load{w.hist}, nullptr, w.table + 1, w.block_sums, w.start, nullptr, st);
  if (rc) return rc;
  if (ns_total > 0) {
    hash_scatter_kernel<<<(int)ceil_div(ns_total, 256), 256, 0, st>>>(s_points, ns_total, w.bucket_of, w.start,
                                                                     w.cursor, w.sorted);
    SE3ET_LAUNCH_CHECK();
  }
  if (nq_total > 0) {
    QueryArgs qa;
    qa.q = q_points; qa.nq = nq_total; qa.q_off = w.q_off; qa.s_off = w.s_off; qa.batch = (int)batch; qa.mins = w.mins;
    qa.sorted = w.sorted; qa.start = w.start; qa.mask = mask; qa.inv_cell = inv_cell; qa.r2 = r2; qa.counts = counts;
    qa.out = out; qa.width = width; qa.ns_total = ns_total; qa.status = status; qa.cloud_max = cloud_max;
    // by cell where it wins (measured on B200, 32 stacked pairs): support sets of 100k points and more (917k points:
    // 1.19 -> 0.57 ms self, 0.36 -> 0.29 ms from the next level; 272k points: 0.35 -> 0.21 ms); smaller support sets
    // have too few cells to fill the GPU one warp per cell (74k points: 0.12 -> 0.13 ms, 21k: 0.04 -> 0.10 ms).
    // se3et_radius_set_mode overrides (tests).
    const int mode = g_radius_mode.load(std::memory_order_relaxed);
    const bool self = q_points == s_points && q_lengths == s_lengths && nq_total == ns_total;
    const bool by_cell = mode == 1 || (mode == 2 && ns_total >= 100000);
    if (!by_cell || ns_total == 0) {
      radius_query_kernel<<<(int)ceil_div(nq_total, kQueryWarps), kQueryWarps * 32, 0, st>>>(qa);
      SE3ET_LAUNCH_CHECK();
      return SE3ET_OK;
    }
    // queries bucketed with the support set's hash (self search: the support table itself)
    const float4* q_sorted = w.sorted;
    const uint32_t* q_start = w.start;
    if (!self) {
      SE3ET_CUDA_CHECK(cudaMemsetAsync(w.q_hist, 0, sizeof(uint32_t) * (w.table + 1), st));
      SE3ET_CUDA_CHECK(cudaMemsetAsync(w.q_cursor, 0, sizeof(uint32_t) * (w.table + 1), st));
      const int qblk = (int)ceil_div(nq_total, 256);
      // cell coordinates in the SUPPORT cloud's frame (w.mins), as the per-query kernel computes them
      hash_count_kernel<<<qblk, 256, 0, st>>>(q_points, nq_total, w.q_off, (int)batch, w.mins, inv_cell, mask,
                                              w.q_bucket_of, w.q_hist);
      SE3ET_LAUNCH_CHECK();
      rc = exclusive_scan(U32Load{w.q_hist}, nullptr, w.table + 1, w.block_sums, w.q_start, nullptr, st);
      if (rc) return rc;
      hash_scatter_kernel<<<qblk, 256, 0, st>>>(q_points, nq_total, w.q_bucket_of, w.q_start, w.q_cursor, w.q_sorted);
      SE3ET_LAUNCH_CHECK();
      q_sorted = w.q_sorted;
      q_start = w.q_start;
    }
    SE3ET_CUDA_CHECK(cudaMemsetAsync(w.fb_count, 0, sizeof(int32_t) * 4, st));
    const size_t smem = (size_t)kCellWarps * kCellWarpBytes;
    SE3ET_ENSURE_SMEM(radius_cell_kernel, smem);
    const int64_t chunks = ceil_div(w.table, 32);
    int64_t blocks = ceil_div(chunks, kCellWarps);
    if (blocks > (int64_t)kNumSMs * 4) blocks = (int64_t)kNumSMs * 4;
    radius_cell_kernel<<<(unsigned)blocks, kCellWarps * 32, smem, st>>>(qa, q_sorted, q_start, (uint32_t)w.table,
                                                                       w.fb_list, w.fb_count);
    SE3ET_LAUNCH_CHECK();
    radius_query_list_kernel<<<kNumSMs * 2, kQueryWarps * 32, 0, st>>>(qa, w.fb_list, w.fb_count);
    SE3ET_LAUNCH_CHECK();
  }
  return SE3ET_OK;
}
