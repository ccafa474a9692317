/* se3et_b200 -- C ABI of the B200-native (sm_100a) SE3ET hot path.
 *
 * Every entry point is `extern "C"`, takes plain DEVICE pointers and sizes plus the
 * CUDA stream to launch on, returns 0 on success or a negative SE3ET_ERR_* code, and
 * never throws.  No torch types cross this boundary; the Python host side
 * (se3et_b200/_lib.py) binds it with ctypes and passes tensor.data_ptr().
 *
 * The library is stateless and re-entrant: all scratch memory is passed in as a
 * workspace sized by the matching *_workspace_bytes() call.  Conditions detected on
 * the device (e.g. voxel grid larger than the workspace) are reported through a
 * caller-provided `int32 status[SE3ET_STATUS_WORDS]` device buffer that the host
 * reads together with the data-dependent output sizes (one sync).
 *
 * Each function cites the reference interface it replaces (paths relative to the
 * reference repository root).
 */
#ifndef SE3ET_B200_H_
#define SE3ET_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* se3et_stream_t; /* cudaStream_t */

enum {
  SE3ET_OK = 0,
  SE3ET_ERR_CUDA = -1,      /* a CUDA runtime call / launch failed (see se3et_last_error) */
  SE3ET_ERR_ARG = -2,       /* invalid argument (null pointer, negative size, unsupported shape) */
  SE3ET_ERR_WORKSPACE = -3, /* workspace too small */
  SE3ET_ERR_UNSUPPORTED = -4
};

/* Device-side status words (int32 each). */
enum {
  SE3ET_STATUS_ERROR = 0,     /* 0 = ok, else SE3ET_DEV_* bit mask */
  SE3ET_STATUS_M_TOTAL = 1,   /* grid_subsample: total number of output points */
  SE3ET_STATUS_MAX_COUNT = 2, /* radius_neighbors: max neighbour count over all queries */
  SE3ET_STATUS_REQ_KCELLS = 3, /* grid_subsample: on GRID_TOO_LARGE, cells needed / 1024 (rounded up, saturating) */
  SE3ET_STATUS_WORDS = 8
};
enum {
  SE3ET_DEV_GRID_TOO_LARGE = 1, /* voxel grid does not fit the workspace bitmap */
  SE3ET_DEV_INDEX_RANGE = 2     /* voxel index below -1 (cannot happen for finite fp32 input) */
};

const char* se3et_last_error(void); /* thread-local text of the last SE3ET_ERR_CUDA */
int se3et_version(void);

/* ------------------------------------------------------------------------------------------
 * grid_subsample
 * replaces: geotransformer.ext.grid_subsampling
 *   (geotransformer/extensions/cpu/grid_subsampling/grid_subsampling.cpp:5-83,
 *    grid_subsampling_cpu.cpp:3-109, grid_subsampling_cpu.h:24-74; pybind.cpp:14-18)
 * For every cloud b (stack mode, `lengths[b]` points each) and every occupied voxel,
 * emits the input point closest to the voxel's fp32 barycentre (ties: lowest input
 * index) and its normal.  Output order is canonical: ascending reference voxel key
 * inside each cloud.  IEEE fp32, no contraction: bit-identical to the reference.
 * s_points / s_normals must have room for n_total rows.  status[M_TOTAL] receives M.
 * ------------------------------------------------------------------------------------------ */
int se3et_grid_subsample_workspace_bytes(int64_t n_total, int64_t batch, int64_t max_cells, size_t* bytes);
int se3et_grid_subsample(const float* points, const int64_t* lengths, const float* normals, int64_t n_total,
                         int64_t batch, float voxel_size, float* s_points, int64_t* s_lengths, float* s_normals,
                         int32_t* status, void* workspace, size_t workspace_bytes, int64_t max_cells,
                         se3et_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * radius_neighbors
 * replaces: geotransformer.ext.radius_neighbors
 *   (geotransformer/extensions/cpu/radius_neighbors/radius_neighbors.cpp:5-76,
 *    radius_neighbors_cpu.cpp:3-91; nanoflann.hpp:249-253,432-440,1279-1289; pybind.cpp:9-13)
 *   and the column truncation of modules/ops/radius_search.py:24-27.
 * For query i of cloud b: all support points j of cloud b with
 *   d2 = ((qx-sx)^2 + (qy-sy)^2) + (qz-sz)^2 < radius*radius     (fp32, strict)
 * ordered by (d2 ascending, j ascending).
 *   counts  (optional, int32[nq])      neighbour count of each query
 *   out     (optional, int64[nq,width]) first `width` neighbours as global support
 *           indices, padded with ns_total
 * status[MAX_COUNT] receives the maximum count; cloud_max (optional, int32[batch]) the maximum
 * per cloud, which is what decides the reference's matrix width when pairs are processed one at a
 * time.  Hashed uniform grid (cell = radius) over the support set + one warp per query.
 * ------------------------------------------------------------------------------------------ */
int se3et_radius_neighbors_workspace_bytes(int64_t nq_total, int64_t ns_total, int64_t batch, size_t* bytes);
int se3et_radius_neighbors(const float* q_points, const float* s_points, const int64_t* q_lengths,
                           const int64_t* s_lengths, int64_t nq_total, int64_t ns_total, int64_t batch,
                           float radius, int32_t* counts, int64_t* out, int64_t width, int32_t* status,
                           int32_t* cloud_max, void* workspace, size_t workspace_bytes, se3et_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * gemm_bf16:  C[M,N] = alpha * A[M,K] * B[N,K]^T (+ bias[N]) (ReLU if act == 1)
 * replaces: every torch.nn.Linear on the path (blocks_epn.py:658 UnaryBlockEPN.mlp,
 *   kpconv/modules.py:75,97, geotransformer.py:278-281,314-315 in/out_proj,
 *   rpe_transformer.py:57-60 / vanilla_transformer.py:58-62 proj_q/k/v/p, output_layer.py:16-20)
 *   and the `kpac,karcd->prd` contraction of KPConvInterSO3.forward (blocks_epn.py:503-506)
 *   after se3et_kpconv_gather has built A.
 * A, B: bf16 row-major with pitches lda / ldb (elements, multiples of 8); N multiple of 16.
 * tcgen05 tensor cores, fp32 accumulation in TMEM.  Outputs fp32 and/or bf16 (either may be NULL).
 * batch > 1: A rows offset by a_batch_rows, B rows by b_batch_rows (0 = shared), C by c_batch_stride.
 * ------------------------------------------------------------------------------------------ */
int se3et_gemm_bf16(const void* a, int64_t lda, const void* b, int64_t ldb, int64_t m, int64_t n, int64_t k,
                    int64_t batch, int64_t a_batch_rows, int64_t b_batch_rows, const float* bias, float alpha,
                    int act, float* out_f32, void* out_bf16, int64_t ldc, int64_t c_batch_stride,
                    se3et_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * E2PN backbone pieces.  Feature tensors are [point, anchor(6), channel]; activations bf16,
 * pre-norm GEMM outputs fp32.  Segments (`seg_offsets`, nseg+1 point offsets) are point-cloud
 * PAIRS: GroupNorm statistics never mix pairs (blocks_epn.py:697-701 normalises ref+src jointly).
 * ------------------------------------------------------------------------------------------ */

/* The octahedral-group index tables the gather kernel has compiled in (kanchor 6, quotient 4,
 * K 15): kidx_rot[k][r] and ridx_rot[a][r] of KPConvInterSO3 (blocks_epn.py:228-332).  The host
 * compares them with the module's buffers and refuses to run on a mismatch. */
int se3et_kpconv_tables(int32_t* kidx_15x6, int32_t* ridx_6x6);

/* kpconv_gather -- first half of KPConvInterSO3.forward (blocks_epn.py:454-506, 334-390):
 *   w[n][k] = max(0, 1 - |s[idx[p][n]] - q[p] - kp[k]| / extent)      (shadow index ns => w = 0)
 *   wf[k][a][c] = sum_n w[n][k] * x[idx[p][n]][a][c]
 *   out[(p*6 + r)][(kc*6 + ridx[a][r])*cin + c] = sum_{k: kidx[k][r] == kc} wf[k][a][c]
 * so that conv = out @ weights.view(36*cin, cout) (one se3et_gemm_bf16).  out: bf16 [nq*6, kpad],
 * kpad >= 36*cin (multiple of 8), padding columns zeroed. */
int se3et_kpconv_gather(const float* q_pts, const float* s_pts, const int64_t* neighbors, int64_t nq, int64_t ns,
                        int64_t h, const void* x_bf16, int64_t cin, const float* kernel_points_15x3, float kp_extent,
                        void* out_bf16, int64_t kpad, se3et_stream_t stream);

/* GroupNormEPN / kpconv GroupNorm (blocks_epn.py:684-701, kpconv/modules.py:33-50): statistics per
 * (segment, group) over rows_per_point * points * channels_per_group values.  stats: double [nseg, groups, 2]. */
int se3et_groupnorm_stats(const float* y, int64_t rows, int64_t channels, int64_t groups, const int64_t* seg_offsets,
                          int64_t nseg, int64_t rows_per_point, double* stats, se3et_stream_t stream);
/* out = LeakyReLU_slope( GN_a(ya) [+ GN_b(yb)] [+ resid] ); slope 1 = no activation.  Covers UnaryBlockEPN,
 * KPConvInterSO3Block, and the residual tail of ResnetBottleneckBlockEPN.forward (blocks_epn.py:833-852). */
int se3et_groupnorm_apply(const float* ya, const double* stats_a, const float* gamma_a, const float* beta_a,
                          const float* yb, const double* stats_b, const float* gamma_b, const float* beta_b,
                          const void* resid_bf16, int64_t rows, int64_t channels, int64_t groups,
                          const int64_t* seg_offsets, int64_t nseg, int64_t rows_per_point, float eps,
                          float leaky_slope, float* out_f32, void* out_bf16, se3et_stream_t stream);
/* max_pool (e2pn/blocks.py:93-110): out[q] = max_n xpad[neighbors[q][n]], zero shadow row. width = 6*C.
 * seg_offsets/seg_width (optional): only the first seg_width[s] columns count for queries of pair s -- the
 * reference's matrix is min(max_count, limit) wide PER PAIR and its zero shadow row enters the max only
 * when a row is shorter than that. */
int se3et_maxpool_nbr(const void* x_bf16, int64_t ns, int64_t width, const int64_t* neighbors, int64_t nq, int64_t h,
                      const int64_t* seg_offsets, const int32_t* seg_width, int64_t nseg, void* out_bf16,
                      se3et_stream_t stream);
/* InvOutBlockEPN / eq->inv pooling (blocks_epn.py:924, conditional_transformer.py:282-283): max over anchors. */
int se3et_anchor_max(const void* x_bf16, int64_t n, int64_t anchors, int64_t channels, void* out_bf16, int64_t out_ld,
                     se3et_stream_t stream);
/* nearest_upsample + cat (kpconv/functional.py:6-22, backbone.py:66-72): out[i] = [xpad[up[i][0]] | y[i]]. */
int se3et_upsample_concat(const void* x_bf16, int64_t nx, int64_t c1, const int64_t* up_idx, int64_t up_ld,
                          const void* y_bf16, int64_t c2, int64_t n, void* out_bf16, se3et_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SE3ET_B200_H_ */
