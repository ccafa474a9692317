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
  SE3ET_STATUS_MAX_LENGTH = 4, /* grid_subsample: largest output cloud (the 2000-superpoint cap is checked without another read-back) */
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
 * time.  Hashed uniform grid (cell = radius) over the support set + one warp per query; large self searches run by
 * CELL instead (round 2): the queries are bucketed with the same hash and a warp stages the 27-cell candidate set of a
 * cell once for all its queries (same result, ~1.6x faster at 900k points).
 * se3et_radius_set_mode: 0 = always per query, 1 = always by cell, 2 = automatic (default); process-wide.
 * ------------------------------------------------------------------------------------------ */
int se3et_radius_neighbors_workspace_bytes(int64_t nq_total, int64_t ns_total, int64_t batch, size_t* bytes);
int se3et_radius_set_mode(int mode);
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

/* gemm_bf16 with the GroupNorm statistics of its fp32 output accumulated in the epilogue:
 * out_f32[M,N] = A[M,K] * B[N,K]^T (+ bias), stats[seg][g] += {sum, sum of squares} over the rows of pair `seg`
 * (row r belongs to point r / rows_per_point) and the channels of group g.  stats is zeroed by the call.
 * replaces: the Linear / KPConv contraction followed by the statistics pass of GroupNormEPN
 *   (blocks_epn.py:658-663, 697-701) without re-reading the activations.
 * out_f32 may be NULL (statistics only).
 * Returns SE3ET_ERR_UNSUPPORTED when N / groups is neither a power of two <= 16 nor a multiple of 16
 * (the host then uses se3et_groupnorm_stats). */
int se3et_gemm_bf16_gnstats(const void* a, int64_t lda, const void* b, int64_t ldb, int64_t m, int64_t n, int64_t k,
                            const float* bias, float* out_f32, int64_t ldc, double* stats,
                            const int64_t* seg_offsets, int64_t nseg, int64_t groups, int64_t rows_per_point,
                            se3et_stream_t stream);

/* Second pass of a Linear + GroupNorm (+ residual) (+ LeakyReLU) block: recomputes the GEMM and applies
 *   out_bf16 = LeakyReLU_slope( (A B^T + bias - mean) * rstd * gamma + beta [+ resid] )
 * in the epilogue, mean / rstd per (pair, group) from `stats` (as produced by se3et_gemm_bf16_gnstats with
 * out_f32 = NULL, which then only accumulates).  The pre-norm activations never reach global memory:
 * UnaryBlockEPN and the residual tail of ResnetBottleneckBlockEPN (blocks_epn.py:639-665, 833-852).
 * resid_bf16 (optional) has the output's shape and pitch.  slope = 1 disables the activation.
 * workspace: 16 * n * nseg bytes (16-byte aligned) for the per-(pair, column) scale / shift table of the streaming kernel
 * (persistent CTAs, eight epilogue warps; used whenever n % 64 == 0); NULL selects the one-tile-per-CTA kernel. */
int se3et_gemm_bf16_gnapply(const void* a, int64_t lda, const void* b, int64_t ldb, int64_t m, int64_t n, int64_t k,
                            const float* bias, const double* stats, const float* gamma, const float* beta, float eps,
                            float leaky_slope, const void* resid_bf16, void* out_bf16, int64_t ldc,
                            const int64_t* seg_offsets, int64_t nseg, int64_t groups, int64_t rows_per_point,
                            void* workspace, size_t workspace_bytes, se3et_stream_t stream);

/* Measurement switch: 0 routes se3et_gemm_bf16_gnapply(_dual) to the one-tile-per-CTA kernels instead of the streaming
 * kernel (persistent CTAs, eight epilogue warps; the default whenever n % 64 == 0). */
int se3et_gemm_set_stream_apply(int on);
/* Measurement switch: 0 routes narrow grouped GEMMs back to the strip (two CTAs per SM) variant. */
int se3et_gemm_set_grouped_small_cta(int on);
/* Measurement switch: widest output tile (64, 128 or 256 columns) of se3et_gemm_bf16 / se3et_gemm_bf16_gnstats. */
int se3et_gemm_set_plain_tile_cap(int bn);
/* Measurement switch: 0 keeps large bf16-output Linears of se3et_gemm_bf16 on the tile kernels (default: streaming kernel). */
int se3et_gemm_set_stream_plain(int on);

/* GroupNorm statistics of y = A W^T + bias as a STREAMING pass (UnaryBlockEPN / ResNet shortcut, blocks_epn.py:639-665,
 * 833-852): persistent tcgen05 GEMM whose epilogue keeps per-thread running sums of y and y^2 across all row tiles of a
 * pair and reduces them once per pair; y is never stored, A is read once.  Same `stats` layout as
 * se3et_gemm_bf16_gnstats ([nseg, groups, 2] {sum, sum sq}, zeroed by the call).  Requires n % 32 == 0, k % 8 == 0 and
 * n / groups a power of two <= 32 or a multiple of 32 (SE3ET_ERR_UNSUPPORTED otherwise). */
int se3et_linear_gnstats_stream(const void* a, int64_t lda, int64_t m, int64_t k, const void* w_bf16, int64_t ldw,
                                int64_t n, const float* bias, const int64_t* seg_offsets, int64_t nseg, int64_t groups,
                                int64_t rows_per_point, double* stats, se3et_stream_t stream);

/* GroupNorm statistics of y = A W^T + bias WITHOUT forming y (UnaryBlockEPN, blocks_epn.py:639-665, when the Linear
 * widens): one pass over A accumulates per pair the Gram matrix A^T A and the column sums (mma.sync, fp32 per CTA, fp64
 * across CTAs), then sum y_j = w_j.s + R b_j and sum y_j^2 = w_j^T G w_j + 2 b_j w_j.s + R b_j^2 per channel in fp64.
 * Same `stats` layout as se3et_gemm_bf16_gnstats ([nseg, groups, 2] {sum, sum sq}).  k must be 32, 64 or 128
 * (SE3ET_ERR_UNSUPPORTED otherwise).  upper_tiles: 0, or an upper bound of sum_pairs ceil(rows / 128).
 * workspace: se3et_linear_gnstats_gram_workspace_bytes(k, nseg). */
int se3et_linear_gnstats_gram_workspace_bytes(int64_t k, int64_t nseg, size_t* bytes);
int se3et_linear_gnstats_gram(const void* a, int64_t lda, int64_t m, int64_t k, const void* w_bf16, int64_t ldw,
                              int64_t n, const float* bias, const int64_t* seg_offsets, int64_t nseg, int64_t groups,
                              int64_t rows_per_point, int64_t upper_tiles, void* workspace, size_t workspace_bytes,
                              double* stats, se3et_stream_t stream);
/* The same for TWO Linears on the same input in one pass over A (the block input of ResnetBottleneckBlockEPN feeds unary1
 * and the shortcut Linear, blocks_epn.py:833-852): the Gram matrix is built once, finalised twice. */
int se3et_linear_gnstats_gram2(const void* a, int64_t lda, int64_t m, int64_t k, const void* w1_bf16, int64_t ldw1,
                               int64_t n1, const float* bias1, int64_t groups1, double* stats1, const void* w2_bf16,
                               int64_t ldw2, int64_t n2, const float* bias2, int64_t groups2, double* stats2,
                               const int64_t* seg_offsets, int64_t nseg, int64_t rows_per_point, void* workspace,
                               size_t workspace_bytes, se3et_stream_t stream);

/* Tail of ResnetBottleneckBlockEPN (blocks_epn.py:833-852) in one kernel:
 *   out_bf16 = LeakyReLU_slope( GroupNorm_1(A1 B1^T + bias1) + GroupNorm_2(A2 B2^T + bias2) )
 * unary2 on the conv branch (A1: [m, k1]) and the shortcut unary (A2: [m, k2]); both Linears are recomputed into two TMEM
 * accumulators, their statistics come from se3et_gemm_bf16_gnstats passes (out_f32 = NULL).  B1: [n, k1], B2: [n, k2]
 * (nn.Linear weights).  tile_n = 0 picks the output tile width (32 / 64 / 128 forces it). */
int se3et_gemm_bf16_gnapply_dual(const void* a1, int64_t lda1, const void* b1, int64_t ldb1, int64_t k1,
                                 const float* bias1, const double* stats1, const float* gamma1, const float* beta1,
                                 const void* a2, int64_t lda2, const void* b2, int64_t ldb2, int64_t k2,
                                 const float* bias2, const double* stats2, const float* gamma2, const float* beta2,
                                 int64_t m, int64_t n, float eps, float leaky_slope, void* out_bf16, int64_t ldc,
                                 const int64_t* seg_offsets, int64_t nseg, int64_t groups, int64_t rows_per_point,
                                 int tile_n, void* workspace, size_t workspace_bytes, se3et_stream_t stream);

/* Grouped variant: problem g (blockIdx.z) reads A rows [a_row0, a_row0 + m_rows) and B rows
 * [b_row0, b_row0 + n) of the flat operands, groups = int64 [num_groups][6] {a_row0, b_row0, m_rows, c_off,
 * ldc, unused} on the device; max_m = largest m_rows.  transposed = 1 stores
 * out_f32[c_off + col * ldc + row] for col < n_valid (n itself is rounded up to the UMMA tile by the caller).
 * replaces: the `q . proj_p(embedding)` term of RPEMultiHeadAttention (rpe_transformer.py:60,71), computed
 * as embedding[n] @ (W_p^T q[n]) per query n without materialising proj_p(embedding). */
int se3et_gemm_grouped_bf16(const void* a, int64_t lda, int64_t a_rows_total, const void* b, int64_t ldb,
                            int64_t b_rows_total, const int64_t* groups, int64_t num_groups, int64_t max_m, int64_t n,
                            int64_t n_valid, int64_t k, float alpha, float* out_f32, int64_t ldc, int transposed,
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

/* kpconv_fused -- the whole KPConvInterSO3.forward (blocks_epn.py:454-546, 334-390) in one kernel:
 *   out[(p, r)][d] = sum_{k, a, c} ( sum_n w[p][n][k] x[idx[p][n]][a][c] ) * W[kidx[k][r]][ridx[a][r]][c][d]
 * The gathered operand is built per 16-point tile in shared memory (mma.sync on the 16-row basis of the
 * octahedral index tables) and contracted by tcgen05 tensor cores with fp32 accumulation in TMEM; it never
 * touches global memory.
 *   x_bf16  [ns, 6, cin] bf16;  out [nq * 6, cout] fp32 (pre-norm, as KPConvInterSO3 returns it), or bf16 when
 *           out_bf16 != 0 (the model's inference path: the GroupNorm passes that follow read half the bytes)
 *   w_bf16  [cout, 36 * cin] bf16, K-major, K index = (chunk * 36 + kc * 6 + a') * 16 + c for input channel
 *           chunk * 16 + c of weights[kc][a'][.][d]  (se3et_b200/modules/e2pn.py:KPConvInterSO3._w_fused)
 *   stats   optional double [nseg, groups, 2]: GroupNorm sums of the output per pair (zeroed by the call)
 * Requires cin % 16 == 0, cout % 16 == 0, h <= 48 (h <= 40 when cout % 128 == 0); otherwise
 * SE3ET_ERR_UNSUPPORTED and the host uses se3et_kpconv_gather + se3et_gemm_bf16(_gnstats). */
int se3et_kpconv_fused(const float* q_pts, const float* s_pts, const int64_t* neighbors, int64_t nq, int64_t ns,
                       int64_t h, const void* x_bf16, int64_t cin, const void* w_bf16, int64_t cout,
                       const float* kernel_points_15x3, float kp_extent, void* out, int out_bf16, double* stats,
                       const int64_t* seg_offsets, int64_t nseg, int64_t groups, se3et_stream_t stream);

/* kpconv_rows -- the whole KPConvInterSO3.forward (blocks_epn.py:454-546, 334-390) in one kernel, UMMA rows = query
 * points: per (16-channel chunk, input anchor a) the basis products B[p][beta][a][c] (kpconv_tables.cuh) of 128 / 96
 * points are stored once in shared memory and the 36 (output anchor r, class kc) K = 16 tcgen05.mma pick their basis
 * slab and their weight slice W[kc][ridx[a][r]] by descriptor; accumulators: 6 x 32 / 64 fp32 columns of TMEM per point,
 * the basis weights of the tile's points are parked in the remaining TMEM columns.
 *   x_bf16  [ns, 6, cin] bf16;  out [nq * 6, cout] fp32 (pre-norm, as KPConvInterSO3 returns it) or bf16 (out_bf16 != 0)
 *   w_rows_bf16 [cout, 216 * cin] bf16, K-major, K index = (((chunk * 6 + a) * 36 + r * 6 + kc) * 16 + c'),
 *           value weights[kc][ridx[a][r]][chunk * 16 + (c' ^ 8 * flip[r * 6 + kc])][d]; the slot / flip tables come
 *           from se3et_kpconv_rows_layout (se3et_b200/modules/e2pn.py:KPConvInterSO3._w_rows)
 * Requires cin % 16 == 0, cout % 32 == 0, h <= 48, ns * 6 * cin / 8 < 2^32; otherwise SE3ET_ERR_UNSUPPORTED and the
 * host uses se3et_kpconv_fused.  Shadow neighbours (index >= ns) contribute zero (their weight is zero; their gathered
 * row is row 0, which must be finite). */
int se3et_kpconv_rows(const float* q_pts, const float* s_pts, const int64_t* neighbors, int64_t nq, int64_t ns,
                      int64_t h, const void* x_bf16, int64_t cin, const void* w_rows_bf16, int64_t cout,
                      const float* kernel_points_15x3, float kp_extent, void* out, int out_bf16, se3et_stream_t stream);

/* Weight layout tables of se3et_kpconv_rows: src_slot[a][r * 6 + kc] = kc * 6 + ridx[a][r] (the weights[kc][a'] slice
 * step a needs for (r, kc)); flip[r * 6 + kc] = 1 when that slice's two 8-channel halves are stored swapped. */
int se3et_kpconv_rows_layout(int32_t* src_slot_6x36, int32_t* flip_36);

/* kpconv_cin1 -- KPConvInterSO3.forward for the first backbone layer (lifted input, one channel per anchor): K is
 * only 36, so the whole convolution runs on CUDA cores, one warp per query point.  x_bf16 [ns, 6]; w_36xcout fp32
 * [(kc, a'), cout] = weights[kc][a'][0][:]; out_f32 [nq * 6, cout]; optional GroupNorm statistics as for
 * se3et_kpconv_fused.  cout 32 or 64, h <= 40; otherwise SE3ET_ERR_UNSUPPORTED (gather + GEMM path).
 * lifted = 1: the input is the LiftBlockEPN output (blocks_epn.py:993-1004), identical for the six anchors, and x_bf16
 * is [ns] (one value per support point): 16 basis products and anchor-summed weights per point. */
int se3et_kpconv_cin1(const float* q_pts, const float* s_pts, const int64_t* neighbors, int64_t nq, int64_t ns,
                      int64_t h, const void* x_bf16, const float* w_36xcout, int64_t cout,
                      const float* kernel_points_15x3, float kp_extent, float* out_f32, double* stats,
                      const int64_t* seg_offsets, int64_t nseg, int64_t groups, int lifted, se3et_stream_t stream);

/* kpconv_lift -- the lifted first layer (se3et_kpconv_cin1 with lifted = 1) rebuilt around a thread per query point:
 * KPConvInterSO3.forward (blocks_epn.py:454-546, 334-390) on the LiftBlockEPN output (blocks_epn.py:993-1004), whose
 * value f[ns] (bf16) is the same for the six anchors.  16 basis products per point in registers, then
 * out[p][(r, d)] = D[p][0..15] x Bm[16][6 * cout] on mma.sync (bf16 hi + lo operands, fp32-grade result).
 *   out       [nq * 6, cout] fp32, or bf16 when out_bf16 != 0 (the statistics are taken before the rounding)
 *   stats     optional double [nseg, groups, 2] as for se3et_kpconv_fused
 *   workspace se3et_kpconv_lift_workspace_bytes(ns): the support points packed as {x, y, z, f}
 * cout 32 or 64, h <= 48, ns < 2^31; otherwise SE3ET_ERR_UNSUPPORTED (the host uses se3et_kpconv_cin1). */
int64_t se3et_kpconv_lift_workspace_bytes(int64_t ns);
int se3et_kpconv_lift(const float* q_pts, const float* s_pts, const int64_t* neighbors, int64_t nq, int64_t ns,
                      int64_t h, const void* f_bf16, const float* w_36xcout, int64_t cout,
                      const float* kernel_points_15x3, float kp_extent, void* out, int out_bf16, double* stats,
                      const int64_t* seg_offsets, int64_t nseg, int64_t groups, void* workspace,
                      int64_t workspace_bytes, se3et_stream_t stream);

/* Diagnostics: {registers, static smem bytes, max threads per block, local bytes, max dynamic smem} of the fused
 * kernel instantiation for output tile width bn (16, 32, 64 or 128). */
int se3et_kpconv_fused_attrs(int bn, int* out5);

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
/* Two GroupNorm + LeakyReLU stages back to back without the intermediate tensor (KPConvInterSO3Block.norm followed
 * by the enclosing block's norm, blocks_epn.py:737-741 with 790-794 / 841-843):
 *   f = LeakyReLU(GN_1(y)), out = LeakyReLU(GN_2(f)).   apply = 0: accumulate the statistics of f into stats2
 * (zeroed by the call); apply = 1: recompute f and write out_bf16 using stats2; apply = 2: accumulate the statistics of
 * y itself into stats2 (stats1 / gamma / beta unused) -- the first norm's statistics as a streaming pass.
 * y is fp32 [rows, channels], or bf16 when y_bf16 != 0 (what the conv kernels write with out_bf16).
 * channels / 4 a power of two <= 256. */
int se3et_groupnorm_double(const void* y, int y_bf16, const double* stats1, const float* gamma1, const float* beta1,
                           double* stats2, const float* gamma2, const float* beta2, int64_t rows, int64_t channels,
                           int64_t groups, const int64_t* seg_offsets, int64_t nseg, int64_t rows_per_point, float eps,
                           float leaky_slope, int apply, void* out_bf16, se3et_stream_t stream);

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

/* ------------------------------------------------------------------------------------------
 * Superpoint transformer.  Clouds are flat: cloud b owns points [cloud_offsets[b], cloud_offsets[b+1]).
 * Equivariant states are stored [point, anchor(6), channel] (the reference's (B, A, N, C) transposed back;
 * every per-row op is layout independent).
 * ------------------------------------------------------------------------------------------ */

/* GeometricStructureEmbedding.get_embedding_indices (geotransformer.py:69-99): for every ordered pair (n, m) of a
 * cloud, out_idx4[emb_offsets[b] + n*n_b + m] = {dist/sigma_d, angle_0, angle_1, angle_2} * (180/(sigma_a*pi)),
 * angles against the 3 nearest neighbours of n (ties: lower index).  angle_k must be 3. */
int se3et_geo_embed_indices(const float* points, const int64_t* cloud_offsets, int64_t nclouds, int64_t total_points,
                            int64_t max_cloud, const int64_t* emb_offsets, float sigma_d, float sigma_a,
                            int64_t angle_k, float* out_idx4, se3et_stream_t stream);
/* GeometricStructureEmbedding.forward (geotransformer.py:101-115) + SinusoidalPositionalEmbedding
 * (positional_embedding.py:18-34): out[r] = W_d emb(idx4[r].x) + max_k W_a emb(idx4[r].yzw) + bias_sum, bf16
 * [rows, channels].  The sinusoids are generated inside the kernel as the tcgen05 A operand. */
int se3et_geo_embed_project(const float* idx4, int64_t rows, int64_t channels, const void* w_d_bf16,
                            const void* w_a_bf16, const float* bias_sum, void* out_bf16, se3et_stream_t stream);
/* The same embedding assembled from tabulated projections: W emb(u) + b is a function of the scalar u alone, so
 * table_d[i] = W_d emb(i * step) + b_d and table_a[i] = W_a emb(i * step) + b_a (bf16 [n, channels], built by the host
 * once per weight version, se3et_b200/modules/transformer.py) give
 *   out[r] = table_d[round(idx4[r].x / step)] + max_k table_a[round(idx4[r].{y,z,w} / step)]     (indices clamped).
 * With step = 1/512 the tabulation error is below the bf16 rounding of the result.  channels in {64, 128, 256}. */
int se3et_geo_embed_lookup(const float* idx4, int64_t rows, int64_t channels, const void* table_d_bf16, int64_t nd,
                           const void* table_a_bf16, int64_t na, float step, void* out_bf16, se3et_stream_t stream);
/* Fused attention: RPEMultiHeadAttention equivariant branch (rpe_transformer.py:56-131, bias = q.proj_p(emb)) and
 * MultiHeadAttention with 4-D value (vanilla_transformer.py:58-85).  problems: int64 [num_problems][5]
 * {q_start, n_q, kv_start, n_kv, bias_off}.  Strides in elements: *_pt per point, *_an per anchor (0 = none).
 * bias (optional, fp32): [bias_off + ((i*A + a)*H + h)*n_kv + m].  out row = (q_start + i)*A + a, pitch ldo. */
int se3et_flash_attention(const void* q, int64_t q_pt, int64_t q_an, const void* k, int64_t k_pt, int64_t k_an,
                          const void* v, int64_t v_pt, int64_t v_an, const float* bias, const int64_t* problems,
                          int64_t num_problems, int64_t max_q, int64_t anchors, int64_t heads, int64_t head_dim,
                          float scale, void* out_bf16, int64_t ldo, se3et_stream_t stream);
/* ---- SE3ET-E: MultiHeadAttentionEQ in the modes a_soft / r_soft (vanilla_transformer.py:87-870) ----------------
 * The local scores S[a,e,h,n,m] = q_a . k_e / sqrt(c) are never materialised: the global anchor statistics come
 * from anchor_pair_stats, the per-(a, e) attentions from se3et_flash_attention (one launch per key anchor e), their
 * weighted sum over e from anchor_mix.  problems = the attention problem table {q_start, n_q, kv_start, n_kv, -}. */

/* g[p][a][e] = sum_{n,m} f( q_a[n] . k_e[m] * scale ), scale = 1 / (heads * sqrt(head_dim)) (head-mean of the local
 * scores, vanilla_transformer.py:380-431); positive: 0 'sq', 1 'softplus', 2 'sigmoid', 3 'relu', 4 'abs'.
 * g is zeroed by the call; channels in {64, 128, 256}. */
int se3et_anchor_pair_stats(const void* q_bf16, int64_t q_pt, int64_t q_an, const void* k_bf16, int64_t k_pt,
                            int64_t k_an, const int64_t* problems, int64_t num_problems, int64_t max_q,
                            int64_t anchors, int64_t channels, float scale, int positive, float* g,
                            se3et_stream_t stream);
/* w[p][a][e]: r_soft = 0: g / (n m) normalised over e (a_soft, :466-476); r_soft = 1: rotation weights
 * attn_r[r] = mean_a g[a][perms[r][a]] normalised over r (:560-575; optional output [p][num_rotations]) folded to
 * w[a][e] = sum_{r : perms[r][a] == e} attn_r[r].  perms = trace_idx_ori, int32 [num_rotations][anchors]. */
int se3et_anchor_mix_weights(const float* g, const int64_t* problems, int64_t num_problems, const int32_t* perms,
                             int64_t num_rotations, int64_t anchors, int r_soft, float* w, float* attn_r,
                             se3et_stream_t stream);
/* out[(n, a)][c] = sum_e w[cloud(n)][a][e] * in[e * stride_e + n * stride_n + a * stride_a + c]  (bf16; strides in
 * elements; cloud_offsets: nclouds + 1 point offsets of the query side).  Serves the sum over key anchors of the
 * per-(a, e) attention outputs (:812-818) and eq2inv_soft (conditional_transformer.py:209-249). */
int se3et_anchor_mix(const void* in_bf16, int64_t stride_e, int64_t stride_n, int64_t stride_a, const float* w,
                     const int64_t* cloud_offsets, int64_t nclouds, int64_t anchors, int64_t channels,
                     int64_t n_points, void* out_bf16, se3et_stream_t stream);
/* Equivariant (spherical-harmonics, l = 1) score term of the equivariant self attention (rpe_transformer.py:76-79,
 * geotransformer.py:57-67): bias[bias_off + ((i A + a) H + h) n + m] += c1 * (anchors[a]^T unit(p_i - p_m)) .
 * u[((q_start + i) A + a) ldu + 3 h ..], u = q W_eq[:, 1:4] per head.  The l = 0 part is constant along m. */
int se3et_sh_bias_add(const float* points, const int64_t* problems, int64_t num_problems, int64_t max_n,
                      const float* u, int64_t ldu, const float* anchors_Ax3x3, int64_t anchors, int64_t heads, float c1,
                      float* bias, se3et_stream_t stream);

/* LayerNorm(x + resid[row / resid_div]) (rpe_transformer.py:161-163, vanilla_transformer.py:908-911 with the
 * (N, C) -> (A, N, C) residual broadcast, output_layer.py:16-22). x fp32 [rows, channels], resid bf16 or NULL. */
int se3et_add_layernorm(const float* x, const void* resid_bf16, int64_t resid_div, int64_t rows, int64_t channels,
                        const float* gamma, const float* beta, float eps, float* out_f32, void* out_bf16,
                        se3et_stream_t stream);

/* linear_add_layernorm -- `linear` / `squeeze` + residual + LayerNorm of the transformer layers in one kernel
 * (rpe_transformer.py:163-175, vanilla_transformer.py:905-913, output_layer.py:17-22):
 *   out = LayerNorm(resid[row / resid_div] + a W^T + bias),  a bf16 [m, k], W bf16 [n, k], out bf16 [m, n].
 * The fp32 Linear output never reaches memory.  n must be 256 (the hidden width of every SE3ET variant) and k a
 * multiple of 8; otherwise SE3ET_ERR_UNSUPPORTED and the host uses se3et_gemm_bf16 + se3et_add_layernorm. */
int se3et_linear_add_layernorm(const void* a_bf16, int64_t lda, const void* w_bf16, int64_t ldw, int64_t m, int64_t n,
                               int64_t k, const float* bias, const void* resid_bf16, int64_t resid_div,
                               const float* gamma, const float* beta, float eps, void* out_bf16, se3et_stream_t stream);
/* F.normalize(x, p=2, dim=1) (experiments/se3eti.3dmatch/model.py:156-157). */
int se3et_l2_normalize_rows(const float* x, int64_t rows, int64_t channels, float eps, float* out,
                            se3et_stream_t stream);

/* SuperPointMatching.forward (superpoint_matching.py:13-55) for num_pairs pairs at once.
 * problems: int64 [num_pairs][5] {ref_start, n_ref, src_start, n_src, e_off}; masks uint8 (1 = keep) or NULL.
 * e_workspace: fp32, 16-byte aligned, se3et_superpoint_matching_workspace_floats(...) floats: the score matrices
 * (e_total = sum of n_ref*n_src floats) followed by the candidate lists of the selection (eight CTAs per pair each
 * select their slice's top-k, one CTA per pair merges them); row_sums / col_sums: fp32 per ref / src superpoint.
 * Outputs per pair: num_correspondences (ref index, src index, score) sorted by (score desc, flat index asc),
 * indices local to the pair, -1 padded; counts[p] = min(num_correspondences, #unmasked entries). */
int se3et_superpoint_matching_workspace_floats(int64_t num_pairs, int64_t num_correspondences, int64_t e_total,
                                               int64_t* floats);
int se3et_superpoint_matching(const float* ref_feats, const float* src_feats, int64_t channels,
                              const uint8_t* ref_masks, const uint8_t* src_masks, const int64_t* problems,
                              int64_t num_pairs, int64_t max_ref, int64_t max_src, int64_t num_correspondences,
                              int dual_normalization, float* e_workspace, int64_t e_total, float* row_sums,
                              float* col_sums, int64_t* ref_idx, int64_t* src_idx, float* scores, int32_t* counts,
                              se3et_stream_t stream);

/* point_to_node_partition -- geotransformer/modules/ops/pointcloud_partition.py:60-107 (called twice per pair from
 * experiments/se3eti.3dmatch/model.py:109-114), batched over `batch` stacked clouds: cloud b owns point_lengths[b]
 * consecutive rows of points [n_points, 3] and node_lengths[b] rows of nodes [n_nodes, 3] (both lengths on the device).
 *   point_to_node    [n_points]           node (cloud-local index) nearest to each point; squared distance
 *                                         (x2 - 2 x.y) + y2 clamped at 0 as pairwise_distance.py:27-30, fp32 in a fixed
 *                                         operation order without FMA; ties -> lowest node index
 *   node_masks       [n_nodes]            1 when the node received at least one point
 *   node_sizes       [n_nodes] or NULL    number of points per node (return_count = True)
 *   node_knn_indices [n_nodes, limit]     the node's nearest assigned points (cloud-local), ascending (distance, index),
 *                                         padded with the cloud's point count
 *   node_knn_masks   [n_nodes, limit]     1 for real entries
 * workspace: se3et_point_to_node_partition_workspace_bytes(n_points, batch), 256-byte aligned. */
int se3et_point_to_node_partition_workspace_bytes(int64_t n_points, int64_t batch, size_t* bytes);
int se3et_point_to_node_partition(const float* points, const int64_t* point_lengths, int64_t n_points,
                                  const float* nodes, const int64_t* node_lengths, int64_t n_nodes, int64_t batch,
                                  int64_t point_limit, int64_t* point_to_node, uint8_t* node_masks, int64_t* node_sizes,
                                  int64_t* node_knn_indices, uint8_t* node_knn_masks, void* workspace,
                                  size_t workspace_bytes, se3et_stream_t stream);

/* log_optimal_transport -- LearnableLogOptimalTransport.forward (geotransformer/modules/sinkhorn/learnable_sinkhorn.py:
 * 21-66, called from experiments/se3eti.3dmatch/model.py:202-205): log-domain Sinkhorn with a dustbin row / column.
 *   scores [batch, num_row, num_col] fp32; row_masks / col_masks [batch, num_row] / [batch, num_col] (1 = valid) or NULL;
 *   alpha: device pointer to the learnable dustbin score; out [batch, num_row + 1, num_col + 1] fp32.
 * One CTA per matrix, all iterations in shared memory (the reference: 2 * num_iterations logsumexp launches). */
int se3et_log_optimal_transport(const float* scores, const uint8_t* row_masks, const uint8_t* col_masks,
                                const float* alpha, int64_t batch, int64_t num_row, int64_t num_col,
                                int64_t num_iterations, float* out, se3et_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fine-stage registration: LocalGlobalRegistration.forward
 * (geotransformer/modules/geotransformer/local_global_registration.py:49-235; call site
 * experiments/se3eti.3dmatch/model.py:208-224) with weighted_procrustes
 * (geotransformer/modules/registration/procrustes.py:6-73), for every patch correspondence of every pair
 * of a launch sequence.  Two stages share one workspace (se3et_lgr_workspace_bytes); between them the host
 * turns `counts` into `corr_offsets` (exclusive scan, a device-side cumsum: no sync).
 * ------------------------------------------------------------------------------------------ */
int se3et_lgr_workspace_bytes(int64_t num_patches, int64_t k_points, int64_t topk, size_t* bytes);

/* Stage 1 (compute_correspondence_matrix, :49-84, mutual = True, no dustbin): log_scores fp32 [B, ld, ld], the
 * top-left k_points x k_points block is the patch's log-likelihood matrix; exp, top-k along rows and columns with
 * (score desc, index asc) order, confidence threshold, masks uint8 [B, k_points].  counts int32 [B]: correspondences
 * per patch (kept, in torch.nonzero's row-major order, in workspace slots).  k_points <= 128, topk <= 4. */
int se3et_lgr_correspondences(const float* log_scores, int64_t ld, const uint8_t* ref_masks, const uint8_t* src_masks,
                              int64_t num_patches, int64_t k_points, int64_t topk, float confidence_threshold,
                              void* workspace, size_t workspace_bytes, int32_t* counts, se3et_stream_t stream);

/* Stage 2 (local_to_global_registration, :137-205): per patch with >= correspondence_threshold correspondences a
 * weighted Procrustes transform and its support over all correspondences of the pair; per pair the first best
 * transform starts num_refinement_steps rounds of inlier-weighted Procrustes (no patch qualifies: all correspondences
 * start it).  knn points fp32 [B, k_points, 3]; patch_offsets int64 [num_pairs + 1]; corr_offsets int64 [B + 1].
 * Outputs: compacted correspondences fp32 [sum counts, 3] x 2 and scores [sum counts] (capacity B * topk * k_points
 * suffices), transforms fp32 [num_pairs, 4, 4]. */
int se3et_lgr_register(const float* ref_knn_points, const float* src_knn_points, const int64_t* patch_offsets,
                       int64_t num_pairs, int64_t num_patches, int64_t k_points, int64_t topk, float acceptance_radius,
                       int64_t correspondence_threshold, int64_t num_refinement_steps, void* workspace,
                       size_t workspace_bytes, const int32_t* counts, const int64_t* corr_offsets,
                       float* out_ref_points, float* out_src_points, float* out_scores, float* out_transforms,
                       se3et_stream_t stream);

/* weighted_procrustes (procrustes.py:6-73) for `batch` point sets of n points: src / ref fp32 [batch, n, 3], weights
 * fp32 [batch, n] or NULL; out fp32 [batch, 4, 4] mapping src onto ref.  The rotation is the proper rotation the
 * reference builds from its SVD (V diag(1, 1, det) U^T), found from Horn's quaternion matrix in fp64. */
int se3et_weighted_procrustes(const float* src_points, const float* ref_points, const float* weights, int64_t batch,
                              int64_t n, float weight_thresh, float eps, float* out_transforms, se3et_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SE3ET_B200_H_ */
