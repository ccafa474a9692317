// Host interface of the streaming Linear + GroupNorm apply kernel (gemm_stream.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace se3et {

struct StreamNorm {
  const double* stats;  // [nseg, groups, 2] {sum, sum sq} of the fp32 Linear output (bias included)
  const float* gamma;
  const float* beta;
  const float* bias;    // nullable
};

bool gemm_stream_supported(int64_t n, int64_t k1, int64_t k2, int64_t ldc);

// k2 = 0: one Linear (resid optional); k2 > 0: two Linears summed (resid must be null)
int gemm_stream_gnapply(const void* a1, int64_t lda1, const void* b1, int64_t ldb1, int64_t k1, const StreamNorm& n1,
                        const void* a2, int64_t lda2, const void* b2, int64_t ldb2, int64_t k2, const StreamNorm& n2,
                        int64_t m, int64_t n, float eps, float slope, const void* resid, void* out, int64_t ldc,
                        const int64_t* seg_off, int64_t nseg, int64_t groups, int64_t rpp, void* workspace,
                        size_t workspace_bytes, cudaStream_t st);
size_t gemm_stream_workspace_bytes(int64_t n, int64_t nseg);

// out_bf16 = act(alpha * A B^T + bias): plain Linear with a bf16 output on the streaming kernel (slope 0 = ReLU, 1 = none)
int gemm_stream_plain(const void* a, int64_t lda, const void* b, int64_t ldb, int64_t m, int64_t n, int64_t k,
                      const float* bias, float alpha, float slope, void* out, int64_t ldc, cudaStream_t st);
bool gemm_stream_plain_enabled();

}  // namespace se3et
