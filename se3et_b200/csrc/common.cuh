// Shared helpers for the se3et_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/se3et_b200.h"

namespace se3et {

void set_last_error(const char* what, cudaError_t err);

#define SE3ET_CUDA_CHECK(expr)                      \
  do {                                              \
    cudaError_t _e = (expr);                        \
    if (_e != cudaSuccess) {                        \
      ::se3et::set_last_error(#expr, _e);           \
      return SE3ET_ERR_CUDA;                        \
    }                                               \
  } while (0)

#define SE3ET_LAUNCH_CHECK() SE3ET_CUDA_CHECK(cudaGetLastError())

constexpr int kNumSMs = 148;  // B200
constexpr int kMaxDevices = 32;

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device attribute: remembered per (call site, device) so that a
// process driving several GPUs configures every one of them; host threads may race here (atomics, idempotent call).
#define SE3ET_ENSURE_SMEM(kernel, bytes)                                                                            \
  do {                                                                                                              \
    static std::atomic<int> _cfg[::se3et::kMaxDevices];                                                             \
    int _dev = 0;                                                                                                   \
    SE3ET_CUDA_CHECK(cudaGetDevice(&_dev));                                                                         \
    const bool _in = _dev >= 0 && _dev < ::se3et::kMaxDevices;                                                      \
    if (!_in || _cfg[_dev].load(std::memory_order_relaxed) < (int)(bytes)) {                                        \
      SE3ET_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));   \
      if (_in) _cfg[_dev].store((int)(bytes), std::memory_order_relaxed);                                           \
    }                                                                                                               \
  } while (0)

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Carves a workspace into aligned sub-buffers; `fits()` tells whether it overflowed.
struct Carver {
  char* base;
  size_t size, off;
  Carver(void* p, size_t n) : base(static_cast<char*>(p)), size(n), off(0) {}
  template <typename T>
  T* take(size_t count) {
    off = align_up(off, 256);
    T* r = reinterpret_cast<T*>(base + off);
    off += count * sizeof(T);
    return r;
  }
  bool fits() const { return off <= size; }
};

// ---- device helpers -------------------------------------------------------------------------

// index of the stack-mode segment containing global row i (offsets has nseg+1 entries)
__device__ __forceinline__ int segment_of(const int64_t* __restrict__ offsets, int nseg, int64_t i) {
  int lo = 0, hi = nseg - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (offsets[mid] <= i) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// order-preserving float <-> uint32 mapping (for atomicMin/Max on floats)
__device__ __forceinline__ uint32_t float_to_ordered(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace se3et
