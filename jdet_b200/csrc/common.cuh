// common.cuh — shared helpers for the jdet_b200 C-ABI kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>

#define JDET_API extern "C" __attribute__((visibility("default")))

// error codes returned by every entry point: 0 = ok, >0 = cudaError_t, <0 = argument error
#define JDET_ERR_BAD_ARG (-1)
#define JDET_ERR_WORKSPACE (-2)
#define JDET_ERR_UNSUPPORTED (-3)

#define JDET_RETURN_IF_CUDA(expr)                 \
  do {                                            \
    cudaError_t _e = (expr);                      \
    if (_e != cudaSuccess) return (int)_e;        \
  } while (0)

static inline size_t jdet_align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
static inline int jdet_ceil_div(int a, int b) { return (a + b - 1) / b; }

namespace jdet {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31; }

// streaming 16-B store that does not allocate in L1 (outputs are write-once)
__device__ __forceinline__ void st_stream_v4(float* p, float a, float b, float c, float d) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b),
               "f"(c), "f"(d)
               : "memory");
}
__device__ __forceinline__ void st_stream(float* p, float a) {
  asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory");
}

}  // namespace jdet
