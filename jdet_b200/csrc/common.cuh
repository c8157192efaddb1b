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

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs (compile-time default; launch sizing asks the device: num_sms())

// SM count of the current device, cached per device (grid sizing of the persistent kernels)
inline int num_sms() {
  static int cache[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return kNumSMs;
  if (cache[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kNumSMs;
    cache[dev] = n;
  }
  return cache[dev];
}

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

// 16-B read-only load that yields zeros when !live.  volatile: a run of these stays a run of loads in the
// emitted code (the compiler otherwise interleaves each load with its consumers to save registers, which
// serialises the memory latency).
__device__ __forceinline__ float4 ldg_v4_if(const float* p, bool live) {
  float4 v;
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.s32 q, %5, 0;\n\t"
      "mov.b32 %0, 0;\n\tmov.b32 %1, 0;\n\tmov.b32 %2, 0;\n\tmov.b32 %3, 0;\n\t"
      "@q ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
      : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
      : "l"(p), "r"((int)live));
  return v;
}

// 16-B read-only load into v when live; v keeps its value otherwise (no zero-fill: the caller owns the initial state).
// volatile for the same reason as above.
__device__ __forceinline__ void ldg_v4_keep(float4& v, const float* p, bool live) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.s32 q, %5, 0;\n\t"
      "@q ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];\n\t}"
      : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)
      : "l"(p), "r"((int)live));
}

// ---- mbarrier / bulk-async-copy helpers (sm_90+ PTX; SASS: SYNCS.*, UBLKCP) -------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
// global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// shared -> global as one bulk-async group; bulk_s2g_wait_read(): the source smem may be overwritten again
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_s2g_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

}  // namespace jdet
