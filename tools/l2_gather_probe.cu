// l2_gather_probe.cu — ceiling probe for the RoIAlign gather: random 1-KB (or 512-B) chunk reads from a
// channel-last map that sits in L2.  Prints TB/s for a few (warps/SM, loads in flight, working set) points.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/_build/l2probe tools/l2_gather_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int SLOTS, int NQ, bool BYPASS>
__global__ void probe(const float4* __restrict__ buf, unsigned npix, int iters, float* out) {
  const int lane = threadIdx.x & 31;
  unsigned s = (blockIdx.x * blockDim.x + threadIdx.x) / 32 * 2654435761u + 12345u;
  float acc = 0.f;
  for (int it = 0; it < iters; it++) {
    float4 v[SLOTS][NQ];
#pragma unroll
    for (int k = 0; k < SLOTS; k++) {
      s = s * 1664525u + 1013904223u;
      const unsigned pix = (s >> 8) % npix;
      const float4* p = buf + (size_t)pix * 64 + lane;          // 64 float4 = 1 KB per pixel
#pragma unroll
      for (int u = 0; u < NQ; u++) {
        if (BYPASS) asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[k][u].x), "=f"(v[k][u].y), "=f"(v[k][u].z), "=f"(v[k][u].w) : "l"(p + 32 * u));
        else v[k][u] = __ldg(p + 32 * u);
      }
    }
#pragma unroll
    for (int k = 0; k < SLOTS; k++)
#pragma unroll
      for (int u = 0; u < NQ; u++) acc += v[k][u].x + v[k][u].y + v[k][u].z + v[k][u].w;
  }
  if (acc == 123.456f) out[0] = acc;
}

template <int SLOTS, int NQ, bool BYPASS>
void run(const float4* buf, unsigned npix, int ctas_per_sm, int threads, float* out) {
  const int iters = 64;
  const int grid = 148 * ctas_per_sm * 8;      // 8 waves
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  probe<SLOTS, NQ, BYPASS><<<grid, threads>>>(buf, npix, iters, out);   // warm L2
  cudaEventRecord(a);
  probe<SLOTS, NQ, BYPASS><<<grid, threads>>>(buf, npix, iters, out);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b);
  const double bytes = (double)grid * (threads / 32) * iters * SLOTS * NQ * 512.0;
  printf("ws %4u MB  %s  slots %d x %4d B  warps/SM %2d : %.2f TB/s  (%s)\n", npix / 1024, BYPASS ? "L1-bypass" : "L1-alloc ", SLOTS, NQ * 512,
         ctas_per_sm * threads / 32, bytes / (ms * 1e-3) / 1e12, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  float4* buf; float* out;
  const size_t bytes = (size_t)96 << 20;
  cudaMalloc(&buf, bytes); cudaMalloc(&out, 4);
  cudaMemset(buf, 0, bytes);
  for (unsigned mb : {16u, 64u, 96u}) {
    const unsigned npix = mb * 1024;
    run<4, 2, true>(buf, npix, 4, 256, out);
    run<8, 2, true>(buf, npix, 4, 256, out);
    run<8, 1, true>(buf, npix, 4, 256, out);
    run<8, 2, true>(buf, npix, 2, 256, out);
    run<8, 2, true>(buf, npix, 8, 256, out);
    run<16, 1, true>(buf, npix, 8, 256, out);
    run<8, 2, false>(buf, npix, 4, 256, out);
  }
  return 0;
}
