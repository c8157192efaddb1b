// relayout.cu — NCHW -> channel-last (N, HW, C) re-layout shared by the staged RoIAlign path and the
// tcgen05 AlignConv producer.  32x32 tiles through padded smem; both sides 128-B coalesced.
#include "common.cuh"

namespace jdet {

__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ in, float* __restrict__ out,
                                                            int C, int HW) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  const float* src = in + (size_t)b * C * HW;
  float* dst = out + (size_t)b * C * HW;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int c = c0 + ty + 8 * k, p = p0 + tx;
    if (c < C && p < HW) tile[ty + 8 * k][tx] = __ldg(src + (size_t)c * HW + p);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const int p = p0 + ty + 8 * k, c = c0 + tx;
    if (c < C && p < HW) dst[(size_t)p * C + c] = tile[tx][ty + 8 * k];
  }
}

void launch_nchw_to_nhwc(const float* in, float* out, int B, int C, int HW, cudaStream_t st) {
  dim3 g(jdet_ceil_div(HW, 32), jdet_ceil_div(C, 32), B);
  nchw_to_nhwc_kernel<<<g, 256, 0, st>>>(in, out, C, HW);
}

}  // namespace jdet
