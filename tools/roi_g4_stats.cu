// roi_g4_stats.cu — where do the warps of roi_gather4_kernel spend their cycles?  Builds the RoIAlign translation unit with
// JDET_G4_STATS (clock64 counters per role) and runs cfg2 through the C ABI; prints cycles per SM per role.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -DJDET_G4_STATS -Ijdet_b200/csrc -o tools/_build/roi_g4_stats tools/roi_g4_stats.cu jdet_b200/csrc/relayout.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
unsigned long long* g_jdet_g4_stats = nullptr;
#include "../jdet_b200/csrc/roi_align_rotated.cu"

int main() {
  const int B = 1, C = 256, H = 256, W = 256, R = 2048, PH = 7, PW = 7, sr = 2;
  std::vector<float> rois(R * 6);
  srand(1);
  auto u = [] { return rand() / (float)RAND_MAX; };
  for (int i = 0; i < R; i++) {
    float L = 8.f * powf(32.f, u()), ar = powf(8.f, u()), w = L, h = L / ar;
    if (u() < 0.5f) std::swap(w, h);
    rois[i * 6 + 0] = 0; rois[i * 6 + 1] = u() * 1024; rois[i * 6 + 2] = u() * 1024; rois[i * 6 + 3] = w; rois[i * 6 + 4] = h;
    rois[i * 6 + 5] = (u() - 0.5f) * 3.14159f;
  }
  float *d_in, *d_rois, *d_out; void* ws;
  const size_t nin = (size_t)B * C * H * W, nout = (size_t)R * C * PH * PW;
  cudaMalloc(&d_in, nin * 4); cudaMalloc(&d_rois, rois.size() * 4); cudaMalloc(&d_out, nout * 4);
  cudaMemset(d_in, 0, nin * 4);
  cudaMemcpy(d_rois, rois.data(), rois.size() * 4, cudaMemcpyHostToDevice);
  const size_t wsb = jdet_roi_align_rotated_workspace_bytes(B, C, H, W, R, PH, PW, sr);
  cudaMalloc(&ws, wsb);
  cudaMalloc(&g_jdet_g4_stats, 64 * 8);
  for (int it = 0; it < 3; it++) {
    cudaMemset(g_jdet_g4_stats, 0, 64 * 8);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    int rc = jdet_roi_align_rotated(1, d_in, B, C, H, W, d_rois, R, PH, PW, 0.25f, sr, d_out, ws, wsb, nullptr);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    unsigned long long h[8];
    cudaMemcpy(h, g_jdet_g4_stats, sizeof(h), cudaMemcpyDeviceToHost);
    printf("rc %d  %.1f us  (%s)\n", rc, ms * 1e3, cudaGetErrorString(cudaGetLastError()));
    const char* names[8] = {"producer: wait table", "producer: wait ring space", "producer: issue", "-", "consumer: wait table", "consumer: wait chunk",
                            "consumer: compute", "consumer: release + item barrier"};
    for (int i = 0; i < 8; i++) printf("  %-34s %9.0f cycles per warp (sum over the launch / SMs / warps of the role)\n", names[i], h[i] / 148.0 / (i < 4 ? 8 : 16));
  }
  return 0;
}
