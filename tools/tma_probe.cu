// tma_probe.cu — which asynchronous fill path can feed the RoIAlign gather?  (numbers quoted in DESIGN.md)
// One producer thread per CTA keeps a ring of shared-memory stages full from an L2-resident 256x256x256 fp32 map:
//   mode 0  cp.async.bulk, one 1-D copy per pixel slab (UBLKCP), channel-last map
//   mode 1  cp.async.bulk.tensor.2d tile::gather4: four arbitrary pixel rows of the (pixels, C) view per instruction
//   mode 2  cp.async.bulk.tensor.3d tile box (CS, BX, BY) of the channel-last map viewed as (C, W, H)
//   mode 3  cp.async.bulk.tensor.3d tile box (BX, BY, CS) of the NCHW map viewed as (W, H, C)
// Prints fill throughput (TB/s) and checks the bytes that landed against the map.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/_build/tma_probe tools/tma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int Wd = 256, Hd = 256, Cd = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_gather4(void* dst, const CUtensorMap* map, int col, int r0, int r1, int r2, int r3, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
__device__ __forceinline__ void tma_tile3(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

struct Params { int mode, cs, bx, by, px_per_stage, stages, iters, warps; };

__global__ void __launch_bounds__(512) probe(const __grid_constant__ CUtensorMap map, const float* __restrict__ buf, Params P, float* dump) {
  extern __shared__ __align__(1024) unsigned char smem_all[];
  __shared__ uint64_t full_all[16][8];
  const uint32_t stage_bytes = (uint32_t)P.px_per_stage * P.cs * 4u;
  const int warp = threadIdx.x >> 5;
  uint64_t* full = full_all[warp];
  unsigned char* smem = smem_all + (size_t)warp * P.stages * stage_bytes;
  if ((threadIdx.x & 31) == 0 && warp < P.warps) {
    for (int s = 0; s < P.stages; s++) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0 && warp < P.warps) {
    unsigned rng = (blockIdx.x * 16 + warp) * 2654435761u + 12345u;
    for (int it = 0; it < P.iters; it++) {
      const int s = it % P.stages;
      if (it >= P.stages) mbar_wait(&full[s], ((it / P.stages) - 1) & 1);
      unsigned char* dst = smem + (size_t)s * stage_bytes;
      mbar_expect_tx(&full[s], stage_bytes);
      rng = rng * 1664525u + 1013904223u;
      const int bx0 = (rng >> 8) % (Wd - 32), by0 = (rng >> 18) % (Hd - 32);
      const int c0 = ((rng >> 4) % (Cd / P.cs)) * P.cs;
      if (P.mode == 0) {
        for (int p = 0; p < P.px_per_stage; p++) {
          rng = rng * 1664525u + 1013904223u;
          const int x = bx0 + ((rng >> 8) & 15), y = by0 + ((rng >> 16) & 15);
          bulk_g2s(dst + (size_t)p * P.cs * 4, buf + ((size_t)y * Wd + x) * Cd + c0, P.cs * 4, &full[s]);
        }
      } else if (P.mode == 1) {
        for (int p = 0; p < P.px_per_stage; p += 4) {
          int r[4];
          for (int j = 0; j < 4; j++) {
            rng = rng * 1664525u + 1013904223u;
            r[j] = (by0 + ((rng >> 16) & 15)) * Wd + bx0 + ((rng >> 8) & 15);
          }
          tma_gather4(dst + (size_t)p * P.cs * 4, &map, c0, r[0], r[1], r[2], r[3], &full[s]);
        }
      } else if (P.mode == 2) {
        const int per = P.bx * P.by;
        for (int p = 0, k = 0; p < P.px_per_stage; p += per, k++)
          tma_tile3(dst + (size_t)p * P.cs * 4, &map, c0, bx0 + (k & 3) * P.bx, by0 + (k >> 2) * P.by, &full[s]);
      } else {
        const int per = P.bx * P.by;
        for (int p = 0, k = 0; p < P.px_per_stage; p += per, k++)
          tma_tile3(dst + (size_t)p * P.cs * 4, &map, bx0 + (k & 1) * P.bx, by0 + (k >> 1) * P.by, c0, &full[s]);
      }
    }
    for (int it = P.iters; it < P.iters + P.stages; it++) {   // drain
      const int s = it % P.stages;
      if (it >= P.stages) mbar_wait(&full[s], ((it / P.stages) - 1) & 1);
    }
  }
  __syncthreads();
  if (dump && blockIdx.x == 0) {   // last stage written by iteration iters-1, re-derived by the host from the same LCG
    const int s = (P.iters - 1) % P.stages;
    const float* src = reinterpret_cast<const float*>(smem_all + (size_t)s * stage_bytes);
    for (int i = threadIdx.x; i < (int)(stage_bytes / 4); i += blockDim.x) dump[i] = src[i];
  }
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeFn get_encode() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  if (!fn) { printf("no cuTensorMapEncodeTiled\n"); exit(1); }
  return (EncodeFn)fn;
}

static float value_nhwc(int y, int x, int c) { return (float)(((y * Wd + x) * 31 + c * 7) % 8191) * 0.125f; }

int main(int argc, char** argv) {
  const size_t n = (size_t)Wd * Hd * Cd;
  std::vector<float> h_nhwc(n), h_nchw(n);
  for (int y = 0; y < Hd; y++) for (int x = 0; x < Wd; x++) for (int c = 0; c < Cd; c++) {
    const float v = value_nhwc(y, x, c);
    h_nhwc[((size_t)y * Wd + x) * Cd + c] = v;
    h_nchw[((size_t)c * Hd + y) * Wd + x] = v;
  }
  float *d_nhwc, *d_nchw, *d_dump;
  CK(cudaMalloc(&d_nhwc, n * 4)); CK(cudaMalloc(&d_nchw, n * 4)); CK(cudaMalloc(&d_dump, 256 * 1024));
  CK(cudaMemcpy(d_nhwc, h_nhwc.data(), n * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_nchw, h_nchw.data(), n * 4, cudaMemcpyHostToDevice));
  EncodeFn enc = get_encode();
  int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));

  struct Cfg { int mode, cs, bx, by, px, stages, ctas; int g4box; int warps; };
  std::vector<Cfg> cfgs;
  if (argc > 2 && atoi(argv[2]) == 3) {   // NCHW box sweep: which (BX, BY, CS) boxes does the tensor-map path take?
    for (int cs : {1, 2, 4, 8, 16, 32})
      for (int bx : {32, 64}) cfgs.push_back({3, cs, bx, 8, bx * 8 * 2, 2, 1, 0, 4});
  } else {
  for (int warps : {1, 2, 4, 8, 16}) {
    const int ctas = 1;
    cfgs.push_back({0, 128, 0, 0, 16, 2, ctas, 0, warps});    // bulk 512 B per op
    cfgs.push_back({1, 128, 0, 0, 8, 2, ctas, 1, warps});     // gather4 2 KB per op (4 KB stages)
    cfgs.push_back({1, 64, 0, 0, 16, 2, ctas, 1, warps});     // gather4 1 KB per op
    cfgs.push_back({2, 128, 4, 4, 16, 2, ctas, 0, warps});    // tile 4x4x128ch 8 KB per op
    cfgs.push_back({2, 64, 4, 4, 16, 2, ctas, 0, warps});     // tile 4x4x64ch 4 KB per op
    cfgs.push_back({2, 64, 2, 2, 16, 2, ctas, 0, warps});     // tile 2x2x64ch 1 KB per op
  }
  }
  const int only = argc > 1 ? atoi(argv[1]) : -1;
  if (only >= (int)cfgs.size()) return 3;
  for (int ci = 0; ci < (int)cfgs.size(); ci++) {
    if (only >= 0 && ci != only) continue;
    const Cfg& c = cfgs[ci];
    CUtensorMap map;
    memset(&map, 0, sizeof(map));
    CUresult r = CUDA_SUCCESS;
    if (c.mode == 1) {
      cuuint64_t gdim[2] = {(cuuint64_t)Cd, (cuuint64_t)Wd * Hd};
      cuuint64_t gstr[1] = {(cuuint64_t)Cd * 4};
      cuuint32_t box[2] = {(cuuint32_t)c.cs, (cuuint32_t)c.g4box};
      cuuint32_t es[2] = {1, 1};
      r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_nhwc, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else if (c.mode == 2) {
      cuuint64_t gdim[3] = {(cuuint64_t)Cd, (cuuint64_t)Wd, (cuuint64_t)Hd};
      cuuint64_t gstr[2] = {(cuuint64_t)Cd * 4, (cuuint64_t)Cd * Wd * 4};
      cuuint32_t box[3] = {(cuuint32_t)c.cs, (cuuint32_t)c.bx, (cuuint32_t)c.by};
      cuuint32_t es[3] = {1, 1, 1};
      r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d_nhwc, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else if (c.mode == 3) {
      cuuint64_t gdim[3] = {(cuuint64_t)Wd, (cuuint64_t)Hd, (cuuint64_t)Cd};
      cuuint64_t gstr[2] = {(cuuint64_t)Wd * 4, (cuuint64_t)Wd * Hd * 4};
      cuuint32_t box[3] = {(cuuint32_t)c.bx, (cuuint32_t)c.by, (cuuint32_t)c.cs};
      cuuint32_t es[3] = {1, 1, 1};
      r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d_nchw, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) { printf("mode %d cs %d: encode failed (%d)\n", c.mode, c.cs, (int)r); continue; }
    Params P{c.mode, c.cs, c.bx, c.by, c.px, c.stages, 2000, c.warps};
    const size_t stage_bytes = (size_t)c.px * c.cs * 4, smem = stage_bytes * c.stages * c.warps;
    const int grid = sms * c.ctas;
    CK(cudaMemset(d_dump, 0xff, 256 * 1024));
    probe<<<grid, 32 * c.warps, smem>>>(map, c.mode == 3 ? d_nchw : d_nhwc, P, d_dump);   // warm
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d cs %d box %dx%d: launch failed: %s\n", c.mode, c.cs, c.bx, c.by, cudaGetErrorString(e)); return 1; }
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    probe<<<grid, 32 * c.warps, smem>>>(map, c.mode == 3 ? d_nchw : d_nhwc, P, nullptr);
    cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double bytes = (double)grid * c.warps * P.iters * stage_bytes;
    const double ops = (double)grid * c.warps * P.iters * (c.mode == 0 ? c.px : c.mode == 1 ? c.px / 4 : c.px / (c.bx * c.by));
    // verify block 0's last stage
    std::vector<float> h(stage_bytes / 4);
    CK(cudaMemcpy(h.data(), d_dump, stage_bytes, cudaMemcpyDeviceToHost));
    unsigned rng = 0 * 2654435761u + 12345u;
    long bad = 0, total = 0;
    for (int it = 0; it < P.iters; it++) {
      rng = rng * 1664525u + 1013904223u;
      const int bx0 = (rng >> 8) % (Wd - 32), by0 = (rng >> 18) % (Hd - 32);
      const int c0 = ((rng >> 4) % (Cd / c.cs)) * c.cs;
      const bool last = it == P.iters - 1;
      if (c.mode == 0 || c.mode == 1) {
        for (int p = 0; p < c.px; p++) {
          rng = rng * 1664525u + 1013904223u;
          const int x = bx0 + ((rng >> 8) & 15), y = by0 + ((rng >> 16) & 15);
          if (last) for (int ch = 0; ch < c.cs; ch++) { total++; bad += h[(size_t)p * c.cs + ch] != value_nhwc(y, x, c0 + ch); }
        }
      } else if (last && c.mode == 2) {
        const int per = c.bx * c.by;
        for (int p = 0, k = 0; p < c.px; p += per, k++)
          for (int yy = 0; yy < c.by; yy++) for (int xx = 0; xx < c.bx; xx++) for (int ch = 0; ch < c.cs; ch++) {
            total++;
            bad += h[(size_t)(p + yy * c.bx + xx) * c.cs + ch] != value_nhwc(by0 + (k >> 2) * c.by + yy, bx0 + (k & 3) * c.bx + xx, c0 + ch);
          }
      } else if (last) {
        const int per = c.bx * c.by;
        for (int p = 0, k = 0; p < c.px; p += per, k++)
          for (int ch = 0; ch < c.cs; ch++) for (int yy = 0; yy < c.by; yy++) for (int xx = 0; xx < c.bx; xx++) {
            total++;
            bad += h[(size_t)p * c.cs + ((size_t)ch * c.by + yy) * c.bx + xx] != value_nhwc(by0 + (k >> 1) * c.by + yy, bx0 + (k & 1) * c.bx + xx, c0 + ch);
          }
      }
    }
    printf("mode %d cs %3d box %2dx%2d issuers/SM %2d px/stage %3d stages %d ctas/SM %d (%3zu KB in flight/CTA): %6.2f TB/s  %7.1f Mops/s/SM (%.1f clk/op @1.9GHz)  mismatches %ld/%ld\n",
           c.mode, c.cs, c.bx, c.by, c.warps * c.ctas, c.px, c.stages, c.ctas, smem / 1024, bytes / (ms * 1e-3) / 1e12, ops / (ms * 1e-3) / 1e6 / sms,
           1.9e9 / (ops / (ms * 1e-3) / sms), bad, total);
    fflush(stdout);
  }
  return 0;
}
