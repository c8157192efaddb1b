// feature_refine.cu — rotated_feature_align ("feature_refine" in JDet) forward for sm_100a.
//
// Replaces feature_refine_forward_kernel + launch (/root/reference/python/jdet/ops/fr.py:114-165,
// 234-240):  out[n,c,h,w] = in[n,c,h,w] + sum_{i<points} bilinear(in[n,c], py_i, px_i),
// points in {1,5}; bbox (n,h,w,5): bbox[0]*scale is the ROW, bbox[1]*scale the COLUMN (:133-134).
//
// Reference shape of work: one thread per (n,c,h,w): the box is re-read and the 1/5 sample points
// (cosf/sinf included) re-derived for every channel, 256x redundantly at the bench shape.
// Here a thread owns one pixel: it decodes its box once into taps+weights held in registers, then
// walks a slab of channels.  Consecutive lanes are consecutive pixels, so the centre read and the
// store are 128-B coalesced and the taps of neighbouring pixels fall in the same few lines of the
// plane (L1 hits).  HBM traffic ~ one read + one write of the feature map: an HBM-bound stream.
//
// Layout in HBM: features (N,C,H,W) fp32, boxes (N,H,W,5) fp32, out (N,C,H,W) fp32.
#include <algorithm>
#include "common.cuh"

namespace jdet {

struct Tap4 {
  int o00, o01, o10, o11;
  float w1, w2, w3, w4;
};

// fr.py:18-67
__device__ __forceinline__ Tap4 fr_tap(float y, float x, int H, int W) {
  Tap4 t;
  if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W || !(y == y) || !(x == x)) {
    t.o00 = -1; t.o01 = t.o10 = t.o11 = 0;          // sample contributes exactly 0 (fr.py:24-26)
    t.w1 = t.w2 = t.w3 = t.w4 = 0.f;
    return t;
  }
  if (y <= 0.f) y = 0.f;
  if (x <= 0.f) x = 0.f;
  int yl = (int)y, xl = (int)x, yh, xh;
  if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
  if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
  const float ly = y - (float)yl, lx = x - (float)xl;
  const float hy = 1.f - ly, hx = 1.f - lx;
  t.o00 = yl * W + xl; t.o01 = yl * W + xh; t.o10 = yh * W + xl; t.o11 = yh * W + xh;
  t.w1 = __fmul_rn(hy, hx); t.w2 = __fmul_rn(hy, lx); t.w3 = __fmul_rn(ly, hx); t.w4 = __fmul_rn(ly, lx);
  return t;
}

// grid = (pixel tiles of 256, channel slabs, N)
template <int POINTS>
__global__ void __launch_bounds__(256, 3) feature_refine_kernel(const float* __restrict__ feat, const float* __restrict__ boxes,
                                                              int C, int H, int W, float spatial_scale, int ch_per_cta,
                                                              float* __restrict__ out) {
  const int HW = H * W;
  const int p = blockIdx.x * 256 + threadIdx.x;
  const int n = blockIdx.z;
  const int c0 = blockIdx.y * ch_per_cta, c1 = min(C, c0 + ch_per_cta);
  if (p >= HW) return;
  const float* bb = boxes + ((size_t)n * HW + p) * 5;
  const float roi_y = __fmul_rn(bb[0], spatial_scale), roi_x = __fmul_rn(bb[1], spatial_scale);
  Tap4 taps[POINTS];
  taps[0] = fr_tap(roi_y, roi_x, H, W);
  if (POINTS > 1) {   // fr.py:139-153
    const float rw = __fmul_rn(bb[2], spatial_scale), rh = __fmul_rn(bb[3], spatial_scale), ra = bb[4];
    const float w2 = rw * 0.5f, h2 = rh * 0.5f;
    const float ca = cosf(ra), sa = sinf(ra);
    const float wx = __fmul_rn(ca, w2), wy = __fmul_rn(sa, w2), hx = __fmul_rn(-sa, h2), hy = __fmul_rn(ca, h2);
    taps[1 % POINTS] = fr_tap(__fadd_rn(__fadd_rn(roi_y, wy), hy), __fadd_rn(__fadd_rn(roi_x, wx), hx), H, W);
    taps[2 % POINTS] = fr_tap(__fadd_rn(__fsub_rn(roi_y, wy), hy), __fadd_rn(__fsub_rn(roi_x, wx), hx), H, W);
    taps[3 % POINTS] = fr_tap(__fsub_rn(__fsub_rn(roi_y, wy), hy), __fsub_rn(__fsub_rn(roi_x, wx), hx), H, W);
    taps[4 % POINTS] = fr_tap(__fsub_rn(__fadd_rn(roi_y, wy), hy), __fsub_rn(__fadd_rn(roi_x, wx), hx), H, W);
  }
  const float* plane = feat + ((size_t)n * C + c0) * HW;
  float* dst = out + ((size_t)n * C + c0) * HW + p;
#pragma unroll 2
  for (int c = c0; c < c1; c++) {
    float v = __ldg(plane + p);
#pragma unroll
    for (int i = 0; i < POINTS; i++) {
      const Tap4& t = taps[i];
      if (t.o00 >= 0)
        v += t.w1 * __ldg(plane + t.o00) + t.w2 * __ldg(plane + t.o01) + t.w3 * __ldg(plane + t.o10) +
           t.w4 * __ldg(plane + t.o11);
    }
    st_stream(dst, v);
    plane += HW;
    dst += HW;
  }
}

// points == 1, W % 4 == 0: a thread owns 4 consecutive pixels, so the centre read and the store are
// 16-B vectors (4x the bytes in flight per load instruction: the op is an HBM stream whose speed is
// set by memory-level parallelism); the 4 x 4 taps stay scalar L1 gathers.
__global__ void __launch_bounds__(256) feature_refine_p1_vec4_kernel(const float* __restrict__ feat,
                                                                     const float* __restrict__ boxes, int C, int H, int W,
                                                                     float spatial_scale, int ch_per_cta,
                                                                     float* __restrict__ out) {
  const int HW = H * W;
  const int p = (blockIdx.x * 256 + threadIdx.x) * 4;
  const int n = blockIdx.z;
  const int c0 = blockIdx.y * ch_per_cta, c1 = min(C, c0 + ch_per_cta);
  if (p >= HW) return;
  Tap4 t[4];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const float* bb = boxes + ((size_t)n * HW + p + j) * 5;
    t[j] = fr_tap(__fmul_rn(__ldg(bb), spatial_scale), __fmul_rn(__ldg(bb + 1), spatial_scale), H, W);
  }
  const float* plane = feat + ((size_t)n * C + c0) * HW;
  float* dst = out + ((size_t)n * C + c0) * HW + p;
#pragma unroll 2
  for (int c = c0; c < c1; c++) {
    const float4 ctr = __ldg(reinterpret_cast<const float4*>(plane + p));
    float v[4] = {ctr.x, ctr.y, ctr.z, ctr.w};
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (t[j].o00 >= 0)
        v[j] += t[j].w1 * __ldg(plane + t[j].o00) + t[j].w2 * __ldg(plane + t[j].o01) + t[j].w3 * __ldg(plane + t[j].o10) +
                t[j].w4 * __ldg(plane + t[j].o11);
    st_stream_v4(dst, v[0], v[1], v[2], v[3]);
    plane += HW;
    dst += HW;
  }
}

// backward (fr.py:167-232): grad_in = grad_out + scatter(grad_out * w) — thread = pixel, channel walk.
// grad_in must already hold a copy of grad_out (the identity term); taps are float atomics as in the reference.
template <int POINTS>
__global__ void __launch_bounds__(256, 2) feature_refine_bwd_kernel(const float* __restrict__ grad_out,
                                                                     const float* __restrict__ boxes, int C, int H, int W,
                                                                     float spatial_scale, int ch_per_cta,
                                                                     float* __restrict__ grad_in) {
  const int HW = H * W;
  const int p = blockIdx.x * 256 + threadIdx.x;
  const int n = blockIdx.z;
  const int c0 = blockIdx.y * ch_per_cta, c1 = min(C, c0 + ch_per_cta);
  if (p >= HW) return;
  const float* bb = boxes + ((size_t)n * HW + p) * 5;
  const float roi_y = __fmul_rn(bb[0], spatial_scale), roi_x = __fmul_rn(bb[1], spatial_scale);
  Tap4 taps[POINTS];
  taps[0] = fr_tap(roi_y, roi_x, H, W);
  if (POINTS > 1) {
    const float rw = __fmul_rn(bb[2], spatial_scale), rh = __fmul_rn(bb[3], spatial_scale), ra = bb[4];
    const float w2 = rw * 0.5f, h2 = rh * 0.5f;
    const float ca = cosf(ra), sa = sinf(ra);
    const float wx = __fmul_rn(ca, w2), wy = __fmul_rn(sa, w2), hx = __fmul_rn(-sa, h2), hy = __fmul_rn(ca, h2);
    taps[1 % POINTS] = fr_tap(__fadd_rn(__fadd_rn(roi_y, wy), hy), __fadd_rn(__fadd_rn(roi_x, wx), hx), H, W);
    taps[2 % POINTS] = fr_tap(__fadd_rn(__fsub_rn(roi_y, wy), hy), __fadd_rn(__fsub_rn(roi_x, wx), hx), H, W);
    taps[3 % POINTS] = fr_tap(__fsub_rn(__fsub_rn(roi_y, wy), hy), __fsub_rn(__fsub_rn(roi_x, wx), hx), H, W);
    taps[4 % POINTS] = fr_tap(__fsub_rn(__fadd_rn(roi_y, wy), hy), __fsub_rn(__fadd_rn(roi_x, wx), hx), H, W);
  }
  const float* src = grad_out + ((size_t)n * C + c0) * HW + p;
  float* plane = grad_in + ((size_t)n * C + c0) * HW;
  for (int c = c0; c < c1; c++) {
    const float g = __ldg(src);
#pragma unroll
    for (int i = 0; i < POINTS; i++) {
      const Tap4& t = taps[i];
      if (t.o00 >= 0) {
        atomicAdd(plane + t.o00, g * t.w1);
        atomicAdd(plane + t.o01, g * t.w2);
        atomicAdd(plane + t.o10, g * t.w3);
        atomicAdd(plane + t.o11, g * t.w4);
      }
    }
    src += HW;
    plane += HW;
  }
}

// ---- TMA-staged kernel --------------------------------------------------------------------------
// Full-width row bands of an NCHW plane are contiguous, so the band a CTA works on (its rows plus kHalo
// rows above and below) is ONE cp.async.bulk (UBLKCP) per channel into a shared-memory ring guarded by
// mbarriers: kStages channels are in flight per CTA with no registers tied up, which is what an HBM
// stream with a short gather needs.  Centre reads and taps are then served from shared memory (lanes are
// consecutive pixels: conflict-light); a sample whose taps leave the band falls back to global loads.
// CTA = (row band, channel slab, image), one thread per pixel of the band.
// the 1 or 5 sample points of a box (fr.py:128-153), decoded into taps
template <int POINTS>
__device__ __forceinline__ void fr_decode(const float* __restrict__ bb, float spatial_scale, int H, int W, Tap4 (&taps)[POINTS]) {
  const float roi_y = __fmul_rn(__ldg(bb), spatial_scale), roi_x = __fmul_rn(__ldg(bb + 1), spatial_scale);
  taps[0] = fr_tap(roi_y, roi_x, H, W);
  if (POINTS > 1) {
    const float rw = __fmul_rn(__ldg(bb + 2), spatial_scale), rh = __fmul_rn(__ldg(bb + 3), spatial_scale), ra = __ldg(bb + 4);
    const float w2 = rw * 0.5f, h2 = rh * 0.5f;
    const float ca = cosf(ra), sa = sinf(ra);
    const float wx = __fmul_rn(ca, w2), wy = __fmul_rn(sa, w2), hx = __fmul_rn(-sa, h2), hy = __fmul_rn(ca, h2);
    taps[1 % POINTS] = fr_tap(__fadd_rn(__fadd_rn(roi_y, wy), hy), __fadd_rn(__fadd_rn(roi_x, wx), hx), H, W);
    taps[2 % POINTS] = fr_tap(__fadd_rn(__fsub_rn(roi_y, wy), hy), __fadd_rn(__fsub_rn(roi_x, wx), hx), H, W);
    taps[3 % POINTS] = fr_tap(__fsub_rn(__fsub_rn(roi_y, wy), hy), __fsub_rn(__fsub_rn(roi_x, wx), hx), H, W);
    taps[4 % POINTS] = fr_tap(__fsub_rn(__fadd_rn(roi_y, wy), hy), __fsub_rn(__fadd_rn(roi_x, wx), hx), H, W);
  }
}

namespace fr_tma {

constexpr int kHalo = 4;
constexpr int kMaxStages = 8;

constexpr int kConsumerWarps = 8;
constexpr int kConsumers = kConsumerWarps * 32;

// CTA = 8 consumer warps + 1 producer warp; grid = (row bands, channel chunks, N).  A band is 256*PPT / W
// full-width rows (+ kHalo rows either side): contiguous in an NCHW plane, so one cp.async.bulk per channel
// stages it.  The producer lane only waits on empty[] and issues copies; the consumers never block on a
// refill, and `stages` channels per CTA x several CTAs per SM are in flight.  Each consumer thread owns PPT
// pixels (rows 256/W apart): their POINTS tap sets and weights live in registers for the whole channel walk
// (points = 1: PPT = 4; points = 5: PPT = 2); a sample whose taps leave the staged rows reads them from global.
template <int POINTS, int PPT>
__device__ __forceinline__ void fr_tma_body(const float* __restrict__ feat, const float* __restrict__ boxes, int C, int H, int W,
                                            float spatial_scale, int rows_per_band, int ch_per_cta, int stages,
                                            float* __restrict__ out, int band, int chunk, int n) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int HW = H * W;
  const int c0 = chunk * ch_per_cta, c1 = min(C, c0 + ch_per_cta);
  const int r0 = band * rows_per_band;                             // first output row of this band
  const int lo = max(0, r0 - kHalo), hi = min(H, r0 + rows_per_band + kHalo);   // staged rows [lo, hi)
  const uint32_t band_bytes = (uint32_t)((hi - lo) * W) * 4u;
  const int stage_elems = (rows_per_band + 2 * kHalo) * W;
  float* ring = reinterpret_cast<float*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)stages * stage_elems);
  uint64_t* empty = full + kMaxStages;
  const int tid = threadIdx.x, lane = tid & 31;
  const int nch = c1 - c0;

  if (tid == 0) {
    for (int s = 0; s < stages; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], kConsumerWarps); }
    mbar_init_fence();
  }
  __syncthreads();

  if (tid >= kConsumers) {                                         // ---- producer warp
    if (lane == 0) {
      const float* src0 = feat + ((size_t)n * C + c0) * HW + (size_t)lo * W;
      uint32_t stage = 0, phase = 0;
      for (int c = 0; c < nch; c++) {
        if (c >= stages) mbar_wait(&empty[stage], phase ^ 1);      // consumers released the previous occupant
        mbar_expect_tx(&full[stage], band_bytes);
        bulk_g2s(ring + (size_t)stage * stage_elems, src0 + (size_t)c * HW, band_bytes, &full[stage]);
        if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1; }
      }
    }
    return;
  }

  // ---- consumers: this thread's pixels and their taps (band-relative offsets when every corner is staged)
  Tap4 taps[PPT][POINTS];
  int pc[PPT];                                                     // centre, band-relative; -1: pixel not in the map
  unsigned staged[PPT];                                            // bit i: sample i reads the smem band
  const int blo = lo * W, bhi = hi * W;
  const int band_px = min(rows_per_band, H - r0) * W;
#pragma unroll
  for (int j = 0; j < PPT; j++) {
    const int i = tid + j * kConsumers;                            // pixel index inside the band
    pc[j] = -1;
    staged[j] = 0u;
#pragma unroll
    for (int k = 0; k < POINTS; k++) taps[j][k].o00 = -1;
    if (i < band_px) {
      const int p = r0 * W + i;
      fr_decode<POINTS>(boxes + ((size_t)n * HW + p) * 5, spatial_scale, H, W, taps[j]);
#pragma unroll
      for (int k = 0; k < POINTS; k++) {
        Tap4& t = taps[j][k];
        if (t.o00 >= blo && t.o11 < bhi) {                         // o00 is the smallest, o11 the largest offset
          staged[j] |= 1u << k;
          t.o00 -= blo; t.o01 -= blo; t.o10 -= blo; t.o11 -= blo;
        }
      }
      pc[j] = p - blo;
    }
  }
  float* dst = out + ((size_t)n * C + c0) * HW + blo;
  const float* gplane = feat + ((size_t)n * C + c0) * HW;

  uint32_t stage = 0, phase = 0;
  for (int c = 0; c < nch; c++) {
    mbar_wait(&full[stage], phase);
    const float* sp = ring + (size_t)stage * stage_elems;
    float v[PPT];
#pragma unroll
    for (int j = 0; j < PPT; j++) {
      v[j] = 0.f;
      if (pc[j] >= 0) {
        v[j] = sp[pc[j]];
#pragma unroll
        for (int k = 0; k < POINTS; k++) {
          const Tap4& t = taps[j][k];
          if (staged[j] >> k & 1u) v[j] += t.w1 * sp[t.o00] + t.w2 * sp[t.o01] + t.w3 * sp[t.o10] + t.w4 * sp[t.o11];
          else if (t.o00 >= 0)
            v[j] += t.w1 * __ldg(gplane + t.o00) + t.w2 * __ldg(gplane + t.o01) + t.w3 * __ldg(gplane + t.o10) + t.w4 * __ldg(gplane + t.o11);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < PPT; j++)
      if (pc[j] >= 0) st_stream(dst + pc[j], v[j]);
    __syncwarp();                                                  // every lane issued its stores, so its smem reads returned
    if (lane == 0) mbar_arrive(&empty[stage]);
    dst += HW;
    gplane += HW;
    if (++stage == (uint32_t)stages) { stage = 0; phase ^= 1; }
  }
}

template <int POINTS, int PPT>
__global__ void __launch_bounds__(kConsumers + 32) feature_refine_tma_kernel(const float* __restrict__ feat,
                                                                             const float* __restrict__ boxes, int C, int H,
                                                                             int W, float spatial_scale, int rows_per_band,
                                                                             int ch_per_cta, int stages, float* __restrict__ out) {
  fr_tma_body<POINTS, PPT>(feat, boxes, C, H, W, spatial_scale, rows_per_band, ch_per_cta, stages, out, blockIdx.x, blockIdx.y, blockIdx.z);
}

// Every FPN level of a head in ONE launch (FeatureRefineModule applies FR to each level, fr.py:339-346): a 1-D grid of
// (level, image, channel chunk, row band) items, level 0 first.  At cfg4 the five per-level launches took 130 us for 361 MB
// while level 0 alone streams at 4.3 TB/s; the coarse levels (8 x 8 ... 32 x 32 maps) are launch-bound on their own.
constexpr int kMaxLevels = 8;
struct FrLevel { const float* feat; const float* boxes; float* out; int H, W; float scale; int rows, cpc, stages, bands, chunks, item_begin; };
struct FrLevels { FrLevel lv[kMaxLevels]; int n, C; };
template <int POINTS, int PPT>
__global__ void __launch_bounds__(kConsumers + 32) feature_refine_tma_multi_kernel(const __grid_constant__ FrLevels L) {
  int l = 0;
#pragma unroll
  for (int i = 1; i < kMaxLevels; i++)
    if (i < L.n && (int)blockIdx.x >= L.lv[i].item_begin) l = i;
  // a static index per case: the level's scalars stay constant-bank operands (a dynamically indexed parameter struct costs
  // 24 more registers and the third resident CTA per SM)
#define JDET_FR_CASE(I)                                                                                                \
  case I: {                                                                                                            \
    int item = blockIdx.x - L.lv[I].item_begin;                                                                        \
    const int band = item % L.lv[I].bands; item /= L.lv[I].bands;                                                      \
    const int chunk = item % L.lv[I].chunks;                                                                           \
    fr_tma_body<POINTS, PPT>(L.lv[I].feat, L.lv[I].boxes, L.C, L.lv[I].H, L.lv[I].W, L.lv[I].scale, L.lv[I].rows, L.lv[I].cpc,  \
                             L.lv[I].stages, L.lv[I].out, band, chunk, item / L.lv[I].chunks);                         \
    break;                                                                                                             \
  }
  switch (l) {
    JDET_FR_CASE(0) JDET_FR_CASE(1) JDET_FR_CASE(2) JDET_FR_CASE(3) JDET_FR_CASE(4) JDET_FR_CASE(5) JDET_FR_CASE(6) JDET_FR_CASE(7)
  }
#undef JDET_FR_CASE
}

}  // namespace fr_tma

}  // namespace jdet

namespace jdet {
// band / stage / channel-chunk geometry of the TMA-staged path for one map; false: the map does not qualify.
static bool fr_tma_config(const float* features, int N, int C, int H, int W, int levels, fr_tma::FrLevel* out) {
  using namespace fr_tma;
  const int band_pixels = kConsumers * 4;
  if (W % 4 != 0 || W > band_pixels || ((uintptr_t)features & 15) != 0) return false;
  const int rows = max(1, min(H, band_pixels / W));
  const int stage_elems = (rows + 2 * kHalo) * W;
  int stages = (int)((64 * 1024) / ((size_t)stage_elems * 4));   // ~64 KB of copies in flight per CTA
  stages = stages > kMaxStages ? kMaxStages : stages;
  if (stages < 2) return false;
  const int bands = jdet_ceil_div(H, rows);
  int cpc = C;
  const long long want = 148 * 6;
  while (cpc > 4 * stages && (long long)bands * jdet_ceil_div(C, cpc) * N < want) cpc = (cpc + 1) / 2;
  out->H = H; out->W = W; out->rows = rows; out->stages = stages; out->cpc = cpc; out->bands = bands; out->chunks = jdet_ceil_div(C, cpc);
  return true;
}
}  // namespace jdet

// jdet.ops.fr.feature_refine(features, best_rbboxes, spatial_scale, points) (ops/fr.py:255-273)
JDET_API int jdet_feature_refine(const float* features, const float* best_rbboxes, int N, int C, int H, int W,
                                 int points, float spatial_scale, float* output, void* stream) {
  using namespace jdet;
  if (N < 0 || C < 0 || H < 0 || W < 0 || (points != 1 && points != 5)) return JDET_ERR_BAD_ARG;
  if ((size_t)N * C * H * W == 0) return 0;
  if (!features || !best_rbboxes || !output) return JDET_ERR_BAD_ARG;
  if (N > 65535) return JDET_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int HW = H * W;
  // TMA-staged path: full-width row bands (contiguous in NCHW), one thread per pixel of the band
  // TMA-staged path: full-width row bands (contiguous in NCHW)
  // (measured on B200, cfg4: points = 1: 0.130 ms vs 0.203 ms for the 16-B-vector register gather; points = 5 with
  //  PPT = 2 (110 registers, 2 CTAs/SM): 1.14 ms vs 0.68 ms for the register gather — 21 dependent smem reads per
  //  pixel-channel on 16 warps per SM are latency-bound, so points = 5 stays on feature_refine_kernel<5>)
  fr_tma::FrLevel cfg;
  if (points == 1 && fr_tma_config(features, N, C, H, W, 1, &cfg)) {
    using namespace fr_tma;
    {
      const int rows = cfg.rows, stages = cfg.stages, cpc = cfg.cpc, bands = cfg.bands;
      const int stage_elems = (rows + 2 * kHalo) * W;
      const size_t smem = (size_t)stages * stage_elems * 4 + 2 * kMaxStages * sizeof(uint64_t);
      dim3 g(bands, jdet_ceil_div(C, cpc), N);
#define JDET_LAUNCH_FR_TMA(P, T)                                                                                       \
  do {                                                                                                                 \
    JDET_RETURN_IF_CUDA(cudaFuncSetAttribute(feature_refine_tma_kernel<P, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    /* without this the driver picks the smallest carve-out that fits ONE block (ncu: occupancy_limit_shared_mem = 1) */ \
    JDET_RETURN_IF_CUDA(cudaFuncSetAttribute(feature_refine_tma_kernel<P, T>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)); \
    feature_refine_tma_kernel<P, T><<<g, kConsumers + 32, smem, st>>>(features, best_rbboxes, C, H, W, spatial_scale, rows, cpc, stages, output); \
  } while (0)
      JDET_LAUNCH_FR_TMA(1, 4);
#undef JDET_LAUNCH_FR_TMA
      return (int)cudaGetLastError();
    }
  }
  const int ptiles = jdet_ceil_div(HW, 256);
  // enough CTAs to fill 148 SMs several times over, but keep slabs long enough to amortise the box decode
  int ch_per_cta = C;
  while (ch_per_cta > 16 && (long long)ptiles * jdet_ceil_div(C, ch_per_cta) * N < 148 * 8) ch_per_cta = (ch_per_cta + 1) / 2;
  dim3 grid(ptiles, jdet_ceil_div(C, ch_per_cta), N);
  // (A/B on B200, level 0 of cfg4: this 16-B-vector kernel 121 us; scalar kernel with 8 channels in flight 178 us)
  if (points == 1 && W % 4 == 0 && (((uintptr_t)features | (uintptr_t)output) & 15) == 0) {
    const int vt = jdet_ceil_div(HW / 4, 256);
    int cpc = C;
    while (cpc > 16 && (long long)vt * jdet_ceil_div(C, cpc) * N < 148 * 8) cpc = (cpc + 1) / 2;
    dim3 vgrid(vt, jdet_ceil_div(C, cpc), N);
    feature_refine_p1_vec4_kernel<<<vgrid, 256, 0, st>>>(features, best_rbboxes, C, H, W, spatial_scale, cpc, output);
  } else if (points == 1) feature_refine_kernel<1><<<grid, 256, 0, st>>>(features, best_rbboxes, C, H, W, spatial_scale, ch_per_cta, output);
  else             feature_refine_kernel<5><<<grid, 256, 0, st>>>(features, best_rbboxes, C, H, W, spatial_scale, ch_per_cta, output);
  return (int)cudaGetLastError();
}

// feature_refine on every FPN level of a head in one call (FeatureRefineModule.execute, ops/fr.py:339-346, calls FR per level):
// the levels that qualify for the TMA-staged path (points == 1, W % 4 == 0) share ONE launch; the others take the per-level
// kernels.  features / best_rbboxes / outputs / Hs / Ws / scales: HOST arrays of nlevels (<= 8) entries.
JDET_API int jdet_feature_refine_multi(const float* const* features, const float* const* best_rbboxes, int nlevels, int N, int C,
                                       const int* Hs, const int* Ws, const float* scales, int points, float* const* outputs,
                                       void* stream) {
  using namespace jdet;
  if (nlevels <= 0 || nlevels > fr_tma::kMaxLevels || N < 0 || C < 0 || !features || !best_rbboxes || !Hs || !Ws || !scales || !outputs ||
      (points != 1 && points != 5))
    return JDET_ERR_BAD_ARG;
  if (N > 65535) return JDET_ERR_UNSUPPORTED;
  fr_tma::FrLevels L{};
  L.C = C;
  long long items = 0;
  size_t smem = 0;
  for (int l = 0; l < nlevels; l++) {
    if (Hs[l] < 0 || Ws[l] < 0) return JDET_ERR_BAD_ARG;
    if ((size_t)N * C * Hs[l] * Ws[l] == 0) continue;
    if (!features[l] || !best_rbboxes[l] || !outputs[l]) return JDET_ERR_BAD_ARG;
    fr_tma::FrLevel v;
    if (points == 1 && fr_tma_config(features[l], N, C, Hs[l], Ws[l], nlevels, &v)) {
      v.feat = features[l]; v.boxes = best_rbboxes[l]; v.out = outputs[l]; v.scale = scales[l]; v.item_begin = (int)items;
      items += (long long)v.bands * v.chunks * N;
      smem = std::max(smem, (size_t)v.stages * (v.rows + 2 * fr_tma::kHalo) * v.W * 4 + 2 * fr_tma::kMaxStages * sizeof(uint64_t));
      L.lv[L.n++] = v;
    } else {
      const int e = jdet_feature_refine(features[l], best_rbboxes[l], N, C, Hs[l], Ws[l], points, scales[l], outputs[l], stream);
      if (e) return e;
    }
  }
  if (L.n == 0) return 0;
  if (items > 0x7fffffffLL) return JDET_ERR_UNSUPPORTED;
  using namespace fr_tma;
  JDET_RETURN_IF_CUDA(cudaFuncSetAttribute(feature_refine_tma_multi_kernel<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  JDET_RETURN_IF_CUDA(cudaFuncSetAttribute(feature_refine_tma_multi_kernel<1, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
  feature_refine_tma_multi_kernel<1, 4><<<(int)items, kConsumers + 32, smem, (cudaStream_t)stream>>>(L);
  return (int)cudaGetLastError();
}

// backward of jdet_feature_refine w.r.t. features: FeatureRefineFunction.grad (ops/fr.py:266-271, 242-252)
JDET_API int jdet_feature_refine_backward(const float* grad_output, const float* best_rbboxes, int N, int C, int H, int W,
                                          int points, float spatial_scale, float* grad_input, void* stream) {
  using namespace jdet;
  if (N < 0 || C < 0 || H < 0 || W < 0 || (points != 1 && points != 5)) return JDET_ERR_BAD_ARG;
  const size_t total = (size_t)N * C * H * W;
  if (total == 0) return 0;
  if (!grad_output || !best_rbboxes || !grad_input || N > 65535) return JDET_ERR_BAD_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  JDET_RETURN_IF_CUDA(cudaMemcpyAsync(grad_input, grad_output, total * sizeof(float), cudaMemcpyDeviceToDevice, st));
  const int HW = H * W;
  const int ptiles = jdet_ceil_div(HW, 256);
  int ch_per_cta = C;
  while (ch_per_cta > 16 && (long long)ptiles * jdet_ceil_div(C, ch_per_cta) * N < 148 * 8) ch_per_cta = (ch_per_cta + 1) / 2;
  dim3 grid(ptiles, jdet_ceil_div(C, ch_per_cta), N);
  if (points == 1) feature_refine_bwd_kernel<1><<<grid, 256, 0, st>>>(grad_output, best_rbboxes, C, H, W, spatial_scale, ch_per_cta, grad_input);
  else             feature_refine_bwd_kernel<5><<<grid, 256, 0, st>>>(grad_output, best_rbboxes, C, H, W, spatial_scale, ch_per_cta, grad_input);
  return (int)cudaGetLastError();
}
